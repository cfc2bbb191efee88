"""What follows "Sampling token time" at config 2 (100 samples, L = 256): batched structure decode + PDB writing.
`gpurun -- python tools/decode_timing.py`"""
import sys, tempfile, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from esmdiff_b200.decoder import decode_to_pdb, load_decoder, pdb_models_text  # noqa: E402

dec = load_decoder(None)
g = torch.Generator().manual_seed(0)
N, L = 100, 256
tok = torch.randint(0, 4096, (N, L), generator=g)
seq = "".join("ACDEFGHIKLMNPQRSTVWY"[i % 20] for i in range(L))
out = Path(tempfile.mkdtemp()) / "x.pdb"
for it in range(3):
    torch.cuda.synchronize()
    t0 = time.time()
    bb, pl = decode_to_pdb(dec, tok, seq, out)
    t1 = time.time()
    print(f"decode_to_pdb N={N} L={L}: {t1 - t0:.3f} s total ({out.stat().st_size / 1e6:.1f} MB file)")
