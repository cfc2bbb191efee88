"""Two-GPU smoke of the CLI under torchrun: tiny checkpoint, a PDB with backbone coordinates, `--mode ddpm --mask_ids`
with the random-init encoder and decoder.  `gpurun --gpus 2 -- python tools/cli_n2_smoke.py`."""
import os
import subprocess
import sys
import tempfile
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from test_abi_and_host import write_run_dir  # noqa: E402
from test_encoder import BPTI, V, write_backbone_pdb  # noqa: E402
from oracle import esm3_ref  # noqa: E402

tiny = dict(d_model=256, n_heads=4, v_heads=64, n_layers=2)
tmp = Path(tempfile.mkdtemp())
net, emb = esm3_ref.build_reference_model(esm3_ref.Esm3Dims(**tiny), seed=5)
extra = "\n".join(f"    {k}: {v}" for k, v in tiny.items())
ckpt = write_run_dir(tmp / "run", "dir", net_extra=extra, hidden=256, module=esm3_ref.full_state_dict(net, emb))
(tmp / "targets").mkdir()
write_backbone_pdb(tmp / "targets" / "bpti.pdb", BPTI, V.synthetic_backbone(len(BPTI), seed=4))
cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
       "--master-port", "29611", "-m", "esmdiff_b200.sample_esmdiff", "--input", str(tmp / "targets"), "--ckpt", str(ckpt),
       "--output", str(tmp / "out"), "--mode", "ddpm", "--num_steps", "6", "--num_samples", "5", "--seed", "1",
       "--mask_ids", "3,4,5,6", "--encoder_ckpt", "random", "--decoder_ckpt", "random"]
r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
print(r.stdout[-2000:])
print(r.stderr[-1500:])
assert r.returncode == 0, r.returncode
files = list((tmp / "out").glob("step6_*_N5_*/bpti.pdb"))
assert len(files) == 1, files
text = files[0].read_text()
assert text.count("MODEL ") == 5 and text.rstrip().endswith("END")
print("2-GPU CLI smoke ok:", files[0], text.count("ATOM"), "atoms")
