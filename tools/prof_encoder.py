"""One encode of a 256-residue chain + one forward of the sampling network with structure coordinates (B = 100, T = 258),
for an ncu launch list of the inpainting front end's kernels:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/enc_launches.csv python tools/prof_encoder.py"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from esmdiff_b200.encoder import load_encoder  # noqa: E402
from esmdiff_b200.engine import Dims, Engine  # noqa: E402
from esmdiff_b200.synthetic import random_state_dict  # noqa: E402
from oracle.vqvae_enc_ref import synthetic_backbone  # noqa: E402

enc = load_encoder(None)
bb = synthetic_backbone(256, seed=1)[None].cuda()
enc.encode(bb)
torch.cuda.synchronize()
d = Dims(n_layers=1)
e = Engine(d)
e.load_state_dict(random_state_dict(d, device="cuda", seed=0, full=True))
B, T = 100, 258
g = torch.Generator().manual_seed(0)
seq = torch.randint(4, 24, (B, T), generator=g).cuda()
xt = torch.randint(0, 4096, (B, T), generator=g).cuda()
coords = torch.full((B, T, 3, 3), float("nan"))
coords[:, 1:-1] = synthetic_backbone(T - 2, seed=2)
coords[:, 2:34] = float("inf")
e.set_structure_coords(coords)
e.forward_sigma(seq, xt, 0.5)
e.synchronize()
print("done")
