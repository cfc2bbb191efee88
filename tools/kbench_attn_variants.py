"""Timing + parity of the resident attention kernel for alternative builds of the library (ESMDIFF_LIB)."""
import os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))
from kbench import timeit, engine, dev  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
g = torch.Generator(device="cuda").manual_seed(8)
tag = os.environ.get("ESMDIFF_LIB", "product").split("/")[-1]
e = engine()
line = f"{tag:14s}"
for (B, T, H) in [(3, 258, 24)]:
    D = H * 64
    qkv = (torch.randn(B * T, 3 * D, device=dev, generator=g) * 1.5).bfloat16()
    q, k, v = [z.view(B, T, H, 64).transpose(1, 2) for z in qkv.float().chunk(3, -1)]
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * T, D)
    got = e.op_attention(qkv, B, T, H)
    line += f" rel_fro {((got.float() - ref).norm() / ref.norm()).item():.2e} |"
for (B, T, H) in [(100, 258, 24), (100, 256, 24), (13, 258, 24), (32, 514, 24)]:
    D = H * 64
    qkv = torch.randn(B * T, 3 * D, device=dev, generator=g).bfloat16()
    sumsq = (torch.rand(B * T, 2 * D // 128, device=dev, generator=g) * 200 + 20).contiguous()
    ms = timeit(lambda: e.op_attention(qkv, B, T, H), flush=flush, n=30)
    ms2 = timeit(lambda: e.op_attention(qkv, B, T, H, qk_sumsq=sumsq), flush=flush, n=30)
    line += f" B={B} T={T}: {ms * 1e3:6.1f} / +ln {ms2 * 1e3:6.1f} us |"
print(line, flush=True)
e.close()
