"""Timing of the residual GEMMs with the LayerNorm by-products (epilogue 6: x += acc / s, bf16 copy, statistics) for
alternative builds of the library (ESMDIFF_LIB)."""
import os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))
from kbench import timeit, engine, dev  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
g = torch.Generator(device="cuda").manual_seed(7)
tag = os.environ.get("ESMDIFF_LIB", "product").split("/")[-1]
e = engine()
D, F = 1536, 4096
line = f"{tag:16s}"
for M in (25800, 3354):
    x = torch.randn(M, D, device=dev, generator=g)
    xb = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
    stats = torch.zeros(M, 16, 2, device=dev)
    att = torch.randn(M, D, device=dev, generator=g).bfloat16()
    hb = torch.randn(M, F, device=dev, generator=g).bfloat16()
    wo = (torch.randn(D, D, device=dev, generator=g) / D ** 0.5).bfloat16()
    w2 = (torch.randn(D, F, device=dev, generator=g) / F ** 0.5).bfloat16()
    t_o = timeit(lambda: e.op_gemm_ln(6, att, wo, x, scale=1.1547, stats_out=stats, xb_out=xb), flush=flush, n=30)
    t_2 = timeit(lambda: e.op_gemm_ln(6, hb, w2, x, scale=1.1547, stats_out=stats, xb_out=xb), flush=flush, n=30)
    line += f" M={M}: out_proj {t_o * 1e3:6.1f} us ({2.0 * M * D * D / t_o / 1e9:5.0f} TF)  w2 {t_2 * 1e3:6.1f} us ({2.0 * M * D * F / t_2 / 1e9:5.0f} TF) |"
print(line, flush=True)
e.close()
