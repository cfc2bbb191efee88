"""Timing of the QKV GEMM (LN-folded store epilogue 5 vs the q/k-LN + RoPE epilogue 8) for alternative builds of the
library (ESMDIFF_LIB), e.g. with the TMA ring depth capped (-DESMDIFF_STAGES_CAP=n)."""
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
D = 1536
line = f"{tag:16s}"
for (B, T) in [(100, 258), (13, 258)]:
    M = B * T
    x = torch.randn(M, D, device=dev, generator=g)
    xs = x.view(M, D // 128, 128)
    stats = torch.stack([xs.mean(-1), ((xs - xs.mean(-1, keepdim=True)) ** 2).sum(-1)], -1).contiguous()
    xb = x.bfloat16()
    w = torch.randn(3 * D, D, device=dev, generator=g) / D ** 0.5
    ones = torch.ones(D, device=dev)
    wf, cs, bs = e.op_fold_layernorm(w, ones, None)
    wfc, csc, bsc = e.op_fold_layernorm(w, ones, None, center_rows=2 * D, center_block=D)
    qkv = torch.empty(M, 3 * D, dtype=torch.bfloat16, device=dev)
    gam = torch.ones(2 * D, device=dev)
    t5 = timeit(lambda: e.op_gemm_ln(5, xb, wf, qkv, bias=bs, stats_in=stats, colsum=cs), flush=flush, n=30)
    t8 = timeit(lambda: e.op_gemm_qkv_rope(xb, wfc, bsc, stats, csc, gam, T, 2 * D), flush=flush, n=30)
    w1 = (torch.randn(8192, D, device=dev, generator=g) / D ** 0.5)
    w1f, c1, b1 = e.op_fold_layernorm(w1, ones, ones, swiglu_hidden=4096)
    h = torch.empty(M, 4096, dtype=torch.bfloat16, device=dev)
    t7 = timeit(lambda: e.op_gemm_ln(7, xb, w1f, h, bias=b1, stats_in=stats, colsum=c1), flush=flush, n=30)
    line += f" B={B}: plain {t5 * 1e3:6.1f} rope {t8 * 1e3:6.1f} w1 {t7 * 1e3:6.1f} us |"
print(line, flush=True)
e.close()
