"""Launches each hot kernel a few times at the benchmark shape (B=100, T=258) for an ncu capture:

    ncu --set full --import-source on --clock-control none -k regex:'attention_resident|gemm_bf16' \
        -o gpurun_out/prof python tools/prof_kernels.py

Order of launches (each kernel twice): QKV GEMM with the q/k-LN + RoPE epilogue (epilogue 8), QKV GEMM
with the plain LN-folded store (5), attention with folded 1/std, attention plain, out_proj residual +
LN by-products (6), W2 residual + LN by-products (6), W1 SwiGLU (7).
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from esmdiff_b200.engine import Dims, Engine  # noqa: E402

dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(7)
e = Engine(Dims())
B, T, D, H, F = 100, 258, 1536, 24, 4096
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
wo = (torch.randn(D, D, device=dev, generator=g) / D ** 0.5).bfloat16()
w2 = (torch.randn(D, F, device=dev, generator=g) / F ** 0.5).bfloat16()
w1 = torch.randn(2 * F, D, device=dev, generator=g) / D ** 0.5
w1f, c1, b1 = e.op_fold_layernorm(w1, ones, None, swiglu_hidden=F)
att = torch.randn(M, D, device=dev, generator=g).bfloat16()
hb = torch.randn(M, F, device=dev, generator=g).bfloat16()
xr = x.clone()
st2 = torch.zeros(M, 16, 2, device=dev)          # capacity for 96- or 128-column spans
xb2 = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
hout = torch.empty(M, F, dtype=torch.bfloat16, device=dev)
for rep in range(2):
    qkv2, sumsq = e.op_gemm_qkv_rope(xb, wfc, bsc, stats, csc, gam, T, 2 * D)
    e.op_gemm_ln(5, xb, wf, qkv, bias=bs, stats_in=stats, colsum=cs)
    e.op_attention(qkv2, B, T, H, qk_sumsq=sumsq)
    e.op_attention(qkv, B, T, H)
    e.op_gemm_ln(6, att, wo, xr, scale=1.1547, stats_out=st2, xb_out=xb2)
    e.op_gemm_ln(6, hb, w2, xr, scale=1.1547, stats_out=st2, xb_out=xb2)
    e.op_gemm_ln(7, xb, w1f, hout, bias=b1, stats_in=stats, colsum=c1)
e.synchronize()
print("done")
e.close()
