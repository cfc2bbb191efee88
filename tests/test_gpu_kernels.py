"""GPU: every kernel behind the C ABI against its reference.

* sampler kernels (integer / index work): bit-exact token ids against the reference-pinned golden
  vectors and the oracle (near-tie exemption: rows whose best/second race score differ by < 1e-5
  relative, where 1-2 ulp of libm difference may swap the argmax -- counted and bounded).
* floating-point kernels (GEMM epilogues, LayerNorm, q/k-LN + RoPE, attention): against a plain
  PyTorch fp32 reference of the same op on the same bf16-rounded operands.  Tolerances: fp32
  outputs 2e-5 rel-Frobenius (fp32 accumulation-order noise), bf16 outputs 4e-3 rel-Frobenius and
  1.2e-2 of the tensor's max per element (bf16 has 8 mantissa bits: 2^-9 = 2e-3 rounding).
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import mdlm_ref
from oracle.esm3_ref import apply_rotary, rotary_tables

pytestmark = pytest.mark.gpu
MASK = 4096
DEV = "cuda"


def rel_fro(got, ref):
    got, ref = got.float(), ref.float()
    return float((got - ref).norm() / ref.norm().clamp_min(1e-20))


def check(got, ref, fro, elem=None):
    assert not bool(torch.isnan(got.float()).any())
    r = rel_fro(got, ref)
    assert r < fro, f"rel_fro {r:.3e} >= {fro}"
    if elem is not None:
        worst = float((got.float() - ref.float()).abs().max() / ref.float().abs().max())
        assert worst < elem, f"max elementwise error {worst:.3e} of max >= {elem}"


# ---------------------------------------------------------------------------------------------
# GEMM
# ---------------------------------------------------------------------------------------------
GEMM_SHAPES = [(128, 256, 64), (77, 384, 128), (256, 512, 1536), (1000, 4608, 1536), (300, 1536, 4096),
               (1, 256, 64), (129, 264, 192), (16254, 1536, 1536)]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_store_bf16(engine, M, N, K):
    g = torch.Generator(device=DEV).manual_seed(M * 7 + N)
    a = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    w = (torch.randn(N, K, device=DEV, generator=g) / K ** 0.5).bfloat16()
    out = torch.full((M, N), 7.0, dtype=torch.bfloat16, device=DEV)
    engine.op_gemm(0, a, w, out)
    engine.synchronize()
    check(out, a.float() @ w.float().T, 4e-3, 1.2e-2)


def test_gemm_row_col_k_mapping(engine):
    """Structured operands: any mistake in the swizzle / descriptor / TMEM lane mapping shows up
    as an exact mismatch (values are small integers, exactly representable)."""
    M, N, K = 256, 512, 128
    for k0 in (0, 7, 8, 16, 33, 63, 64, 127):
        a = torch.zeros(M, K, device=DEV); a[:, k0] = (torch.arange(M, device=DEV) % 17).float()      # products <= 96: exact in bf16
        w = torch.zeros(N, K, device=DEV); w[:, k0] = (torch.arange(N, device=DEV) % 13).float() - 6
        out = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
        engine.op_gemm(0, a.bfloat16(), w.bfloat16(), out)
        engine.synchronize()
        assert torch.equal(out.float(), a @ w.T), k0


def test_gemm_epilogues(engine):
    g = torch.Generator(device=DEV).manual_seed(5)
    M, N, K = 1000, 1536, 1536
    a = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    w = (torch.randn(N, K, device=DEV, generator=g) / K ** 0.5).bfloat16()
    ref = a.float() @ w.float().T
    x0 = torch.randn(M, N, device=DEV, generator=g) * 30
    x = x0.clone()
    engine.op_gemm(1, a, w, x, scale=1.1547005)
    engine.synchronize()
    check(x, x0 + ref / 1.1547005, 2e-5)                       # residual add, fp32 stream
    bias = torch.randn(N, device=DEV, generator=g)
    out = torch.empty(M, N, device=DEV)
    engine.op_gemm(3, a, w, out, bias=bias)
    engine.synchronize()
    check(out, F.gelu(ref + bias), 2e-5)                       # bias + exact-erf GELU
    Nv = 4101                                                  # ragged N: TMA zero-fill + masked store
    w3 = (torch.randn(Nv, K, device=DEV, generator=g) / K ** 0.5).bfloat16()
    b3 = torch.randn(Nv, device=DEV, generator=g)
    out = torch.full((M, Nv + 3), -5.0, device=DEV)[:, :Nv]    # row stride != N, guard columns
    engine.op_gemm(4, a, w3, out, bias=b3)
    engine.synchronize()
    check(out, a.float() @ w3.float().T + b3, 2e-5)
    assert bool((out.as_strided((M, 3), (Nv + 3, 1), Nv) == -5.0).all()), "wrote past column N"
    Fh = 4096                                                  # SwiGLU with interleaved W1 rows
    w1 = torch.randn(2 * Fh, K, device=DEV, generator=g) / K ** 0.5
    w1i = engine.op_convert_bf16(w1.contiguous(), swiglu_hidden=Fh)
    out = torch.empty(M, Fh, dtype=torch.bfloat16, device=DEV)
    engine.op_gemm(2, a, w1i, out)
    engine.synchronize()
    z = a.float() @ w1.bfloat16().float().T
    check(out, F.silu(z[:, :Fh]) * z[:, Fh:], 4e-3, 1.2e-2)


def test_gemm_linearity_full_size(engine):
    """Size-independent property at the BASELINE config-2 shape (M = 63*258): the residual
    epilogue is linear, so applying it twice from x0 equals x0 + 2*acc/scale."""
    g = torch.Generator(device=DEV).manual_seed(6)
    M, N, K = 63 * 258, 1536, 4096
    a = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    w = (torch.randn(N, K, device=DEV, generator=g) / K ** 0.5).bfloat16()
    x = torch.zeros(M, N, device=DEV)
    engine.op_gemm(1, a, w, x, scale=2.0)
    engine.synchronize()
    once = x.clone()
    engine.op_gemm(1, a, w, x, scale=2.0)
    engine.synchronize()
    assert torch.equal(x, once + once)                          # deterministic and exactly linear
    check(once, (a.float() @ w.float().T) / 2.0, 2e-5)


def _combine_stats(stats, D, span=128):
    """(mean, M2) partials over `span`-column spans (densely packed, D / span per row) -> row mean,
    biased variance (Chan et al.)."""
    M = stats.shape[0]
    st = stats.reshape(-1)[: M * (D // span) * 2].reshape(M, D // span, 2)      # dense: D / span entries per row
    mean_i, m2_i = st[..., 0].double(), st[..., 1].double()
    mean = mean_i.mean(-1)
    m2 = m2_i.sum(-1) + float(span) * ((mean_i - mean[:, None]) ** 2).sum(-1)
    return mean, m2 / D


@pytest.mark.parametrize("M,D,F_h", [(1000, 1536, 4096), (130, 256, 768), (16254, 1536, 4096)])
def test_gemm_layernorm_folded_epilogues(engine, M, D, F_h):
    """The block pre-LayerNorms folded through the GEMMs (gemm.cuh epilogues 5/6/7) against
    torch: residual update + bf16 copy + partial statistics, then LN(x) @ W^T with gamma folded
    into W, non-trivial gamma/beta, a row offset and an outlier channel (real ESM3 streams carry
    |x| ~ 1e4 in single channels)."""
    g = torch.Generator(device=DEV).manual_seed(11)
    a = torch.randn(M, D, device=DEV, generator=g).bfloat16()
    wo = (torch.randn(D, D, device=DEV, generator=g) / D ** 0.5).bfloat16()
    x0 = torch.randn(M, D, device=DEV, generator=g) * 3 + 0.7
    x0[:, 77] += 900.0                                           # outlier channel
    x0[::7] -= 2.5                                               # row-dependent mean
    x = x0.clone()
    xb = torch.empty(M, D, dtype=torch.bfloat16, device=DEV)
    stats = torch.zeros(M, (D + 95) // 96, 2, device=DEV)         # capacity for either span (include/esmdiff_b200.h)
    engine.op_gemm_ln(6, a, wo, x, scale=1.1547005, stats_out=stats, xb_out=xb)
    engine.synchronize()
    ref_x = x0 + (a.float() @ wo.float().T) / 1.1547005
    check(x, ref_x, 2e-5)
    assert torch.equal(xb, x.bfloat16())                         # the copy is the rounded new stream
    mean, var = _combine_stats(stats, D, engine.stats_span)
    assert float((mean - x.double().mean(-1)).abs().max()) < 1e-4
    ref_var = x.double().var(-1, unbiased=False)
    assert float(((var - ref_var).abs() / ref_var).max()) < 1e-5

    gamma = 1.0 + 0.3 * torch.randn(D, device=DEV, generator=g)
    beta = 0.2 * torch.randn(D, device=DEV, generator=g)
    wq = torch.randn(3 * D, D, device=DEV, generator=g) / D ** 0.5
    ln = F.layer_norm(x, (D,), gamma, beta, 1e-5)
    wq_f, cq, bq = engine.op_fold_layernorm(wq, gamma, beta)
    assert torch.equal(wq_f, (wq * gamma).bfloat16())
    qkv = torch.empty(M, 3 * D, dtype=torch.bfloat16, device=DEV)
    engine.op_gemm_ln(5, xb, wq_f, qkv, bias=bq, stats_in=stats, colsum=cq)
    engine.synchronize()
    check(qkv, ln @ wq.T, 6e-3, 3e-2)                            # bf16 operands AND bf16 output

    w1 = torch.randn(2 * F_h, D, device=DEV, generator=g) / D ** 0.5
    w1_f, c1, b1 = engine.op_fold_layernorm(w1, gamma, beta, swiglu_hidden=F_h)
    h = torch.empty(M, F_h, dtype=torch.bfloat16, device=DEV)
    engine.op_gemm_ln(7, xb, w1_f, h, bias=b1, stats_in=stats, colsum=c1)
    engine.synchronize()
    z = ln @ w1.T
    check(h, F.silu(z[:, :F_h]) * z[:, F_h:], 8e-3, 3e-2)
    # against the stand-alone LayerNorm kernel + plain epilogue on the same inputs (the two product
    # variants, ESMDIFF_LN=separate): same operands up to where the bf16 rounding is applied
    xn = engine.op_layernorm(x, gamma, beta)
    qkv2 = torch.empty_like(qkv)
    engine.op_gemm(0, xn, wq.bfloat16(), qkv2)
    engine.synchronize()
    assert rel_fro(qkv, ln @ wq.T) < 1.5 * rel_fro(qkv2, ln @ wq.T) + 1e-4


@pytest.mark.parametrize("M", [3354, 1000, 130, 25800])
def test_residual_gemm_192_wide_tiles(M):
    """The residual GEMMs can run 192-wide tiles (N = 1536 as 8 column tiles: 84 -> 112 tiles of 3/4 the
    length at 13 samples, 2 waves either way); the statistics they leave are then per 96 columns and the
    LayerNorm-folded consumers combine 16 spans instead of 12.  Same checks as the 256-wide form, plus
    bit-equality of the updated stream between the two tile widths (same MMA order along K)."""
    import os
    from esmdiff_b200.engine import Dims, Engine
    engs = {}
    for bn in ("256", "192"):
        os.environ["ESMDIFF_RESID_BN"] = bn
        try:
            engs[bn] = Engine(Dims())
        finally:
            os.environ.pop("ESMDIFF_RESID_BN")
    g = torch.Generator(device=DEV).manual_seed(21)
    D, F_h = 1536, 4096
    a = torch.randn(M, D, device=DEV, generator=g).bfloat16()
    wo = (torch.randn(D, D, device=DEV, generator=g) / D ** 0.5).bfloat16()
    h = torch.randn(M, F_h, device=DEV, generator=g).bfloat16()
    w2 = (torch.randn(D, F_h, device=DEV, generator=g) / F_h ** 0.5).bfloat16()
    x0 = torch.randn(M, D, device=DEV, generator=g) * 3 + 0.7
    x0[:, 77] += 900.0
    gamma = 1.0 + 0.3 * torch.randn(D, device=DEV, generator=g)
    beta = 0.2 * torch.randn(D, device=DEV, generator=g)
    wq = torch.randn(3 * D, D, device=DEV, generator=g) / D ** 0.5
    res = {}
    for bn, eng in engs.items():
        x = x0.clone()
        xb = torch.empty(M, D, dtype=torch.bfloat16, device=DEV)
        stats = torch.zeros(M, 16, 2, device=DEV)
        eng.op_gemm_ln(6, a, wo, x, scale=1.1547005, stats_out=stats, xb_out=xb)        # out_proj: K = 1536
        eng.op_gemm_ln(6, h, w2, x, scale=1.1547005, stats_out=stats, xb_out=xb)        # W2: K = 4096
        eng.synchronize()
        ref_x = x0 + (a.float() @ wo.float().T) / 1.1547005 + (h.float() @ w2.float().T) / 1.1547005
        check(x, ref_x, 2e-5)
        assert torch.equal(xb, x.bfloat16())
        assert eng.stats_span == (96 if bn == "192" else 128)
        mean, var = _combine_stats(stats, D, eng.stats_span)
        assert float((mean - x.double().mean(-1)).abs().max()) < 1e-4
        ref_var = x.double().var(-1, unbiased=False)
        assert float(((var - ref_var).abs() / ref_var).max()) < 1e-5
        wq_f, cq, bq = eng.op_fold_layernorm(wq, gamma, beta)
        qkv = torch.empty(M, 3 * D, dtype=torch.bfloat16, device=DEV)
        eng.op_gemm_ln(5, xb, wq_f, qkv, bias=bq, stats_in=stats, colsum=cq)             # consumer reads the producer's spans
        plain = x0.clone()
        eng.op_gemm(1, a, wo, plain, scale=1.1547005)                                   # residual epilogue without statistics
        eng.synchronize()
        check(qkv, F.layer_norm(x, (D,), gamma, beta, 1e-5) @ wq.T, 6e-3, 3e-2)
        res[bn] = (x, qkv, plain)
    assert torch.equal(res["256"][0], res["192"][0]) and torch.equal(res["256"][2], res["192"][2])
    assert rel_fro(res["192"][1], res["256"][1]) < 3e-3
    for eng in engs.values():
        eng.close()


def test_residual_epilogue_statistics_with_large_row_offset(engine):
    """The residual epilogue accumulates plain (sum, sum of squares) per 128-column span; the claimed
    loss is ~1e-7 (1 + (mean/std)^2) relative.  Pin it at mean/std = 50 (far beyond a LayerNorm
    input): variance within 1e-3, i.e. rstd within 5e-4, below the bf16 rounding of the operands."""
    g = torch.Generator(device=DEV).manual_seed(12)
    M, D = 777, 1536
    a = torch.zeros(M, D, device=DEV).bfloat16()                 # acc = 0: the statistics are those of x0
    wo = torch.zeros(D, D, device=DEV).bfloat16()
    x0 = 50.0 + torch.randn(M, D, device=DEV, generator=g)
    x = x0.clone()
    xb = torch.empty(M, D, dtype=torch.bfloat16, device=DEV)
    stats = torch.zeros(M, (D + 95) // 96, 2, device=DEV)
    engine.op_gemm_ln(6, a, wo, x, scale=1.0, stats_out=stats, xb_out=xb)
    engine.synchronize()
    assert torch.equal(x, x0)
    mean, var = _combine_stats(stats, D, engine.stats_span)
    ref_var = x0.double().var(-1, unbiased=False)
    assert float((mean - x0.double().mean(-1)).abs().max()) < 1e-4
    assert float(((var - ref_var).abs() / ref_var).max()) < 1e-3


# ---------------------------------------------------------------------------------------------
# row kernels
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("D", [1536, 256])
def test_layernorm(engine, D):
    g = torch.Generator(device=DEV).manual_seed(1)
    x = torch.randn(999, D, device=DEV, generator=g) * 3 + 0.5
    x[5, 7] = 12640.0                                           # ESM3-scale outlier
    w = torch.randn(D, device=DEV, generator=g)
    b = torch.randn(D, device=DEV, generator=g)
    check(engine.op_layernorm(x, w, b), F.layer_norm(x, (D,), w, b), 4e-3, 1.2e-2)
    check(engine.op_layernorm(x, w, None), F.layer_norm(x, (D,), w, None), 4e-3, 1.2e-2)


@pytest.mark.parametrize("B,T,D", [(3, 60, 1536), (2, 258, 256), (1, 1026, 512)])
def test_qk_norm_rope(engine, B, T, D):
    g = torch.Generator(device=DEV).manual_seed(2)
    H = D // 64
    qkv = torch.randn(B * T, 3 * D, device=DEV, generator=g).bfloat16()
    qw = torch.randn(D, device=DEV, generator=g)
    kw = torch.randn(D, device=DEV, generator=g)
    q, k, v = qkv.float().chunk(3, -1)
    cos, sin = (z.to(DEV) for z in rotary_tables(T))
    qr = apply_rotary(F.layer_norm(q, (D,), qw).view(B, T, H, 64), cos, sin).reshape(B * T, D)
    kr = apply_rotary(F.layer_norm(k, (D,), kw).view(B, T, H, 64), cos, sin).reshape(B * T, D)
    got = engine.op_qk_norm_rope(qkv.clone(), qw, kw, B, T)
    engine.synchronize()
    check(got[:, :D], qr, 4e-3, 1.5e-2)
    check(got[:, D:2 * D], kr, 4e-3, 1.5e-2)
    assert torch.equal(got[:, 2 * D:].float(), v)               # v third untouched


@pytest.mark.parametrize("B,T,H", [(1, 64, 1), (1, 128, 1), (2, 60, 4), (2, 130, 4), (3, 258, 24),
                                   (1, 514, 4), (1, 1026, 2), (5, 1, 2), (2, 129, 3)])
def test_attention(engine, B, T, H):
    g = torch.Generator(device=DEV).manual_seed(3)
    D = H * 64
    qkv = (torch.randn(B * T, 3 * D, device=DEV, generator=g) * 1.5).bfloat16()
    q, k, v = [z.view(B, T, H, 64).transpose(1, 2) for z in qkv.float().chunk(3, -1)]
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * T, D)
    got = engine.op_attention(qkv, B, T, H)
    engine.synchronize()
    check(got, ref, 5e-3, 2e-2)


@pytest.mark.parametrize("B,T,D", [(3, 60, 1536), (2, 258, 256), (5, 130, 512), (1, 1026, 256), (63, 258, 1536)])
def test_qkv_epilogue_with_qk_layernorm_and_rope(engine, B, T, D):
    """gemm.cuh epilogue 8 + the attention kernels' 1/std factors == LayerNorm(x) -> Linear ->
    q_ln / k_ln -> RoPE -> SDPA in torch fp32 (esm MultiHeadAttention), with non-trivial gamma /
    beta everywhere.  Checked in two stages: (1) the epilogue's bf16 q', k' times the rstd recovered
    from its own partial sums against torch's normalised + rotated q, k (and v, and the partial
    sums themselves against torch); (2) attention on that layout against SDPA."""
    g = torch.Generator(device=DEV).manual_seed(13 + T)
    M, H = B * T, D // 64
    x = torch.randn(M, D, device=DEV, generator=g) * 2 + 0.3
    x[:, 5] += 40.0
    gamma = 1.0 + 0.3 * torch.randn(D, device=DEV, generator=g)
    beta = 0.2 * torch.randn(D, device=DEV, generator=g)
    qw = 1.0 + 0.3 * torch.randn(D, device=DEV, generator=g)
    kw = 1.0 + 0.3 * torch.randn(D, device=DEV, generator=g)
    w = torch.randn(3 * D, D, device=DEV, generator=g) / D ** 0.5
    w[:D] += 0.02                                                   # q rows with a clear column mean
    # statistics + bf16 copy of x as the producers leave them (per 128-column span: mean, M2)
    xs = x.view(M, D // 128, 128)
    stats = torch.stack([xs.mean(-1), ((xs - xs.mean(-1, keepdim=True)) ** 2).sum(-1)], -1).contiguous()
    xb = x.bfloat16()
    wf, cs, bs = engine.op_fold_layernorm(w, gamma, beta, center_rows=2 * D, center_block=D)
    wc = w.clone()
    wc[:D] -= wc[:D].mean(0, keepdim=True)
    wc[D:2 * D] -= wc[D:2 * D].mean(0, keepdim=True)
    pair = lambda z: torch.cat([z[:2 * D].view(2 * H, 2, 32, D).transpose(1, 2).reshape(2 * D, D), z[2 * D:]])
    assert rel_fro(wf, pair(wc * gamma)) < 3e-3                     # centred, gamma folded, partners interleaved
    gam = torch.cat([qw, kw]).view(2 * H, 2, 32).transpose(1, 2).reshape(2 * D).contiguous()
    qkv, sumsq = engine.op_gemm_qkv_rope(xb, wf, bs, stats, cs, gam, T, 2 * D)
    engine.synchronize()
    # torch reference, fp32 throughout
    y = F.layer_norm(x, (D,), gamma, beta, 1e-5) @ w.T
    q, k, v = y.chunk(3, -1)
    cos, sin = (z.to(DEV) for z in rotary_tables(T))
    qr = apply_rotary(F.layer_norm(q, (D,), qw).view(B, T, H, 64), cos, sin).reshape(M, D)
    kr = apply_rotary(F.layer_norm(k, (D,), kw).view(B, T, H, 64), cos, sin).reshape(M, D)
    ssq_q = ((q - q.mean(-1, keepdim=True)) ** 2).sum(-1)
    ssq_k = ((k - k.mean(-1, keepdim=True)) ** 2).sum(-1)
    nsp = D // 128
    got_q, got_k = sumsq[:, :nsp].sum(-1), sumsq[:, nsp:].sum(-1)
    # bf16 operands: each sum of squares carries the operand rounding (~2^-9 / sqrt(D) per term, coherent part small)
    assert float(((got_q - ssq_q).abs() / ssq_q).max()) < 5e-3
    assert float(((got_k - ssq_k).abs() / ssq_k).max()) < 5e-3
    rq = torch.rsqrt(got_q / D + 1e-5)[:, None]
    rk = torch.rsqrt(got_k / D + 1e-5)[:, None]
    # q', k' are stored with the rotary partners (d, d + 32) of every head adjacent: undo for the comparison
    unpair = lambda z: z.view(M, H, 32, 2).transpose(-1, -2).reshape(M, D)
    check(unpair(qkv[:, :D].float() * rq), qr, 6e-3, 3e-2)
    check(unpair(qkv[:, D:2 * D].float() * rk), kr, 6e-3, 3e-2)
    check(qkv[:, 2 * D:], v, 6e-3, 3e-2)
    att = engine.op_attention(qkv, B, T, H, qk_sumsq=sumsq)
    engine.synchronize()
    ref = F.scaled_dot_product_attention(qr.view(B, T, H, 64).transpose(1, 2), kr.view(B, T, H, 64).transpose(1, 2),
                                         v.view(B, T, H, 64).transpose(1, 2)).transpose(1, 2).reshape(M, D)
    check(att, ref, 1e-2, 4e-2)
    # the same attention from the epilogue's OWN rounded operands: isolates the attention kernel
    qn = (qkv[:, :D].float() * rq).view(B, T, H, 64).transpose(1, 2)
    kn = (qkv[:, D:2 * D].float() * rk).bfloat16().float().view(B, T, H, 64).transpose(1, 2)
    ref2 = F.scaled_dot_product_attention(qn, kn, qkv[:, 2 * D:].float().view(B, T, H, 64).transpose(1, 2))
    check(att, ref2.transpose(1, 2).reshape(M, D), 6e-3, 2.5e-2)


@pytest.mark.parametrize("B,T,H", [(13, 258, 24), (2, 514, 4), (1, 642, 2), (3, 130, 4)])
def test_attention_query_range_split_is_bit_identical(B, T, H):
    """Two CTAs per (sample, head), each walking half of the query tiles (chosen per launch when the grid
    is just over a multiple of the resident CTA count), must give exactly the one-CTA result: a query
    row's arithmetic does not depend on which CTA owns it."""
    import os
    from esmdiff_b200.engine import Dims, Engine
    g = torch.Generator(device=DEV).manual_seed(31)
    D = H * 64
    qkv = (torch.randn(B * T, 3 * D, device=DEV, generator=g) * 1.5).bfloat16()
    sumsq = (torch.rand(B * T, 2 * D // 128, device=DEV, generator=g) + 0.5) * 128
    outs = []
    for mode in ("1", "2"):
        os.environ["ESMDIFF_ATTN_QSPLIT"] = mode
        try:
            eng = Engine(Dims(d_model=256, n_heads=4, v_heads=8, n_layers=1))
        finally:
            os.environ.pop("ESMDIFF_ATTN_QSPLIT")
        outs.append((eng.op_attention(qkv, B, T, H), eng.op_attention(qkv, B, T, H, qk_sumsq=sumsq)))
        eng.synchronize()
        eng.close()
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    q, k, v = [z.view(B, T, H, 64).transpose(1, 2) for z in qkv.float().chunk(3, -1)]
    check(outs[1][0], F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * T, D), 5e-3, 2e-2)


def test_two_contexts_on_one_device_share_function_attributes(engine):
    """The dynamic shared-memory limit of a kernel belongs to (function, device): a second context that
    needs less must not lower it under the first (seen in round 2: 'invalid argument' at T = 386 after
    another engine had launched T = 258), and a context must not assume another one raised it."""
    from esmdiff_b200.engine import Dims, Engine
    g = torch.Generator(device=DEV).manual_seed(9)
    H, D = 2, 128
    big = (torch.randn(514, 3 * D, device=DEV, generator=g)).bfloat16()
    small = (torch.randn(60, 3 * D, device=DEV, generator=g)).bfloat16()
    a = engine.op_attention(big, 1, 514, H)
    other = Engine(Dims(d_model=256, n_heads=4, v_heads=8, n_layers=1))
    other.op_attention(small, 1, 60, H)
    other.synchronize()
    b = engine.op_attention(big, 1, 514, H)                     # larger request again, from the first context
    c = other.op_attention(big, 1, 514, H)                      # and from the one that only ever asked for less
    engine.synchronize()
    other.synchronize()
    assert torch.equal(a, b) and torch.equal(a, c)
    other.close()


def test_attention_permutation_property_full_size(engine):
    """softmax attention is equivariant to permuting the keys/values of a sample: at the config-2
    shape (B=63,T=258,H=24), shuffling the kv rows of every sample leaves the output unchanged
    up to summation order."""
    g = torch.Generator(device=DEV).manual_seed(4)
    B, T, H = 63, 258, 24
    D = H * 64
    qkv = torch.randn(B, T, 3 * D, device=DEV, generator=g).bfloat16()
    perm = torch.randperm(T, device=DEV, generator=g)
    shuffled = qkv.clone()
    shuffled[:, :, D:] = qkv[:, perm, D:]
    a = engine.op_attention(qkv.view(B * T, -1), B, T, H)
    b = engine.op_attention(shuffled.view(B * T, -1).contiguous(), B, T, H)
    engine.synchronize()
    check(b, a, 4e-3, 2e-2)


# ---------------------------------------------------------------------------------------------
# sampler: bit-exact vs the reference-pinned golden vectors
# ---------------------------------------------------------------------------------------------
def test_sampler_full_vector_bit_exact(engine, golden_dir):
    g = np.load(golden_dir / "sampler_full.npz")
    logits, u, x = (torch.from_numpy(g[k]).to(DEV) for k in ("logits", "u", "x_t"))
    got = engine.sample_step(x.clone(), logits.contiguous(), u.contiguous(), float(g["mc_t"]), float(g["mc_s"]))
    engine.synchronize()
    assert np.array_equal(got.cpu().numpy(), g["x_next"])
    lp = engine.logits_parameterization(logits.clone(), x)
    engine.synchronize()
    want = g["logp_masked_rows"]
    np.testing.assert_allclose(lp[x == MASK].cpu().numpy()[:, ::97], want, rtol=0, atol=4e-6)


def test_sampler_seeded_vectors(engine, golden_dir):
    from golden_cases import seeded_cases
    n = excused = rows = 0
    for c, logits, u, x in seeded_cases(golden_dir):
        B = x.shape[0]
        got = engine.sample_step(x.clone().to(DEV), logits.to(DEV), u.to(DEV), float(c["mc_t"]), float(c["mc_s"]))
        engine.synchronize()
        mct = torch.full((B, 1, 1), float(c["mc_t"]))
        mcs = torch.full((B, 1, 1), float(c["mc_s"]))
        lp = mdlm_ref.logits_parameterization(logits.clone(), x)
        excused += mdlm_ref.assert_ids_match(got.cpu(), torch.from_numpy(c["x_next"]), lp, mct, mcs, u)
        rows += x.numel()
        # logsumexp of the fused kernel against the reference's
        lpg = engine.logits_parameterization(logits.clone().to(DEV), x.to(DEV))
        engine.synchronize()
        m = x == MASK
        if bool(m.any()):
            np.testing.assert_allclose(lpg.cpu()[m][:, :4096].numpy(), lp[m][:, :4096].numpy(), rtol=1e-6, atol=3e-5)
        k = ~m
        if bool(k.any()):       # unmasked rows: -1e6 everywhere, 0 at x
            rowsk = lpg.cpu()[k]
            assert bool((rowsk.gather(1, x[k][:, None]) == 0).all())
            assert float(rowsk.sum(-1).max()) == pytest.approx(-1e6 * 4100, rel=1e-6)
        n += 1
    assert n == 6 and excused <= max(1, rows // 2000), excused


def test_sampler_denoise_argmax_and_identity(engine):
    g = torch.Generator().manual_seed(8)
    B, T = 3, 50
    logits = torch.randn(B, T, 4101, generator=g) * 2
    logits[..., MASK] = 50.0                                    # mask logit must never win (-1e6)
    x = torch.randint(0, 4096, (B, T), generator=g)
    x[torch.rand(B, T, generator=g) < 0.5] = MASK
    lp = mdlm_ref.logits_parameterization(logits.clone(), x)
    got = engine.denoise_argmax(x.clone().to(DEV), logits.to(DEV))
    engine.synchronize()
    assert torch.equal(got.cpu(), lp.argmax(-1))
    assert int((got == MASK).sum()) == 0
    # nothing masked -> identity, logits never read
    xu = torch.randint(0, 4096, (B, T), generator=g).to(DEV)
    u = torch.rand(B, T, 4101, generator=g).to(DEV)
    assert torch.equal(engine.sample_step(xu.clone(), logits.to(DEV), u, 0.5, 0.4), xu)


def test_sampler_philox_statistics_full_size(engine):
    """Library Philox stream at the config-2 shape: the chance that a masked row stays masked is
    mc_s/mc_t (model.py:602-603) and, when it unmasks, ids follow softmax(logits)."""
    B, T = 63, 258
    g = torch.Generator(device=DEV).manual_seed(9)
    logits = torch.zeros(B, T, 4101, device=DEV)
    logits[..., :4] = torch.tensor([3.0, 2.0, 1.0, 0.0], device=DEV) + 6.0
    x = torch.full((B, T), MASK, device=DEV)
    mc_t, mc_s = 0.8, 0.6
    got = engine.sample_step(x.clone(), logits, None, mc_t, mc_s, seed=11, step=3)
    engine.synchronize()
    stay = float((got == MASK).float().mean())
    assert abs(stay - mc_s / mc_t) < 0.015, stay
    p = torch.softmax(torch.cat([logits[0, 0, :MASK], logits[0, 0, MASK + 1:]]), -1)
    un = got[got != MASK]
    for i in range(3):
        assert abs(float((un == i).float().mean()) - float(p[i])) < 0.02
    # deterministic in (seed, step); different step -> different draw
    again = engine.sample_step(x.clone(), logits, None, mc_t, mc_s, seed=11, step=3)
    other = engine.sample_step(x.clone(), logits, None, mc_t, mc_s, seed=11, step=4)
    engine.synchronize()
    assert torch.equal(again, got) and not torch.equal(other, got)
    assert int(got.min()) >= 0 and int(got.max()) <= 4100


def test_embedding_index_error(tiny_pair):
    _, _, eng = tiny_pair
    seq = torch.tensor([[0, 5, 70, 2]])                         # 70 >= sequence vocab 64
    xt = torch.full((1, 4), MASK)
    with pytest.raises(IndexError):
        eng.forward(seq, xt)
        eng.synchronize()
