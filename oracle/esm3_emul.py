"""bf16-emulating mode of the network oracle (ORACLE -- tests only).

``oracle/esm3_ref.py`` restates the reference network in fp32 (what the reference itself computes:
checkpoint_utils.py:59-72 never casts).  The CUDA path computes the same network with bf16
tensor-core operands and fp32 accumulation, so it differs from the fp32 oracle by bf16 rounding
noise (measured 4e-3 relative on logits after 48 blocks) -- which says nothing about whether a
KERNEL is wrong.  This module evaluates the same weights with a rounding to bf16 at exactly the
points where the product rounds, and nowhere else, so that what is left
between it and the CUDA result is accumulation order, approximate exp2/rcp/rsqrt (<= 2 ulp fp32)
and the bf16 rounding flips those cause.  Every rounding point has a switch: with all of them off
the result equals ``esm3_ref`` up to fp32 rounding (asserted in tests/test_oracle_golden.py), and
switching them on one at a time gives the error budget quoted in DESIGN.md.

Rounding points of the product (kernel, file):
  weights     GEMM weights bf16; LayerNorm gamma folded in first, q/k rows centred   (elementwise.cuh fold kernel)
  act         A operands: bf16 copy of the raw residual stream, attention output, SwiGLU output,
              final-norm / head LayerNorm outputs                                      (gemm.cuh epilogues, layernorm kernel)
  qkv         q' = rope(gamma_q * (q - mean q)), k' likewise, v: stored bf16          (gemm.cuh QKV epilogue)
  k_prescale  NOT a rounding point of the product (both attention kernels apply k_ln's 1/std to the fp32
              scores); the switch stays to show what rescaling the K tile in shared memory -- the first
              version of the fused kernel -- would add to the error budget
  p           softmax numerators bf16, relative to a lazily raised running maximum,
              64-key tiles; the row sum uses the unrounded values                      (attention_resident.cuh)
The statistics of q_ln / k_ln come from the fp32 accumulators (not from rounded q, k), the
residual stream, all LayerNorm statistics and all accumulators are fp32, as in the product.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, replace

import torch
import torch.nn.functional as F

from . import esm3_ref

LOG2E = 1.4426950408889634
RESCALE_LOG2 = 8.0          # attention_resident.cuh: the running max is raised when exceeded by > 2^8


@dataclass(frozen=True)
class Rounding:
    weights: bool = True
    act: bool = True
    qkv: bool = True
    k_prescale: bool = False
    p: bool = True
    f64: bool = False       # accumulate the GEMMs in float64: an fp32-ulp-level perturbation, used to
                            # measure how far two implementations with IDENTICAL rounding points drift

    @staticmethod
    def none():
        return Rounding(False, False, False, False, False)

    @staticmethod
    def product(T: int = 0):
        """What libesmdiff_b200 does (the same at every sequence length T)."""
        return Rounding(k_prescale=False)


def _r(x, on):
    return x.to(torch.bfloat16).to(torch.float32) if on else x


def _mm(a, w, f64):
    """a [..., K] @ w[N, K]^T with fp32 (or float64) accumulation."""
    if f64:
        return (a.double() @ w.double().T).float()
    return a @ w.T


def _stats(x, eps):
    mean = x.mean(-1, keepdim=True)
    var = x.var(-1, unbiased=False, keepdim=True)
    return mean, torch.rsqrt(var + eps)


def folded_linear(x, ln, W, rd: Rounding, center_blocks: int = 0, eps: float = 1e-5):
    """LayerNorm folded through the Linear as the QKV / W1 epilogues evaluate it:
    ``rstd * (bf16(x) (gamma.W)^T - mean * colsum) + beta W^T``; the first ``center_blocks``
    row blocks of d_model rows have their column means removed (q_ln / k_ln centring)."""
    D = x.shape[-1]
    Wc = W
    if center_blocks:
        Wc = W.clone()
        for i in range(center_blocks):
            Wc[i * D:(i + 1) * D] -= Wc[i * D:(i + 1) * D].mean(0, keepdim=True)
    Wf = _r(Wc * ln.weight[None, :], rd.weights)
    c = Wf.sum(1)
    b = (Wc * ln.bias[None, :]).sum(1) if ln.bias is not None else torch.zeros_like(c)
    mean, rstd = _stats(x, eps)
    acc = _mm(_r(x, rd.act), Wf, rd.f64)
    return acc * rstd + ((-rstd * mean) * c + b)


def attention(qp, kp, v, ssq_q, ssq_k, rd: Rounding, eps: float = 1e-5):
    """qp, kp, v: (B, T, H, 64), un-normalised rotated q', k' and v as stored by the QKV epilogue;
    ssq_*: (B, T) sums of squares of the centred q / k rows.  Online softmax exactly as
    attention_resident.cuh walks it (64-key tiles, lazy running max, bf16 P, fp32 row sum);
    the T mod 128 <= 2 trailing query rows follow the CUDA-core path (fp32 P, exact max)."""
    B, T, H, dh = qp.shape
    D = H * dh
    rstd_q = torch.rsqrt(ssq_q / D + eps)                    # (B, T)
    rstd_k = torch.rsqrt(ssq_k / D + eps)
    if rd.k_prescale:
        kk = _r(kp * rstd_k[:, :, None, None], rd.qkv)
        col = None
    else:
        kk, col = kp, rstd_k
    q = qp.permute(0, 2, 1, 3)
    k = kk.permute(0, 2, 1, 3)
    vv = v.permute(0, 2, 1, 3)
    S = q @ k.transpose(-1, -2)                              # (B, H, T, T) fp32
    if col is not None:
        S = S * col[:, None, None, :]
    sc = (0.125 * LOG2E) * rstd_q[:, None, :, None]          # per query row, log2 domain
    n_left = T % 128 if (T > 128 and 0 < T % 128 <= 2) else 0
    thresh = RESCALE_LOG2 / sc
    O = torch.zeros(B, H, T, dh)
    l = torch.zeros(B, H, T, 1)
    m_run = None
    for j0 in range(0, T, 64):
        s = S[..., j0:j0 + 64]
        mx = s.amax(-1, keepdim=True)
        if m_run is None:
            m_run = mx
        else:
            need = mx > m_run + thresh
            m_new = torch.where(need, mx, m_run)
            f = torch.exp2((m_run - m_new) * sc)
            O, l, m_run = O * f, l * f, m_new
        p = torch.exp2(s * sc - m_run * sc)
        l = l + p.sum(-1, keepdim=True)
        pr = _r(p, rd.p)
        if n_left:
            pr[..., T - n_left:, :] = p[..., T - n_left:, :]
        O = O + pr @ vv[..., j0:j0 + 64, :]
    out = O * (1.0 / l)
    return out.permute(0, 2, 1, 3).reshape(B, T, D)


def geom_branch(ga, x, frames, rd: Rounding):
    """Block 0's geometric attention as run_geom_attention (esmdiff_b200.cu) evaluates it: s_norm output, proj
    weight, proj output and the attention output are bf16 (tensor-core operands / their store epilogues); the
    rotation into the global frame, the scores, the softmax and the weighted sum are fp32 (geom.cuh)."""
    D = x.shape[-1]
    ns = _r(F.layer_norm(x, (D,), ga.s_norm.weight, None, 1e-5), rd.act)
    p = _r(_mm(ns, _r(ga.proj.weight, rd.weights), rd.f64), rd.act)
    att = _r(ga.attention(p, *frames), rd.act)
    return _mm(att, _r(ga.out_proj.weight, rd.weights), rd.f64)


def block_forward(blk, x, cos, sin, rd: Rounding, n_heads: int, frames=None):
    B, T, D = x.shape
    a = blk.attn
    y = folded_linear(x, a.layernorm_qkv[0], a.layernorm_qkv[1].weight, rd, center_blocks=2)
    qc, kc, v = y.chunk(3, dim=-1)
    ssq_q, ssq_k = (qc * qc).sum(-1), (kc * kc).sum(-1)
    dh = D // n_heads
    qp = _r(esm3_ref.apply_rotary((qc * a.q_ln.weight).view(B, T, n_heads, dh), cos, sin), rd.qkv)
    kp = _r(esm3_ref.apply_rotary((kc * a.k_ln.weight).view(B, T, n_heads, dh), cos, sin), rd.qkv)
    vv = _r(v, rd.qkv).view(B, T, n_heads, dh)
    att = _r(attention(qp, kp, vv, ssq_q, ssq_k, rd), rd.act)
    inv = 1.0 / torch.tensor(blk.scale, dtype=torch.float32)
    x = x + _mm(att, _r(a.out_proj.weight, rd.weights), rd.f64) * inv
    # block 0's geometric attention is an exact zero without coordinates (esm3_ref.GeomAttnParamsRef)
    if frames is not None and getattr(blk, "with_geom", False):
        x = x + geom_branch(blk.geom_attn, x, frames, rd) * inv
    f = blk.ffn
    y = folded_linear(x, f[0], f[1].weight, rd)
    g, u = y.chunk(2, dim=-1)
    h = _r(F.silu(g) * u, rd.act)
    return x + _mm(h, _r(f[3].weight, rd.weights), rd.f64) * inv


@torch.no_grad()
def forward(net: esm3_ref.CustomizedESM3Ref, structure_tokens, sequence_tokens, auxiliary_embeddings=None,
            rd: Rounding | None = None, structure_coords=None) -> esm3_ref.NetOutput:
    """``CustomizedESM3Ref.forward`` with the product's rounding points."""
    B, T = structure_tokens.shape
    frames = None
    if structure_coords is not None:
        frames = esm3_ref.geom_ref.build_affine3d_from_coordinates(structure_coords[..., :3, :].expand(B, T, 3, 3))
    rd = Rounding.product(T) if rd is None else rd
    dims = net.dims
    st = net.force_special_structure_ids(structure_tokens, sequence_tokens)
    x = net.encoder(sequence_tokens, st)
    if auxiliary_embeddings is not None:
        x = x + auxiliary_embeddings
    cos, sin = esm3_ref.rotary_tables(T, dims.d_head)
    for blk in net.transformer.blocks:
        x = block_forward(blk, x, cos, sin, rd, dims.n_heads, frames)
    emb = x
    head = net.output_heads.structure_head
    xn = _r(F.layer_norm(x, (dims.d_model,), net.transformer.norm.weight, None, 1e-5), rd.act)
    h0 = F.gelu(_mm(xn, _r(head[0].weight, rd.weights), rd.f64) + head[0].bias)
    h1 = _r(F.layer_norm(h0, (dims.d_model,), head[2].weight, head[2].bias, 1e-5), rd.act)
    logits = _mm(h1, _r(head[3].weight, rd.weights), rd.f64) + head[3].bias
    return esm3_ref.NetOutput(structure_logits=logits, embeddings=emb)


def error_budget(net, structure_tokens, sequence_tokens, aux, metric):
    """{rounding point: metric(emulated with only that point on, fp32 oracle)} plus 'all' and the
    drift between two all-on evaluations that differ only in accumulation precision ('f64')."""
    ref = forward(net, structure_tokens, sequence_tokens, aux, Rounding.none())
    T = structure_tokens.shape[1]
    out = {}
    for name in ("weights", "act", "qkv", "k_prescale", "p"):
        rd = replace(Rounding.none(), **{name: True, **({"qkv": True} if name == "k_prescale" else {})})
        out[name] = metric(forward(net, structure_tokens, sequence_tokens, aux, rd), ref)
    full = forward(net, structure_tokens, sequence_tokens, aux, Rounding.product(T))
    out["all"] = metric(full, ref)
    out["f64_drift"] = metric(forward(net, structure_tokens, sequence_tokens, aux,
                                      replace(Rounding.product(T), f64=True)), full)
    return out
