"""Per-kernel micro-benchmarks on the B200 (CUDA events, warm-up, inputs larger than L2 where the
real step's are).  `gpurun -- python tools/kbench.py [gemm attn rows]`.  Tuning aid, not a test."""
from __future__ import annotations

import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from esmdiff_b200.engine import Dims, Engine  # noqa: E402

dev = torch.device("cuda")


def timeit(fn, n=20, warm=5, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def engine(**env):
    for k, v in env.items():
        os.environ[k] = v
    e = Engine(Dims())
    for k in env:
        os.environ.pop(k)
    return e


def bench_gemm(flush):
    g = torch.Generator(device="cuda").manual_seed(4)
    engs = {"auto": engine(), "bn256": engine(ESMDIFF_RESID_BN="256"), "bn192": engine(ESMDIFF_RESID_BN="192")}
    for M in (25800, 3354, 3096, 12900, 6450):
        for (N, K, epi, name) in [(4608, 1536, 0, "qkv"), (1536, 1536, 1, "out_proj"), (8192, 1536, 2, "w1_swiglu"),
                                  (1536, 4096, 1, "w2")]:
            a = torch.randn(M, K, device=dev, generator=g).bfloat16()
            w = (torch.randn(N, K, device=dev, generator=g) / K ** 0.5).bfloat16()
            if epi == 0:
                out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
            elif epi == 2:
                out = torch.empty(M, N // 2, dtype=torch.bfloat16, device=dev)
            else:
                out = torch.zeros(M, N, device=dev)
            line = f"  gemm {name:10s} M={M:6d} N={N:5d} K={K:5d}:"
            for tag, e in engs.items():
                if epi != 1 and tag != "auto":
                    continue
                ms = timeit(lambda: e.op_gemm(epi, a, w, out, scale=1.1547), flush=flush)
                line += f"  {tag} {ms * 1e3:7.1f} us {2 * M * N * K / ms / 1e9:7.1f} TF"
            ms = timeit(lambda: torch.matmul(a, w.T), flush=flush)
            line += f"  | cuBLAS bf16 {ms * 1e3:7.1f} us {2 * M * N * K / ms / 1e9:7.1f} TF"
            print(line, flush=True)
    for e in engs.values():
        e.close()


def bench_attn(flush):
    g = torch.Generator(device="cuda").manual_seed(5)
    engs = {"resident": engine(ESMDIFF_ATTN="resident"), "stream": engine(ESMDIFF_ATTN="stream")}
    # correctness first
    for (B, T, H) in [(1, 64, 1), (2, 60, 4), (2, 130, 4), (3, 258, 24), (1, 514, 4), (2, 129, 3), (5, 1, 2), (1, 700, 2)]:
        D = H * 64
        qkv = (torch.randn(B * T, 3 * D, device=dev, generator=g) * 1.5).bfloat16()
        q, k, v = [z.view(B, T, H, 64).transpose(1, 2) for z in qkv.float().chunk(3, -1)]
        ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * T, D)
        line = f"  attention B={B} T={T} H={H}:"
        for tag, e in engs.items():
            try:
                got = e.op_attention(qkv, B, T, H)
                e.synchronize()
                err = (got.float() - ref).norm() / ref.norm()
                mx = (got.float() - ref).abs().max()
                line += f"  {tag} rel_fro={err.item():.2e} max_abs={mx.item():.2e} nan={int(torch.isnan(got.float()).sum())}"
            except Exception as ex:       # noqa: BLE001
                line += f"  {tag} FAILED {ex}"
        print(line, flush=True)
    # large-logit case: exercises the lazy rescale
    B, T, H = 2, 258, 4
    D = H * 64
    qkv = (torch.randn(B * T, 3 * D, device=dev, generator=g)).bfloat16()
    qkv[:, :2 * D] *= 6.0
    q, k, v = [z.view(B, T, H, 64).transpose(1, 2) for z in qkv.float().chunk(3, -1)]
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * T, D)
    for tag, e in engs.items():
        got = e.op_attention(qkv, B, T, H)
        e.synchronize()
        print(f"  attention peaked logits {tag}: rel_fro={((got.float() - ref).norm() / ref.norm()).item():.2e}", flush=True)
    for (B, T, H) in [(100, 258, 24), (32, 386, 24), (32, 514, 24), (64, 514, 24), (32, 642, 24), (32, 766, 24), (100, 130, 24)]:
        qkv = torch.randn(B * T, 3 * H * 64, device=dev, generator=g).bfloat16()
        line = f"  attention B={B} T={T} H={H}:"
        for tag, e in engs.items():
            ms = timeit(lambda: e.op_attention(qkv, B, T, H), flush=flush)
            line += f"  {tag} {ms * 1e3:7.1f} us {4 * B * H * T * T * 64 / ms / 1e9:6.1f} TF"
        print(line, flush=True)
    for e in engs.values():
        e.close()


def bench_attn_ln(flush):
    """Resident attention, one vs two softmax threads per query row, with the q_ln / k_ln 1/std applied to
    the scores (the product's form): parity against torch on the same inputs, then time."""
    g = torch.Generator(device="cuda").manual_seed(8)
    engs = {"resident": engine(ESMDIFF_ATTN="resident"), "tiles": engine(ESMDIFF_ATTN="tiles"),
            "nofold": engine(ESMDIFF_ATTN_FOLD="0")}
    for (B, T, H) in [(1, 64, 2), (2, 60, 4), (2, 130, 4), (3, 258, 24), (1, 514, 4), (2, 129, 3), (5, 1, 2), (1, 700, 2),
                      (2, 33, 2), (2, 100, 2), (2, 48, 2), (1, 386, 2), (2, 17, 20), (2, 65, 4), (3, 132, 4), (2, 67, 20), (2, 68, 4),
                      (2, 69, 4), (2, 259, 4), (1, 514, 24), (2, 194, 4)]:
        D = H * 64
        nspan = D // 128 if D % 256 == 0 else 0
        qkv = (torch.randn(B * T, 3 * D, device=dev, generator=g) * 1.5).bfloat16()
        line = f"  attention_ln B={B} T={T} H={H}:"
        for with_ln in ((False, True) if nspan else (False,)):
            q, k, v = [z.view(B, T, H, 64).transpose(1, 2) for z in qkv.float().chunk(3, -1)]
            sumsq = None
            if with_ln:
                sumsq = (torch.rand(B * T, 2 * nspan, device=dev, generator=g) * 200 + 20).contiguous()
                rq = torch.rsqrt(sumsq[:, :nspan].sum(-1) / D + 1e-5).view(B, 1, T, 1)
                rk = torch.rsqrt(sumsq[:, nspan:].sum(-1) / D + 1e-5).view(B, 1, T, 1)
                q, k = q * rq, k * rk
            ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * T, D)
            for tag, e in engs.items():
                got = e.op_attention(qkv, B, T, H, qk_sumsq=sumsq)
                e.synchronize()
                err = (got.float() - ref).norm() / ref.norm()
                line += f"  {tag}{'+ln' if with_ln else ''} {err.item():.2e}/{int(torch.isnan(got.float()).sum())}"
        print(line, flush=True)
    B, T, H = 2, 258, 4
    D = H * 64
    qkv = (torch.randn(B * T, 3 * D, device=dev, generator=g)).bfloat16()
    qkv[:, :2 * D] *= 6.0
    qkv[::7, D:2 * D] *= 3.0                 # a few dominant keys: exercises the lazy rescale in both halves
    q, k, v = [z.view(B, T, H, 64).transpose(1, 2) for z in qkv.float().chunk(3, -1)]
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * T, D)
    for tag, e in engs.items():
        got = e.op_attention(qkv, B, T, H)
        e.synchronize()
        print(f"  attention peaked logits {tag}: rel_fro={((got.float() - ref).norm() / ref.norm()).item():.2e}", flush=True)
    for (B, T, H) in [(100, 258, 24), (100, 256, 24), (13, 258, 24), (25, 258, 24), (32, 514, 24), (64, 514, 24),
                      (32, 766, 24), (100, 130, 24), (100, 258, 20)]:
        D = H * 64
        nspan = D // 128
        qkv = torch.randn(B * T, 3 * D, device=dev, generator=g).bfloat16()
        sumsq = (torch.rand(B * T, 2 * nspan, device=dev, generator=g) * 200 + 20).contiguous()
        line = f"  attention B={B} T={T} H={H}:"
        for tag, e in engs.items():
            ms = timeit(lambda: e.op_attention(qkv, B, T, H), flush=flush)
            ms2 = timeit(lambda: e.op_attention(qkv, B, T, H, qk_sumsq=sumsq), flush=flush)
            line += f"  {tag} {ms * 1e3:6.1f} us, +ln {ms2 * 1e3:6.1f} us ({4 * B * H * T * T * 64 / ms2 / 1e9:5.0f} TF)"
        print(line, flush=True)
    for e in engs.values():
        e.close()


def bench_rows(flush):
    g = torch.Generator(device="cuda").manual_seed(6)
    e = engine()
    for M, B in ((16254, 63), (25800, 100)):
        x = torch.randn(M, 1536, device=dev, generator=g)
        w = torch.ones(1536, device=dev)
        ms = timeit(lambda: e.op_layernorm(x, w, w), flush=flush)
        print(f"  layernorm M={M}: {ms * 1e3:.1f} us  {M * 1536 * 6 / ms / 1e6:.0f} GB/s", flush=True)
        qkv = torch.randn(M, 3 * 1536, device=dev, generator=g).bfloat16()
        ms = timeit(lambda: e.op_qk_norm_rope(qkv, w, w, B, 258), flush=flush)
        print(f"  qk_norm_rope M={M}: {ms * 1e3:.1f} us  {M * 1536 * 2 * 4 / ms / 1e6:.0f} GB/s", flush=True)
        logits = torch.randn(B, 258, 4101, device=dev, generator=g)
        xx = torch.full((B, 258), 4096, device=dev)
        ms = timeit(lambda: e.sample_step(xx.fill_(4096), logits, None, 0.7, 0.6, seed=1), flush=flush)
        print(f"  sample_step philox all-masked M={M}: {ms * 1e3:.1f} us  {M * 4101 * 4 / ms / 1e6:.0f} GB/s", flush=True)
    e.close()


def bench_qk(flush):
    """q_ln / k_ln + RoPE folded into the QKV epilogue + attention (default) against the stand-alone
    kernel chain (LN-folded store epilogue -> qk_layernorm_rope -> attention), kernel by kernel."""
    g = torch.Generator(device="cuda").manual_seed(7)
    e = engine()
    D, H = 1536, 24
    for (B, T) in [(100, 258), (13, 258), (32, 514), (16, 1026)]:
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
        t_plain = timeit(lambda: e.op_gemm_ln(5, xb, wf, qkv, bias=bs, stats_in=stats, colsum=cs), flush=flush)
        t_rope = timeit(lambda: e.op_gemm_qkv_rope(xb, wfc, bsc, stats, csc, gam, T, 2 * D), flush=flush)
        t_row = timeit(lambda: e.op_qk_norm_rope(qkv, ones, ones, B, T), flush=flush)
        qkv2, sumsq = e.op_gemm_qkv_rope(xb, wfc, bsc, stats, csc, gam, T, 2 * D)
        runs = ""
        for run in ("0", "2", "3", "6", "9"):
            er = engine(ESMDIFF_QKV_RUN=run)
            tr_ = timeit(lambda: er.op_gemm_qkv_rope(xb, wfc, bsc, stats, csc, gam, T, 2 * D), flush=flush)
            runs += f" run={run}: {tr_ * 1e3:.1f} us"
            er.close()
        print(f"  B={B} T={T}: QKV + rope epilogue by tile schedule (default = contiguous ranges {t_rope * 1e3:.1f} us):{runs}", flush=True)
        t_att = timeit(lambda: e.op_attention(qkv, B, T, H), flush=flush)
        t_att_ln = timeit(lambda: e.op_attention(qkv2, B, T, H, qk_sumsq=sumsq), flush=flush)
        fl = 2.0 * M * 3 * D * D
        print(f"  B={B} T={T}: qkv gemm plain {t_plain * 1e3:.1f} us ({fl / t_plain / 1e9:.0f} TF) | +rope epilogue "
              f"{t_rope * 1e3:.1f} us ({fl / t_rope / 1e9:.0f} TF) | qk_norm_rope kernel {t_row * 1e3:.1f} us | attention "
              f"{t_att * 1e3:.1f} us | attention+rstd {t_att_ln * 1e3:.1f} us || chain separate "
              f"{(t_plain + t_row + t_att) * 1e3:.1f} us, fused {(t_rope + t_att_ln) * 1e3:.1f} us", flush=True)
    e.close()


def bench_enc(flush):
    """Inpainting front end: the fp32 structure encoder (once per target) and what live geometric attention in
    block 0 adds to a forward of the sampling network."""
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    from esmdiff_b200.encoder import load_encoder
    from esmdiff_b200.synthetic import random_state_dict
    from oracle.vqvae_enc_ref import synthetic_backbone
    enc = load_encoder(None)
    d = enc.dims
    for L in (58, 256, 512, 1024):
        bb = synthetic_backbone(L, seed=1)[None].to(dev)
        ms = timeit(lambda: enc.encode(bb, return_aux=True), n=10, warm=3)
        M = L * min(d.knn, L)
        fl = 2.0 * M * d.n_layers * (d.d_model * 15 * d.v_heads + 3 * d.v_heads * d.d_model + 3 * d.d_model * d.ffn_hidden)
        print(f"  encoder L={L:5d}: {ms:7.3f} ms  ({M} neighbourhood rows, {fl / 1e9:.1f} GFLOP in the fp32 GEMMs -> "
              f"{fl / ms / 1e9:.1f} TFLOP/s overall)", flush=True)
    e = Engine(Dims())
    e.load_state_dict(random_state_dict(Dims(), device="cuda", seed=0, full=True))
    for (B, T) in [(100, 258), (13, 258), (32, 514)]:
        g = torch.Generator().manual_seed(0)
        seq = torch.randint(4, 24, (B, T), generator=g).to(dev)
        xt = torch.randint(0, 4096, (B, T), generator=g).to(dev)
        coords = torch.full((B, T, 3, 3), float("nan"))
        coords[:, 1:-1] = synthetic_backbone(T - 2, seed=2)
        coords[:, 2:34] = float("inf")
        logits = torch.empty(B, T, 4101, device=dev)
        t0 = timeit(lambda: e.forward_sigma(seq, xt, 0.5, logits_out=logits), n=6, warm=2)
        e.set_structure_coords(coords)
        t1 = timeit(lambda: e.forward_sigma(seq, xt, 0.5, logits_out=logits), n=6, warm=2)
        e.set_structure_coords(None)
        keys = B * T * T * 256 * 36.0                       # bytes of rotated key / value vectors the attention walks / QPT
        print(f"  forward B={B} T={T}: {t0:7.3f} ms, with structure_coords {t1:7.3f} ms (+{(t1 - t0) * 1e3:.0f} us, "
              f"+{(t1 / t0 - 1) * 100:.2f} %); key-vector reads {keys / 8 / 1e9:.2f} GB at QPT = 8", flush=True)
    e.close()


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    which = sys.argv[1:] or ["attn", "gemm", "rows"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > L2 (126 MB)
    table = dict(gemm=bench_gemm, attn=bench_attn, attn_ln=bench_attn_ln, rows=bench_rows, qk=bench_qk, enc=bench_enc)
    for wname in which:
        print(f"=== {wname} ===", flush=True)
        table[wname](flush)
