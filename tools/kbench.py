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
    engs = {"auto": engine()}
    for M in (16254, 9546, 25800):
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
    engs = {"resident": engine(), "stream": engine(ESMDIFF_ATTN="stream")}
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


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    which = sys.argv[1:] or ["attn", "gemm", "rows"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > L2 (126 MB)
    table = dict(gemm=bench_gemm, attn=bench_attn, rows=bench_rows)
    for wname in which:
        print(f"=== {wname} ===", flush=True)
        table[wname](flush)
