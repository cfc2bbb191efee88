"""GPU bring-up diagnostics: every kernel against a torch reference, with error metrics printed.
Run on the B200 box (`gpurun -- python tools/gpu_diag.py`).  Not a test; tests live in tests/."""
from __future__ import annotations

import sys
import time
import traceback
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from esmdiff_b200.engine import Dims, Engine  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda")


def stat(name, got, ref):
    got, ref = got.float(), ref.float()
    d = (got - ref).abs()
    denom = ref.abs().max().clamp_min(1e-12)
    rel_fro = (got - ref).norm() / ref.norm().clamp_min(1e-12)
    bad = (d > 0.02 * denom).float().mean()
    print(f"  {name:46s} max_abs={d.max().item():.4e} rel_max={(d.max() / denom).item():.3e} "
          f"rel_fro={rel_fro.item():.3e} frac_bad={bad.item():.4f} nan={int(torch.isnan(got).sum())}", flush=True)
    return rel_fro.item()


def section(title):
    print(f"\n=== {title} ===", flush=True)


def run(fn, eng):
    try:
        fn(eng)
        eng.synchronize()
    except Exception as e:      # keep going: one broken kernel must not hide the others
        print(f"  !! {fn.__name__} raised {type(e).__name__}: {e}", flush=True)
        traceback.print_exc()


def diag_gemm_structured(eng):
    section("gemm structured (row/col/k mapping)")
    M, N, K = 128, 256, 64
    for k0 in (0, 7, 8, 16, 33, 63):
        a = torch.zeros(M, K, device=dev); a[:, k0] = torch.arange(M, device=dev) % 64
        w = torch.zeros(N, K, device=dev); w[:, k0] = 1.0
        out = torch.zeros(M, N, dtype=torch.bfloat16, device=dev)
        eng.op_gemm(0, a.bfloat16(), w.bfloat16(), out); eng.synchronize()
        r1 = stat(f"k0={k0} rows", out, (a @ w.T))
        a = torch.zeros(M, K, device=dev); a[:, k0] = 1.0
        w = torch.zeros(N, K, device=dev); w[:, k0] = torch.arange(N, device=dev) % 64
        out = torch.zeros(M, N, dtype=torch.bfloat16, device=dev)
        eng.op_gemm(0, a.bfloat16(), w.bfloat16(), out); eng.synchronize()
        r2 = stat(f"k0={k0} cols", out, (a @ w.T))
        if (r1 > 1e-2 or r2 > 1e-2) and k0 == 0:
            print("   first rows of out:", out[:4, :8].float().tolist())


def diag_gemm_random(eng):
    section("gemm random, all epilogues")
    g = torch.Generator(device="cuda").manual_seed(0)
    for (M, N, K) in [(128, 256, 64), (128, 256, 256), (256, 512, 1536), (1000, 4608, 1536),
                      (16254, 1536, 1536), (300, 1536, 4096), (77, 384, 128)]:
        a = torch.randn(M, K, device=dev, generator=g).bfloat16()
        w = (torch.randn(N, K, device=dev, generator=g) / K ** 0.5).bfloat16()
        ref = a.float() @ w.float().T
        out = torch.full((M, N), 7.0, dtype=torch.bfloat16, device=dev)
        eng.op_gemm(0, a, w, out); eng.synchronize()
        stat(f"store_bf16 M={M} N={N} K={K}", out, ref)
    M, N, K = 1000, 1536, 1536
    a = torch.randn(M, K, device=dev, generator=g).bfloat16()
    w = (torch.randn(N, K, device=dev, generator=g) / K ** 0.5).bfloat16()
    ref = a.float() @ w.float().T
    x0 = torch.randn(M, N, device=dev, generator=g)
    x = x0.clone()
    eng.op_gemm(1, a, w, x, scale=1.1547005); eng.synchronize()
    stat("resid_f32", x, x0 + ref / 1.1547005)
    bias = torch.randn(N, device=dev, generator=g)
    out = torch.empty(M, N, device=dev)
    eng.op_gemm(3, a, w, out, bias=bias); eng.synchronize()
    stat("bias_gelu_f32", out, torch.nn.functional.gelu(ref + bias))
    Nv = 4101
    w3 = (torch.randn(Nv, K, device=dev, generator=g) / K ** 0.5).bfloat16()
    b3 = torch.randn(Nv, device=dev, generator=g)
    out = torch.full((M, Nv), -5.0, device=dev)
    eng.op_gemm(4, a, w3, out, bias=b3); eng.synchronize()
    stat("bias_f32 N=4101", out, a.float() @ w3.float().T + b3)
    F = 4096
    w1 = torch.randn(2 * F, K, device=dev, generator=g) / K ** 0.5
    w1i = eng.op_convert_bf16(w1.contiguous(), swiglu_hidden=F)
    out = torch.empty(M, F, dtype=torch.bfloat16, device=dev)
    eng.op_gemm(2, a, w1i, out); eng.synchronize()
    z = a.float() @ w1.bfloat16().float().T
    stat("swiglu_bf16", out, torch.nn.functional.silu(z[:, :F]) * z[:, F:])


def diag_rows(eng):
    section("layernorm / qk_norm_rope")
    g = torch.Generator(device="cuda").manual_seed(1)
    for D in (1536, 256):
        x = torch.randn(999, D, device=dev, generator=g) * 3 + 0.5
        w = torch.randn(D, device=dev, generator=g); b = torch.randn(D, device=dev, generator=g)
        stat(f"layernorm D={D} w+b", eng.op_layernorm(x, w, b), torch.nn.functional.layer_norm(x, (D,), w, b))
        stat(f"layernorm D={D} w", eng.op_layernorm(x, w, None), torch.nn.functional.layer_norm(x, (D,), w, None))
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    from oracle.esm3_ref import apply_rotary, rotary_tables
    for (B, T, D) in [(3, 60, 1536), (2, 258, 256)]:
        H = D // 64
        qkv = torch.randn(B * T, 3 * D, device=dev, generator=g).bfloat16()
        qw = torch.randn(D, device=dev, generator=g); kw = torch.randn(D, device=dev, generator=g)
        q, k, v = qkv.float().chunk(3, -1)
        cos, sin = rotary_tables(T)
        cos, sin = cos.to(dev), sin.to(dev)
        qr = apply_rotary(torch.nn.functional.layer_norm(q, (D,), qw).view(B, T, H, 64), cos, sin).reshape(B * T, D)
        kr = apply_rotary(torch.nn.functional.layer_norm(k, (D,), kw).view(B, T, H, 64), cos, sin).reshape(B * T, D)
        got = eng.op_qk_norm_rope(qkv.clone(), qw, kw, B, T)
        stat(f"qk_norm_rope q B={B} T={T} D={D}", got[:, :D], qr)
        stat(f"qk_norm_rope k", got[:, D:2 * D], kr)
        stat(f"qk_norm_rope v untouched", got[:, 2 * D:], v)


def diag_attention(eng):
    section("attention")
    g = torch.Generator(device="cuda").manual_seed(2)
    for (B, T, H) in [(1, 64, 1), (1, 128, 1), (2, 60, 4), (2, 130, 4), (3, 258, 24), (1, 514, 4), (1, 1026, 2)]:
        D = H * 64
        qkv = (torch.randn(B * T, 3 * D, device=dev, generator=g) * 1.5).bfloat16()
        q, k, v = [z.view(B, T, H, 64).transpose(1, 2) for z in qkv.float().chunk(3, -1)]
        ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * T, D)
        got = eng.op_attention(qkv, B, T, H); eng.synchronize()
        stat(f"attention B={B} T={T} H={H}", got, ref)


def diag_sampler(eng):
    section("sampler")
    g = torch.Generator(device="cuda").manual_seed(3)
    B, T, V = 4, 60, 4101
    logits = torch.randn(B, T, V, device=dev, generator=g) * 3
    u = torch.rand(B, T, V, device=dev, generator=g)
    x = torch.randint(0, 4096, (B, T), device=dev, generator=g)
    x[torch.rand(B, T, device=dev, generator=g) < 0.6] = 4096
    mc_t, mc_s = 0.72, 0.68
    lg = logits.clone(); lg[..., 4096] += -1e6
    logp = lg - torch.logsumexp(lg, -1, keepdim=True)
    keep = x != 4096
    lp = logp.clone(); lp[keep] = -1e6; lp[keep, x[keep]] = 0
    got_lp = eng.logits_parameterization(logits.clone(), x)
    stat("logits_parameterization", got_lp.clamp_min(-50), lp.clamp_min(-50))
    q = lp.exp() * (mc_t - mc_s); q[..., 4096] = mc_s
    cand = (q / (1e-10 - (u + 1e-10).log())).argmax(-1)
    want = torch.where(keep, x, cand)
    got = eng.sample_step(x.clone(), logits, u, mc_t, mc_s); eng.synchronize()
    print(f"  sample_step mismatches: {(got != want).sum().item()} / {B * T}")
    got = eng.denoise_argmax(x.clone(), logits); eng.synchronize()
    print(f"  denoise_argmax mismatches: {(got != torch.where(keep, x, lp.argmax(-1))).sum().item()} / {B * T}")
    got = eng.sample_step(x.clone(), logits, None, mc_t, mc_s, seed=5, step=2); eng.synchronize()
    print(f"  philox sample_step: ids in range = {bool(((got >= 0) & (got <= 4100)).all())}, "
          f"still masked = {(got == 4096).float().mean().item():.3f} (expect ~ {0.6 * mc_s / mc_t:.3f})")


def diag_forward(eng_unused):
    section("tiny model forward vs oracle")
    from oracle import esm3_ref
    dims_o = esm3_ref.Esm3Dims(d_model=256, n_heads=4, v_heads=8, n_layers=2)
    net, emb = esm3_ref.build_reference_model(dims_o, seed=0)
    eng = Engine(Dims(d_model=256, n_heads=4, v_heads=8, n_layers=2))
    eng.load_state_dict(esm3_ref.full_state_dict(net, emb))
    g = torch.Generator().manual_seed(0)
    B, T = 3, 70
    seq = torch.cat([torch.zeros(B, 1, dtype=torch.long), torch.randint(4, 24, (B, T - 2), generator=g),
                     torch.full((B, 1), 2)], 1)
    xt = torch.randint(0, 4096, (B, T), generator=g)
    xt[torch.rand(B, T, generator=g) < 0.5] = 4096
    sigma = 1.234
    cond_ref = emb(torch.tensor([sigma]))[0]
    cond = eng.time_embed(sigma); eng.synchronize()
    stat("time_embed", cond.cpu(), cond_ref.detach())
    ref = net(structure_tokens=xt, sequence_tokens=seq, auxiliary_embeddings=cond_ref[None, None].expand(B, T, -1))
    logits, embd = eng.forward(seq, xt, aux=cond, want_embeddings=True); eng.synchronize()
    stat("embeddings (pre-norm residual)", embd.cpu(), ref.embeddings)
    stat("structure_logits", logits.cpu(), ref.structure_logits)
    l2 = eng.forward_sigma(seq, xt, sigma); eng.synchronize()
    stat("forward_sigma vs forward", l2, logits)
    eng.close()


def diag_perf(eng):
    section("rough kernel timings (CUDA events, 10 iters)")
    g = torch.Generator(device="cuda").manual_seed(4)
    M = 16254
    def timeit(fn, n=10):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    for (N, K, epi) in [(4608, 1536, 0), (1536, 1536, 1), (8192, 1536, 2), (1536, 4096, 1)]:
        a = torch.randn(M, K, device=dev, generator=g).bfloat16()
        w = (torch.randn(N, K, device=dev, generator=g) / K ** 0.5).bfloat16()
        if epi == 0: out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        elif epi == 2: out = torch.empty(M, N // 2, dtype=torch.bfloat16, device=dev)
        else: out = torch.zeros(M, N, device=dev)
        ms = timeit(lambda: eng.op_gemm(epi, a, w, out, scale=1.1547))
        print(f"  gemm epi={epi} M={M} N={N} K={K}: {ms:.3f} ms  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
        ms = timeit(lambda: torch.matmul(a, w.T))
        print(f"  cuBLAS (torch.matmul bf16) same shape: {ms:.3f} ms  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
    B, T, H = 63, 258, 24
    qkv = torch.randn(B * T, 3 * H * 64, device=dev, generator=g).bfloat16()
    ms = timeit(lambda: eng.op_attention(qkv, B, T, H))
    print(f"  attention B={B} T={T} H={H}: {ms:.3f} ms  {4 * B * H * T * T * 64 / ms / 1e9:.1f} TFLOP/s")
    x = torch.randn(M, 1536, device=dev, generator=g); w = torch.ones(1536, device=dev)
    ms = timeit(lambda: eng.op_layernorm(x, w, w))
    print(f"  layernorm M={M}: {ms:.3f} ms  {M * 1536 * 6 / ms / 1e6:.0f} GB/s")
    ms = timeit(lambda: eng.op_qk_norm_rope(qkv, w, w, B, T))
    print(f"  qk_norm_rope: {ms:.3f} ms  {M * 1536 * 2 * 4 / ms / 1e6:.0f} GB/s")
    logits = torch.randn(B, T, 4101, device=dev, generator=g)
    u = torch.rand(B, T, 4101, device=dev, generator=g)
    xx = torch.full((B, T), 4096, device=dev)
    ms = timeit(lambda: eng.sample_step(xx.fill_(4096), logits, u, 0.7, 0.6))
    print(f"  sample_step all-masked: {ms:.3f} ms  {M * 4101 * 8 / ms / 1e6:.0f} GB/s")


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), torch.__version__, flush=True)
    t0 = time.time()
    eng = Engine(Dims())
    which = sys.argv[1:] or ["structured", "random", "rows", "attention", "sampler", "forward", "perf"]
    table = dict(structured=diag_gemm_structured, random=diag_gemm_random, rows=diag_rows,
                 attention=diag_attention, sampler=diag_sampler, forward=diag_forward, perf=diag_perf)
    for w in which:
        run(table[w], eng)
    print(f"\ndiag done in {time.time() - t0:.1f}s, launches={eng.launch_count}")
