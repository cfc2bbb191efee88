import sys, torch
sys.path.insert(0, '/root/repo')
from tools.kbench import timeit, engine, dev
e = engine()
g = torch.Generator(device="cuda").manual_seed(4)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for M in (4096, 9546, 25800):
    for (N, K, name) in [(1536, 1536, "out_proj"), (1536, 4096, "w2")]:
        a = torch.randn(M, K, device=dev, generator=g).bfloat16()
        w = (torch.randn(N, K, device=dev, generator=g) / K ** 0.5).bfloat16()
        x = torch.zeros(M, N, device=dev)
        xb = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        st = torch.zeros(M, N // 128, 2, device=dev)
        o16 = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        for tag, fn in (("resid", lambda: e.op_gemm(1, a, w, x, scale=1.15)),
                        ("resid_ln", lambda: e.op_gemm_ln(6, a, w, x, scale=1.15, stats_out=st, xb_out=xb)),
                        ("store", lambda: e.op_gemm(0, a, w, o16))):
            t_cold = timeit(fn, flush=flush)
            t_warm = timeit(fn, flush=None)
            print(f"{name} M={M} {tag:9s}: cold {t_cold*1e3:7.1f} us {2*M*N*K/t_cold/1e9:7.1f} TF | L2-warm {t_warm*1e3:7.1f} us {2*M*N*K/t_warm/1e9:7.1f} TF", flush=True)
