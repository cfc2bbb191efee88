"""Benchmark of the ESMDiff ddpm sampling path (BASELINE.json metric: structure-tokens/s).

    python bench.py [--gpus N --steps K --warmup W] [--workload config2|config3] [--scaling strong|weak]
                    [--impl reference]

One "step" = one complete sampling job of the workload.  Default workload = BASELINE config 2: a
synthetic L=256 protein, random-init ESM3-open-sized weights, num_steps=25 (+1 noise-removal
forward), num_samples=100 (the reference splits them [63, 37] only to fit a 32-80 GB GPU,
sample_esmdiff.py:181-194; `--chunks reference` keeps that list; samples are i.i.d.).
`--workload config3` = BASELINE config 3: L=512, num_steps=50, num_samples=256.  The timed window is
the reference's own (sample_esmdiff.py:177 -> :223, "Sampling token time").
structure-tokens/s = num_samples * L / window time.

N > 1 (torchrun, one rank per GPU): conformation samples are independent (the reference `repeat`s
one row, sample_esmdiff.py:186,190), so the ONE job is sharded: rank r samples its contiguous
share of the num_samples (12 or 13 of 100 at N = 8; 32 of 256 for config 3) and the final tokens
are all-gathered inside the timed region -- `"scaling": "strong"`, value = num_samples * L /
max-over-ranks time.  `--scaling weak` runs the whole job on every rank instead (replicas; value =
N * num_samples * L / time).  Weights are generated on rank 0 and broadcast once over NCCL before
the timed region.  No per-step collective exists on this path.

`value`   : device-resident inputs, CUDA events on the launching stream, max over ranks, per-launch
            profiling OFF.
`e2e`     : the same job through the public host API (esmdiff_b200.sampling.sample_structure_tokens
            -> MaskedDiffusionLanguageModeling.ddpm_sample) from pinned HOST token buffers to
            HOST int64 tokens, copies inside the timed region.
`roofline`: the dominant kernel family (tcgen05 GEMM, all epilogues) -- algorithmic FLOPs of its
            launches / their CUDA-event durations, measured live in a SEPARATE profiled pass of the
            same job right after the timed region (esmdiff_profile_*: an event pair around every
            launch costs ~1 %, so it is kept out of `value`); peak from MEASURED_PEAKS.json
            (sustained bf16) else the fallback.  `whole_job_frac` = algorithmic FLOPs of the job /
            the timed (unprofiled) window / peak.
`cpu_baseline` / `--impl reference`: the CPU restatement of the reference path (oracle/: the
            reference's sampler ops + an fp32 PyTorch ESM3 restatement; the esm package itself is
            not installable offline) on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

EPS = 1e-5
WORKLOADS = {      # BASELINE.json configs[1] (the metric's configuration) and configs[2]
    "config2": {"L": 256, "samples": 100, "steps": 25,
                "name": "config2: synthetic L=256 protein, random-init ESM3-open dims (d=1536, 48 layers, 24 heads, "
                        "V=4101), num_steps=25 + noise removal, num_samples=100"},
    "config3": {"L": 512, "samples": 256, "steps": 50,
                "name": "config3: synthetic L=512 protein, random-init ESM3-open dims, num_steps=50 + noise removal, "
                        "num_samples=256"},
}
FALLBACK_PEAK_TFLOPS = 1400.0       # B200_PROFILING.md: sustained cuBLAS bf16 under the 1 kW cap
FALLBACK_PEAK_GBS = 6650.0


def forward_flops(B: int, T: int) -> float:
    """GEMMs + attention matmuls of one forward (ESM3-open dims)."""
    return float(B) * T * (48 * (56_623_104 + 6144 * T) + 17_316_864)


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel family (the tcgen05 GEMMs: QKV, out_proj, W1, W2 --
    one launch each per block) from the newest committed `ncu --set full` summary
    (tools/ncu_summary.py -> profiles/*_ncu_full.json).  None when no capture is committed."""
    files = sorted((ROOT / "profiles").glob("*_ncu_full.json"))
    if not files:
        return None, None
    d = json.loads(files[-1].read_text())
    g = [k["dram_bytes"] for k in d["kernels"] if "gemm_bf16_tn_kernel" in k["kernel"]]
    return (sum(g) / len(g), files[-1].name) if g else (None, None)


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), float(d["hbm_gbs"]), "measured"
    return FALLBACK_PEAK_TFLOPS, FALLBACK_PEAK_GBS, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.thread.join(timeout=5)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle restatement of the reference path on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_reference(wl: dict, steps: int, warmup: int, budget_s: float, emit_line: bool):
    """The reference path (its own sampler ops + the fp32 restated network, oracle/) on all host
    cores.  BASELINE.md section 3 asks for one forward + one sampler update at the reference's real
    chunk shape (B = 63 at T = 258, ~35 s per pass on 16 cores); a bench line has (steps + warmup)
    timed iterations to fit into `budget_s`, so the batch is the largest B <= the reference's chunk
    size whose pass fits the per-iteration budget (calibrated on a B = 1 pass), and the deviation
    is stated in `sample`.  tokens/s = B * L / ((num_steps + 1) * mean pass time)."""
    from esmdiff_b200.engine import Dims
    from esmdiff_b200.sampling import chunk_sizes
    from esmdiff_b200.synthetic import random_state_dict
    from esmdiff_b200.tokenization import synthetic_sequence_tokens
    from oracle import esm3_ref, mdlm_ref

    L, n_steps = wl["L"], wl["steps"]
    T = L + 2
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = random_state_dict(Dims(), device="cpu", seed=0, full=True)
    net, emb = esm3_ref.build_from_state_dict(esm3_ref.Esm3Dims(), sd)
    del sd
    row = synthetic_sequence_tokens(L, seed=0)
    sampler = mdlm_ref.SamplerRef(net, emb)
    ts, dt = mdlm_ref.time_grid(n_steps, EPS)
    torch.manual_seed(123)

    def one_update(i, x, seq):
        sigma, mc_t, mc_s = mdlm_ref.move_chances(ts[i] * torch.ones(x.shape[0], 1), dt)
        logp, _ = sampler.log_p_x0(x, seq, sigma)
        return mdlm_ref.ddpm_update_tail(logp, x, mc_t, mc_s, torch.rand_like(logp))

    with torch.no_grad():
        x1 = torch.full((1, T), 4096, dtype=torch.int64)
        one_update(0, x1, row[None])                                # warms the allocator / thread pool
        t0 = time.perf_counter()
        one_update(0, x1, row[None])
        t1 = time.perf_counter() - t0                               # one B = 1 pass
    n_iter = max(1, steps + warmup)
    ref_chunk = chunk_sizes(T, wl["samples"])[0]
    B = int(max(1, min(ref_chunk, budget_s / n_iter / max(t1, 1e-3))))
    seq = row[None].repeat(B, 1)
    x = torch.full((B, T), 4096, dtype=torch.int64)
    times = []
    with torch.no_grad():
        for it in range(n_iter):
            t0 = time.perf_counter()
            x = one_update(min(it, n_steps - 1), x, seq)
            if it >= warmup:
                times.append(time.perf_counter() - t0)
    per_pass = sum(times) / len(times)
    value = B * L / (per_pass * (n_steps + 1))                      # 25 updates + 1 noise-removal forward
    sample = (f"B={B} of the reference's {ref_chunk}-sample chunk at T={T} (largest batch whose pass fits the time "
              f"budget; BASELINE.md asks for B={ref_chunk}), 1 forward+update pass per timed iteration, "
              f"{len(times)} iterations, fp32 torch CPU; tokens/s = B*{L} / ({n_steps + 1} x mean pass time)")
    base = {"value": value, "unit": "tokens/s", "cores": cores, "kind": "port", "sample": sample}
    if not emit_line:
        return base
    line = {"impl": "reference", "metric": "structure_tokens_per_sec", "value": value, "unit": "tokens/s",
            "n_gpus": 0, "steps": steps, "warmup": warmup, "ms_per_step": per_pass * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(wl), "cpu_baseline": base,
            "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return base


def workload_config(wl, chunks=None, n_gpus=1, rng="philox", scaling="strong", per_rank=None):
    from esmdiff_b200.sampling import chunk_sizes
    T = wl["L"] + 2
    return {"workload": wl["name"], "L": wl["L"], "T": T, "num_samples": wl["samples"], "num_steps": wl["steps"],
            "samples_per_gpu": per_rank if per_rank is not None else [wl["samples"]],
            "chunks_rank0": chunks, "reference_chunks": chunk_sizes(T, wl["samples"]), "uniforms": rng,
            "parallelism": (f"ONE job, samples sharded over {n_gpus} GPU(s)" if scaling == "strong"
                            else f"independent replicas of the job x{n_gpus}"),
            "l2": "inputs larger than L2 (2.7 GB of bf16 weights streamed per forward)"}


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def gpu_bench(args):
    # NCCL prints its version banner on stdout; the contract is ONE JSON line there
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    from esmdiff_b200 import distributed as D
    from esmdiff_b200.engine import Dims, Engine
    from esmdiff_b200.model import MaskedDiffusionLanguageModeling
    from esmdiff_b200.noise_utils import LogLinearNoise
    from esmdiff_b200.sampling import chunk_sizes, chunk_sizes_b200, sample_structure_tokens
    from esmdiff_b200.synthetic import random_state_dict
    from esmdiff_b200.tokenization import synthetic_sequence_tokens

    wl = WORKLOADS[args.workload]
    L, n_samples, n_steps = wl["L"], wl["samples"], wl["steps"]
    T = L + 2
    rank, world, local = D.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dims = Dims()
    eng = Engine(dims, device=local)
    sd = random_state_dict(dims, device=dev, seed=0) if rank == 0 else None
    sd = D.broadcast_state_dict(sd, dev)                            # the one NCCL weight broadcast
    eng.load_state_dict(sd)
    del sd
    torch.cuda.empty_cache()

    class Net:                                                      # CustomizedESM3 surface over the engine
        engine, device, output_heads = eng, dev, None

    model = MaskedDiffusionLanguageModeling(net=Net(), noise_schedule=LogLinearNoise(), sigma_embedder=None,
                                            time_conditioning=True, noise_removal=True, rng=args.rng)
    # the job's samples owned by this rank: a contiguous share (strong) or all of them (weak replicas)
    strong = args.scaling == "strong"
    if args.shard_of > 1:
        # one-GPU study of the strong-scaling shard: rank 0's share of an args.shard_of-way split, alone
        assert world == 1, "--shard-of is a single-GPU projection"
        shards = [D.shard_samples(n_samples, args.shard_of, 0)]
    elif strong:
        shards = [D.shard_samples(n_samples, world, r) for r in range(world)]
    else:
        shards = [(0, n_samples)] * world
    first, count = shards[rank]
    counts = [c for _, c in shards]
    chunks = chunk_sizes_b200(T, count) if args.chunks == "b200" else chunk_sizes(T, count)
    seq_host = synthetic_sequence_tokens(L, seed=0).pin_memory()
    seq_dev = seq_host.to(dev)
    sigma, mc_t, mc_s = model._schedule(n_steps, EPS, 1.0, dev)

    def job_resident(step_idx):
        outs = []
        for ci, bs in enumerate(chunks):
            batch = seq_dev[None].expand(bs, T).contiguous()
            # RNG contract of the sharded job: seed = base + first sample index of the chunk
            outs.append(eng.ddpm_sample(batch, None, n_steps, sigma, mc_t, mc_s,
                                        seed=100003 * (step_idx + 7) + first + sum(chunks[:ci]) + 1009 * rank * (not strong)))
        tok = torch.cat(outs)[:, 1:-1].contiguous()
        return D.gather_tokens(tok, counts)

    def job_e2e(step_idx):
        torch.manual_seed(123 + step_idx + 1000 * rank)
        tok, _ = sample_structure_tokens(model, seq_host, count, n_steps, eps=EPS, chunks=chunks,
                                         verbose=False)
        tok = D.gather_tokens(tok, counts)
        return tok.to("cpu", non_blocking=False)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident timing (per-launch profiling off) --------------------------------------
    for w in range(args.warmup):
        job_resident(-1 - w)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for k in range(args.steps):
        tok = job_resident(k)
    e1.record()
    barrier()
    eng.synchronize()
    launches = eng.launch_count - l0
    ms_total = D.max_over_ranks(e0.elapsed_time(e1), dev)
    clk = clocks.stop() if rank == 0 else None
    total = sum(counts)
    assert tok.shape == (total, L) and int((tok == 4096).sum()) == 0

    # ---- the same job once more with an event pair around every launch: per-kernel table -------
    eng.profile(True)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    job_resident(args.steps)
    p1.record()
    eng.synchronize()
    eng.profile(False)
    prof = eng.profile_read()
    prof_ms = p0.elapsed_time(p1)

    # ---- end to end through the host API ------------------------------------------------------
    for w in range(min(args.warmup, 1)):
        job_e2e(-1 - w)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        tok_host = job_e2e(k)
    torch.cuda.synchronize(dev)
    e2e_s = D.max_over_ranks(time.perf_counter() - t0, dev)
    if world > 1:
        torch.distributed.barrier()
    assert not tok_host.is_cuda and tok_host.shape == (total, L)

    tokens_per_step = total * L
    ms_per_step = ms_total / args.steps
    value = tokens_per_step / (ms_per_step * 1e-3)
    e2e_value = tokens_per_step / (e2e_s / args.steps)

    peak_tf, peak_gbs, how = measured_peaks()
    traffic, traffic_src = ncu_traffic()
    gemm = [prof[k] for k in ("gemm_store_bf16", "gemm_resid_f32", "gemm_swiglu", "gemm_bias_gelu", "gemm_bias")]
    g_ms, g_fl, g_n = (sum(x[i] for x in gemm) for i in range(3))
    achieved = g_fl / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    # algorithmic FLOPs of what THIS rank computed per step; the whole job is the sum over ranks
    fwd_flops = (n_steps + 1) * sum(forward_flops(b, T) for b in chunks)
    job_flops = D.sum_over_ranks(fwd_flops, dev)
    kernels = {}
    for name, (ms, work, n) in prof.items():
        if n == 0:
            continue
        tensor = name.startswith("gemm") or name == "attention"
        kernels[name] = {"launches": n, "ms_total": round(ms, 3), "share_of_step": round(ms / prof_ms, 4),
                         ("tflops" if tensor else "gbs"): round(work / (ms * 1e-3) / (1e12 if tensor else 1e9), 1)}
    whole = job_flops / (ms_per_step * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": "gemm_bf16_tn_kernel (tcgen05, all epilogues)",
                "achieved": round(achieved, 1), "peak": peak_tf, "unit": "TFLOP/s",
                "frac": round(achieved / peak_tf, 4), "peak_source": f"{how} bf16 sustained",
                "launches": g_n, "avg_launch_ms": round(g_ms / max(g_n, 1), 4),
                "flops_per_launch": g_fl / max(g_n, 1), "share_of_step": round(g_ms / prof_ms, 4),
                "traffic": traffic, "traffic_source": traffic_src,
                "whole_job_frac": round(whole / (peak_tf * world), 4),
                "whole_job_tflops": round(whole, 1), "algorithmic_tflop_per_step": round(job_flops / 1e12, 1),
                "profiled_pass": {"rank": 0, "ms": round(prof_ms, 2), "timed_ms_per_step": round(ms_per_step, 2)},
                "kernels": kernels}

    if rank == 0:
        line = {"metric": "structure_tokens_per_sec", "value": round(value, 1), "unit": "tokens/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": round(ms_per_step, 2), "higher_is_better": True,
                "scaling": "strong" if strong else "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": workload_config(wl, chunks, world, args.rng, args.scaling, counts), "clocks": clk,
                "e2e": {"value": round(e2e_value, 1), "unit": "tokens/s",
                        "h2d_bytes_per_step": int(count * T * 8),
                        "d2h_bytes_per_step": int(tok_host.numel() * 8),
                        "api": "esmdiff_b200.sampling.sample_structure_tokens (pinned host tokens in, host "
                               "int64 tokens out), uniforms=" + args.rng},
                "gpu_launches": int(launches), "roofline": roofline}
        if args.shard_of > 1:
            line["projection"] = (f"rank 0's shard ({count} samples) of a {args.shard_of}-GPU strong-scaling run, measured "
                                  f"alone on one GPU; x{args.shard_of} = {round(value * n_samples / count, 1)} tokens/s if every "
                                  "rank took this long")
        if world == 1 and not args.no_cpu_baseline and args.shard_of <= 1:
            line["cpu_baseline"] = cpu_reference(wl, 3, 1, 24.0, emit_line=False)
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chunks", default="b200", choices=["b200", "reference"],
                    help="b200: batch list sized for 180 GB (one batch of 100 here); reference: the "
                         "reference's 32-80 GB-GPU memory guard, sample_esmdiff.py:181-194 -> [63, 37]")
    ap.add_argument("--rng", default="philox", choices=["philox", "torch"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS),
                    help="config2 (default; the configuration BASELINE.json's metric is quoted on) or config3")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: strong = shard the ONE job's samples over the ranks (default); weak = every "
                         "rank runs the whole job (replicas)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--shard-of", type=int, default=1,
                    help="(study) run only rank 0's share of an N-way strong-scaling split on ONE GPU")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return                                                   # rank 0 alone runs the CPU arm
        cpu_reference(WORKLOADS[args.workload], args.steps, args.warmup, 150.0, emit_line=True)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != world and world == 1 and args.gpus > 1:
        # convenience: relaunch under torchrun when called as plain `python bench.py --gpus N`
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", __file__] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    gpu_bench(args)


if __name__ == "__main__":
    main()
