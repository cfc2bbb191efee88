"""BASELINE config 5: batch sweep L in {128, 256, 512, 1024} x num_samples in {1, 8, 64, 512},
25 steps + noise removal, random-init ESM3-open-sized weights -> roofline table.

    gpurun -- python tools/sweep.py > profiles/<round>_sweep.md     (one GPU, ~3 min)
    gpurun --gpus 8 -- python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29541 tools/sweep.py > profiles/<round>_sweep_n8.md
Under torchrun the samples of every cell are sharded over the ranks exactly as bench.py / the CLI do
(distributed.shard_samples; ranks without a sample idle), the time is the max over ranks and the
tokens are all-gathered inside the timed window; per-kernel columns are rank 0's.

Per cell: structure-tokens/s (device-resident inputs, CUDA events, after a 2-step warm run of the cell),
whole-job TFLOP/s on the algorithmic FLOPs of SURVEY.md 8(d) and its fraction of the measured
sustained bf16 peak, share of the time in the tcgen05 GEMMs / attention, HBM GB/s of the sampling
kernel.  The reference's CPU path beside it is bench.py's cpu_baseline (one number per box)."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from esmdiff_b200 import distributed as D  # noqa: E402
from esmdiff_b200.engine import Dims, Engine  # noqa: E402
from esmdiff_b200.sampling import chunk_sizes_b200  # noqa: E402
from esmdiff_b200.synthetic import random_state_dict  # noqa: E402
from esmdiff_b200.tokenization import synthetic_sequence_tokens  # noqa: E402


def fwd_flops(B, T):
    return float(B) * T * (48 * (56_623_104 + 6144 * T) + 17_316_864)


def main():
    import os
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    rank, world, local = D.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    out = print if rank == 0 else (lambda *a, **k: None)
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    dims = Dims()
    eng = Engine(dims, device=local)
    eng.load_state_dict(D.broadcast_state_dict(random_state_dict(dims, device=dev, seed=0) if rank == 0 else None, dev))
    steps = 25
    sched = eng.schedule(steps)
    out(f"B200 x{world} sweep, {steps} steps + noise removal, bf16 tcgen05 path; peak = {peak} TFLOP/s per GPU (measured "
        f"sustained bf16); samples of a cell sharded over the {world} GPU(s)\n")
    out("| L | samples | batches (rank 0) | ms | tokens/s | TFLOP/s | % of peak (all GPUs) | GEMM share | attention share | GEMM TFLOP/s | attention TFLOP/s | sampler GB/s |")
    out("|---:|---:|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    for L in (128, 256, 512, 1024):
        T = L + 2
        seq = synthetic_sequence_tokens(L, seed=0).to(dev)
        for N in (1, 8, 64, 512):
            spans = [D.shard_samples(N, world, r) for r in range(world)]
            first, count = spans[rank]
            counts = [c for _, c in spans]
            chunks = chunk_sizes_b200(T, count) if count else []

            def job(seed):
                outs = [eng.ddpm_sample(seq[None].expand(b, T).contiguous(), None, steps, *sched, seed=seed + first + i)
                        for i, b in enumerate(chunks)]
                tok = torch.cat(outs) if outs else torch.empty(0, T, dtype=torch.int64, device=dev)
                return D.gather_tokens(tok, counts)

            # warm every cell with a 2-step run: the first call at a larger B*T re-allocates the library
            # workspace (cudaFree + cudaMalloc of GBs: seconds of host time), builds TMA descriptors and,
            # for small batches, captures the CUDA graph -- one-time costs, not throughput
            warm = eng.schedule(2)
            for i, b in enumerate(chunks):
                eng.ddpm_sample(seq[None].expand(b, T).contiguous(), None, 2, *warm, seed=i)
            torch.cuda.synchronize()
            if world > 1:
                torch.distributed.barrier()
            eng.profile(True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            tok = job(7)
            e1.record()
            torch.cuda.synchronize()
            eng.synchronize()
            eng.profile(False)
            ms = D.max_over_ranks(e0.elapsed_time(e1), dev)
            prof = eng.profile_read()
            assert tok.shape == (N, T) and int((tok == 4096).sum()) == 0
            flops = D.sum_over_ranks((steps + 1) * sum(fwd_flops(b, T) for b in chunks), dev)
            if rank != 0:
                continue
            g = [prof[k] for k in prof if k.startswith("gemm")]
            g_ms, g_fl = sum(x[0] for x in g), sum(x[1] for x in g)
            a_ms, a_fl, _ = prof["attention"]
            s_ms, s_b, _ = prof["sampler"]
            print(f"| {L} | {N} | {chunks if len(chunks) < 4 else str(len(chunks)) + ' x ' + str(chunks[0])} | {ms:.1f} | "
                  f"{N * L / ms * 1e3:.0f} | {flops / ms / 1e9:.0f} | {100 * flops / ms / 1e9 / (peak * world):.1f} | "
                  f"{100 * g_ms / ms:.1f} | {100 * a_ms / ms:.1f} | {g_fl / g_ms / 1e9:.0f} | {a_fl / a_ms / 1e9:.0f} | "
                  f"{s_b / s_ms / 1e6:.0f} |", flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()
