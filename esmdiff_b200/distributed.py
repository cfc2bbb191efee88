"""Multi-GPU sharding of independent conformation samples.

Samples are i.i.d. given the sequence (the reference just ``repeat``s one row,
sample_esmdiff.py:186,190), so the path shards with no per-step communication: one process per
GPU, contiguous ranges of samples per rank, one ``broadcast`` of the weights from rank 0 at start
and one ``all_gather`` of the final int64 tokens.  Uniform streams: each rank seeds
``base_seed + first_sample_index`` before sampling (the documented multi-GPU contract; a single
global generator walked across ranks would serialise them).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None):
    """torchrun-style env (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*).  Returns (rank, world, local)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def shard_samples(num_samples: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous [start, start+count) of the sample index range owned by ``rank``."""
    base, rem = divmod(num_samples, world)
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def broadcast_state_dict(sd: dict | None, device, src: int = 0) -> dict:
    """Rank ``src`` holds the weights; everyone gets them with one flat broadcast per dtype."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return sd
    meta = [None]
    if dist.get_rank() == src:
        meta[0] = [(k, tuple(v.shape), v.dtype) for k, v in sd.items()]
    dist.broadcast_object_list(meta, src=src)
    total = sum(int(torch.tensor(s).prod()) if len(s) else 1 for _, s, _ in meta[0])
    flat = torch.empty(total, dtype=torch.float32, device=device)
    if dist.get_rank() == src:
        off = 0
        for k, s, _ in meta[0]:
            n = sd[k].numel()
            flat[off:off + n] = sd[k].reshape(-1).to(device, torch.float32)
            off += n
    dist.broadcast(flat, src=src)
    out, off = {}, 0
    for k, s, dt in meta[0]:
        n = 1
        for z in s:
            n *= z
        out[k] = flat[off:off + n].view(s)
        off += n
    return out


def gather_tokens(local_tokens: torch.Tensor, counts: list[int]) -> torch.Tensor:
    """all_gather of ragged (count_r, L) int64 token blocks -> (sum counts, L) on every rank."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local_tokens
    L = local_tokens.shape[1]
    cap = max(counts)
    pad = torch.zeros(cap, L, dtype=torch.int64, device=local_tokens.device)
    pad[: local_tokens.shape[0]] = local_tokens
    bufs = [torch.empty_like(pad) for _ in counts]
    dist.all_gather(bufs, pad)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)


def max_over_ranks(value: float, device=None) -> float:
    """MAX all-reduce of one float (device-timed milliseconds in bench.py)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64,
                     device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64,
                     device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
