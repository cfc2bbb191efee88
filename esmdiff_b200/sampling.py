"""Token-level sampling driver: what ``ddpm_sample_by_esm`` (reference
slm/sample_esmdiff.py:137-233) does between ``start_t = time()`` (:177) and
"Sampling token time" (:223) -- chunk list, inpainting prior, sampler calls, concat, BOS/EOS strip.
"""
from __future__ import annotations

from time import time

import torch

from .tokenization import STRUCTURE_MASK

N_MAX_RESIDUE_SQUARE = 200 * 200 * 105      # sample_esmdiff.py:146


def chunk_sizes(T: int, num_samples: int, n_max_residue_square: int = N_MAX_RESIDUE_SQUARE) -> list[int]:
    """The reference's batch list (sample_esmdiff.py:181-194): ``cap // T^2`` samples per full
    chunk, one residual chunk with whatever is left (which can exceed the per-chunk size: e.g.
    T=1026, N=512 -> 128 x [3] + [128]).  Kept verbatim in behaviour because the chunk list fixes
    how the uniform stream is consumed."""
    target = T * T * num_samples
    n_batch = target // n_max_residue_square
    batch_size = n_max_residue_square // int(T * T)
    bsz = [batch_size] * n_batch
    if target % n_max_residue_square > 0:
        bsz.append(num_samples - sum(bsz))
    assert sum(bsz) == num_samples, f"{sum(bsz)} != {num_samples}"
    return bsz


MAX_TOKENS_PER_BATCH = 1 << 18              # B*T rows per ddpm_sample call on a 180 GB B200


def chunk_sizes_b200(T: int, num_samples: int, max_tokens: int = MAX_TOKENS_PER_BATCH) -> list[int]:
    """Batch list sized for 180 GB of HBM instead of the reference's ``B*T^2 <= 4.2M`` guard (a
    32-80 GB-GPU memory heuristic): all samples of a target in one batch while B*T stays below
    ``max_tokens`` (2^18 rows = 4.3 GB of fp32 logits + ~3 GB of activations), else equal chunks.
    Samples are i.i.d. given the sequence, so the chunking does not change what is sampled; it
    only changes how a *torch* uniform stream would be consumed, which is why the torch-RNG
    parity mode keeps :func:`chunk_sizes`."""
    per = max(1, max_tokens // T)
    n_chunks = -(-num_samples // per)
    base, rem = divmod(num_samples, n_chunks)
    return [base + (1 if i < rem else 0) for i in range(n_chunks)]


def build_prior(structure_tokens: torch.Tensor, batch: int, mask_ids=None, filled_ids=None,
                total_size=None):
    """``input_prior`` (sample_esmdiff.py:197-209).  ``mask_ids`` index TOKEN positions (BOS = 0),
    exactly as the reference does."""
    if mask_ids is not None:
        prior = structure_tokens[None, :].repeat(batch, 1)
        for idx in mask_ids:
            prior[:, idx] = STRUCTURE_MASK
        return prior
    if filled_ids is not None:
        prior = structure_tokens[None, :].repeat(batch, 1)
        for idx in range(total_size):
            if idx not in filled_ids:
                prior[:, idx] = STRUCTURE_MASK
        return prior
    return None


@torch.no_grad()
def sample_structure_tokens(model, sequence_tokens_singleton: torch.Tensor, num_samples: int,
                            num_steps: int, eps: float = 1e-5, structure_tokens=None, mask_ids=None,
                            filled_ids=None, total_size=None, sample_max_t: float = 1.0,
                            chunks: list[int] | None = None, verbose: bool = True):
    """Returns (tokens int64 (num_samples, L) without BOS/EOS, seconds) -- the metric's window."""
    T = sequence_tokens_singleton.size(0)
    start_t = time()
    if chunks is None:
        # torch-RNG parity mode consumes the uniform stream chunk by chunk exactly like the
        # reference; the library-Philox mode is chunking-independent and batches for the B200
        chunks = (chunk_sizes(T, num_samples) if getattr(model, "rng", "torch") == "torch"
                  else chunk_sizes_b200(T, num_samples))
    bsz = chunks
    if verbose:
        print(f"Total {num_samples} samples will be generated in batchs {bsz}...")
    outs = []
    for bs in bsz:
        batch = sequence_tokens_singleton[None, :].repeat(bs, 1)
        prior = build_prior(structure_tokens, bs, mask_ids, filled_ids, total_size)
        outs.append(model.ddpm_sample(num_steps=num_steps, sequence_tokens=batch, eps=eps,
                                      input_prior=prior, sample_max_t=sample_max_t))
    tokens = torch.cat(outs, dim=0)[:, 1:-1]
    if tokens.is_cuda:
        torch.cuda.synchronize(tokens.device)
    elapsed = time() - start_t
    if verbose:
        print(f"Sampling token time: {elapsed:.2f}s")
    return tokens, elapsed


@torch.no_grad()
def sample_structure_tokens_sharded(model, sequence_tokens_singleton: torch.Tensor, num_samples: int,
                                    num_steps: int, rank: int = 0, world: int = 1, seed: int | None = None,
                                    **kw):
    """The same job over ``world`` ranks (one process per GPU, torchrun): samples are i.i.d. given the
    sequence (the reference ``repeat``s one row, sample_esmdiff.py:186,190), so rank r samples its
    contiguous share ``distributed.shard_samples(num_samples, world, r)`` with no per-step
    communication and the int64 tokens are all-gathered once at the end.  Every rank
    returns the full (num_samples, L) tensor and the max-over-ranks window time.
    RNG contract: ``seed`` given -> each rank seeds torch with ``seed + first_sample_index`` before
    sampling (identical tokens for any later re-run with the same world size); None -> whatever
    state the process's generator is in, as in the reference."""
    from . import distributed as D
    spans = [D.shard_samples(num_samples, world, r) for r in range(world)]
    first, count = spans[rank]
    if seed is not None:
        torch.manual_seed(int(seed) + first)
    L = sequence_tokens_singleton.size(0) - 2
    if count > 0:
        tokens, elapsed = sample_structure_tokens(model, sequence_tokens_singleton, count, num_steps, **kw)
    else:
        dev = getattr(model, "device", "cpu")
        tokens, elapsed = torch.empty(0, L, dtype=torch.int64, device=dev), 0.0
    if world > 1:
        tokens = D.gather_tokens(tokens, [c for _, c in spans])
        elapsed = D.max_over_ranks(elapsed, tokens.device if tokens.is_cuda else None)
    return tokens, elapsed
