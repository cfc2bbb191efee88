"""Property tests (hypothesis) of the host logic around the path: the reference's chunk arithmetic, the B200 batch
list, the multi-GPU sample split, esm's cosine unmasking schedule and the vectorised PDB number formatting.  CPU only."""
import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from esmdiff_b200.decoder import _fixed_width, _int_width
from esmdiff_b200.distributed import shard_samples
from esmdiff_b200.gibbs import gibbs_chunk_sizes, unmask_schedule
from esmdiff_b200.sampling import build_prior, chunk_sizes, chunk_sizes_b200

CAP = 200 * 200 * 105


def _reference_chunks(T, N, cap=CAP):
    """sample_esmdiff.py:181-194 restated literally (the residual chunk takes whatever is left); None where the
    reference's own assert fires (target a multiple of the cap with samples left over)."""
    target = T * T * N
    n_batch, residual, bs = target // cap, target % cap, cap // (T * T)
    bsz = [bs] * n_batch
    if residual > 0:
        bsz.append(N - sum(bsz))
    return bsz if sum(bsz) == N else None


@settings(max_examples=300, deadline=None)
@given(T=st.integers(3, 2050), N=st.integers(1, 1024))
def test_chunk_list_is_the_reference_arithmetic(T, N):
    want = _reference_chunks(T, N)
    if want is None:
        with pytest.raises(AssertionError):
            chunk_sizes(T, N)
        return
    got = chunk_sizes(T, N)
    assert got == want and sum(got) == N
    if len(got) > 1:
        assert all(b == got[0] for b in got[:-1]) and got[0] * T * T <= CAP      # full chunks respect the memory guard
    assert sum(gibbs_chunk_sizes(T, N)) == N or _reference_chunks(T, N) is None


@settings(max_examples=200, deadline=None)
@given(T=st.integers(3, 4100), N=st.integers(1, 5000))
def test_b200_batch_list_partitions_evenly(T, N):
    got = chunk_sizes_b200(T, N)
    assert sum(got) == N and max(got) - min(got) <= 1 and min(got) >= 1
    assert max(got) * T <= max(1 << 18, T)                    # at most 2^18 token rows per call (one sample always fits)


@settings(max_examples=200, deadline=None)
@given(N=st.integers(0, 5000), world=st.integers(1, 16))
def test_shard_samples_is_a_contiguous_partition(N, world):
    spans = [shard_samples(N, world, r) for r in range(world)]
    assert spans[0][0] == 0 and sum(c for _, c in spans) == N
    for (s0, c0), (s1, _) in zip(spans, spans[1:]):
        assert s1 == s0 + c0
    counts = [c for _, c in spans]
    assert max(counts) - min(counts) <= 1 and counts == sorted(counts, reverse=True)


@settings(max_examples=200, deadline=None)
@given(steps=st.integers(1, 64), total=st.integers(0, 1100))
def test_unmask_schedule_reveals_everything_once(steps, total):
    ks = unmask_schedule(steps, total)
    assert len(ks) == (min(steps, total) if total > 0 else steps)
    assert all(k >= 0 for k in ks) and sum(ks) == total
    # the number still masked follows esm's cosine: non-increasing, zero after the last step
    left = total - np.cumsum(ks)
    assert (np.diff(left) <= 0).all() and (left[-1] == 0 if len(ks) else True)


@settings(max_examples=100, deadline=None)
@given(L=st.integers(3, 40), B=st.integers(1, 5), data=st.data())
def test_build_prior_masks_token_positions(L, B, data):
    ids = data.draw(st.lists(st.integers(0, L + 1), unique=True, max_size=L))
    tok = torch.arange(L + 2)
    prior = build_prior(tok, B, mask_ids=ids)
    assert prior.shape == (B, L + 2)
    for p in range(L + 2):
        assert bool((prior[:, p] == (4096 if p in ids else p)).all())
    assert build_prior(tok, B) is None


f32 = st.floats(min_value=-999.0, max_value=9999.0, allow_nan=False, width=32)


@settings(max_examples=300, deadline=None)
@given(xs=st.lists(f32, min_size=1, max_size=50))
def test_fixed_width_formatting_is_pythons(xs):
    a = np.array(xs, dtype=np.float32)
    got = _fixed_width(a, 8, 3)
    for row, v in zip(got, a):
        assert bytes(row).decode() == f"{float(v):8.3f}", (float(v), bytes(row))
    b = np.clip(np.abs(a), 0, 99.0).astype(np.float32)
    for row, v in zip(_fixed_width(b, 6, 2), b):
        assert bytes(row).decode() == f"{float(v):6.2f}"
    n = np.abs(a).astype(np.int64) % 10000
    for row, v in zip(_int_width(n, 5), n):
        assert bytes(row).decode() == f"{int(v):5d}"
