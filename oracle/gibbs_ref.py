"""TEST INFRASTRUCTURE -- CPU restatement of the esm==3.0.4 iterative structure-track sampler behind
``--mode gibbs`` (reference slm/sample_esmdiff.py:66-130: ``iterative_sampling_raw(esm3_model,
proteins=[ESMProtein(sequence, coordinates)] * bs, configs=[GenerationConfig(track="structure",
num_steps, temperature, top_p)] * bs)``).

PARITY UNPINNED: ``esm`` (requirements.txt:30, esm==3.0.4) is not vendored in /root/reference, not
installed and not fetchable; the functions below restate the published code of
``esm/utils/sampling.py`` (top_p_logits, sample_logits, _compute_track_metadata) and
``esm/utils/generation.py`` (iterative_sampling_tokens,
_get_iterative_sampling_mask_for_prompt_and_step) from memory, anchored on the reference call site
above (its arguments, the defaults it leaves untouched: schedule "cosine", strategy "entropy",
temperature_annealing off, invalid_ids empty) and on torch's own documented behaviour
(``torch.multinomial(p, 1)`` = ``argmax(p / Exp(1))``, ATen MultinomialKernel).  Only tests/,
``__graft_entry__.smoke()`` and bench.py's cpu_baseline may import this module.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

STRUCTURE_MASK = 4096
N_CODES = 4096                     # ids >= 4096 are the special tokens (MASK, EOS, BOS, PAD, CHAINBREAK)


def top_p_logits(logits: torch.Tensor, top_p: float) -> torch.Tensor:
    """esm.utils.sampling.top_p_logits: sorted softmax, cumulative sum <= top_p stays (first always)."""
    batch_dims = logits.size()[:-1]
    logits = logits.reshape(-1, logits.shape[-1]).clone()
    sorted_logits, sorted_indices = torch.sort(logits, dim=-1, descending=True)
    cumsum = sorted_logits.softmax(-1).cumsum(-1)
    keep = cumsum <= top_p
    keep[:, 0] = True
    rows, _ = torch.where(~keep)
    logits[rows, sorted_indices[~keep]] = torch.finfo(logits.dtype).min
    return logits.reshape(*batch_dims, -1)


def filtered_logits(logits: torch.Tensor, top_p: float, n_valid: int = N_CODES) -> torch.Tensor:
    """The row sample_logits draws from: top-p on the raw logits, then invalid ids -> -inf."""
    out = top_p_logits(logits, top_p) if top_p < 1.0 else logits.clone()
    out[..., n_valid:] = -torch.inf
    return out


def sample_and_entropy(logits: torch.Tensor, temperature: float, top_p: float, noise: torch.Tensor | None = None,
                       generator: torch.Generator | None = None):
    """(ids, entropy) per row.  ``noise``: Exp(1) draws of the logits' shape -> argmax(p / noise), the
    arithmetic of torch.multinomial(p, 1); None -> torch.multinomial itself."""
    fl = filtered_logits(logits, top_p)
    probs = F.softmax(fl / temperature, dim=-1)
    flat = probs.reshape(-1, probs.shape[-1])
    if noise is None:
        ids = torch.multinomial(flat, 1, generator=generator).squeeze(1)
    else:
        ids = (flat / noise.reshape(flat.shape)).argmax(-1)
    # _compute_track_metadata: Categorical(probs=exp(log_softmax(filtered logits))).entropy()
    p1 = fl.log_softmax(-1).exp()
    ent = torch.distributions.Categorical(probs=p1).entropy()
    return ids.reshape(logits.shape[:-1]), ent


def cosine_schedule(t: torch.Tensor) -> torch.Tensor:
    return torch.cos(t * math.pi * 0.5)


def num_to_unmask(step: int, num_steps: int, total_to_sample: int, still_masked: int) -> int:
    """_get_iterative_sampling_mask_for_prompt_and_step: positions revealed at 0-based ``step``."""
    perc = cosine_schedule(torch.tensor((step + 1) / num_steps))
    after = int((perc * torch.tensor(total_to_sample) + 0.1).int())
    if step + 1 == num_steps:
        after = 0
    return still_masked - after


def unmask_schedule(num_steps: int, total_to_sample: int) -> list[int]:
    num_steps = min(num_steps, total_to_sample) if total_to_sample > 0 else num_steps
    ks, still = [], total_to_sample
    for t in range(num_steps):
        k = max(num_to_unmask(t, num_steps, total_to_sample, still), 0)
        ks.append(k)
        still -= k
    return ks


def gibbs_step(x: torch.Tensor, logits: torch.Tensor, k: int, temperature: float, top_p: float,
               noise: torch.Tensor | None = None, generator=None) -> torch.Tensor:
    """One decoding step on (B, T) tokens: sample everywhere, keep the k lowest-entropy masked positions."""
    ids, ent = sample_and_entropy(logits, temperature, top_p, noise, generator)
    B, T = x.shape
    mask = x == STRUCTURE_MASK
    mask[:, 0] = False
    mask[:, -1] = False
    out = x.clone()
    for b in range(B):
        e = ent[b].masked_fill(~mask[b], torch.inf)
        kk = min(k, int(mask[b].sum()))
        if kk <= 0:
            continue
        # lowest entropy first, ties by position (torch.topk leaves tie order unspecified)
        order = sorted(range(T), key=lambda i: (float(e[i]), i))[:kk]
        idx = torch.tensor(order)
        out[b, idx] = ids[b, idx]
    return out


def iterative_sampling_structure(forward, seq: torch.Tensor, prior: torch.Tensor, num_steps: int, temperature: float,
                                 top_p: float, generator=None) -> torch.Tensor:
    """``forward(seq, x) -> logits (B, T, V)``; prior (B, T) with MASK at the positions to sample."""
    x = prior.clone()
    inner = x[:, 1:-1] == STRUCTURE_MASK
    total = int(inner[0].sum())
    for t, k in enumerate(unmask_schedule(num_steps, total)):
        x = gibbs_step(x, forward(seq, x), k, temperature, top_p, generator=generator)
    return x
