"""CPU restatement of the MDLM sampler half of the ddpm path (ORACLE -- tests only).

Follows, operation for operation (same torch ops in the same order, so results are
bit-identical to the reference on CPU -- checked by ``oracle/make_golden.py`` against the
reference's own code and frozen in ``tests/golden/sampler_*.npz``):

* ``LogLinearNoise.total_noise``          reference slm/utils/noise_utils.py:205-206
* ``logits_parameterization``             reference slm/models/model.py:527-533
* ``_sample_categorical``                 reference slm/models/model.py:24-28
* ``_ddpm_update`` (move chances + tail)  reference slm/models/model.py:583-607
* ``ddpm_sample`` (time grid, loop, noise removal)   reference slm/models/model.py:543-581
* ``_model_wrapper`` (time conditioning)  reference slm/models/model.py:464-492
* chunk list of ``ddpm_sample_by_esm``    reference slm/sample_esmdiff.py:181-194
* inpainting prior                        reference slm/sample_esmdiff.py:197-201
"""
from __future__ import annotations

import torch

MASK = 4096
NEG_INF = -1000000.0
N_MAX_RESIDUE_SQUARE = 200 * 200 * 105      # sample_esmdiff.py:146


def total_noise(t: torch.Tensor, eps: float = 1e-3) -> torch.Tensor:
    return -torch.log1p(-(1 - eps) * t)


def move_chances(t: torch.Tensor, dt: float, eps: float = 1e-3):
    """t: (B,1) fp32.  Returns (sigma_t (B,), mc_t (B,1,1), mc_s (B,1,1))."""
    sigma_t = total_noise(t, eps).squeeze(-1)
    sigma_s = total_noise(t - dt, eps).squeeze(-1)
    mc_t = (1 - torch.exp(-sigma_t))[:, None, None]
    mc_s = (1 - torch.exp(-sigma_s))[:, None, None]
    return sigma_t, mc_t, mc_s


def logits_parameterization(logits: torch.Tensor, xt: torch.Tensor) -> torch.Tensor:
    """Mutates ``logits`` in place exactly as the reference does, returns log p(x0)."""
    logits[:, :, MASK] += NEG_INF
    logits = logits - torch.logsumexp(logits, dim=-1, keepdim=True)
    keep = xt != MASK
    logits[keep] = NEG_INF
    logits[keep, xt[keep]] = 0
    return logits


def race_argmax(q: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
    """``_sample_categorical`` with the uniforms passed in (reference draws
    ``u = torch.rand_like(q)``)."""
    g = 1e-10 - (u + 1e-10).log()
    return (q / g).argmax(dim=-1)


def ddpm_update_tail(log_p_x0, x, mc_t, mc_s, u):
    q = log_p_x0.exp() * (mc_t - mc_s)
    q[:, :, MASK] = mc_s[:, :, 0]
    cand = race_argmax(q, u)
    keep = (x != MASK).to(x.dtype)
    return keep * x + (1 - keep) * cand


def time_grid(num_steps: int, eps: float = 1e-5, sample_max_t: float = 1.0):
    ts = torch.linspace(sample_max_t, eps, num_steps + 1)
    dt = (1 - eps) / num_steps
    return ts, dt


def chunk_sizes(T: int, num_samples: int, cap: int = N_MAX_RESIDUE_SQUARE) -> list[int]:
    """Batch list of ``ddpm_sample_by_esm``; T = tokens incl. BOS/EOS."""
    target = T * T * num_samples
    n_full = target // cap
    per = cap // int(T * T)
    sizes = [per] * n_full
    if target % cap > 0:
        sizes.append(num_samples - sum(sizes))
    assert sum(sizes) == num_samples, f"{sum(sizes)} != {num_samples}"
    return sizes


def inpainting_prior(structure_tokens: torch.Tensor, batch: int, mask_ids) -> torch.Tensor:
    """``input_prior`` of sample_esmdiff.py:197-201.  ``mask_ids`` index TOKEN positions here
    (BOS is position 0) although the same ids index residues when the sequence is masked
    (models/utils.py:117-123) -- the reference's off-by-one is reproduced, not fixed."""
    prior = structure_tokens[None, :].repeat(batch, 1)
    for idx in mask_ids:
        prior[:, idx] = MASK
    return prior


class SamplerRef:
    """ddpm_sample driven by any ``net(structure_tokens=, sequence_tokens=,
    auxiliary_embeddings=, labels=None)`` callable and a sigma embedder."""

    def __init__(self, net, sigma_embedder, time_conditioning=True, noise_removal=True,
                 noise_eps: float = 1e-3, record=None, uniform_fn=None, device="cpu"):
        self.device = torch.device(device)     # the reference's ``self.device`` (model.py:555,561)
        self.net, self.sigma_embedder = net, sigma_embedder
        self.time_conditioning, self.noise_removal = time_conditioning, noise_removal
        self.noise_eps = noise_eps
        self.record = record                  # optional list receiving per-step dicts
        self.uniform_fn = uniform_fn or torch.rand_like

    def log_p_x0(self, xt, sequence_tokens, sigma):
        if not self.time_conditioning:
            sigma = torch.zeros_like(sigma)
        cond = self.sigma_embedder(sigma.to(self.device, torch.float32))
        cond = torch.tile(cond[:, None, :], (1, xt.shape[1], 1))
        out = self.net(structure_tokens=xt, sequence_tokens=sequence_tokens,
                       auxiliary_embeddings=cond, labels=None)
        raw = out.structure_logits
        rec = raw.clone() if self.record is not None else None
        return logits_parameterization(raw, xt), rec

    @torch.no_grad()
    def ddpm_sample(self, sequence_tokens, num_steps, eps=1e-5, input_prior=None,
                    sample_max_t=1.0):
        if input_prior is None:
            x = (MASK * torch.ones(*sequence_tokens.shape, dtype=torch.int64)).to(self.device)
            assert sample_max_t == 1.0
        else:
            x = input_prior.clone().to(self.device)
            assert x.shape == sequence_tokens.shape
        sequence_tokens = sequence_tokens.to(self.device)
        ts, dt = time_grid(num_steps, eps, sample_max_t)
        for i in range(num_steps):
            t = (ts[i] * torch.ones(x.shape[0], 1)).to(self.device)
            sigma_t, mc_t, mc_s = move_chances(t, dt, self.noise_eps)
            logp, raw = self.log_p_x0(x, sequence_tokens, sigma_t)
            u = self.uniform_fn(logp)       # same shape/dtype as q_xs (model.py:25-27)
            x_next = ddpm_update_tail(logp, x, mc_t, mc_s, u)
            if self.record is not None:
                self.record.append(dict(step=i, x_t=x.clone(), raw_logits=raw, u=u,
                                        mc_t=float(mc_t[0, 0, 0]), mc_s=float(mc_s[0, 0, 0]),
                                        sigma_t=float(sigma_t[0]), x_next=x_next.clone()))
            x = x_next
        if self.noise_removal:
            t = (ts[-1] * torch.ones(x.shape[0], 1)).to(self.device)
            sigma_t = total_noise(t, self.noise_eps).squeeze(-1)
            logp, raw = self.log_p_x0(x, sequence_tokens, sigma_t)
            x_final = logp.argmax(dim=-1)
            if self.record is not None:
                self.record.append(dict(step=num_steps, x_t=x.clone(), raw_logits=raw,
                                        sigma_t=float(sigma_t[0]), x_next=x_final.clone()))
            x = x_final
        return x


def race_top2_gap(log_p_x0, mc_t, mc_s, u):
    """Relative gap between the best and second-best race score q/g per row, (B,T) fp64.
    Parity harness only: a CUDA/CPU token-id mismatch is excused only on rows where this gap is
    below 1e-5, i.e. where 1-2 ulp of libm difference in exp/log can swap the argmax
    (SURVEY.md 7 "libm parity in K-sample")."""
    q = log_p_x0.exp() * (mc_t - mc_s)
    q[:, :, MASK] = mc_s[:, :, 0]
    g = 1e-10 - (u + 1e-10).log()
    top = (q / g).double().topk(2, dim=-1).values
    return (top[..., 0] - top[..., 1]) / top[..., 0].clamp_min(1e-300)


def assert_ids_match(got, want, log_p_x0, mc_t, mc_s, u, rtol: float = 1e-5):
    """Bit-exact except on documented near-ties.  Returns the number of excused rows."""
    bad = got != want
    if not bool(bad.any()):
        return 0
    gap = race_top2_gap(log_p_x0, mc_t, mc_s, u)
    hard = bad & (gap >= rtol)
    assert not bool(hard.any()), (
        f"{int(hard.sum())} token ids differ from the oracle on rows that are not near-ties "
        f"(min gap {float(gap[bad].min()):.3e})")
    return int(bad.sum())
