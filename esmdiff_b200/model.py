"""``MaskedDiffusionLanguageModeling`` for the B200 path: the reference's sampler surface
(slm/models/model.py:316-383 constructor, :464-492 ``_model_wrapper``, :527-533
``logits_parameterization``, :543-607 ``ddpm_sample`` / ``_ddpm_update``) with the arithmetic in
``libesmdiff_b200.so``.

Differences that do not change results:
  * the time embedding is evaluated once per step, not once per sample (all rows share sigma);
  * logits_parameterization + q_xs + Gumbel-race argmax + token blend are ONE kernel and rows that
    are already unmasked are skipped (their update is the identity, model.py:606-607);
  * the whole schedule (sigma, move chances) is evaluated before the loop with the reference's
    own torch ops, so nothing synchronises the stream inside the loop.
Uniforms: ``rng="torch"`` draws ``torch.rand(B, T, 4101)`` per step from torch's CUDA generator --
the same call, shape and order as the reference's ``torch.rand_like(q_xs)`` (model.py:25-27), so
with ``torch.manual_seed(s)`` both consume the same stream.  ``rng="philox"`` uses the library's
counter-based generator inside the sampling kernel (no 4101-wide uniform tensor in HBM); its key
is drawn once per ``ddpm_sample`` call from torch's CPU generator.
"""
from __future__ import annotations

import torch
from torch import nn

from . import noise_utils
from .engine import MASK
from .net import CustomizedESM3, TimestepEmbedder


def _sample_categorical(categorical_probs, engine=None):
    """Reference model.py:24-28 (Gumbel-max written as a ratio race).  Provided for API parity;
    the fused path never materialises ``categorical_probs``."""
    g = 1e-10 - (torch.rand_like(categorical_probs) + 1e-10).log()
    return (categorical_probs / g).argmax(dim=-1)


class MaskedDiffusionLanguageModeling(nn.Module):
    def __init__(self, net: CustomizedESM3 = None, optimizer=None, scheduler=None, compile=False,
                 noise_schedule: noise_utils.Noise = None, sigma_embedder: nn.Module = None,
                 time_conditioning: bool = False, change_of_variables: bool = False,
                 importance_sampling: bool = False, condition_dropout: float = 0.0,
                 condition_mask_rate: float = 0.5, sequence_prediction: bool = False, T: int = 0,
                 sampling_eps: float = 1e-3, antithetic_sampling: bool = True,
                 noise_removal: bool = True, structure_only: bool = False,
                 coupled_condition_mask: bool = False, rng: str = "torch", **unused):
        super().__init__()
        if noise_schedule is None:
            print("Using default noise schedule: CosineNoise(eps=1e-3)")
            noise_schedule = noise_utils.CosineNoise(eps=1e-3)
        if sequence_prediction:
            raise NotImplementedError("sequence_prediction is off on the ddpm path (mdlm.yaml:45)")
        assert not (change_of_variables and importance_sampling)
        assert rng in ("torch", "philox")
        self.net = net
        self.noise = noise_schedule
        self.sigma_embedder = sigma_embedder
        self.time_conditioning = time_conditioning
        self.T = T
        self.sampling_eps = sampling_eps
        self.noise_removal = noise_removal
        self.antithetic_sampling = antithetic_sampling
        self.change_of_variables = change_of_variables
        self.importance_sampling = importance_sampling
        self.condition_dropout = condition_dropout
        self.condition_mask_rate = condition_mask_rate
        self.structure_only = structure_only
        self.coupled_condition_mask = coupled_condition_mask
        self.condition_mask_index = 32                 # C.SEQUENCE_MASK_TOKEN
        self.sequence_prediction = False
        self.rng = rng
        self.vocab_size = 4101
        self.mask_index = MASK
        self.neg_infinity = -1000000.0
        if net is not None:
            net.engine.set_time_conditioning(time_conditioning)      # the library zeroes sigma when off (model.py:538-539)

    # -- module plumbing ------------------------------------------------------------------------
    @property
    def device(self):
        return self.net.device

    @property
    def engine(self):
        return self.net.engine

    def to(self, *a, **k):
        if self.sigma_embedder is not None:
            self.sigma_embedder.to(self.device)
        return self

    def load_state_dict(self, state_dict, strict=True):
        """Keys of the DeepSpeed ['module'] dict: ``net.*`` and ``sigma_embedder.*``."""
        te = {k[len("sigma_embedder."):]: v for k, v in state_dict.items() if k.startswith("sigma_embedder.")}
        if self.sigma_embedder is not None and te:
            self.sigma_embedder.load_state_dict(te, strict=strict)
        if strict:
            unexpected = [k for k in state_dict if not k.startswith(("net.", "sigma_embedder."))]
            if unexpected:
                raise RuntimeError(f"Error(s) in loading state_dict: Unexpected key(s): {unexpected[:5]}")
        self.engine.load_state_dict(state_dict, strict=strict)
        return self

    # -- reference-shaped pieces ------------------------------------------------------------------
    def _process_sigma(self, sigma):
        if sigma.ndim > 1:
            sigma = sigma.squeeze(-1)
        if not self.time_conditioning:
            sigma = torch.zeros_like(sigma)
        assert sigma.ndim == 1, sigma.shape
        return sigma

    def _sample_prior(self, *batch_dims):
        return self.mask_index * torch.ones(*batch_dims, dtype=torch.int64)

    def logits_parameterization(self, logits, xt):
        """In place, like the reference (model.py:527-533)."""
        return self.engine.logits_parameterization(logits, xt.to(self.device).contiguous())

    def _model_wrapper(self, xt, sequence_tokens=None, sigma=None, shield_special_tokens=False):
        """log p(x0 | xt): time embedding + forward + logits_parameterization (model.py:464-492).
        ``sigma``: tensor (B,) / (B,1) with one value per sample; the path shares it across the
        batch (ddpm_sample builds it as a constant vector, model.py:571-572)."""
        xt = xt.to(self.device).contiguous()
        if sequence_tokens is None:
            sequence_tokens = torch.full_like(xt, 32)      # sequence mask id, the reference's default (net.py:411)
        if sigma is not None:
            sigma = self._process_sigma(torch.as_tensor(sigma))
            s0 = float(sigma.reshape(-1)[0])
            if sigma.numel() > 1 and not bool((sigma == sigma.reshape(-1)[0]).all()):
                cond = self.sigma_embedder(sigma.to(self.device, torch.float32))      # per-sample sigma
                cond = cond[:, None, :].expand(-1, xt.shape[1], -1)
                logits, _ = self.engine.forward(sequence_tokens, xt, aux=cond)
            else:
                logits = self.engine.forward_sigma(sequence_tokens, xt, s0)
        else:
            logits, _ = self.engine.forward(sequence_tokens, xt, aux=None)
        logits = self.engine.logits_parameterization(logits, xt)
        if shield_special_tokens:
            logits[..., 4096:4101] += self.neg_infinity
        return logits, None

    # -- training-side reuse, forward half ----------------------------------------
    def q_xt(self, x, move_chance, condition_seq=None, non_moving_mask=None):
        """Noisy sample x_t (reference model.py:494-513): every token moves to MASK with its sample's chance."""
        move_indices = torch.rand(*x.shape, device=x.device) < move_chance
        if non_moving_mask is not None:
            move_indices = move_indices & (~non_moving_mask)
        xt = torch.where(move_indices, self.mask_index, x)
        if self.coupled_condition_mask and condition_seq is not None:
            condition_seq = torch.where(move_indices, self.condition_mask_index, condition_seq)
        return xt, condition_seq

    def _sample_t(self, n, device):
        """reference model.py:518-526 (antithetic: one stratum per sample)."""
        _eps_t = torch.rand(n, device=device)
        if self.antithetic_sampling:
            offset = torch.arange(n, device=device) / n
            _eps_t = (_eps_t / n + offset) % 1
        t = (1 - self.sampling_eps) * _eps_t + self.sampling_eps
        if self.importance_sampling:
            return self.noise.importance_sampling_transformation(t)
        return t

    @torch.no_grad()
    def model_step(self, batch, training=False):
        """The diffusion NELBO of one batch (reference model.py:386-462) with the forward on the CUDA path:
        ``batch`` = {"structure_tokens", "sequence_tokens", "mask"[, "non_moving_mask"]} int64 / bool (B, L).
        Forward half only -- what ``validation_step`` / ``test_step`` need (model.py:157-196 call it and log the
        loss); there is no backward on this path, so ``training=True`` only switches the reference's condition
        dropout / masking on.  Random draws (t, condition mask, q_xt) are made on ``batch``'s device with torch's
        generator in the reference's order, so a seeded CPU batch sees the reference's own noise."""
        from random import random
        labels = batch["structure_tokens"].detach().clone()
        x0 = batch["structure_tokens"].detach().clone()
        condition_seq = batch["sequence_tokens"]
        if self.condition_dropout > 0 and training:
            if random() < self.condition_dropout:
                condition_seq = None
        if self.condition_mask_rate > 0 and condition_seq is not None and training:
            mask = (torch.rand_like(condition_seq, dtype=torch.float) < self.condition_mask_rate) & (condition_seq != 1)
            condition_seq = torch.where(mask, self.condition_mask_index, condition_seq)
        loss_mask = batch["mask"] * (labels != 4099)              # C.STRUCTURE_PAD_TOKEN
        t = self._sample_t(x0.shape[0], x0.device)
        if self.T > 0:
            t = (t * self.T).to(torch.int) / self.T
            t += (1 / self.T)
        if self.change_of_variables:
            net_conditioning = t[:, None]
            f_T = torch.log1p(- torch.exp(- self.noise.sigma_max))
            f_0 = torch.log1p(- torch.exp(- self.noise.sigma_min))
            move_chance = torch.exp(f_0 + t * (f_T - f_0))[:, None]
        else:
            sigma, dsigma = self.noise(t)
            net_conditioning = sigma[:, None]
            move_chance = 1 - torch.exp(-sigma[:, None])
        if self.structure_only:
            condition_seq = None
        xt, condition_seq = self.q_xt(x0, move_chance, condition_seq=condition_seq,
                                      non_moving_mask=batch.get("non_moving_mask", None))
        logits, _ = self._model_wrapper(xt, None if condition_seq is None else condition_seq.to(self.device),
                                        net_conditioning)
        log_p_theta = torch.gather(logits, -1, x0.to(self.device)[:, :, None]).squeeze(-1)
        if self.change_of_variables or self.importance_sampling:
            loss = log_p_theta * torch.log1p(- torch.exp(- self.noise.sigma_min))
        else:
            loss = - log_p_theta * (dsigma / torch.expm1(sigma)).to(self.device)[:, None]
        loss_mask = loss_mask.to(self.device)
        loss = (loss * loss_mask).sum() / loss_mask.sum()
        return loss, {"nelbo": loss.detach().clone(), "xt": xt, "t": t}

    def _schedule(self, num_steps, eps, sample_max_t, device):
        """sigma_t, move_chance_t, move_chance_s for every step and sigma at the last grid point,
        with the reference's ops (model.py:564-567, 584-593), evaluated once."""
        ts = torch.linspace(sample_max_t, eps, num_steps + 1, device=device)
        dt = (1 - eps) / num_steps
        t = ts[:-1, None]
        sigma_t = self.noise(t)[0].squeeze(-1)
        sigma_s = self.noise(t - dt)[0].squeeze(-1)
        mc_t = 1 - torch.exp(-sigma_t)
        mc_s = 1 - torch.exp(-sigma_s)
        sigma_last = self.noise(ts[-1:, None])[0].reshape(-1)
        sig = self._process_sigma(torch.cat([sigma_t, sigma_last]))
        return sig.cpu().tolist(), mc_t.cpu().tolist(), mc_s.cpu().tolist()

    @torch.no_grad()
    def ddpm_sample(self, sequence_tokens, num_steps=None, eps=1e-5, input_prior=None, sample_max_t=1.0,
                    seed=None):
        """Generate samples (reference model.py:543-581).  int64 (B,T) in -> int64 (B,T) on device."""
        if num_steps is None:
            print("Using by default num_steps: 1000")
            num_steps = 1000
        if input_prior is None:
            x = self._sample_prior(*sequence_tokens.shape).to(self.device)
            assert sample_max_t == 1.0, f"sample_max_t has to be 1.0 when input_prior is None"
        else:
            print(f"Using input_prior: {input_prior.shape}")
            x = input_prior.to(self.device).clone()
            assert x.shape == sequence_tokens.shape, \
                f"Invalid input_prior shape: {x.shape} v.s. (seq) {sequence_tokens.shape}"
        x = x.contiguous()
        seq = sequence_tokens.to(self.device).contiguous()
        sigma, mc_t, mc_s = self._schedule(num_steps, eps, sample_max_t, self.device)
        eng = self.engine
        if self.rng == "philox":
            if seed is None:
                # one draw from torch's CPU generator per call: reproducible under
                # torch.manual_seed and different for every chunk of a run
                seed = int(torch.randint(0, 2 ** 62, (1,)).item())
            out = eng.ddpm_sample(seq, x, num_steps, sigma, mc_t, mc_s, seed=seed,
                                  noise_removal=self.noise_removal)
            eng.synchronize()          # surfaces out-of-range ids (IndexError) / watchdog trips of the loop
            return out
        B, T = x.shape
        V = self.vocab_size
        logits = torch.empty(B, T, V, dtype=torch.float32, device=self.device)
        for i in range(num_steps):
            eng.forward_sigma(seq, x, sigma[i], logits_out=logits)
            u = torch.rand(B, T, V, dtype=torch.float32, device=self.device)   # == rand_like(q_xs)
            eng.sample_step(x, logits, u, mc_t[i], mc_s[i])
        if self.noise_removal:
            eng.forward_sigma(seq, x, sigma[num_steps], logits_out=logits)
            eng.denoise_argmax(x, logits)
        eng.synchronize()
        return x

    def _ddpm_update(self, x, t, sequence_tokens, dt):
        """One reverse step (reference model.py:583-607) for callers that drive the loop themselves."""
        sigma_t = self.noise(t)[0].reshape(-1)
        sigma_s = self.noise(t - dt)[0].reshape(-1)
        mc_t = float((1 - torch.exp(-sigma_t))[0])
        mc_s = float((1 - torch.exp(-sigma_s))[0])
        x = x.to(self.device).clone().contiguous()
        sig = float(self._process_sigma(sigma_t)[0])
        logits = self.engine.forward_sigma(sequence_tokens.to(self.device), x, sig)
        u = torch.rand(*logits.shape, dtype=torch.float32, device=self.device)
        return self.engine.sample_step(x, logits, u, mc_t, mc_s)
