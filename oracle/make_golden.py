"""Regenerate tests/golden/ from the reference's OWN code (build container only).

    python -m oracle.make_golden

Every expected value written here is produced by code imported verbatim from
/root/reference (``oracle/ref_loader.py``): ``MaskedDiffusionLanguageModeling``
(``logits_parameterization``, ``_ddpm_update``, ``_sample_categorical``, ``ddpm_sample``),
``LogLinearNoise`` and ``TimestepEmbedder``; tokenizer pins come from the reference's
``data/dummy_train_data/*.pth``.  The restated oracle (``oracle/mdlm_ref.py``) is asserted
bit-identical to the reference on every case before anything is saved.

Fixtures
--------
sampler_full.npz      one small sampler step with every array stored (logits, u, x_t -> x_next)
sampler_seeded.npz    larger sampler steps; logits/u are regenerated from stored seeds with
                      ``torch.Generator`` (checksums stored to detect generator drift)
schedule.npz          time grid, sigma, move chances for num_steps in {10, 25, 50}
timestep_embedder.npz TimestepEmbedder outputs for seeded weights
trajectory_tiny.npz   a full 25-step ddpm_sample of the reference sampler driving the tiny
                      oracle net (d=256, 4 heads, 2 layers): x_t per step, final ids
tokenizer_pins.json   sequence <-> token ids, BOS/EOS ids, structure-code range (from *.pth)
chunks.json           chunk lists of sample_esmdiff.py:181-194 for the BASELINE configs
"""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import torch

from . import esm3_ref, mdlm_ref, ref_loader

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"
V = 4101
MASK = 4096


class _FixedLogitsNet(torch.nn.Module):
    """Stands in for ``net`` so the reference sampler sees prescribed logits."""

    def __init__(self):
        super().__init__()
        self.logits = None
        self.output_heads = type("H", (), {"sequence_head": None})()

    def forward(self, structure_tokens, sequence_tokens, auxiliary_embeddings, labels=None):
        return type("O", (), {"structure_logits": self.logits.clone(), "sequence_logits": None})()


def seeded_case(seed: int, B: int, T: int, frac_masked: float, scale: float):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(B, T, V, generator=g) * scale
    u = torch.rand(B, T, V, generator=g)
    x = torch.randint(0, 4096, (B, T), generator=g)
    x = torch.where(torch.rand(B, T, generator=g) < frac_masked, torch.full_like(x, MASK), x)
    return logits, u, x


def reference_step(ref_model, logits, u, x, t_scalar: float, dt: float):
    """One ``_ddpm_update`` of the reference with prescribed logits and uniforms."""
    ref_model.net.logits = logits
    real = torch.rand_like
    torch.rand_like = lambda q, *a, **k: u.clone()      # only patch of the reference: its RNG draw
    try:
        t = t_scalar * torch.ones(x.shape[0], 1)
        seq = torch.zeros_like(x)
        out = ref_model._ddpm_update(x.clone(), t, sequence_tokens=seq, dt=dt)
    finally:
        torch.rand_like = real
    return out


def main():
    assert ref_loader.available(), "needs /root/reference"
    OUT.mkdir(parents=True, exist_ok=True)
    model_mod, noise_mod, TE = ref_loader.load()

    # ---- schedule -------------------------------------------------------------------------
    noise = noise_mod.LogLinearNoise()
    sched = {}
    for n in (10, 25, 50):
        ts = torch.linspace(1.0, 1e-5, n + 1)
        dt = (1 - 1e-5) / n
        t = ts[:-1, None]
        sig_t, sig_s = noise(t)[0].squeeze(-1), noise(t - dt)[0].squeeze(-1)
        sched[f"ts_{n}"] = ts.numpy()
        sched[f"sigma_t_{n}"] = sig_t.numpy()
        sched[f"mc_t_{n}"] = (1 - torch.exp(-sig_t)).numpy()
        sched[f"mc_s_{n}"] = (1 - torch.exp(-sig_s)).numpy()
        sched[f"sigma_final_{n}"] = noise(ts[-1:])[0].numpy()
        # oracle restatement must agree bit for bit
        ts_o, dt_o = mdlm_ref.time_grid(n)
        s_o, mct_o, mcs_o = mdlm_ref.move_chances(ts_o[:-1, None], dt_o)
        assert torch.equal(ts_o, ts) and torch.equal(s_o, sig_t)
        assert torch.equal(mct_o[:, 0, 0], 1 - torch.exp(-sig_t))
        assert torch.equal(mcs_o[:, 0, 0], 1 - torch.exp(-sig_s))
    np.savez(OUT / "schedule.npz", **sched)

    # ---- timestep embedder ----------------------------------------------------------------
    torch.manual_seed(7)
    te = TE(1536).eval()
    sig = torch.tensor([6.9067545, 3.2, 0.5, 0.04078, 1.0e-5], dtype=torch.float32)
    with torch.no_grad():
        out = te(sig)
    mine = esm3_ref.TimestepEmbedderRef(1536).eval()
    mine.load_state_dict(te.state_dict())
    with torch.no_grad():
        assert torch.equal(mine(sig), out)
    np.savez(OUT / "timestep_embedder.npz", seed=7, sigma=sig.numpy(), out=out.numpy(),
             w0_checksum=float(te.mlp[0].weight.double().sum()),
             w2_checksum=float(te.mlp[2].weight.double().sum()))

    # ---- sampler steps ---------------------------------------------------------------------
    ref_model = ref_loader.build_reference_sampler(_FixedLogitsNet(), te)
    ts25, dt25 = mdlm_ref.time_grid(25)

    def one(seed, B, T, frac, scale, step):
        logits, u, x = seeded_case(seed, B, T, frac, scale)
        t_scalar = float(ts25[step])
        x_next = reference_step(ref_model, logits, u, x, t_scalar, dt25)
        # reference log-probs (for lse parity) from its own logits_parameterization
        logp = ref_model.logits_parameterization(logits=logits.clone(), xt=x)
        # oracle restatement, bit for bit
        _, mct, mcs = mdlm_ref.move_chances(ts25[step] * torch.ones(B, 1), dt25)
        lp_o = mdlm_ref.logits_parameterization(logits.clone(), x)
        assert torch.equal(lp_o, logp)
        xo = mdlm_ref.ddpm_update_tail(lp_o, x, mct, mcs, u)
        assert torch.equal(xo, x_next), "oracle != reference"
        return logits, u, x, x_next, logp, float(mct[0, 0, 0]), float(mcs[0, 0, 0])

    logits, u, x, x_next, logp, mct, mcs = one(11, 1, 6, 0.7, 3.0, 3)
    np.savez(OUT / "sampler_full.npz", logits=logits.numpy(), u=u.numpy(), x_t=x.numpy(),
             x_next=x_next.numpy(), mc_t=np.float32(mct), mc_s=np.float32(mcs), step=3,
             logp_masked_rows=logp[x == MASK].numpy()[:, ::97])

    cases = []
    for (seed, B, T, frac, scale, step) in [
            (21, 2, 60, 1.0, 1.0, 0), (22, 4, 60, 0.6, 4.0, 7), (23, 3, 130, 0.3, 0.3, 15),
            (24, 2, 258, 0.05, 8.0, 24), (25, 1, 33, 0.0, 1.0, 12), (26, 5, 17, 0.9, 20.0, 20)]:
        logits, u, x, x_next, logp, mct, mcs = one(seed, B, T, frac, scale, step)
        cases.append(dict(seed=seed, B=B, T=T, frac=frac, scale=scale, step=step,
                          mc_t=mct, mc_s=mcs, x_t=x.numpy(), x_next=x_next.numpy(),
                          logits_sum=float(logits.double().sum()), u_sum=float(u.double().sum()),
                          lse=torch.logsumexp(
                              torch.cat([logits[..., :MASK], logits[..., MASK + 1:]], -1), -1).numpy()))
    np.savez(OUT / "sampler_seeded.npz", n=len(cases),
             **{f"{k}_{i}": np.asarray(v) for i, c in enumerate(cases) for k, v in c.items()})

    # ---- full trajectory on the tiny oracle net ---------------------------------------------
    dims = esm3_ref.Esm3Dims(d_model=256, n_heads=4, v_heads=8, n_layers=2)
    net, emb = esm3_ref.build_reference_model(dims, seed=0)
    te_small = TE(256).eval()
    te_small.load_state_dict(emb.state_dict())
    ref_small = ref_loader.build_reference_sampler(net, te_small)
    g = torch.Generator().manual_seed(0)
    L = 20
    seq = torch.cat([torch.tensor([0]), torch.randint(4, 24, (L,), generator=g), torch.tensor([2])])
    seqs = seq[None].repeat(3, 1)
    torch.manual_seed(123)
    x_ref = ref_small.ddpm_sample(sequence_tokens=seqs, num_steps=25, eps=1e-5,
                                  input_prior=None, sample_max_t=1.0)
    rec = []
    torch.manual_seed(123)
    x_or = mdlm_ref.SamplerRef(net, emb, record=rec).ddpm_sample(seqs, 25)
    assert torch.equal(x_ref, x_or), "oracle trajectory != reference trajectory"
    # inpainting variant (cfg4 shape of the problem): prior with positions 1..8 masked
    prior_codes = torch.randint(0, 4096, (L + 2,), generator=g)
    prior_codes[0], prior_codes[-1] = 4098, 4097
    prior = mdlm_ref.inpainting_prior(prior_codes, 3, list(range(1, 9)))
    torch.manual_seed(321)
    x_ref_inp = ref_small.ddpm_sample(sequence_tokens=seqs, num_steps=25, eps=1e-5,
                                      input_prior=prior.clone(), sample_max_t=1.0)
    torch.manual_seed(321)
    x_or_inp = mdlm_ref.SamplerRef(net, emb).ddpm_sample(seqs, 25, input_prior=prior.clone())
    assert torch.equal(x_ref_inp, x_or_inp)
    np.savez(OUT / "trajectory_tiny.npz", seq=seqs.numpy(), x_final=x_ref.numpy(),
             x_t=np.stack([r["x_t"].numpy() for r in rec]),
             x_next=np.stack([r["x_next"].numpy() for r in rec]),
             sigma_t=np.array([r["sigma_t"] for r in rec], dtype=np.float32),
             logits_abs_sum=np.array([float(r["raw_logits"].double().abs().sum()) for r in rec]),
             prior=prior.numpy(), x_final_inpaint=x_ref_inp.numpy(),
             dims=np.array([dims.d_model, dims.n_heads, dims.v_heads, dims.n_layers]),
             weight_seed=0, sample_seed=123, inpaint_seed=321)

    # ---- tokenizer pins from the reference's own .pth dumps ----------------------------------
    pins = {"aa_to_id": {}, "files": {}}
    for p in sorted((ref_loader.REF_ROOT / "data/dummy_train_data").glob("*.pth")):
        d = torch.load(p, map_location="cpu", weights_only=False)
        toks, s = d["sequence_tokens"].tolist(), d["sequence"]
        assert len(toks) == len(s) + 2
        for ch, tk in zip(s, toks[1:-1]):
            assert pins["aa_to_id"].setdefault(ch, tk) == tk
        st = d["structure_tokens"]
        pins["files"][p.name] = dict(
            L=len(s), seq_bos=toks[0], seq_eos=toks[-1], struct_bos=int(st[0]),
            struct_eos=int(st[-1]), struct_code_max=int(st[1:-1].max()),
            emb_shape=list(d["embeddings"].shape), logits_shape=list(d["structure_logits"].shape),
            emb_absmax=float(d["embeddings"].abs().max()))
    (OUT / "tokenizer_pins.json").write_text(json.dumps(pins, indent=1, sort_keys=True))

    chunks = {f"T{T}_N{N}": mdlm_ref.chunk_sizes(T, N)
              for T, N in [(60, 4), (258, 100), (514, 32), (514, 256), (130, 64), (1026, 512),
                           (1026, 8), (130, 1)]}
    (OUT / "chunks.json").write_text(json.dumps(chunks, indent=1, sort_keys=True))
    print("golden written to", OUT)
    for f in sorted(OUT.iterdir()):
        print(f"  {f.name:28s} {f.stat().st_size:>9d} B")


if __name__ == "__main__":
    main()
