"""GPU: the whole ddpm path through the reference-shaped host API, against the oracle.

Parity tiers:
  T1  fused sampling kernel == reference `_ddpm_update` tail given identical (logits, u, x):
      token ids bit-exact (documented near-tie exemption, counted).
  T2  forward vs the CPU oracle, teacher-forced on the oracle's x_t.  The contract (BASELINE.json
      north_star "logits within 1e-3 rel bf16", pinned down by SURVEY.md 8c T2 / section 7 as the
      POST-logits_parameterization log-probs of masked rows): rel-Frobenius error of the masked-row
      log-probs < 1e-3, asserted against BOTH oracles:
        * oracle/esm3_ref.py  -- fp32, what the reference computes;
        * oracle/esm3_emul.py -- the same weights with a bf16 rounding at exactly the product's
          rounding points, which separates kernel error from bf16 operand noise.
      Raw logits are reported and bounded at <= 2x their measured error (PARITY table below).
      Two implementations with IDENTICAL rounding points that differ only at fp32-ulp level drift
      3e-3 apart on raw logits after 48 blocks (tests/test_oracle_golden.py measures that floor on
      CPU), so raw logits cannot be held to 1e-3 against either oracle by any bf16 kernel; the
      log-prob metric can, and is.
  T3  free-running agreement with the oracle trajectory: reported, not asserted to be 100 %
      (bf16 vs fp32 argmax near-ties diverge; even reference-GPU vs reference-CPU would).
  T4  a reference-style sampler (torch ops, the oracle's restatement of model.py:543-607 --
      pinned bit-for-bit to the reference in the build container) driving the CUDA-backed
      ``CustomizedESM3`` reproduces the fused kernels step by step.
Full-size checks (BASELINE config 2 / 4 shapes) use size-independent properties.
"""
import numpy as np
import pytest
import torch

from oracle import esm3_emul, esm3_ref, mdlm_ref

pytestmark = pytest.mark.gpu
MASK = 4096
DEV = "cuda"


def rel_fro(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20))


# PARITY bounds: contract 1e-3 on masked-row log-probs; everything else <= 2x what the B200 measured
# (profiles/r2_parity.md holds the measured values these come from).
LP_REL_MAX = 6e-4             # masked-row log-probs, rel-Frobenius (contract: 1e-3).  Measured <= 2.95e-4 vs the fp32
                              # oracle, <= 2.14e-4 vs the bf16-emulating oracle
RAW_REL_FP32_MAX = 8.5e-3     # raw logits vs fp32 oracle         (measured 3.9e-3 .. 4.22e-3: bf16 operand noise)
RAW_REL_EMUL_MAX = 6.4e-3     # raw logits vs emulating oracle    (measured 2.5e-3 .. 3.18e-3: the drift floor of two
                              # implementations with equal rounding points, 3.1e-3 in the CPU float64-vs-fp32 probe)
EMB_REL_FP32_MAX = 4.8e-3     # pre-final-norm residual stream vs fp32 oracle (measured <= 2.38e-3)


def masked_logp(logits, xt):
    return mdlm_ref.logits_parameterization(logits.clone(), xt)[xt == MASK][:, :4096]


def parity(logits, xt, ref_logits, emul_logits, emb=None, ref_emb=None):
    """All parity figures of one forward; asserts the bounds above."""
    logits, xt = logits.float().cpu(), xt.cpu()
    lp = masked_logp(logits, xt)
    lp_ref, lp_emu = masked_logp(ref_logits, xt), masked_logp(emul_logits, xt)
    mx = lambda t: float(t.abs().max()) if t.numel() else 0.0            # late steps: no masked row left
    out = {"lp_rel_fp32": rel_fro(lp, lp_ref), "lp_rel_emul": rel_fro(lp, lp_emu),
           "lp_maxabs_fp32": mx(lp - lp_ref), "lp_maxabs_emul": mx(lp - lp_emu),
           "raw_rel_fp32": rel_fro(logits, ref_logits), "raw_rel_emul": rel_fro(logits, emul_logits),
           "emul_vs_fp32_raw": rel_fro(emul_logits, ref_logits),
           "top1_fp32": float((lp.argmax(-1) == lp_ref.argmax(-1)).float().mean()) if lp.numel() else 1.0,
           "top1_emul": float((lp.argmax(-1) == lp_emu.argmax(-1)).float().mean()) if lp.numel() else 1.0}
    if emb is not None:
        out["emb_rel_fp32"] = rel_fro(emb.float().cpu(), ref_emb)
        assert out["emb_rel_fp32"] < EMB_REL_FP32_MAX, out
    assert out["lp_rel_fp32"] < LP_REL_MAX and out["lp_rel_emul"] < LP_REL_MAX, out
    assert out["raw_rel_fp32"] < RAW_REL_FP32_MAX and out["raw_rel_emul"] < RAW_REL_EMUL_MAX, out
    return out


def fmt(d):
    return ", ".join(f"{k} {v:.2e}" if v < 0.5 else f"{k} {v:.3f}" for k, v in d.items())


def make_seq(B, T, seed=0):
    g = torch.Generator().manual_seed(seed)
    row = torch.cat([torch.zeros(1, dtype=torch.long), torch.randint(4, 24, (T - 2,), generator=g),
                     torch.full((1,), 2)])
    return row[None].repeat(B, 1)


# ---------------------------------------------------------------------------------------------
# tiny model: golden trajectory, teacher-forced
# ---------------------------------------------------------------------------------------------
def test_teacher_forced_trajectory_tiny(tiny_pair, golden_dir):
    net, emb, eng = tiny_pair
    g = np.load(golden_dir / "trajectory_tiny.npz")
    seq = torch.from_numpy(g["seq"])
    rec = []
    torch.manual_seed(int(g["sample_seed"]))
    x_final = mdlm_ref.SamplerRef(net, emb, record=rec).ddpm_sample(seq, 25)
    assert np.array_equal(x_final.numpy(), g["x_final"])            # oracle == reference-made fixture
    assert np.array_equal(np.stack([r["x_t"].numpy() for r in rec]), g["x_t"])
    B = seq.shape[0]
    excused = 0
    worst = {}
    agree = []
    for r in rec[:-1]:
        x_t = r["x_t"]
        logits = eng.forward_sigma(seq, x_t.to(DEV), r["sigma_t"])
        eng.synchronize()
        cond = emb(torch.tensor([r["sigma_t"]]))[0][None, None].expand(B, seq.shape[1], -1)
        emu = esm3_emul.forward(net, x_t, seq, cond).structure_logits
        for k, v in parity(logits, x_t, r["raw_logits"], emu).items():                     # T2
            worst[k] = min(worst.get(k, 1.0), v) if k.startswith("top1") else max(worst.get(k, 0.0), v)
        mct = torch.full((B, 1, 1), r["mc_t"])
        mcs = torch.full((B, 1, 1), r["mc_s"])
        # T1: the kernel on the ORACLE's logits and uniforms
        got = eng.sample_step(x_t.clone().to(DEV), r["raw_logits"].to(DEV).contiguous(), r["u"].to(DEV),
                              r["mc_t"], r["mc_s"])
        eng.synchronize()
        lp = mdlm_ref.logits_parameterization(r["raw_logits"].clone(), x_t)
        excused += mdlm_ref.assert_ids_match(got.cpu(), r["x_next"], lp, mct, mcs, r["u"])
        # T3 (reported): the kernel on ITS OWN logits
        own = eng.sample_step(x_t.clone().to(DEV), logits, r["u"].to(DEV), r["mc_t"], r["mc_s"])
        eng.synchronize()
        m = x_t == MASK
        agree.append(float((own.cpu()[m] == r["x_next"][m]).float().mean()) if bool(m.any()) else 1.0)
    last = rec[-1]                                                                          # noise removal
    got = eng.denoise_argmax(last["x_t"].clone().to(DEV), last["raw_logits"].to(DEV).contiguous())
    eng.synchronize()
    assert torch.equal(got.cpu(), last["x_next"])
    print(f"\n[T2 tiny, worst over 25 steps] {fmt(worst)}; [T1] near-tie rows excused {excused}; "
          f"[T3] per-step agreement on own logits min {min(agree):.3f} mean {sum(agree) / len(agree):.3f}")
    assert excused <= 2
    assert sum(agree) / len(agree) > 0.9


def test_reference_style_sampler_drives_cuda_net(tiny_pair):
    """T4: the sampler half in plain torch ops (oracle restatement of the reference's
    ddpm_sample) calls ``net(structure_tokens=, sequence_tokens=, auxiliary_embeddings=,
    labels=None)`` on the CUDA-backed CustomizedESM3 mirror exactly like model.py:475-480."""
    from esmdiff_b200.net import ESMOutput
    net_o, emb_o, eng = tiny_pair

    class Net(torch.nn.Module):          # CustomizedESM3.forward surface over the shared engine
        def forward(self, structure_tokens, labels=None, mask=None, sequence_tokens=None, *,
                    auxiliary_embeddings=None):
            assert labels is None
            logits, _ = eng.forward(sequence_tokens, structure_tokens, aux=auxiliary_embeddings)
            return ESMOutput(sequence_logits=None, structure_logits=logits)

    te = esm3_ref.TimestepEmbedderRef(256).to(DEV).eval()
    te.load_state_dict(emb_o.state_dict())
    seq = make_seq(3, 50, seed=4)
    rec = []
    torch.manual_seed(5)
    sampler = mdlm_ref.SamplerRef(Net(), te, record=rec, device=DEV)
    out = sampler.ddpm_sample(seq, 12)
    assert out.shape == (3, 50) and out.is_cuda and int((out == MASK).sum()) == 0
    excused = 0
    for r in rec[:-1]:
        B = r["x_t"].shape[0]
        got = eng.sample_step(r["x_t"].clone(), r["raw_logits"].contiguous(), r["u"], r["mc_t"], r["mc_s"])
        eng.synchronize()
        mct = torch.full((B, 1, 1), r["mc_t"])
        mcs = torch.full((B, 1, 1), r["mc_s"])
        lp = mdlm_ref.logits_parameterization(r["raw_logits"].cpu().clone(), r["x_t"].cpu())
        excused += mdlm_ref.assert_ids_match(got.cpu(), r["x_next"].cpu(), lp, mct, mcs, r["u"].cpu())
        # aux as a full (B,T,d) tensor (what the reference passes) == one shared vector
        l2 = eng.forward_sigma(seq, r["x_t"], r["sigma_t"])
        eng.synchronize()
        # (the torch MLP and the library's time-embedding kernel sum in different orders: a
        # 1e-7 difference in cond flips a few bf16 roundings downstream)
        assert rel_fro(l2, r["raw_logits"]) < 5e-3
    assert excused <= 2


def test_host_mirror_model_loop_matches_stepwise(tiny_pair):
    """MaskedDiffusionLanguageModeling.ddpm_sample (rng='torch') == driving forward_sigma +
    sample_step by hand with the same torch CUDA generator stream; rng='philox' == the C loop."""
    from esmdiff_b200.model import MaskedDiffusionLanguageModeling
    from esmdiff_b200.noise_utils import LogLinearNoise
    _, emb_o, eng = tiny_pair

    class NetShim:
        engine = eng
        device = eng.device
        output_heads = None

    m = MaskedDiffusionLanguageModeling(net=NetShim(), noise_schedule=LogLinearNoise(), sigma_embedder=None,
                                        time_conditioning=True, noise_removal=True)
    seq = make_seq(2, 33, seed=7)
    torch.manual_seed(21)
    a = m.ddpm_sample(seq, num_steps=9)
    sigma, mc_t, mc_s = m._schedule(9, 1e-5, 1.0, DEV)
    torch.manual_seed(21)
    x = torch.full((2, 33), MASK, device=DEV)
    for i in range(9):
        logits = eng.forward_sigma(seq, x, sigma[i])
        u = torch.rand(2, 33, 4101, device=DEV)
        eng.sample_step(x, logits, u, mc_t[i], mc_s[i])
    logits = eng.forward_sigma(seq, x, sigma[9])
    eng.denoise_argmax(x, logits)
    eng.synchronize()
    assert torch.equal(a, x)
    # assertion behaviour of the reference (model.py:556,562)
    with pytest.raises(AssertionError):
        m.ddpm_sample(seq, num_steps=3, sample_max_t=0.5)
    with pytest.raises(AssertionError):
        m.ddpm_sample(seq, num_steps=3, input_prior=torch.full((2, 30), MASK))
    m.rng = "philox"
    b1 = m.ddpm_sample(seq, num_steps=9, seed=3)
    b2 = eng.ddpm_sample(seq, None, 9, sigma, mc_t, mc_s, seed=3)
    b3 = m.ddpm_sample(seq, num_steps=9, seed=4)
    eng.synchronize()
    assert torch.equal(b1, b2) and not torch.equal(b1, b3)
    host = eng.ddpm_sample_host(seq.contiguous(), None, 9, seed=3)
    # host entry point evaluates the schedule in C floats (<= 2 ulp from torch's): same ids unless
    # a uniform lands within 1e-7 of a move-chance boundary
    assert float((host.to(DEV) == b1).float().mean()) > 0.98


def test_model_step_nelbo_matches_reference_golden(tiny_pair, golden_dir):
    """SURVEY.md 8f row 4, forward half: the diffusion NELBO of one batch (``model_step``, reference model.py:386-462
    -- what validation_step / test_step log) with the forward on the CUDA path, against the value the reference's OWN
    model_step returned around the fp32 oracle network (tests/golden/model_step.npz).  The draws (t, x_t) are replayed
    bit for bit from torch's CPU generator; the loss is a masked mean of log-probs: 2e-4 relative."""
    from esmdiff_b200 import noise_utils
    from esmdiff_b200.model import MaskedDiffusionLanguageModeling
    from esmdiff_b200.net import TimestepEmbedder
    net, emb, eng = tiny_pair
    g = np.load(golden_dir / "model_step.npz")

    class _Net:                                     # the fixture's engine behind the mirror's network surface
        engine, device = eng, eng.device

    te = TimestepEmbedder(256)
    te.load_state_dict(emb.state_dict())
    m = MaskedDiffusionLanguageModeling(net=_Net(), noise_schedule=noise_utils.LogLinearNoise(), sigma_embedder=te.to(DEV),
                                        time_conditioning=True, condition_mask_rate=0.0)
    batch = {k: torch.from_numpy(g[k]) for k in ("structure_tokens", "sequence_tokens", "mask")}
    for name, kw in (("plain", {}), ("discrete_T", {"T": 50}), ("importance", {"importance_sampling": True}),
                     ("change_of_variables", {"change_of_variables": True})):
        m.T, m.importance_sampling, m.change_of_variables = 0, False, False
        for k, v in kw.items():
            setattr(m, k, v)
        torch.manual_seed(11)
        loss, bd = m.model_step(batch, training=False)
        want = float(g[f"{name}_loss"])
        assert np.array_equal(bd["xt"].numpy(), g[f"{name}_xt"]) and np.array_equal(bd["t"].numpy(), g[f"{name}_t"])
        print(f"[model_step {name}] nelbo {float(loss):.5f} vs reference {want:.5f} (rel {abs(float(loss) - want) / abs(want):.2e})")
        assert abs(float(loss) - want) / abs(want) < 2e-4     # measured <= 4.1e-5 on B200


def test_inpainting_tiny(tiny_pair):
    from esmdiff_b200.sampling import build_prior
    _, _, eng = tiny_pair
    T = 60
    seq = make_seq(4, T, seed=8)
    g = torch.Generator().manual_seed(1)
    st = torch.randint(0, 4096, (T,), generator=g)
    st[0], st[-1] = 4098, 4097
    prior = build_prior(st, 4, mask_ids=list(range(1, 33)))
    out = eng.ddpm_sample(seq, prior, 25, *eng.schedule(25), seed=2)
    eng.synchronize()
    out = out.cpu()
    known = prior != MASK
    assert torch.equal(out[known], prior[known])                   # never resampled (model.py:606-607)
    assert int((out == MASK).sum()) == 0
    assert len({tuple(r.tolist()) for r in out[:, 1:33]}) > 1      # samples differ


def test_trained_like_layernorm_weights_both_variants():
    """Default init has gamma = 1, beta = 0 in every LayerNorm, which would leave the LayerNorm
    folding (gemm.cuh) and the q_ln/k_ln weights untested end to end: perturb them all, then the
    default path (pre-LNs folded through the GEMMs, q_ln / k_ln + RoPE folded into the QKV
    epilogue and attention), the variant with the stand-alone q/k-LN + RoPE kernel
    (ESMDIFF_QK=separate) and the all-stand-alone-kernels variant (ESMDIFF_LN=separate) must all
    match the fp32 oracle, and the default one the bf16-emulating oracle."""
    import os
    from conftest import TINY
    from esmdiff_b200.engine import Dims, Engine
    net, emb = esm3_ref.build_reference_model(esm3_ref.Esm3Dims(**TINY), seed=3)
    g = torch.Generator().manual_seed(4)
    with torch.no_grad():
        for name, prm in net.named_parameters():
            if prm.dim() == 1 and ("layernorm" in name or "ffn.0" in name or "q_ln" in name or "k_ln" in name
                                   or "norm" in name or "structure_head.2" in name):
                if name.endswith("weight"):
                    prm.mul_(1.0 + 0.25 * torch.randn(prm.shape, generator=g))
                else:
                    prm.add_(0.2 * torch.randn(prm.shape, generator=g))
    sd = esm3_ref.full_state_dict(net, emb)
    B, T = 3, 70
    seq = make_seq(B, T, seed=5)
    xt = torch.randint(0, 4096, (B, T), generator=g)
    xt[torch.rand(B, T, generator=g) < 0.6] = MASK
    with torch.no_grad():
        cond = emb(torch.tensor([0.7]))[0]
        ref = net(structure_tokens=xt, sequence_tokens=seq, auxiliary_embeddings=cond[None, None].expand(B, T, -1))
        emu = esm3_emul.forward(net, xt, seq, cond[None, None].expand(B, T, -1))
    errs = {}
    for mode, env in (("fused", {}), ("qk_separate", {"ESMDIFF_QK": "separate"}), ("ln_separate", {"ESMDIFF_LN": "separate"})):
        os.environ.update(env)
        try:
            eng = Engine(Dims(**TINY))
        finally:
            for k in env:
                os.environ.pop(k)
        eng.load_state_dict(sd)
        logits, embd = eng.forward(seq, xt.to(DEV), aux=eng.time_embed(0.7), want_embeddings=True)
        eng.synchronize()
        lp, lp_ref = masked_logp(logits.cpu(), xt), masked_logp(ref.structure_logits, xt)
        errs[mode] = {"emb_rel_fp32": rel_fro(embd.cpu(), ref.embeddings), "raw_rel_fp32": rel_fro(logits.cpu(), ref.structure_logits),
                      "lp_rel_fp32": rel_fro(lp, lp_ref)}
        if mode == "fused":          # the emulating oracle restates the default (fused) product path
            errs[mode].update(parity(logits, xt, ref.structure_logits, emu.structure_logits, embd, ref.embeddings))
        eng.close()
    print("\n[LayerNorm variants, trained-like gamma/beta] " + "; ".join(f"{m}: {fmt(e)}" for m, e in errs.items()))
    for e in errs.values():
        assert e["emb_rel_fp32"] < EMB_REL_FP32_MAX and e["raw_rel_fp32"] < RAW_REL_FP32_MAX and e["lp_rel_fp32"] < LP_REL_MAX
    # folding q_ln / k_ln into the epilogue removes one bf16 rounding of q and k: it must not be worse
    assert errs["fused"]["raw_rel_fp32"] < 1.5 * errs["ln_separate"]["raw_rel_fp32"] + 1e-3


# ---------------------------------------------------------------------------------------------
# ESM3-open-sized model (d=1536, 48 layers, 24 heads)
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def full_model():
    from esmdiff_b200.engine import Dims, Engine
    from esmdiff_b200.synthetic import random_state_dict
    dims = Dims()
    sd = random_state_dict(dims, device=DEV, seed=0, full=True)
    eng = Engine(dims)
    eng.load_state_dict(sd)
    yield eng, sd
    eng.close()


@pytest.fixture(scope="module")
def full_oracle(full_model):
    """fp32 oracle modules holding the same tensors as the engine (built once: 1.4 B parameters)."""
    _, sd = full_model
    return esm3_ref.build_from_state_dict(esm3_ref.Esm3Dims(), sd)


BPTI = "RPDFCLEPPYTGPCKARIIRYFYNAKAGLCQTFVYGGCRAKRNNFKSAEDCMRTCGGA"      # data/targets/bpti/bpti.pdb, 58 residues


def _case(name):
    """(seq (B,T), x_t (B,T), sigma) of the BASELINE configurations, teacher-forced mid-trajectory."""
    from esmdiff_b200.sampling import build_prior
    from esmdiff_b200.tokenization import tokenize_sequence
    g = torch.Generator().manual_seed(2)
    if name == "config1_bpti_B4_T60":
        seq = tokenize_sequence(BPTI)[None].repeat(4, 1)
    elif name == "config3_B1_T514":
        seq = make_seq(1, 514, seed=3)
    elif name == "config5_B1_T1026":
        seq = make_seq(1, 1026, seed=5)
    else:
        seq = make_seq(2, 258, seed=1)
    B, T = seq.shape
    if name == "config4_inpaint_B2_T258":
        # sample_esmdiff.py:197-201 + models/utils.py:117-123: token positions 1..32 masked in the prior,
        # residues 1..32 ('_' -> id 32) masked in the sequence (the reference's off-by-one, kept)
        seq = seq.clone()
        seq[:, 2:34] = 32
        st = torch.randint(0, 4096, (T,), generator=g)
        st[0], st[-1] = 4098, 4097
        xt = build_prior(st, B, mask_ids=list(range(1, 33)))
        sigma = 6.9                                   # first step of the grid: t = 1 (model.py:564-567)
    else:
        xt = torch.randint(0, 4096, (B, T), generator=g)
        xt[torch.rand(B, T, generator=g) < 0.5] = MASK
        sigma = 0.9
    return seq, xt, sigma


@pytest.mark.parametrize("name", ["config1_bpti_B4_T60", "config2_B2_T258", "config3_B1_T514",
                                  "config4_inpaint_B2_T258", "config5_B1_T1026"])
def test_full_size_forward_vs_oracles(full_model, full_oracle, name):
    """T2 at the real architecture (d=1536, 48 blocks, 24 heads) and at the sequence lengths of
    every BASELINE configuration: T = 60 (one partial query tile), T = 258 (two query tiles + the
    two CUDA-core trailing rows, five K/V tiles, 8-9 GEMM row tiles) and T = 514 (four query tiles,
    nine K/V tiles); config 4's partially masked prior with masked sequence residues; config 5's longest chain,
    T = 1026: the STREAMING attention kernel (K/V no longer resident in shared memory) inside the full forward."""
    eng, _ = full_model
    net, emb = full_oracle
    seq, xt, sigma = _case(name)
    B, T = seq.shape
    with torch.no_grad():
        cond = emb(torch.tensor([sigma]))[0][None, None].expand(B, T, -1)
        ref = net(structure_tokens=xt, sequence_tokens=seq, auxiliary_embeddings=cond)
        emu = esm3_emul.forward(net, xt, seq, cond)
    logits, embd = eng.forward(seq, xt.to(DEV), aux=eng.time_embed(sigma), want_embeddings=True)
    eng.synchronize()
    out = parity(logits, xt, ref.structure_logits, emu.structure_logits, embd, ref.embeddings)
    print(f"\n[T2 full size {name}] {fmt(out)}")


def test_full_size_forward_with_structure_coords(full_model, full_oracle):
    """SURVEY.md 8a A6 live: ``structure_coords`` given (net.py:385, 433-441) -> block 0's geometric attention
    (v_heads = 256, T = 258 keys) contributes; residues 1..32 without coordinates (gibbs / inpainting prompt,
    sample_esmdiff.py:92-98) and the BOS / EOS positions are frameless.  Same bounds as the other T2 cases, and
    the branch must matter (otherwise the comparison says nothing) and must switch off again exactly."""
    from oracle import vqvae_enc_ref
    eng, sd = full_model
    net, emb = full_oracle
    seq, xt, sigma = _case("config2_B2_T258")
    B, T = seq.shape
    coords = torch.full((B, T, 3, 3), float("nan"))
    coords[:, 1:-1] = vqvae_enc_ref.synthetic_backbone(T - 2, seed=7)
    coords[:, 2:34] = float("inf")
    g = torch.Generator().manual_seed(5)
    ga = net.transformer.blocks[0].geom_attn                     # non-trivial per-head scales on both sides
    wd, wr = torch.randn(256, generator=g), torch.randn(256, generator=g)
    with torch.no_grad():
        ga.distance_scale_per_head.copy_(wd)
        ga.rotation_scale_per_head.copy_(wr)
    eng.set_weight("net.transformer.blocks.0.geom_attn.distance_scale_per_head", wd)
    eng.set_weight("net.transformer.blocks.0.geom_attn.rotation_scale_per_head", wr)
    eng.finalize()
    with torch.no_grad():
        cond = emb(torch.tensor([sigma]))[0][None, None].expand(B, T, -1)
        ref = net(structure_tokens=xt, sequence_tokens=seq, auxiliary_embeddings=cond, structure_coords=coords)
        ref0 = net(structure_tokens=xt, sequence_tokens=seq, auxiliary_embeddings=cond)
        emu = esm3_emul.forward(net, xt, seq, cond, structure_coords=coords)
    base, _ = eng.forward(seq, xt.to(DEV), aux=eng.time_embed(sigma))
    eng.set_structure_coords(coords)
    logits, embd = eng.forward(seq, xt.to(DEV), aux=eng.time_embed(sigma), want_embeddings=True)
    eng.set_structure_coords(None)
    again, _ = eng.forward(seq, xt.to(DEV), aux=eng.time_embed(sigma))
    eng.synchronize()
    out = parity(logits, xt, ref.structure_logits, emu.structure_logits, embd, ref.embeddings)
    effect = rel_fro(ref.structure_logits, ref0.structure_logits)
    print(f"\n[T2 full size structure_coords B2_T258] {fmt(out)}, effect of the branch on the logits {effect:.3f}")
    assert effect > 10 * out["raw_rel_fp32"]
    assert torch.equal(again, base)


def test_config2_full_run_properties(full_model):
    """BASELINE config 2 in full (L=256, 100 samples, 25 steps, reference chunk list [63, 37])."""
    from esmdiff_b200.sampling import chunk_sizes
    eng, _ = full_model
    T = 258
    row = make_seq(1, T, seed=0)[0]
    sched = eng.schedule(25)
    outs = []
    for ci, bs in enumerate(chunk_sizes(T, 100)):
        outs.append(eng.ddpm_sample(row[None].repeat(bs, 1), None, 25, *sched, seed=100 + ci))
    eng.synchronize()
    tok = torch.cat(outs)[:, 1:-1].cpu()
    assert tok.shape == (100, 256)
    assert int(tok.min()) >= 0 and int(tok.max()) <= 4100 and int((tok == MASK).sum()) == 0
    assert len({tuple(r.tolist()) for r in tok}) == 100            # i.i.d. samples, all distinct
    again = eng.ddpm_sample(row[None].repeat(63, 1), None, 25, *sched, seed=100)
    eng.synchronize()
    assert torch.equal(again, outs[0])                              # deterministic in the seed


def test_config4_inpainting_full_size(full_model):
    """BASELINE config 4: mask_ids 1..32 on L=256 (token positions, the reference's off-by-one)."""
    from esmdiff_b200.sampling import build_prior
    eng, _ = full_model
    T = 258
    row = make_seq(1, T, seed=0)[0].clone()
    row[2:34] = 32                                                  # residues 1..32 -> '_' (mask id)
    g = torch.Generator().manual_seed(1)
    st = torch.randint(0, 4096, (T,), generator=g)
    st[0], st[-1] = 4098, 4097
    prior = build_prior(st, 37, mask_ids=list(range(1, 33)))
    out = eng.ddpm_sample(row[None].repeat(37, 1), prior, 25, *eng.schedule(25), seed=5).cpu()
    eng.synchronize()
    known = prior != MASK
    assert int(known.sum()) == 37 * (T - 32)
    assert torch.equal(out[known], prior[known]) and int((out == MASK).sum()) == 0


def test_config3_shape_properties(full_model):
    """BASELINE config 3 shape (L=512, T=514, num_steps=50) on one GPU's shard, reduced to 6 samples
    so the test stays short: size-independent properties of the sampler (every mask resolved, ids in
    range, fixed BOS/EOS forcing, determinism in the seed, i.i.d. samples distinct) at the longer
    sequence (nine 64-key tiles, four 128-row query tiles + the two CUDA-core leftover rows)."""
    eng, _ = full_model
    T, N, steps = 514, 6, 50
    row = make_seq(1, T, seed=3)[0]
    sched = eng.schedule(steps)
    out = eng.ddpm_sample(row[None].repeat(N, 1), None, steps, *sched, seed=42)
    again = eng.ddpm_sample(row[None].repeat(N, 1), None, steps, *sched, seed=42)
    eng.synchronize()
    assert torch.equal(out, again)
    tok = out.cpu()
    assert tok.shape == (N, T) and int((tok == MASK).sum()) == 0
    assert int(tok.min()) >= 0 and int(tok.max()) <= 4100
    assert len({tuple(r.tolist()) for r in tok}) == N


def test_cuda_graph_replay_matches_eager_launches():
    """Small batches replay each forward from a CUDA graph (esmdiff_b200.cu forward_step); the sampled
    tokens must be identical to the eagerly launched loop (same kernels, same seed), call after call
    (first forward eager, second captured, the rest replayed; a second call reuses the graph)."""
    import os
    from conftest import TINY
    from esmdiff_b200.engine import Dims, Engine
    net, emb = esm3_ref.build_reference_model(esm3_ref.Esm3Dims(**TINY), seed=0)
    sd = esm3_ref.full_state_dict(net, emb)
    B, T, steps = 3, 70, 8
    seq = make_seq(B, T, seed=9).to(DEV)
    outs = {}
    for mode in ("0", "1"):
        os.environ["ESMDIFF_GRAPH"] = mode
        try:
            eng = Engine(Dims(**TINY))
        finally:
            os.environ.pop("ESMDIFF_GRAPH")
        eng.load_state_dict(sd)
        sched = eng.schedule(steps)
        buf = []
        for call in range(3):
            seqs = seq if call < 2 else seq.clone()            # third call: new pointers -> new graph
            buf.append(eng.ddpm_sample(seqs, None, steps, *sched, seed=11 + call).cpu())
        eng.synchronize()
        outs[mode] = (buf, eng.launch_count)
        eng.close()
    for a, b in zip(outs["0"][0], outs["1"][0]):
        assert torch.equal(a, b)
        assert int((a == MASK).sum()) == 0
    assert outs["0"][1] == outs["1"][1]                          # the launch count is graph-agnostic


def test_two_stream_half_batches_are_bit_identical():
    """ESMDIFF_SPLIT_ROWS (off by default): small batches sampled as two independent halves on two
    streams with two workspaces (esmdiff_ddpm_sample; fills the partial tile waves of small GEMMs).  Per-row arithmetic does not
    depend on the batch and the Philox counter uses the row index of the whole batch, so the tokens
    must equal the single-stream loop's bit for bit -- with and without CUDA graphs, with a prior,
    for odd and even batch sizes, and across repeated calls (workspace / graph reuse)."""
    import os
    from conftest import TINY
    from esmdiff_b200.engine import Dims, Engine
    from esmdiff_b200.sampling import build_prior
    net, emb = esm3_ref.build_reference_model(esm3_ref.Esm3Dims(**TINY), seed=0)
    sd = esm3_ref.full_state_dict(net, emb)
    outs = {}
    for mode, env in (("single", {"ESMDIFF_SPLIT_ROWS": "0"}), ("split", {"ESMDIFF_SPLIT_ROWS": "100000"}),
                      ("split_eager", {"ESMDIFF_SPLIT_ROWS": "100000", "ESMDIFF_GRAPH": "0"})):
        os.environ.update(env)
        try:
            eng = Engine(Dims(**TINY))
        finally:
            for k in env:
                os.environ.pop(k)
        eng.load_state_dict(sd)
        res = []
        for (B, T, steps) in ((5, 70, 8), (2, 33, 5), (1, 40, 4), (6, 130, 6)):
            seq = make_seq(B, T, seed=B).to(DEV)
            sched = eng.schedule(steps)
            res.append(eng.ddpm_sample(seq, None, steps, *sched, seed=17).cpu())
            res.append(eng.ddpm_sample(seq, None, steps, *sched, seed=18).cpu())          # reuse: graphs, workspaces
        g = torch.Generator().manual_seed(3)
        st = torch.randint(0, 4096, (70,), generator=g)
        st[0], st[-1] = 4098, 4097
        prior = build_prior(st, 5, mask_ids=list(range(1, 33)))
        res.append(eng.ddpm_sample(make_seq(5, 70, seed=5).to(DEV), prior, 8, *eng.schedule(8), seed=19).cpu())
        logits = eng.forward_sigma(make_seq(3, 50, seed=1), torch.full((3, 50), MASK, device=DEV), 0.5)   # set 0 is active again
        eng.synchronize()
        res.append(logits.cpu())
        outs[mode] = res
        eng.close()
    for a, b, c in zip(outs["single"], outs["split"], outs["split_eager"]):
        assert torch.equal(a, b) and torch.equal(a, c)
        assert a.dtype != torch.int64 or int((a == MASK).sum()) == 0


def test_skip_of_dead_noise_removal_forward_is_exact(monkeypatch):
    """SURVEY.md section 7 step 6 (optional, off by default): with ESMDIFF_SKIP_DENOISE_FORWARD=1 the device loop leaves
    out the noise-removal forward when no MASK is left after the last step -- unmasked rows return themselves
    (model.py:575-579), so the tokens are bit-identical and a forward's worth of launches is saved; with a MASK left
    it runs as usual."""
    from conftest import TINY
    from esmdiff_b200.engine import Dims, Engine
    net, emb = esm3_ref.build_reference_model(esm3_ref.Esm3Dims(**TINY), seed=0)
    sd = esm3_ref.full_state_dict(net, emb)
    seq = make_seq(6, 40, seed=4)
    outs, launches = {}, {}
    for flag in ("0", "1"):
        monkeypatch.setenv("ESMDIFF_SKIP_DENOISE_FORWARD", flag)
        eng = Engine(Dims(**TINY))
        eng.load_state_dict(sd)
        sigma, mc_t, mc_s = eng.schedule(12)
        res = []
        for case, last_mc_s in (("none_left", 0.0), ("some_left", 0.5)):
            mcs = list(mc_s)
            mcs[-1] = last_mc_s                      # chance to stay masked after the last step: 0 -> every row is revealed
            l0 = eng.launch_count
            res.append(eng.ddpm_sample(seq, None, 12, sigma, mc_t, mcs, seed=3).cpu())
            eng.synchronize()
            launches[(flag, case)] = eng.launch_count - l0
        outs[flag] = res
        eng.close()
    assert torch.equal(outs["0"][0], outs["1"][0]) and torch.equal(outs["0"][1], outs["1"][1])
    assert int((outs["1"][0] == MASK).sum()) == 0 and int((outs["1"][1] == MASK).sum()) == 0
    saved = launches[("0", "none_left")] - launches[("1", "none_left")]
    assert saved > 10, launches                                              # one forward (2 blocks + head) not launched
    assert launches[("1", "some_left")] == launches[("0", "some_left")] + 1  # only the counting kernel added
