"""CPU: the oracle restatement against the golden vectors produced by the reference's own code
(oracle/make_golden.py), and -- where /root/reference is present -- against the reference live."""
import json

import numpy as np
import pytest
import torch

from oracle import esm3_emul, esm3_ref, mdlm_ref, ref_loader, vqvae_ref

MASK = 4096


def test_schedule_matches_reference(golden_dir):
    g = np.load(golden_dir / "schedule.npz")
    for n in (10, 25, 50):
        ts, dt = mdlm_ref.time_grid(n)
        sig, mct, mcs = mdlm_ref.move_chances(ts[:-1, None], dt)
        assert np.array_equal(ts.numpy(), g[f"ts_{n}"])
        assert np.array_equal(sig.numpy(), g[f"sigma_t_{n}"])
        assert np.array_equal(mct[:, 0, 0].numpy(), g[f"mc_t_{n}"])
        assert np.array_equal(mcs[:, 0, 0].numpy(), g[f"mc_s_{n}"])
    # SURVEY.md 8a A4 spot values
    assert abs(g["mc_t_25"][0] - 0.9990) < 1e-4 and abs(g["mc_s_25"][0] - 0.9590) < 1e-4
    assert abs(g["mc_t_25"][24] - 0.03997) < 1e-4 and abs(g["mc_s_25"][24] - 1.0e-5) < 2e-6


def test_timestep_embedder_matches_reference(golden_dir):
    g = np.load(golden_dir / "timestep_embedder.npz")
    torch.manual_seed(int(g["seed"]))
    te = esm3_ref.TimestepEmbedderRef(1536).eval()      # same init order as the reference module
    assert abs(float(te.mlp[0].weight.double().sum()) - float(g["w0_checksum"])) < 1e-9
    with torch.no_grad():
        out = te(torch.from_numpy(g["sigma"]))
    assert np.array_equal(out.numpy(), g["out"])


def test_sampler_full_vector(golden_dir):
    g = np.load(golden_dir / "sampler_full.npz")
    logits, u, x = (torch.from_numpy(g[k]) for k in ("logits", "u", "x_t"))
    mct = torch.full((1, 1, 1), float(g["mc_t"]))
    mcs = torch.full((1, 1, 1), float(g["mc_s"]))
    lp = mdlm_ref.logits_parameterization(logits.clone(), x)
    assert np.array_equal(lp[x == MASK].numpy()[:, ::97], g["logp_masked_rows"])
    assert np.array_equal(mdlm_ref.ddpm_update_tail(lp, x, mct, mcs, u).numpy(), g["x_next"])


from golden_cases import seeded_cases  # noqa: E402


def test_sampler_seeded_vectors(golden_dir):
    n = 0
    for c, logits, u, x in seeded_cases(golden_dir):
        B = x.shape[0]
        mct = torch.full((B, 1, 1), float(c["mc_t"]))
        mcs = torch.full((B, 1, 1), float(c["mc_s"]))
        lp = mdlm_ref.logits_parameterization(logits.clone(), x)
        out = mdlm_ref.ddpm_update_tail(lp, x, mct, mcs, u)
        assert np.array_equal(out.numpy(), c["x_next"])
        # edge cases the fixtures cover: nothing masked -> identity; everything masked at step 0
        if float(c["frac"]) == 0.0:
            assert torch.equal(out, x)
        n += 1
    assert n == 6


def test_trajectory_tiny(golden_dir):
    g = np.load(golden_dir / "trajectory_tiny.npz")
    d = [int(v) for v in g["dims"]]
    dims = esm3_ref.Esm3Dims(d_model=d[0], n_heads=d[1], v_heads=d[2], n_layers=d[3])
    net, emb = esm3_ref.build_reference_model(dims, seed=int(g["weight_seed"]))
    seq = torch.from_numpy(g["seq"])
    rec = []
    torch.manual_seed(int(g["sample_seed"]))
    x = mdlm_ref.SamplerRef(net, emb, record=rec).ddpm_sample(seq, 25)
    assert np.array_equal(x.numpy(), g["x_final"])
    assert np.array_equal(np.stack([r["x_t"].numpy() for r in rec]), g["x_t"])
    # masked fraction falls roughly like t
    frac = (g["x_t"] == MASK).mean(axis=(1, 2))
    assert frac[0] == 1.0 and frac[-1] < 0.15 and np.all(np.diff(frac) <= 1e-9)
    # inpainting: known positions are never resampled, masked ones are all filled
    torch.manual_seed(int(g["inpaint_seed"]))
    prior = torch.from_numpy(g["prior"])
    xi = mdlm_ref.SamplerRef(net, emb).ddpm_sample(seq, 25, input_prior=prior.clone())
    assert np.array_equal(xi.numpy(), g["x_final_inpaint"])
    known = prior != MASK
    assert torch.equal(xi[known], prior[known]) and not bool((xi == MASK).any())


def test_tokenizer_pins(golden_dir):
    pins = json.loads((golden_dir / "tokenizer_pins.json").read_text())
    from esmdiff_b200.tokenization import AA_TO_ID, tokenize_sequence
    for aa, tid in pins["aa_to_id"].items():
        assert AA_TO_ID[aa] == tid
    for f in pins["files"].values():
        assert (f["seq_bos"], f["seq_eos"], f["struct_bos"], f["struct_eos"]) == (0, 2, 4098, 4097)
        assert f["struct_code_max"] < 4096 and f["emb_shape"][1] == 1536 and f["logits_shape"][1] == 4096
    t = tokenize_sequence("RPDFCLEPPY")
    assert t[0] == 0 and t[-1] == 2 and t.tolist()[1:4] == [10, 14, 13]


def test_chunk_lists(golden_dir):
    chunks = json.loads((golden_dir / "chunks.json").read_text())
    from esmdiff_b200.sampling import chunk_sizes
    for key, want in chunks.items():
        T, N = (int(v[1:]) for v in key.split("_"))
        assert mdlm_ref.chunk_sizes(T, N) == want == chunk_sizes(T, N)
    assert chunks["T258_N100"] == [63, 37] and chunks["T1026_N512"][-1] == 128     # residual quirk


def test_oracle_net_structure():
    dims = esm3_ref.Esm3Dims()
    assert dims.ffn_hidden == 4096 and abs(dims.residue_scale - 1.1547005) < 1e-6
    tiny = esm3_ref.Esm3Dims(d_model=256, n_heads=4, v_heads=8, n_layers=2)
    net, emb = esm3_ref.build_reference_model(tiny, seed=0)
    sd = esm3_ref.full_state_dict(net, emb)
    # weight ABI
    for k, shp in {"net.transformer.blocks.0.attn.layernorm_qkv.1.weight": (768, 256),
                   "net.transformer.blocks.1.ffn.1.weight": (2 * tiny.ffn_hidden, 256),
                   "net.transformer.blocks.1.ffn.3.weight": (256, tiny.ffn_hidden),
                   "net.transformer.blocks.0.geom_attn.proj.weight": (120, 256),
                   "net.output_heads.structure_head.3.weight": (4101, 256),
                   "net.encoder.structure_tokens_embed.weight": (4101, 256),
                   "sigma_embedder.mlp.0.weight": (256, 256)}.items():
        assert tuple(sd[k].shape) == shp, k
    assert "net.transformer.blocks.1.geom_attn.proj.weight" not in sd
    assert "net.transformer.norm.bias" not in sd and "net.transformer.blocks.0.attn.q_ln.bias" not in sd
    # BOS/EOS forcing and determinism
    seq = torch.tensor([[0, 5, 6, 7, 2]])
    xt = torch.full((1, 5), MASK)
    a = net(structure_tokens=xt, sequence_tokens=seq).structure_logits
    xt2 = xt.clone(); xt2[0, 0] = 17; xt2[0, -1] = 99       # overwritten by the forcing
    b = net(structure_tokens=xt2, sequence_tokens=seq).structure_logits
    assert torch.equal(a, b) and a.shape == (1, 5, 4101)
    assert esm3_ref.forward_flops(1, 258) == 258 * (48 * (56623104 + 6144 * 258) + 17316864)


def test_rotary_matches_flash_attn():
    """Second anchor for the unpinned network half: esm/layers/rotary.py is the
    flash-attn rotary module (rotate-half, non-interleaved, inv_freq = base^(-2i/d), fp32 tables);
    the copy installed in this image (flash_attn/layers/rotary.py:14-33 + RotaryEmbedding's
    inv_freq / cos-sin cache) must agree with the oracle's restatement bit for bit."""
    fr = pytest.importorskip("flash_attn.layers.rotary")
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 37, 3, 64, generator=g)
    cos, sin = esm3_ref.rotary_tables(37, 64)
    inv_freq = 1.0 / (10000.0 ** (torch.arange(0, 64, 2, dtype=torch.float32) / 64))     # RotaryEmbedding._compute_inv_freq
    freqs = torch.outer(torch.arange(37, dtype=torch.float32), inv_freq)                # _update_cos_sin_cache
    assert torch.equal(cos, torch.cos(freqs)) and torch.equal(sin, torch.sin(freqs))
    assert torch.equal(esm3_ref.apply_rotary(x, cos, sin), fr.apply_rotary_emb_torch(x, cos, sin, interleaved=False))
    rot = fr.rotate_half(x)
    assert torch.equal(rot[..., :32], -x[..., 32:]) and torch.equal(rot[..., 32:], x[..., :32])


def _masked_logp(logits, xt):
    return mdlm_ref.logits_parameterization(logits.clone(), xt)[xt == MASK][:, :4096]


def test_bf16_emulating_oracle():
    """oracle/esm3_emul.py: all rounding points off == the fp32 oracle (same algebra: LayerNorms
    folded through the Linears, q_ln / k_ln centring in the weight, 1/std applied to the scores,
    online softmax in 64-key tiles); all on == bf16 noise of the expected size.  Also measures, on
    CPU, how far two implementations with IDENTICAL rounding points drift apart when they differ
    only at fp32-ulp level (float64 vs fp32 accumulation): that drift, not 0, is the floor of any
    product-vs-emulation comparison."""
    dims = esm3_ref.Esm3Dims(d_model=256, n_heads=4, v_heads=8, n_layers=3)
    net, emb = esm3_ref.build_reference_model(dims, seed=2)
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for name, prm in net.named_parameters():            # non-trivial LayerNorm weights and biases
            if prm.dim() == 1 and "head.0.bias" not in name and "head.3.bias" not in name:
                prm.mul_(1 + 0.2 * torch.randn(prm.shape, generator=g)) if name.endswith("weight") \
                    else prm.add_(0.1 * torch.randn(prm.shape, generator=g))
    B, T = 2, 130                                            # T = 128 + 2: exercises the trailing-row path
    seq = torch.cat([torch.zeros(1, dtype=torch.long), torch.randint(4, 24, (T - 2,), generator=g),
                     torch.full((1,), 2)])[None].repeat(B, 1)
    xt = torch.randint(0, 4096, (B, T), generator=g)
    xt[torch.rand(B, T, generator=g) < 0.5] = MASK
    cond = emb(torch.tensor([0.8]))[0][None, None].expand(B, T, -1)
    ref = net(structure_tokens=xt, sequence_tokens=seq, auxiliary_embeddings=cond)
    rel = lambda a, b: float((a - b).norm() / b.norm())
    off = esm3_emul.forward(net, xt, seq, cond, esm3_emul.Rounding.none())
    assert rel(off.structure_logits, ref.structure_logits) < 2e-5 and rel(off.embeddings, ref.embeddings) < 1e-5
    budget = esm3_emul.error_budget(
        net, xt, seq, cond,
        lambda o, r: (rel(o.structure_logits, r.structure_logits),
                      rel(_masked_logp(o.structure_logits, xt), _masked_logp(r.structure_logits, xt))))
    print("\n[bf16 emulation, 3 layers d=256] (raw logits rel, masked log-prob rel) per rounding point:",
          {k: (round(a, 5), round(b, 6)) for k, (a, b) in budget.items()})
    assert 5e-4 < budget["all"][0] < 1.5e-2                 # bf16 operands: a few 1e-3 on raw logits
    assert budget["all"][1] < 1e-3                           # the path's contract metric
    for k in ("weights", "act", "qkv", "k_prescale", "p"):
        assert 0 < budget[k][0] <= 1.2 * budget["all"][0] + 1e-4
    assert budget["f64_drift"][1] < 1e-3                     # ... and it is well-posed: the drift stays below it


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference only exists in the build container")
def test_oracle_equals_reference_live():
    net, emb = esm3_ref.build_reference_model(esm3_ref.Esm3Dims(d_model=128, n_heads=2, v_heads=8, n_layers=1), 3)
    _, _, TE = ref_loader.load()
    te = TE(128).eval(); te.load_state_dict(emb.state_dict())
    ref = ref_loader.build_reference_sampler(net, te)
    seq = torch.tensor([[0, 5, 9, 12, 7, 7, 20, 2]]).repeat(2, 1)
    torch.manual_seed(9)
    a = ref.ddpm_sample(sequence_tokens=seq, num_steps=7, eps=1e-5, input_prior=None, sample_max_t=1.0)
    torch.manual_seed(9)
    b = mdlm_ref.SamplerRef(net, emb).ddpm_sample(seq, 7)
    assert torch.equal(a, b)


def test_decoder_oracle_geometry_and_emulation():
    """oracle/vqvae_ref.py (esm StructureTokenDecoder restated, parity unpinned): the geometry the
    head guarantees for any weights -- ideal N-CA / CA-C bond lengths from BB_COORDINATES, right-handed
    orthonormal frames, |C-O| = |O_VECTOR|, NaN oxygen exactly at BOS / EOS / the last residue, pLDDT
    in [0, 1] -- and the bf16-emulating decode with every rounding off equals the fp32 decode."""
    dims = vqvae_ref.DecoderDimsRef(d_model=256, n_heads=4, n_layers=2)
    dec = vqvae_ref.build_decoder(dims, seed=4)
    g = torch.Generator().manual_seed(5)
    tok = torch.randint(0, 4096, (2, 37), generator=g)
    tok[:, 0], tok[:, -1] = 4098, 4097
    out = dec.decode(tok)
    bb, o = out["bb_pred"], out["oxygen"]
    assert bb.shape == (2, 37, 3, 3) and o.shape == (2, 37, 3) and out["affine"].shape == (2, 37, 23)
    n, ca, c = bb.unbind(-2)
    assert float(((n - ca).norm(dim=-1) - 1.4592).abs().max()) < 1e-3      # |(0.5256, 1.3612, 0)|
    assert float(((c - ca).norm(dim=-1) - 1.5251).abs().max()) < 1e-3
    nan = torch.isnan(o).any(-1)
    want = torch.zeros(2, 37, dtype=torch.bool)
    want[:, 0] = want[:, -1] = want[:, -2] = True
    assert torch.equal(nan, want)
    assert float(((o - c)[~nan].norm(dim=-1) - 1.2311).abs().max()) < 1e-3
    rot = vqvae_ref.graham_schmidt(torch.randn(5, 3, generator=g), torch.randn(5, 3, generator=g))
    eye = torch.eye(3).expand(5, 3, 3)
    assert float((rot.transpose(-1, -2) @ rot - eye).abs().max()) < 1e-5 and float((torch.linalg.det(rot) - 1).abs().max()) < 1e-5
    assert float(out["plddt"].min()) >= 0.0 and float(out["plddt"].max()) <= 1.0
    off = vqvae_ref.decode_emulated(dec, tok, esm3_emul.Rounding.none())
    assert float((off["affine"] - out["affine"]).norm() / out["affine"].norm()) < 2e-5
    on = vqvae_ref.decode_emulated(dec, tok)
    assert 1e-4 < float((on["affine"] - out["affine"]).norm() / out["affine"].norm()) < 2e-2
