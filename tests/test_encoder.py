"""SURVEY.md 8f row 3 -- the VQ-VAE structure ENCODER front end (reference slm/models/utils.py:99-146,
slm/sample_esmdiff.py:166-175, 197-209, 278-284) and the geometric attention it is made of (also block 0 of the
sampling network when coordinates are given, net.py:433-441).

The arithmetic is esm==3.0.4's (not vendored, weights not available offline): oracle/vqvae_enc_ref.py and
oracle/geom_ref.py are restatements, PARITY UNPINNED.  What the CPU tests pin is the restatement's own algebra
(SE(3) invariance of the codes, masking rules, neighbour order); the GPU tests compare the CUDA path with it.
The product computes this path in fp32: codes must be EQUAL except where the two nearest codes are closer than
the fp32 noise of the distance (counted, bounded); pre-quantisation vectors within 2e-4 relative.
"""
import numpy as np
import pytest
import torch

from oracle import geom_ref, vqvae_enc_ref as V

DEV = "cuda"
BPTI = "RPDFCLEPPYTGPCKARIIRYFYNAKAGLCQTFVYGGCRAKRNNFKSAEDCMRTCGGA"
TINY = dict(d_model=128, v_heads=16, n_layers=2, d_out=32, n_codes=256)


def rel_fro(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20))


def _oracle_encoder(dims_kw, seed=0):
    enc = V.build_encoder(V.EncoderDimsRef(**dims_kw), seed=seed)
    with torch.no_grad():          # put the codes where the pre-quantisation vectors of a chain live (a trained codebook
        z = enc.encode(V.synthetic_backbone(64, seed=99)[None], return_all=True)["z"][0]      # does); N(0, 1) codes
        g = torch.Generator().manual_seed(seed)                                               # leave a handful of winners
        e = enc.codebook.embeddings
        e.copy_(z.mean(0) + 1.5 * z.std(0) * torch.randn(e.shape, generator=g))
    return enc


# ---------------------------------------------------------------------------------------------- CPU: the oracle
def test_oracle_frames_and_black_hole():
    bb = V.synthetic_backbone(12, seed=3)[None].repeat(2, 1, 1, 1)
    bb[0, 4] = float("nan")
    bb[0, 7, 1, 2] = float("inf")
    bb[1] = float("nan")
    rot, trans, mask = geom_ref.build_affine3d_from_coordinates(bb)
    assert mask[0].tolist() == [True] * 4 + [False] + [True] * 2 + [False] + [True] * 4 and not mask[1].any()
    eye = torch.eye(3)
    assert torch.allclose(rot @ rot.transpose(-1, -2), eye.expand_as(rot), atol=1e-5)          # proper frames everywhere
    assert torch.allclose(rot[1], eye.expand(12, 3, 3)) and float(trans[1].abs().max()) == 0.0   # no coordinates: identity
    assert torch.equal(rot[0, 4], rot[0, 7]) and torch.equal(trans[0, 4], trans[0, 7])         # the black-hole frame
    # a valid residue: CA is the origin, C on the negative x axis, N in the xy plane
    local = torch.einsum("ji,aj->ai", rot[0, 0], bb[0, 0] - trans[0, 0])
    assert local[1].abs().max() < 1e-5 and local[2, 0] < 0 and local[2, 1:].abs().max() < 1e-5 and abs(float(local[0, 2])) < 1e-5


def test_oracle_knn_order_and_sequence_fallback():
    bb = V.synthetic_backbone(40, seed=1)
    mask = torch.ones(1, 40, dtype=torch.bool)
    mask[0, 10:14] = False
    edges = V.knn_graph(bb[None, :, 1], mask, 16)[0]
    assert edges.shape == (40, 16) and edges[:, 0].tolist() == list(range(40))                  # self first
    d = (bb[:, None, 1] - bb[None, :, 1]).norm(dim=-1)
    i = 20
    valid = [j for j in edges[i].tolist()]
    assert all(mask[0, j] for j in valid)                                                        # 36 valid residues >= 16
    assert valid == sorted(valid, key=lambda j: float(d[i, j]))
    # a frameless residue: neighbours by sequence distance, lower index first on ties
    assert edges[12, :5].tolist() == [12, 11, 13, 10, 14]
    assert V.knn_graph(bb[None, :7, 1], mask[:, :7], 16).shape == (1, 7, 7)                      # L < knn


def test_oracle_codes_are_se3_invariant_and_masking_is_local():
    enc = _oracle_encoder(TINY)
    bb = V.synthetic_backbone(48, seed=1)
    a = enc.encode(bb[None], return_all=True)
    assert len(set(a["codes"][0].tolist())) > 8                                                  # a discriminating test set-up
    R = geom_ref.graham_schmidt(torch.tensor([0.3, -1.0, 0.5]), torch.tensor([1.0, 0.2, -0.4]))
    b = enc.encode((bb @ R.T + torch.tensor([30.0, -12.0, 7.0]))[None], return_all=True)
    assert rel_fro(b["z"], a["z"]) < 1e-4
    assert float((a["codes"] == b["codes"]).float().mean()) > 0.95
    # masking residues 5..9: their codes collapse to the code nearest to pre_vq_proj.bias; residues whose
    # neighbourhoods do not touch them keep their codes
    bb2 = bb.clone()
    bb2[5:10] = float("inf")
    c = enc.encode(bb2[None], return_all=True)
    assert len(set(c["codes"][0, 5:10].tolist())) == 1
    untouched = [i for i in range(48) if not any(5 <= j < 10 for j in a["edges"][0, i].tolist())]
    assert len(untouched) > 5 and all(int(a["codes"][0, i]) == int(c["codes"][0, i]) for i in untouched)
    tok = V.tokenize_structure(enc, bb)
    assert tok.shape == (50,) and int(tok[0]) == 4098 and int(tok[-1]) == 4097 and int(tok[1:-1].max()) < 256


def test_pdb_coordinates_reader(tmp_path):
    from esmdiff_b200.encoder import ATOM37, coordinates_from_pdb, normalize_coordinates
    assert len(ATOM37) == 37 and ATOM37[:5] == ("N", "CA", "C", "CB", "O")
    bb = V.synthetic_backbone(5, seed=0)
    lines = []
    for i in range(5):
        for a, name in enumerate(("N", "CA", "C")):
            x, y, z = bb[i, a].tolist()
            lines.append(f"ATOM  {3 * i + a + 1:5d}  {name:<3s} ALA A{i + 1:4d}    {x:8.3f}{y:8.3f}{z:8.3f}  1.00  0.00           {name[0]:>2s}")
    lines.insert(4, lines[3][:16] + "B" + lines[3][17:30] + "   0.000   0.000   0.000" + lines[3][54:])   # altloc B of residue 2's N
    (tmp_path / "x.pdb").write_text("\n".join(lines) + "\nEND\n")
    seq, c = coordinates_from_pdb(tmp_path / "x.pdb")
    assert seq == "AAAAA" and c.shape == (5, 37, 3)
    assert torch.allclose(c[:, :3], bb, atol=1e-3) and torch.isnan(c[:, 3:]).all()
    n = normalize_coordinates(c)
    assert torch.allclose((n[1:, 1] - n[:-1, 1]).norm(dim=-1), (bb[1:, 1] - bb[:-1, 1]).norm(dim=-1), atol=1e-3)
    assert n[:, 1].mean(0).abs().max() < 1e-3


# ---------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_backbone_frames_kernel():
    import ctypes as C
    from esmdiff_b200 import _lib
    L = _lib.lib()
    bb = torch.stack([V.synthetic_backbone(70, seed=s) for s in range(3)])
    bb[0, 4] = float("nan")
    bb[0, 30:40, 2, 1] = float("inf")
    bb[1] = float("nan")
    bb[2] += 500.0
    rot, trans, mask = geom_ref.build_affine3d_from_coordinates(bb)
    c = bb.to(DEV).contiguous()
    r = torch.empty(3 * 70, 9, device=DEV)
    t = torch.empty(3 * 70, 3, device=DEV)
    m = torch.empty(3 * 70, dtype=torch.uint8, device=DEV)
    assert L.esmdiff_op_backbone_frames(c.data_ptr(), 3, 70, r.data_ptr(), t.data_ptr(), m.data_ptr(), None) == 0
    torch.cuda.synchronize()
    assert torch.equal(m.cpu().bool().view(3, 70), mask)
    assert float((r.cpu().view(3, 70, 3, 3) - rot).abs().max()) < 2e-5
    assert float((t.cpu().view(3, 70, 3) - trans).abs().max()) < 1e-3 * 1e-1


@pytest.mark.gpu
@pytest.mark.parametrize("G,S,H,use_idx,zero", [(37, 16, 128, True, 0), (5, 70, 256, False, 1), (3, 258, 64, False, 1),
                                                 (9, 7, 16, True, 0)])
def test_geometric_attention_kernel(G, S, H, use_idx, zero):
    from esmdiff_b200 import _lib
    L = _lib.lib()
    g = torch.Generator().manual_seed(G * 1000 + S)
    ga = geom_ref.GeometricReasoningRef(64, H, mask_and_zero_frameless=bool(zero))
    with torch.no_grad():
        ga.distance_scale_per_head.copy_(torch.randn(H, generator=g))
        ga.rotation_scale_per_head.copy_(torch.randn(H, generator=g))
    nres = 90 if use_idx else G * S
    bb = V.synthetic_backbone(nres, seed=S)[None]
    bb[0, 3] = float("nan")
    bb[0, 11] = float("nan")
    rot, trans, mask = geom_ref.build_affine3d_from_coordinates(bb)
    rot, trans, mask = rot[0], trans[0], mask[0]
    idx = torch.randint(0, nres, (G, S), generator=g) if use_idx else torch.arange(G * S).view(G, S)
    if use_idx:
        idx[0] = torch.tensor([3, 11] * S)[:S]                           # a group whose keys are all frameless
    p = torch.randn(G, S, 15 * H, generator=g)
    with torch.no_grad():
        want = ga.attention(p, rot[idx], trans[idx], mask[idx])
    pd, out, work = p.to(DEV), torch.empty(G, S, 3 * H, device=DEV), torch.empty(G, S, 15 * H, device=DEV)
    r, t, m = rot.reshape(-1, 9).contiguous().to(DEV), trans.contiguous().to(DEV), mask.to(torch.uint8).to(DEV)
    idx_d = idx.to(torch.int32).to(DEV).contiguous() if use_idx else None
    w_r, w_d = ga.rotation_scale_per_head.detach().to(DEV), ga.distance_scale_per_head.detach().to(DEV)
    rc = L.esmdiff_op_geometric_attention(pd.data_ptr(), r.data_ptr(), t.data_ptr(), m.data_ptr(),
                                          idx_d.data_ptr() if use_idx else None,
                                          w_r.data_ptr(), w_d.data_ptr(), G, S, H, zero,
                                          work.data_ptr(), out.data_ptr(), None)
    assert rc == 0
    torch.cuda.synchronize()
    err = float((out.cpu() - want).abs().max())
    print(f"[geom attention G={G} S={S} H={H}] max abs err {err:.2e}, rel {rel_fro(out.cpu(), want):.2e}")
    assert err < 2e-4 and rel_fro(out.cpu(), want) < 2e-5


def _compare_codes(got, want, dist):
    """ids equal except near-ties of the two best codes (margin below the fp32 noise of the distance)."""
    bad = (got != want).nonzero().flatten().tolist()
    excused = 0
    for i in bad:
        d = dist[i]
        margin = float((d[got[i]] - d[want[i]]).abs())
        assert margin < 2e-4 * float(d[want[i]].abs().clamp_min(1.0)), f"residue {i}: code {int(got[i])} vs {int(want[i])}, margin {margin}"
        excused += 1
    return excused


@pytest.mark.gpu
@pytest.mark.parametrize("name,dims_kw,B,L", [("tiny", TINY, 2, 48), ("tiny_short", TINY, 1, 9),
                                               ("esm3_encoder_v0", dict(), 1, 256), ("esm3_encoder_v0_b2", dict(), 2, 70)])
def test_encoder_vs_oracle(name, dims_kw, B, L):
    from esmdiff_b200.encoder import EncoderDims, StructureTokenEncoder
    ref = _oracle_encoder(dims_kw, seed=1)
    enc = StructureTokenEncoder(EncoderDims(**dims_kw))
    enc.load_state_dict(ref.state_dict())
    bb = torch.stack([V.synthetic_backbone(L, seed=10 + b) for b in range(B)])
    if L > 20:
        bb[0, 5:9] = float("inf")                          # inpainting-style holes (utils.py:121)
        bb[B - 1, L - 3] = float("nan")
    ri = torch.arange(1, L + 1)[None].repeat(B, 1)
    ri[:, L // 2:] += 7                                    # a numbering gap
    want = ref.encode(bb, residue_index=ri, return_all=True)
    got = enc.encode(bb, residue_index=ri, return_aux=True)
    torch.cuda.synchronize()
    edges = got["edges"].cpu().long()
    valid_rows = want["mask"]
    assert torch.equal(edges[valid_rows][:, 0], want["edges"][valid_rows][:, 0])
    same_edges = float((edges == want["edges"]).float().mean())
    assert same_edges > 0.999, same_edges
    z_err = rel_fro(got["z"].cpu(), want["z"])
    excused = _compare_codes(got["codes"].cpu().flatten(), want["codes"].flatten(), want["dist"].flatten(0, 1))
    n_codes = len(set(want["codes"].flatten().tolist()))
    print(f"[encoder {name}] z rel {z_err:.2e}, codes differing at near-ties {excused}/{B * L}, distinct codes {n_codes}")
    assert z_err < 2e-4
    assert excused <= max(1, B * L // 100)
    zq, codes = enc.encode(bb, residue_index=ri)
    assert torch.equal(codes, got["codes"]) and torch.equal(zq.cpu(), ref.codebook.embeddings[codes.cpu()])


@pytest.mark.gpu
def test_encoder_is_se3_invariant_full_size():
    """Size-independent property at the reference's dims: a rigid motion of the chain leaves the codes alone."""
    from esmdiff_b200.encoder import load_encoder
    enc = load_encoder(None, seed=3)
    bb = V.synthetic_backbone(512, seed=5)
    R = geom_ref.graham_schmidt(torch.tensor([0.3, -1.0, 0.5]), torch.tensor([1.0, 0.2, -0.4]))
    a = enc.encode(bb[None], return_aux=True)
    b = enc.encode((bb @ R.T + torch.tensor([30.0, -12.0, 7.0]))[None], return_aux=True)
    assert rel_fro(b["z"], a["z"]) < 1e-4
    assert float((a["codes"] == b["codes"]).float().mean()) > 0.98


THREE = {"R": "ARG", "P": "PRO", "D": "ASP", "F": "PHE", "C": "CYS", "L": "LEU", "E": "GLU", "Y": "TYR", "T": "THR",
         "G": "GLY", "K": "LYS", "A": "ALA", "I": "ILE", "N": "ASN", "Q": "GLN", "V": "VAL", "S": "SER", "M": "MET"}


def write_backbone_pdb(path, seq, bb):
    lines = []
    for i, aa in enumerate(seq):
        for a, name in enumerate(("N", "CA", "C")):
            x, y, z = bb[i, a].tolist()
            lines.append(f"ATOM  {3 * i + a + 1:5d}  {name:<3s} {THREE[aa]} A{i + 1:4d}    {x:8.3f}{y:8.3f}{z:8.3f}  1.00  0.00           {name[0]:>2s}")
    path.write_text("\n".join(lines) + "\nEND\n")


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["ddpm", "gibbs"])
def test_cli_inpainting_from_pdb_coordinates(tmp_path, capsys, mode):
    """``--mask_ids`` end to end from a PDB with backbone coordinates (sample_esmdiff.py:278-294): the VQ-VAE encoder
    gives the prompt's structure tokens; ddpm: prior = tokens with TOKEN positions mask_ids masked (:197-201, the
    reference's off-by-one), every other position comes back unchanged; gibbs: the known residues keep their codes,
    the masked ones are sampled with the known backbone frames live in block 0's geometric attention."""
    from test_abi_and_host import write_run_dir
    from esmdiff_b200 import sample_esmdiff
    from esmdiff_b200.encoder import coordinates_from_pdb, load_encoder, tokenize_structure
    from oracle import esm3_ref
    tiny = dict(d_model=256, n_heads=4, v_heads=64, n_layers=2)
    net, emb = esm3_ref.build_reference_model(esm3_ref.Esm3Dims(**tiny), seed=5)
    sd = esm3_ref.full_state_dict(net, emb)
    extra = "\n".join(f"    {k}: {v}" for k, v in tiny.items())
    ckpt = write_run_dir(tmp_path / "run", "file", net_extra=extra, hidden=tiny["d_model"], module=sd)
    inp = tmp_path / "targets"
    inp.mkdir()
    write_backbone_pdb(inp / "bpti.pdb", BPTI, V.synthetic_backbone(len(BPTI), seed=4))
    out = tmp_path / "out"
    mask_ids = list(range(3, 11))
    sample_esmdiff.main(["--input", str(inp), "--ckpt", str(ckpt), "--output", str(out), "--mode", mode, "--num_steps", "6",
                         "--num_samples", "4", "--seed", "1", "--mask_ids", ",".join(map(str, mask_ids)),
                         "--encoder_ckpt", "random"])
    text = capsys.readouterr().out
    assert "Sampling token time" in text
    saved = list(out.glob("*/bpti.structure_tokens.pt"))
    assert len(saved) == 1
    blob = torch.load(saved[0], weights_only=False)
    tokens = blob["structure_tokens"]
    assert tokens.shape == (4, len(BPTI)) and int(tokens.max()) < 4096 and int(tokens.min()) >= 0
    assert blob["sequence"] == BPTI[:3] + "_" * 8 + BPTI[11:]
    seq, coords = coordinates_from_pdb(inp / "bpti.pdb")
    coords[mask_ids] = float("inf")
    prompt = tokenize_structure(coords, load_encoder(None))[1:-1]       # the same random-init encoder (seed 0)
    if mode == "ddpm":
        kept = [i for i in range(len(BPTI)) if (i + 1) not in mask_ids]   # token position i + 1 <-> residue i
    else:
        kept = [i for i in range(len(BPTI)) if i not in mask_ids]
        assert "Masking 8 residues and inpainting..." in text
    assert torch.equal(tokens[:, kept], prompt[kept][None].expand(4, -1))
    changed = [i for i in range(len(BPTI)) if i not in kept]
    assert len(set(map(tuple, tokens[:, changed].tolist()))) > 1          # the samples differ where they were sampled


@pytest.mark.gpu
def test_protseq_to_data_inpainting_front_end(tmp_path):
    """utils.py:105-146 with mask_ids, from a PDB file: sequence '_' + ids 32 at the masked residues, their
    coordinates inf, structure tokens BOS + codes + EOS equal to the oracle's tokenize_structure."""
    from esmdiff_b200.encoder import EncoderDims, StructureTokenEncoder, pdb_to_data
    from esmdiff_b200.sampling import build_prior
    ref = _oracle_encoder(TINY, seed=2)
    enc = StructureTokenEncoder(EncoderDims(**TINY))
    enc.load_state_dict(ref.state_dict())
    bb = V.synthetic_backbone(len(BPTI), seed=4)
    write_backbone_pdb(tmp_path / "bpti.pdb", BPTI, bb)
    mask_ids = list(range(1, 9))
    data = pdb_to_data(tmp_path / "bpti.pdb", enc, encode_only=True, mask_ids=mask_ids)
    assert data["sequence"] == BPTI[0] + "_" * 8 + BPTI[9:]
    assert data["sequence_tokens"][2:10].tolist() == [32] * 8 and int(data["sequence_tokens"][1]) != 32
    assert torch.isinf(data["coordinates"][1:9]).all() and torch.isfinite(data["coordinates"][0, :3]).all()
    rounded = torch.tensor([[float(f"{v:.3f}") for v in row] for row in bb.reshape(-1, 3).tolist()]).view(-1, 3, 3)
    rounded[1:9] = float("inf")
    want = V.tokenize_structure(ref, rounded)
    got = data["structure_tokens"]
    assert got.shape == want.shape == (len(BPTI) + 2,) and int(got[0]) == 4098 and int(got[-1]) == 4097
    assert int((got != want).sum()) <= 1
    # the prior the sampler starts from: token positions mask_ids set to MASK (the reference's off-by-one, :197-201)
    prior = build_prior(got, 2, mask_ids=mask_ids)
    assert prior.shape == (2, len(BPTI) + 2) and (prior[:, 1:9] == 4096).all() and int(prior[0, 9]) == int(got[9])


@pytest.mark.gpu
def test_network_forward_with_structure_coords_tiny():
    """SURVEY.md 8a A6 live at tiny dims through the host mirror (``CustomizedESM3.forward(structure_coords=)``,
    net.py:385, 433-441): block 0's geometric attention on, against the fp32 and the bf16-emulating oracle."""
    from esmdiff_b200.net import CustomizedESM3
    from oracle import esm3_emul, esm3_ref
    tiny = dict(d_model=256, n_heads=4, v_heads=64, n_layers=2)
    net, emb = esm3_ref.build_reference_model(esm3_ref.Esm3Dims(**tiny), seed=3)
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        net.transformer.blocks[0].geom_attn.distance_scale_per_head.normal_(generator=g)
        net.transformer.blocks[0].geom_attn.rotation_scale_per_head.normal_(generator=g)
    model = CustomizedESM3(**tiny)
    model.load_state_dict(esm3_ref.full_state_dict(net, emb))
    B, T = 3, 70
    seq = torch.randint(4, 24, (B, T), generator=g)
    seq[:, 0], seq[:, -1] = 0, 2
    xt = torch.randint(0, 4096, (B, T), generator=g)
    xt[:, 5:30] = 4096
    coords = torch.full((B, T, 37, 3), float("nan"))                     # an atom37 array: [..., :3, :] is used
    coords[:, 1:-1, :3] = torch.stack([V.synthetic_backbone(T - 2, seed=20 + b) for b in range(B)])
    coords[:, 5:30] = float("inf")
    coords[2] = float("nan")                                             # a sample without any frame: exact zero branch
    with torch.no_grad():
        ref = net(xt, sequence_tokens=seq, structure_coords=coords).structure_logits
        ref0 = net(xt, sequence_tokens=seq).structure_logits
        emu = esm3_emul.forward(net, xt, seq, None, structure_coords=coords).structure_logits
    got = model(xt.to(DEV), sequence_tokens=seq.to(DEV), structure_coords=coords).structure_logits.float().cpu()
    got0 = model(xt.to(DEV), sequence_tokens=seq.to(DEV)).structure_logits.float().cpu()
    e_ref, e_emu, effect = rel_fro(got, ref), rel_fro(got, emu), rel_fro(ref[:2], ref0[:2])
    print(f"[structure_coords tiny] logits rel vs fp32 {e_ref:.2e}, vs emul {e_emu:.2e}, without coords {rel_fro(got0, ref0):.2e}, "
          f"effect of the branch {effect:.3f}")
    assert e_ref < 8.5e-3 and e_emu < 6.4e-3 and effect > 0.05
    assert torch.equal(got[2], got0[2])                                   # frameless sample: bit-identical to no coordinates


@pytest.mark.gpu
def test_structure_coords_error_behaviour(tiny_pair):
    """esmdiff_set_structure_coords: a context whose v_heads cannot run the branch refuses (the fixture's tiny net has
    v_heads = 8: 3 v_heads % 64 != 0); a batch shape other than the one the coordinates were given for fails loudly
    instead of reading frames of other rows; NULL switches the branch off again."""
    from esmdiff_b200._lib import EsmdiffError
    from esmdiff_b200.engine import Dims, Engine
    from esmdiff_b200.synthetic import random_state_dict
    _, _, eng8 = tiny_pair
    with pytest.raises(EsmdiffError, match="v_heads"):
        eng8.set_structure_coords(torch.zeros(2, 10, 3, 3))
    dims = Dims(d_model=256, n_heads=4, v_heads=64, n_layers=2)
    eng = Engine(dims)
    sd = random_state_dict(dims, device=DEV, seed=0, full=False)          # a checkpoint WITHOUT the geom_attn.* keys
    eng.load_state_dict(sd)
    seq = torch.full((2, 20), 5)
    xt = torch.full((2, 20), 4096)
    eng.set_structure_coords(torch.zeros(2, 20, 3, 3))
    with pytest.raises(EsmdiffError, match="geom_attn"):
        eng.forward(seq, xt)
    eng.set_structure_coords(None)
    eng.forward(seq, xt)                                                 # fine again without coordinates
    sd = random_state_dict(dims, device=DEV, seed=0, full=True)
    eng.load_state_dict(sd)
    base, _ = eng.forward(seq, xt)
    eng.set_structure_coords(torch.zeros(2, 20, 3, 3))
    with pytest.raises(EsmdiffError, match="batch shape"):
        eng.forward(seq[:1], xt[:1])
    eng.set_structure_coords(None)
    again, _ = eng.forward(seq, xt)
    eng.synchronize()
    assert torch.equal(again, base)
    eng.close()


def test_sdk_protein_roundtrip_on_cpu(tmp_path):
    """The esm SDK surface mirrors (esmdiff_b200/sdk.py): ESMProtein.to_pdb -> from_pdb gives the backbone back (PDB
    precision), '_' residues are written as UNK / read as X, ESMProteinTensor.to and len behave like esm's."""
    from esmdiff_b200.sdk import ESMProtein, ESMProteinTensor
    bb = V.synthetic_backbone(7, seed=2)
    coords = torch.full((7, 37, 3), float("nan"))
    coords[:, :3] = bb
    coords[:, 4] = bb[:, 2] + torch.tensor([0.5, -1.0, 0.2])
    coords[6, 4] = float("nan")                                           # the last residue has no O (infer_oxygen)
    prot = ESMProtein(sequence="ACD_FGH", coordinates=coords, plddt=torch.linspace(0.1, 0.9, 7))
    assert len(prot) == 7
    prot.to_pdb(tmp_path / "p.pdb")
    text = (tmp_path / "p.pdb").read_text().splitlines()
    assert all(len(ln) == 80 for ln in text) and text[-1].strip() == "END" and text[-2].startswith("TER")
    assert sum(ln.startswith("ATOM") for ln in text) == 7 * 4 - 1
    back = ESMProtein.from_pdb(tmp_path / "p.pdb")
    assert back.sequence == "ACDXFGH"
    assert torch.allclose(back.coordinates[:, :3], bb, atol=1e-3) and torch.isnan(back.coordinates[6, 4]).all()
    assert abs(float(text[0][60:66]) - 0.1) < 0.006                        # pLDDT in the B-factor column
    t = ESMProteinTensor(sequence=torch.arange(9), structure=None)
    assert len(t) == 9 and t.to("cpu").structure is None and torch.equal(t.to("cpu").sequence, t.sequence)


@pytest.mark.gpu
def test_sdk_surface_runs_the_reference_helpers(tmp_path):
    """esm3_model.encode(ESMProtein(sequence, coordinates)) (models/utils.py:136-137) and the reference's per-sample
    decode(structure_tokens, sequence_tokens, esm3_model, save_to) (sample_esmdiff.py:41-61) through the SDK mirror:
    tokens equal protseq_to_data's, the PDB equals MODEL 1 of the batched writer's file."""
    from esmdiff_b200 import sdk
    from esmdiff_b200.decoder import decode_to_pdb, load_decoder
    from esmdiff_b200.encoder import load_encoder, protseq_to_data
    from esmdiff_b200.tokenization import tokenize_sequence
    write_backbone_pdb(tmp_path / "bpti.pdb", BPTI, V.synthetic_backbone(len(BPTI), seed=4))
    prot = sdk.ESMProtein.from_pdb(tmp_path / "bpti.pdb")
    assert prot.sequence == BPTI and prot.coordinates.shape == (len(BPTI), 37, 3)
    esm3 = sdk.ESM3(structure_encoder=load_encoder(None), structure_decoder=load_decoder(None))
    masked = sdk.ESMProtein(sequence=BPTI[:2] + "__" + BPTI[4:], coordinates=prot.coordinates.clone())
    masked.coordinates[2:4] = float("inf")
    toks = esm3.encode(masked)
    want = protseq_to_data(BPTI, esm3.get_structure_encoder(), encode_only=True, coordinates=prot.coordinates, mask_ids=[2, 3])
    assert torch.equal(toks.sequence, want["sequence_tokens"]) and torch.equal(toks.structure, want["structure_tokens"])
    assert toks.coordinates.shape == (len(BPTI) + 2, 37, 3) and torch.isinf(toks.coordinates[0]).all()
    assert len(toks) == len(BPTI) + 2 and toks.to(DEV).structure.is_cuda
    st = toks.structure[1:-1]
    raw = sdk.decode(st, tokenize_sequence(BPTI)[1:-1], esm3, save_to=tmp_path / "one.pdb")
    assert raw.sequence == BPTI and raw.coordinates.shape == (len(BPTI), 37, 3)
    decode_to_pdb(esm3.get_structure_decoder(), st[None], BPTI, tmp_path / "batched.pdb")
    one = [ln for ln in (tmp_path / "one.pdb").read_text().splitlines() if ln.startswith(("ATOM", "TER"))]
    bat = [ln for ln in (tmp_path / "batched.pdb").read_text().splitlines() if ln.startswith(("ATOM", "TER"))]
    assert one == bat and len(one) == 4 * len(BPTI) - 1 + 1


def test_frames_and_atom37_against_installed_openfold_utils():
    """Second anchors for the unpinned restatements, from code that IS installed here: transformers ships the
    OpenFold utilities ESMFold uses, and esm's ``Affine3D.from_graham_schmidt(neg_x_axis, origin, xy_plane)`` states
    it follows AlphaFold's argument convention -- OpenFold's ``Rigid.from_3_points(p_neg_x_axis, origin, p_xy_plane)``
    is that construction.  The oracle's backbone frames (C, CA, N) must equal it; the atom37 order and the
    three-letter table of the PDB reader must equal ``residue_constants``."""
    rc = pytest.importorskip("transformers.models.esm.openfold_utils.residue_constants")
    ru = pytest.importorskip("transformers.models.esm.openfold_utils.rigid_utils")
    from esmdiff_b200.decoder import ONE_TO_THREE
    from esmdiff_b200.encoder import ATOM37
    from esmdiff_b200.tokenization import THREE_TO_ONE
    assert list(ATOM37) == list(rc.atom_types)
    for one, three in rc.restype_1to3.items():
        assert THREE_TO_ONE[three] == one and ONE_TO_THREE[one] == three
    bb = V.synthetic_backbone(50, seed=9)
    rot, trans = geom_ref.backbone_frames(bb)
    rigid = ru.Rigid.from_3_points(bb[:, 2], bb[:, 1], bb[:, 0], eps=1e-12)
    assert torch.allclose(rigid.get_rots().get_rot_mats(), rot, atol=2e-6)
    assert torch.equal(rigid.get_trans(), trans)
    # and the decoder side: Dim6RotStructureHead's frame from (trans, x, y) is the same construction
    from oracle import vqvae_ref
    x, y, t = torch.randn(20, 3), torch.randn(20, 3), torch.randn(20, 3)
    want = ru.Rigid.from_3_points(x + t, t, y + t, eps=1e-12).get_rots().get_rot_mats()
    assert torch.allclose(vqvae_ref.graham_schmidt(t - (x + t), (y + t) - t, 1e-12), want, atol=2e-5)


def test_ideal_backbone_constants_against_literature_positions():
    """esm's BB_COORDINATES (the residue-frame N, CA, C the decoder places, oracle/vqvae_ref.py) are the AlphaFold
    literature positions with the x axis flipped (esm's frame points from C to CA): within 0.01 A of
    residue_constants.rigid_group_atom_positions averaged over the 20 residue types; the carbonyl O vector of
    infer_oxygen has the literature C=O length and direction (psi-frame (0.626, 1.062, 0), y flipped)."""
    rc = pytest.importorskip("transformers.models.esm.openfold_utils.residue_constants")
    from oracle import vqvae_ref
    pos = {a: [] for a in ("N", "CA", "C")}
    o = []
    for res, atoms in rc.rigid_group_atom_positions.items():
        for name, group, xyz in atoms:
            if name in pos and group == 0:
                pos[name].append(xyz)
            if name == "O":
                o.append(xyz)
    lit = torch.tensor([[sum(c) / len(c) for c in zip(*pos[a])] for a in ("N", "CA", "C")])
    lit[:, 0] *= -1
    assert float((torch.tensor(vqvae_ref.BB_COORDINATES) - lit).abs().max()) < 0.012
    o_lit = torch.tensor([sum(c) / len(c) for c in zip(*o)])
    o_esm = torch.tensor(vqvae_ref.O_VECTOR)
    assert abs(float(o_esm.norm() - o_lit.norm())) < 0.01 and float((o_esm - o_lit * torch.tensor([1.0, -1.0, 1.0])).abs().max()) < 0.02
