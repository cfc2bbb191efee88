"""fp32 CPU restatement of the VQ-VAE structure decode step (ORACLE -- tests only).

What this follows
-----------------
* reference call sites: ``decode`` slm/sample_esmdiff.py:41-61 (BOS/EOS added, ``esm3_model.decode``,
  ``to_pdb``), the serial per-sample loop :225-230 and ``merge_pdbfiles`` eval_utils.py:437-492.
* the arithmetic lives in ``esm==3.0.4`` (requirements.txt:30), absent from /root/reference and not
  installed; restated from the published package:
    ``StructureTokenDecoder`` (esm/models/vqvae.py; ESM3_structure_decoder_v0 = d_model 1280,
      20 heads, 30 blocks): ``embed`` -> ``TransformerStack(d, h, 1, n_layers, scale_residue=False,
      n_layers_geom=0)`` -> ``Dim6RotStructureHead(d, 10, predict_torsion_angles=False)``,
      ``plddt_head = RegressionHead(d, 50)`` -> ``CategoricalMixture(...).mean()``;
    ``Dim6RotStructureHead`` (esm/layers/structure_proj.py), ``Affine3D.from_graham_schmidt`` /
      ``_graham_schmidt`` (esm/utils/structure/affine3d.py), ``BB_COORDINATES``;
    ``ProteinChain.infer_oxygen`` (esm/utils/structure/protein_chain.py).
  **PARITY UNPINNED**: no test, fixture or runnable copy of these layers exists in the reference
  tree, and the pretrained decoder weights are not available offline.  Pinned by the reference:
  the decoder's embedding width 1280 (slm/models/net.py:94,102), BOS/EOS ids, the call order and
  the PDB layout ``merge_pdbfiles`` produces.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn.functional as F
from torch import nn

from . import esm3_ref

BB_COORDINATES = [[0.5256, 1.3612, 0.0], [0.0, 0.0, 0.0], [-1.5251, 0.0, 0.0]]     # N, CA, C in the residue frame
O_VECTOR = [0.6240, -1.0613, 0.0103]


@dataclass
class DecoderDimsRef:
    d_model: int = 1280
    n_heads: int = 20
    n_layers: int = 30
    plddt_bins: int = 50
    struct_vocab: int = 4101
    d_head: int = 64
    v_heads: int = 1
    residue_scale: float = 1.0            # scale_residue=False

    @property
    def ffn_hidden(self) -> int:
        return int(((8.0 / 3.0 * self.d_model) + 255) // 256 * 256)


def graham_schmidt(x_axis, xy_plane, eps=1e-12):
    """esm _graham_schmidt: columns [e0, e1, e2]."""
    e1 = xy_plane
    x_axis = x_axis / torch.sqrt((x_axis ** 2).sum(-1, keepdim=True) + eps)
    dot = (x_axis * e1).sum(-1, keepdim=True)
    e1 = e1 - x_axis * dot
    e1 = e1 / torch.sqrt((e1 ** 2).sum(-1, keepdim=True) + eps)
    e2 = torch.cross(x_axis, e1, dim=-1)
    return torch.stack([x_axis, e1, e2], dim=-1)


class Dim6RotStructureHeadRef(nn.Module):
    def __init__(self, d: int, trans_scale_factor: float = 10.0):
        super().__init__()
        self.ffn1 = nn.Linear(d, d)
        self.norm = nn.LayerNorm(d)
        self.proj = nn.Linear(d, 9 + 7 * 2)
        self.trans_scale_factor = trans_scale_factor

    def forward(self, x):
        p = self.proj(self.norm(F.gelu(self.ffn1(x))))
        return p, frames_to_backbone(p, self.trans_scale_factor)


def frames_to_backbone(p, trans_scale_factor=10.0):
    trans, x, y, _ = p.split([3, 3, 3, 14], dim=-1)
    trans = trans * trans_scale_factor
    x = x / (x.norm(dim=-1, keepdim=True) + 1e-5)
    y = y / (y.norm(dim=-1, keepdim=True) + 1e-5)
    # Affine3D.from_graham_schmidt(neg_x_axis = x + trans, origin = trans, xy_plane = y + trans),
    # composed with the identity; the all-False affine_mask keeps every update
    rot = graham_schmidt(trans - (x + trans), (y + trans) - trans, 1e-12)
    local = torch.tensor(BB_COORDINATES, dtype=p.dtype)
    return torch.einsum("...ij,aj->...ai", rot, local) + trans[..., None, :]


def infer_oxygen(bb):
    """bb (..., T, 3, 3) for the residues of ONE chain (no BOS/EOS) -> O (..., T, 3); NaN for the last."""
    n, ca, c = bb.unbind(-2)
    n_next = torch.roll(n, -1, dims=-2).clone()
    n_next[..., -1, :] = float("nan")
    rot = graham_schmidt(c - ca, n_next - c, 1e-10)     # from_graham_schmidt(CA, C, N_next): x = C - CA, plane = N - C
    return torch.einsum("...ij,j->...i", rot, torch.tensor(O_VECTOR, dtype=bb.dtype)) + c


def plddt_mean(logits):
    bins = logits.shape[-1]
    edges = torch.linspace(0, 1, bins + 1, dtype=torch.float32)
    centres = (edges[:-1] + edges[1:]) / 2
    return (logits.float().softmax(-1) @ centres.unsqueeze(1)).squeeze(-1)


class StructureTokenDecoderRef(nn.Module):
    def __init__(self, dims: DecoderDimsRef | None = None):
        super().__init__()
        self.dims = dims or DecoderDimsRef()
        d = self.dims.d_model
        self.embed = nn.Embedding(self.dims.struct_vocab, d)
        self.decoder_stack = nn.Module()
        self.decoder_stack.blocks = nn.ModuleList([esm3_ref.BlockRef(self.dims, False) for _ in range(self.dims.n_layers)])
        self.decoder_stack.norm = nn.LayerNorm(d, bias=False)
        self.affine_output_projection = Dim6RotStructureHeadRef(d)
        self.plddt_head = esm3_ref.regression_head(d, self.dims.plddt_bins) if self.dims.plddt_bins else None

    def trunk(self, structure_tokens):
        x = self.embed(structure_tokens)
        for blk in self.decoder_stack.blocks:
            x = blk(x)
        return self.decoder_stack.norm(x)

    @torch.no_grad()
    def decode(self, structure_tokens):
        """(B, T) with BOS/EOS.  bb_pred (B,T,3,3), oxygen (B,T,3) (NaN at BOS, EOS and the last
        residue), plddt (B,T), affine (B,T,23)."""
        assert bool((structure_tokens[:, 0] == esm3_ref.STRUCTURE_BOS).all())
        assert bool((structure_tokens[:, -1] == esm3_ref.STRUCTURE_EOS).all())
        x = self.trunk(structure_tokens)
        affine, bb = self.affine_output_projection(x)
        o = torch.full(bb.shape[:-2] + (3,), float("nan"))
        o[:, 1:-1] = infer_oxygen(bb[:, 1:-1])
        plddt = plddt_mean(self.plddt_head(x)) if self.plddt_head is not None else None
        return {"bb_pred": bb, "oxygen": o, "plddt": plddt, "affine": affine}


def build_decoder(dims: DecoderDimsRef | None = None, seed: int = 0) -> StructureTokenDecoderRef:
    state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    dec = StructureTokenDecoderRef(dims).eval()
    torch.random.set_rng_state(state)
    return dec


def build_decoder_from_state_dict(dims: DecoderDimsRef, sd: dict) -> StructureTokenDecoderRef:
    with torch.device("meta"):
        dec = StructureTokenDecoderRef(dims)
    keep = {k: v.detach().float().cpu() for k, v in sd.items() if not k.startswith("pairwise_classification_head.")}
    dec.load_state_dict(keep, strict=True, assign=True)
    return dec.eval()


@torch.no_grad()
def decode_emulated(dec: StructureTokenDecoderRef, structure_tokens, rd=None):
    """The same decode with a bf16 rounding at the product's rounding points (oracle/esm3_emul.py)."""
    from . import esm3_emul as E
    B, T = structure_tokens.shape
    rd = E.Rounding.product(T) if rd is None else rd
    x = dec.embed(structure_tokens)
    cos, sin = esm3_ref.rotary_tables(T, dec.dims.d_head)
    for blk in dec.decoder_stack.blocks:
        x = E.block_forward(blk, x, cos, sin, rd, dec.dims.n_heads)
    d = dec.dims.d_model
    xn = E._r(F.layer_norm(x, (d,), dec.decoder_stack.norm.weight, None, 1e-5), rd.act)

    def head(lin0, norm, lin1):
        h = F.gelu(E._mm(xn, E._r(lin0.weight, rd.weights), rd.f64) + lin0.bias)
        h = E._r(F.layer_norm(h, (d,), norm.weight, norm.bias, 1e-5), rd.act)
        return E._mm(h, E._r(lin1.weight, rd.weights), rd.f64) + lin1.bias

    a = dec.affine_output_projection
    affine = head(a.ffn1, a.norm, a.proj)
    bb = frames_to_backbone(affine)
    o = torch.full(bb.shape[:-2] + (3,), float("nan"))
    o[:, 1:-1] = infer_oxygen(bb[:, 1:-1])
    plddt = None
    if dec.plddt_head is not None:
        plddt = plddt_mean(head(dec.plddt_head[0], dec.plddt_head[2], dec.plddt_head[3]))
    return {"bb_pred": bb, "oxygen": o, "plddt": plddt, "affine": affine}
