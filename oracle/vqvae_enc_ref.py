"""TEST INFRASTRUCTURE -- fp32 CPU restatement of the VQ-VAE structure ENCODER front end.

What this follows
-----------------
* reference call sites: ``protseq_to_data`` slm/models/utils.py:105-146 (``mask_ids``: sequence -> '_',
  ``coordinates[idx] = inf``; ``model.encode(ESMProtein(sequence, coordinates))`` -> ``structure_tokens``),
  ``pdb_to_data`` :99-102, its consumers slm/sample_esmdiff.py:166-175, 197-209 (inpainting prior) and
  :278-284 (``ESMProtein.from_pdb(p)`` -> ``prot.coordinates``).
* the arithmetic lives in ``esm==3.0.4`` (requirements.txt:30), absent from /root/reference and not installed;
  restated from the published package:
    ``ESM3.encode`` -> ``tokenize_structure`` (esm/utils/encoding.py): ``ProteinChain.from_atom37`` (residue_index
      1..L), ``to_structure_encoder_inputs``, ``StructureTokenEncoder.encode``, BOS / EOS around the codes;
    ``StructureTokenEncoder`` (esm/models/vqvae.py; ESM3_structure_encoder_v0 = d_model 1024, n_heads 1,
      v_heads 128, n_layers 2, d_out 128, n_codes 4096): ``find_knn_edges`` / ``knn_graph`` (16 nearest by CA
      distance, sequence distance for frameless pairs), ``RelativePositionEmbedding(32, d)``,
      ``TransformerStack(d, n_heads, v_heads, n_layers, n_layers_geom=n_layers, use_plain_attn=False)`` (geometric
      attention + SwiGLU FFN per block, residue scale sqrt(n_layers / 36), final LayerNorm), the query node =
      neighbour 0, ``pre_vq_proj``, ``EMACodebook`` nearest code;
    ``GeometricReasoningOriginalImpl`` and the frames: oracle/geom_ref.py.
  **PARITY UNPINNED**: no test, fixture or runnable copy of these layers exists in the reference tree, and the
  pretrained encoder weights (``esm3_structure_encoder_v0.pth``) are not available offline.  Pinned by the
  reference: the call order above, BOS / EOS = 4098 / 4097 and codes < 4096 (data/dummy_train_data/*.pth).
Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline may import this module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F
from torch import nn

from . import geom_ref

STRUCTURE_BOS, STRUCTURE_EOS = 4098, 4097


@dataclass
class EncoderDimsRef:
    d_model: int = 1024
    v_heads: int = 128
    n_layers: int = 2
    d_out: int = 128
    n_codes: int = 4096
    knn: int = 16
    rel_bins: int = 32

    @property
    def ffn_hidden(self) -> int:
        return int(((8.0 / 3.0 * self.d_model) + 255) // 256 * 256)

    @property
    def residue_scale(self) -> float:
        return math.sqrt(self.n_layers / 36)


def knn_graph(ca, coord_mask, knn: int):
    """esm ``knn_graph`` for one chain without padding: ca (B, L, 3), coord_mask (B, L) ->
    edges (B, L, min(knn, L)) int64: neighbours by CA distance; pairs with a frameless member sort after
    every structural pair, by sequence distance (100 |i - j| + 1e6).  Ties -> lower index first (torch.sort is
    not stable; ties only occur among frameless pairs, which the attention masks out)."""
    B, L, _ = ca.shape
    E = min(knn, L)
    ca = ca.nan_to_num()
    pair_invalid = ~(coord_mask[:, None, :] & coord_mask[:, :, None])
    dists = (ca[:, :, None, :] - ca[:, None, :, :]).norm(dim=-1)
    ar = torch.arange(L)
    seq_d = (ar[:, None] - ar[None, :]).abs().to(dists.dtype) * 1e2 + geom_ref.MAX_SUPPORTED_DISTANCE
    d = torch.where(pair_invalid, seq_d[None].expand(B, L, L), dists)
    _, edges = torch.sort(d, dim=-1, stable=True)
    return edges[..., :E]


class EncoderBlockRef(nn.Module):
    """esm ``UnifiedTransformerBlock(use_geom_attn=True, use_plain_attn=False, ffn swiglu, bias=False)``."""

    def __init__(self, dims: EncoderDimsRef):
        super().__init__()
        d = dims.d_model
        self.geom_attn = geom_ref.GeometricReasoningRef(d, dims.v_heads, mask_and_zero_frameless=False)
        self.ffn = nn.Sequential(nn.LayerNorm(d), nn.Linear(d, 2 * dims.ffn_hidden, bias=False), nn.Identity(),
                                 nn.Linear(dims.ffn_hidden, d, bias=False))
        self.scale = dims.residue_scale

    def forward(self, x, rot, trans, mask):
        x = x + self.geom_attn(x, rot, trans, mask) / self.scale
        a, b = self.ffn[1](self.ffn[0](x)).chunk(2, dim=-1)
        return x + self.ffn[3](F.silu(a) * b) / self.scale


class StructureTokenEncoderRef(nn.Module):
    def __init__(self, dims: EncoderDimsRef | None = None):
        super().__init__()
        self.dims = dims or EncoderDimsRef()
        d = self.dims.d_model
        self.transformer = nn.Module()
        self.transformer.blocks = nn.ModuleList([EncoderBlockRef(self.dims) for _ in range(self.dims.n_layers)])
        self.transformer.norm = nn.LayerNorm(d, bias=False)
        self.pre_vq_proj = nn.Linear(d, self.dims.d_out)
        self.codebook = nn.Module()
        self.codebook.register_buffer("embeddings", torch.randn(self.dims.n_codes, self.dims.d_out))
        self.relative_positional_embedding = nn.Module()
        self.relative_positional_embedding.embedding = nn.Embedding(2 * self.dims.rel_bins + 2, d)

    @torch.no_grad()
    def encode(self, coords: torch.Tensor, residue_index: torch.Tensor | None = None, return_all: bool = False):
        """coords (B, L, >=3, 3) (N, CA, C first; NaN / inf = unknown) -> (z_q (B, L, d_out), codes (B, L) int64)."""
        dims = self.dims
        coords = coords[..., :3, :].float()
        rot, trans, mask = geom_ref.build_affine3d_from_coordinates(coords)
        B, L = mask.shape
        c0 = coords.clone()
        c0[~mask] = 0
        edges = knn_graph(c0[..., 1, :], mask, dims.knn)                          # (B, L, E)
        E = edges.shape[-1]
        bidx = torch.arange(B)[:, None, None]
        k_rot = rot[bidx, edges].reshape(B * L, E, 3, 3)
        k_trans = trans[bidx, edges].reshape(B * L, E, 3)
        k_mask = mask[bidx, edges].reshape(B * L, E)
        res = edges if residue_index is None else residue_index[bidx, edges]
        res = res.reshape(B * L, E)
        diff = (res - res[:, :1]).clamp(-dims.rel_bins, dims.rel_bins) + dims.rel_bins + 1
        z = self.relative_positional_embedding.embedding(diff)                    # (B L, E, d)
        for blk in self.transformer.blocks:
            z = blk(z, k_rot, k_trans, k_mask)
        z = self.transformer.norm(z)
        z = z.view(B, L, E, -1)[:, :, 0, :]
        z = z.masked_fill(~mask[..., None], 0)
        z = self.pre_vq_proj(z)
        e = self.codebook.embeddings
        zf = z.reshape(-1, dims.d_out)
        d = zf.pow(2).sum(1, keepdim=True) + e.pow(2).sum(1) - 2 * zf @ e.t()
        codes = d.argmin(dim=1).view(B, L)
        if return_all:
            return {"z": z, "codes": codes, "edges": edges, "rot": rot, "trans": trans, "mask": mask, "dist": d.view(B, L, -1)}
        return e[codes], codes


def tokenize_structure(enc: StructureTokenEncoderRef, coordinates: torch.Tensor):
    """esm ``tokenize_structure``: atom37 / atom3 coordinates (L, A, 3) of one chain -> int64 (L + 2,) with BOS / EOS."""
    _, codes = enc.encode(normalize_coordinates(coordinates[None, :, :3, :]),
                          residue_index=torch.arange(1, coordinates.shape[0] + 1)[None])
    out = torch.full((coordinates.shape[0] + 2,), STRUCTURE_BOS, dtype=torch.int64)
    out[1:-1] = codes[0]
    out[-1] = STRUCTURE_EOS
    return out


def normalize_coordinates(coords: torch.Tensor) -> torch.Tensor:
    """esm ``normalize_coordinates`` (to_structure_encoder_inputs): express the chain in the frame of its average
    backbone.  The encoder is SE(3)-invariant, so this only conditions the fp32 arithmetic."""
    bb = coords[..., :3, :]
    mask = torch.isfinite(bb).all(-1).all(-1)
    avg = bb.masked_fill(~mask[..., None, None], 0).sum(-3) / (mask.sum(-1)[..., None, None] + 1e-8)
    rot, trans = geom_ref.backbone_frames(avg)                                     # (B,3,3), (B,3)
    return torch.einsum("bji,blaj->blai", rot, coords - trans[:, None, None, :])


def build_encoder(dims: EncoderDimsRef | None = None, seed: int = 0) -> StructureTokenEncoderRef:
    """Deterministic random-init weights; the geometric attention's per-head scales (zeros in esm's initialiser:
    softplus(0) = 0.69 everywhere) and the positional table (esm: std 0.02) are drawn wider so that every term of
    the attention matters in the parity tests."""
    state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    enc = StructureTokenEncoderRef(dims).eval()
    with torch.no_grad():
        for blk in enc.transformer.blocks:
            blk.geom_attn.distance_scale_per_head.normal_(0.0, 1.0)
            blk.geom_attn.rotation_scale_per_head.normal_(0.0, 1.0)
    torch.random.set_rng_state(state)
    return enc


def build_encoder_from_state_dict(dims: EncoderDimsRef, sd: dict) -> StructureTokenEncoderRef:
    with torch.device("meta"):
        enc = StructureTokenEncoderRef(dims)
    keep = {k: v.detach().float().cpu() for k, v in sd.items()
            if not (k.startswith("codebook.") and k != "codebook.embeddings")}
    enc.load_state_dict(keep, strict=True, assign=True)
    return enc.eval()


def synthetic_backbone(L: int, seed: int = 0) -> torch.Tensor:
    """A random self-avoiding-ish CA walk (3.8 A steps) with ideal-ish N / C placed around every CA: (L, 3, 3)."""
    g = torch.Generator().manual_seed(seed)
    steps = F.normalize(torch.randn(L, 3, generator=g), dim=-1)
    for i in range(1, L):                                   # persistence: helices / strands rather than a coil
        steps[i] = F.normalize(0.6 * steps[i - 1] + 0.8 * steps[i], dim=-1)
    ca = torch.cumsum(3.8 * steps, dim=0)
    side = F.normalize(torch.randn(L, 3, generator=g), dim=-1)
    n = ca - 1.46 * F.normalize(steps + 0.5 * side, dim=-1)
    c = ca + 1.52 * F.normalize(torch.roll(steps, -1, 0) - 0.5 * side, dim=-1)
    return torch.stack([n, ca, c], dim=1)
