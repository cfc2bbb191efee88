"""TEST INFRASTRUCTURE -- fp32 CPU restatement of esm's backbone frames and geometric attention.

What this follows
-----------------
* reference call sites: ``build_affine3d_from_coordinates(structure_coords)`` slm/models/net.py:437-441
  (NaN coordinates when none are passed, :433-436), ``TransformerStack(d_model, n_heads, v_heads, n_layers,
  mask_and_zero_frameless=True)`` net.py:337-345 (block 0 carries ``geom_attn``), the stack call
  ``self.transformer(x, sequence_id, affine, affine_mask, chain_id)`` net.py:468.
* the arithmetic lives in ``esm==3.0.4`` (requirements.txt:30), absent from /root/reference and not installed;
  restated from the published package: ``build_affine3d_from_coordinates`` / ``Affine3D.from_graham_schmidt`` /
  ``_graham_schmidt`` (esm/utils/structure/affine3d.py), ``GeometricReasoningOriginalImpl``
  (esm/layers/geom_attention.py).  **PARITY UNPINNED**: no test, fixture or runnable copy of these layers exists
  in the reference tree.  Pinned by the reference: the parameter names and shapes of ``geom_attn.*``
  (``s_norm.weight``, ``proj.weight`` (15 v_heads x d), ``out_proj.weight`` (d x 3 v_heads),
  ``distance_scale_per_head``, ``rotation_scale_per_head``) and the call-site arguments above.
Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline may import this module.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn

MAX_SUPPORTED_DISTANCE = 1e6


def graham_schmidt(x_axis, xy_plane, eps=1e-12):
    """esm ``_graham_schmidt``: rotation matrices with columns [e0, e1, e2]."""
    e1 = xy_plane
    x_axis = x_axis / torch.sqrt((x_axis ** 2).sum(-1, keepdim=True) + eps)
    dot = (x_axis * e1).sum(-1, keepdim=True)
    e1 = e1 - x_axis * dot
    e1 = e1 / torch.sqrt((e1 ** 2).sum(-1, keepdim=True) + eps)
    e2 = torch.cross(x_axis, e1, dim=-1)
    return torch.stack([x_axis, e1, e2], dim=-1)


def backbone_frames(bb):
    """``Affine3D.from_graham_schmidt(C, CA, N)``: (rot (...,3,3), trans (...,3)) of N, CA, C positions."""
    n, ca, c = bb.unbind(-2)
    return graham_schmidt(ca - c, n - ca, 1e-12), ca


def build_affine3d_from_coordinates(coords: torch.Tensor):
    """coords (B, L, 3, 3) = N, CA, C (NaN / inf where unknown) -> rot (B,L,3,3), trans (B,L,3), mask (B,L).
    Residues without a frame get the frame of the average backbone of the valid ones ("black hole"),
    the identity rotation when the sample has no valid residue at all."""
    coord_mask = (torch.isfinite(coords) & (coords < MAX_SUPPORTED_DISTANCE)).all(-1).all(-1)
    coords = coords.clone().float()
    coords[~coord_mask] = 0
    avg = coords.sum(1) / (coord_mask.sum(-1)[..., None, None] + 1e-8)          # (B, 3, 3)
    avg_rot, avg_trans = backbone_frames(avg)
    B, L = coord_mask.shape
    has_any = coord_mask.any(-1)[:, None, None]
    avg_rot = torch.where(has_any, avg_rot, torch.eye(3).expand(B, 3, 3))
    rot, trans = backbone_frames(coords)
    rot = torch.where(coord_mask[..., None, None], rot, avg_rot[:, None].expand(B, L, 3, 3))
    trans = torch.where(coord_mask[..., None], trans, avg_trans[:, None].expand(B, L, 3))
    return rot, trans, coord_mask


class GeometricReasoningRef(nn.Module):
    """esm ``GeometricReasoningOriginalImpl(c_s, v_heads, num_vector_messages=1, mask_and_zero_frameless, bias=False)``."""

    def __init__(self, c_s: int, v_heads: int, mask_and_zero_frameless: bool = True):
        super().__init__()
        self.v_heads = v_heads
        self.mask_and_zero_frameless = mask_and_zero_frameless
        self.s_norm = nn.LayerNorm(c_s, bias=False)
        self.proj = nn.Linear(c_s, 15 * v_heads, bias=False)      # 2 x (q, k) x 3 + v x 3 per head
        self.out_proj = nn.Linear(3 * v_heads, c_s, bias=False)
        self.distance_scale_per_head = nn.Parameter(torch.zeros(v_heads))
        self.rotation_scale_per_head = nn.Parameter(torch.zeros(v_heads))

    def attention(self, p, rot, trans, affine_mask, sequence_id=None, chain_id=None):
        """p = proj(s_norm(s)) (B, S, 15 H) -> rotated-back messages (B, S, 3 H), before out_proj."""
        B, S, _ = p.shape
        H = self.v_heads
        vec_rot, vec_dist = p.split([9 * H, 6 * H], dim=-1)
        vec_rot = torch.einsum("bsij,bshj->bshi", rot, vec_rot.view(B, S, 3 * H, 3))
        q_rot, k_rot, value = vec_rot.split([H, H, H], dim=2)
        vec_dist = torch.einsum("bsij,bshj->bshi", rot, vec_dist.view(B, S, 2 * H, 3)) + trans[:, :, None, :]
        q_dist, k_dist = vec_dist.chunk(2, dim=2)
        # (B, H, Sq, Sk)
        distance_term = (q_dist.permute(0, 2, 1, 3)[:, :, :, None, :] - k_dist.permute(0, 2, 1, 3)[:, :, None, :, :]) \
            .norm(dim=-1) / math.sqrt(3)
        rotation_term = q_rot.permute(0, 2, 1, 3) @ k_rot.permute(0, 2, 3, 1) / math.sqrt(3)
        w_d = F.softplus(self.distance_scale_per_head)[:, None, None]
        w_r = F.softplus(self.rotation_scale_per_head)[:, None, None]
        attn = rotation_term * w_r - distance_term * w_d
        # the bias is the float of the same-sequence mask (+1 for pairs of one sequence, a softmax no-op),
        # finfo.min for frameless keys and for pairs of different chains
        if sequence_id is None:
            sequence_id = torch.zeros(B, S, dtype=torch.int64)
        bias = (sequence_id[:, :, None] == sequence_id[:, None, :])[:, None].float()
        bias = bias.masked_fill(~affine_mask[:, None, None, :], torch.finfo(bias.dtype).min)
        if chain_id is not None:
            bias = bias.masked_fill((chain_id[:, :, None] != chain_id[:, None, :])[:, None], torch.finfo(bias.dtype).min)
        attn = torch.softmax(attn + bias, dim=-1)
        out = attn @ value.permute(0, 2, 1, 3)                                   # (B, H, S, 3), global frame
        out = torch.einsum("bsji,bshj->bshi", rot, out.permute(0, 2, 1, 3))      # rot^T: back to the local frame
        out = out.reshape(B, S, 3 * H)
        if self.mask_and_zero_frameless:
            out = out.masked_fill(~affine_mask[..., None], 0.0)
        return out

    def forward(self, s, rot, trans, affine_mask, sequence_id=None, chain_id=None):
        return self.out_proj(self.attention(self.proj(self.s_norm(s)), rot, trans, affine_mask, sequence_id, chain_id))
