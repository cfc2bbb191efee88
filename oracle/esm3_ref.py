"""fp32 CPU restatement of the network half of the ESMDiff ddpm path (ORACLE -- tests only).

What this follows
-----------------
* ``CustomizedESM3.forward`` wrapper semantics -- default tracks, NaN coordinates, BOS/EOS/
  PAD/CHAINBREAK forcing of structure ids, ``+ auxiliary_embeddings``, call order encoder ->
  transformer -> output heads: reference slm/models/net.py:371-483 (esp. :410-469).
* ``StructureOutputHeads``: reference slm/models/net.py:298-320 (only ``structure_head`` is
  live on this path: ``n_sequence_heads: 0`` in configs/experiment/mdlm.yaml:56-58).
* ``TimestepEmbedder``: reference slm/models/net.py:486-522.
* The layers those call live in the third-party package ``esm==3.0.4``
  (reference requirements.txt:30), which is NOT vendored in /root/reference and not
  installed.  They are restated here from the published esm 3.0.4 algorithm:
  ``EncodeInputs`` (esm/models/esm3.py), ``TransformerStack`` (esm/layers/transformer_stack.py),
  ``UnifiedTransformerBlock`` / ``MultiHeadAttention`` / ``swiglu_ln_ffn``
  (esm/layers/blocks.py, attention.py), ``RotaryEmbedding`` (esm/layers/rotary.py),
  ``RegressionHead`` (esm/layers/regression_head.py).  **PARITY UNPINNED** for this half:
  the reference tree holds no test, golden activation or runnable copy of these layers.
  What *is* pinned: constructor arguments at the call sites (net.py:337-346, :301), the
  2-tuple return of the stack (net.py:468), state-dict key names/shapes and
  the tokenizer ids (data/dummy_train_data/*.pth, see tests/golden/tokenizer_pins.json).

Parameter names equal the esm 3.0.4 names so a real ``release_v0.pt`` ``['module']`` dict
(prefix ``net.``) loads unchanged -- the same key ABI the CUDA library consumes.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F
from torch import nn

from . import geom_ref

# esm.utils.constants.esm3 values used on the path (tokenizer pins in tests/golden)
SEQUENCE_BOS, SEQUENCE_PAD, SEQUENCE_EOS, SEQUENCE_CHAINBREAK, SEQUENCE_MASK = 0, 1, 2, 31, 32
VQVAE_CODEBOOK_SIZE = 4096
STRUCTURE_MASK, STRUCTURE_EOS, STRUCTURE_BOS, STRUCTURE_PAD, STRUCTURE_CHAINBREAK = (
    4096, 4097, 4098, 4099, 4100)
SS8_PAD = SASA_PAD = RESIDUE_PAD = INTERPRO_PAD = 0


@dataclass
class Esm3Dims:
    d_model: int = 1536
    n_heads: int = 24
    v_heads: int = 256
    n_layers: int = 48
    n_structure_heads: int = 4101   # configs/experiment/mdlm.yaml:57
    seq_vocab: int = 64
    struct_vocab: int = 4101        # 4096 codes + 5 specials
    d_head: int = 64

    @property
    def ffn_hidden(self) -> int:
        # esm swiglu_correction_fn(8/3, d): round up to a multiple of 256
        return int(((8.0 / 3.0 * self.d_model) + 255) // 256 * 256)

    @property
    def residue_scale(self) -> float:
        # esm TransformerStack: residue_scaling_factor = sqrt(n_layers / 36)
        return math.sqrt(self.n_layers / 36)


def rbf16(values: torch.Tensor) -> torch.Tensor:
    """esm.utils.misc.rbf(values, 0, 1, n_bins=16)."""
    centers = torch.linspace(0.0, 1.0, 16, dtype=values.dtype)
    z = (values.unsqueeze(-1) - centers) / (1.0 / 16)
    return torch.exp(-(z ** 2))


class EncodeInputsRef(nn.Module):
    """Sum of the eight ESM3 input-track embeddings."""

    def __init__(self, d: int, seq_vocab: int = 64, struct_vocab: int = 4101):
        super().__init__()
        self.sequence_embed = nn.Embedding(seq_vocab, d)
        self.plddt_projection = nn.Linear(16, d)
        self.structure_per_res_plddt_projection = nn.Linear(16, d)
        self.structure_tokens_embed = nn.Embedding(struct_vocab, d)
        self.ss8_embed = nn.Embedding(8 + 3, d)
        self.sasa_embed = nn.Embedding(16 + 3, d)
        self.function_embed = nn.ModuleList(
            [nn.Embedding(260, d // 8, padding_idx=0) for _ in range(8)])
        self.residue_embed = nn.EmbeddingBag(1478, d, mode="sum", padding_idx=0)

    def default_track_vector(self) -> torch.Tensor:
        """Contribution of the six tracks the ddpm path leaves at their defaults
        (net.py:413-431): average_plddt=1, per_res_plddt=0, ss8=0, sasa=0, function=0 (padding
        row -> zeros), residue annotations=0 (padding -> empty bag -> zeros).  The same (d,)
        vector for every position."""
        one = torch.ones(1, dtype=torch.float32)
        zero = torch.zeros(1, dtype=torch.float32)
        v = self.plddt_projection(rbf16(one))[0]
        v = v + self.structure_per_res_plddt_projection(rbf16(zero))[0]
        return v + self.ss8_embed.weight[SS8_PAD] + self.sasa_embed.weight[SASA_PAD]

    def forward(self, sequence_tokens, structure_tokens):
        B, T = structure_tokens.shape
        seq = self.sequence_embed(sequence_tokens)
        # esm order: seq + plddt + per_res_plddt + structure + ss8 + sasa + function + residue
        x = seq + self.plddt_projection(rbf16(torch.ones(1, T)))
        x = x + self.structure_per_res_plddt_projection(rbf16(torch.zeros(1, T)))
        x = x + self.structure_tokens_embed(structure_tokens)
        x = x + self.ss8_embed(torch.full((1, T), SS8_PAD, dtype=torch.long))
        x = x + self.sasa_embed(torch.full((1, T), SASA_PAD, dtype=torch.long))
        func = torch.cat([emb(torch.full((1, T), INTERPRO_PAD, dtype=torch.long))
                          for emb in self.function_embed], -1)
        x = x + func
        res = self.residue_embed(torch.full((T, 16), RESIDUE_PAD, dtype=torch.long))
        x = x + res.view(1, T, -1)
        return x


def rotary_tables(T: int, d_head: int = 64, base: float = 10000.0):
    inv_freq = 1.0 / (base ** (torch.arange(0, d_head, 2, dtype=torch.float32) / d_head))
    t = torch.arange(T, dtype=torch.float32)
    freqs = torch.outer(t, inv_freq)            # (T, d_head/2)
    return torch.cos(freqs), torch.sin(freqs)


def apply_rotary(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """x: (B, T, H, dh); non-interleaved rotate-half (esm/layers/rotary.py)."""
    half = x.shape[-1] // 2
    x1, x2 = x[..., :half], x[..., half:]
    c = cos[None, :, None, :]
    s = sin[None, :, None, :]
    return torch.cat([x1 * c - x2 * s, x2 * c + x1 * s], dim=-1)


class MultiHeadAttentionRef(nn.Module):
    def __init__(self, d: int, n_heads: int):
        super().__init__()
        self.d, self.h = d, n_heads
        self.layernorm_qkv = nn.Sequential(nn.LayerNorm(d), nn.Linear(d, 3 * d, bias=False))
        self.out_proj = nn.Linear(d, d, bias=False)
        self.q_ln = nn.LayerNorm(d, bias=False)     # over the full width, not per head
        self.k_ln = nn.LayerNorm(d, bias=False)

    def forward(self, x):
        B, T, _ = x.shape
        dh = self.d // self.h
        q, k, v = self.layernorm_qkv(x).chunk(3, dim=-1)
        q, k = self.q_ln(q), self.k_ln(k)
        cos, sin = rotary_tables(T, dh)
        q = apply_rotary(q.view(B, T, self.h, dh), cos, sin)
        k = apply_rotary(k.view(B, T, self.h, dh), cos, sin)
        v = v.view(B, T, self.h, dh)
        ctx = F.scaled_dot_product_attention(
            q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))   # no mask: sequence_id None
        return self.out_proj(ctx.transpose(1, 2).reshape(B, T, self.d))


class GeomAttnParamsRef(geom_ref.GeometricReasoningRef):
    """Block 0's geometric attention (oracle/geom_ref.py).

    On the ddpm path coordinates are NaN (net.py:433-441) so ``affine_mask`` is all False and,
    with ``mask_and_zero_frameless=True`` (net.py:339-345), esm zero-fills the attention
    output before a bias-free ``out_proj``: the branch contributes exactly 0.
    With ``frames = (rot, trans, mask)`` of real coordinates it is live.
    """

    def __init__(self, d: int, v_heads: int):
        super().__init__(d, v_heads, mask_and_zero_frameless=True)

    def forward(self, x, frames=None):
        if frames is None:
            zero_ctx = torch.zeros(*x.shape[:-1], self.out_proj.in_features, dtype=x.dtype)
            return self.out_proj(zero_ctx)      # == 0 exactly
        return super().forward(x, *frames)


class SwiGLURef(nn.Module):
    def forward(self, x):
        a, b = x.chunk(2, dim=-1)
        return F.silu(a) * b


class BlockRef(nn.Module):
    def __init__(self, dims: Esm3Dims, with_geom: bool):
        super().__init__()
        d = dims.d_model
        self.attn = MultiHeadAttentionRef(d, dims.n_heads)
        if with_geom:
            self.geom_attn = GeomAttnParamsRef(d, dims.v_heads)
        self.with_geom = with_geom
        self.ffn = nn.Sequential(nn.LayerNorm(d), nn.Linear(d, 2 * dims.ffn_hidden, bias=False),
                                 SwiGLURef(), nn.Linear(dims.ffn_hidden, d, bias=False))
        self.scale = dims.residue_scale

    def forward(self, x, frames=None):
        x = x + self.attn(x) / self.scale
        if self.with_geom:
            x = x + self.geom_attn(x, frames) / self.scale
        x = x + self.ffn(x) / self.scale
        return x


class TransformerStackRef(nn.Module):
    def __init__(self, dims: Esm3Dims):
        super().__init__()
        self.blocks = nn.ModuleList([BlockRef(dims, i < 1) for i in range(dims.n_layers)])
        self.norm = nn.LayerNorm(dims.d_model, bias=False)

    def forward(self, x, frames=None):
        for blk in self.blocks:
            x = blk(x, frames) if blk.with_geom else blk(x)
        return self.norm(x), x


def regression_head(d: int, out: int) -> nn.Sequential:
    return nn.Sequential(nn.Linear(d, d), nn.GELU(), nn.LayerNorm(d), nn.Linear(d, out))


class OutputHeadsRef(nn.Module):
    def __init__(self, d: int, n_structure_heads: int):
        super().__init__()
        self.structure_head = regression_head(d, n_structure_heads)
        self.sequence_head = None


@dataclass
class NetOutput:
    structure_logits: torch.Tensor
    embeddings: torch.Tensor
    sequence_logits: torch.Tensor | None = None


class CustomizedESM3Ref(nn.Module):
    """Oracle for ``CustomizedESM3`` restricted to the tracks the ddpm path feeds."""

    def __init__(self, dims: Esm3Dims | None = None):
        super().__init__()
        self.dims = dims or Esm3Dims()
        d = self.dims.d_model
        self.encoder = EncodeInputsRef(d, self.dims.seq_vocab, self.dims.struct_vocab)
        self.transformer = TransformerStackRef(self.dims)
        self.output_heads = OutputHeadsRef(d, self.dims.n_structure_heads)

    @staticmethod
    def force_special_structure_ids(structure_tokens, sequence_tokens):
        """net.py:445-454."""
        st = structure_tokens.masked_fill(structure_tokens == -1, STRUCTURE_MASK)
        st = st.masked_fill(sequence_tokens == SEQUENCE_BOS, STRUCTURE_BOS)
        st = st.masked_fill(sequence_tokens == SEQUENCE_PAD, STRUCTURE_PAD)
        st = st.masked_fill(sequence_tokens == SEQUENCE_EOS, STRUCTURE_EOS)
        st = st.masked_fill(sequence_tokens == SEQUENCE_CHAINBREAK, STRUCTURE_CHAINBREAK)
        return st

    @torch.no_grad()
    def forward(self, structure_tokens, labels=None, mask=None, sequence_tokens=None, *,
                auxiliary_embeddings=None, structure_coords=None, **unused):
        assert labels is None, "oracle covers the inference branch only (net.py:483)"
        frames = None
        if structure_coords is not None:                      # net.py:437-441
            B, T = structure_tokens.shape
            frames = geom_ref.build_affine3d_from_coordinates(
                structure_coords[..., :3, :].expand(B, T, 3, 3))
        if sequence_tokens is None:
            sequence_tokens = torch.full_like(structure_tokens, SEQUENCE_MASK)
        st = self.force_special_structure_ids(structure_tokens, sequence_tokens)
        x = self.encoder(sequence_tokens, st)
        if auxiliary_embeddings is not None:
            x = x + auxiliary_embeddings
        xn, emb = self.transformer(x, frames)
        return NetOutput(structure_logits=self.output_heads.structure_head(xn), embeddings=emb)


class TimestepEmbedderRef(nn.Module):
    """net.py:486-522: sinusoid(256) -> Linear -> SiLU -> Linear."""

    def __init__(self, hidden_size: int, frequency_embedding_size: int = 256):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(frequency_embedding_size, hidden_size),
                                 nn.SiLU(), nn.Linear(hidden_size, hidden_size))
        self.nfreq = frequency_embedding_size

    @staticmethod
    def features(t: torch.Tensor, dim: int, max_period: float = 10000.0) -> torch.Tensor:
        half = dim // 2
        freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=t.dtype) / half)
        freqs = freqs.to(device=t.device)            # net.py:507-511 does the same
        args = t[:, None] * freqs[None]
        out = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
        if dim % 2:
            out = torch.cat([out, torch.zeros_like(out[:, :1])], dim=-1)
        return out

    def forward(self, t):
        return self.mlp(self.features(t, self.nfreq))


def build_reference_model(dims: Esm3Dims | None = None, seed: int = 0):
    """Deterministic random-init weights: module default init under
    ``torch.manual_seed(seed)`` on CPU fp32.  Returns (net, sigma_embedder)."""
    gen_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    net = CustomizedESM3Ref(dims).eval()
    emb = TimestepEmbedderRef((dims or Esm3Dims()).d_model).eval()
    torch.random.set_rng_state(gen_state)
    return net, emb


def build_from_state_dict(dims: Esm3Dims, sd: dict):
    """Oracle modules holding the tensors of a DeepSpeed-style ``['module']`` dict (keys ``net.*``
    and ``sigma_embedder.*``) without running the default initialisers (meta device + assign)."""
    with torch.device("meta"):
        net = CustomizedESM3Ref(dims)
        emb = TimestepEmbedderRef(dims.d_model)
    net.load_state_dict({k[4:]: v.detach().float().cpu() for k, v in sd.items() if k.startswith("net.")},
                        strict=True, assign=True)
    emb.load_state_dict({k[15:]: v.detach().float().cpu() for k, v in sd.items()
                         if k.startswith("sigma_embedder.")}, strict=True, assign=True)
    return net.eval(), emb.eval()


def full_state_dict(net: nn.Module, emb: nn.Module) -> dict:
    """Keys as in a DeepSpeed ``['module']`` dict (checkpoint_utils.py:62-64)."""
    sd = {f"net.{k}": v for k, v in net.state_dict().items()}
    sd.update({f"sigma_embedder.{k}": v for k, v in emb.state_dict().items()})
    return sd


def forward_flops(B: int, T: int, dims: Esm3Dims | None = None) -> int:
    """Algorithmic FLOPs of one forward: GEMMs + attention matmuls."""
    dm = dims or Esm3Dims()
    d, f, v = dm.d_model, dm.ffn_hidden, dm.n_structure_heads
    per_tok_layer = 2 * d * 3 * d + 2 * d * d + 2 * d * 2 * f + 2 * f * d + 4 * T * d
    head = 2 * d * d + 2 * d * v
    return B * T * (dm.n_layers * per_tok_layer + head)
