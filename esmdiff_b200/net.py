"""Host-side mirror of the reference's network interface for the ddpm path.

``CustomizedESM3`` here has the constructor arguments, ``forward`` signature, state-dict key names
and output object of the reference's ``slm/models/net.py:322-483`` -- but holds no torch
parameters: weights live inside the CUDA library (bf16 GEMM operands, fp32 norms/embeddings) and
``forward`` is one call into ``libesmdiff_b200.so``.  It can therefore be driven unchanged by the
reference's own ``MaskedDiffusionLanguageModeling._model_wrapper`` (model.py:475-480).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
from torch import nn

from .engine import Dims, Engine

ESM3_D_MODEL = 1536


@dataclass
class ESMOutput:
    """Fields of esm.models.esm3.ESMOutput that the path touches (net.py:309-318).  The
    reference aliases one zeros_like(structure_logits) into the five unused logit fields (a wasted
    267 MB memset per forward at B=63,T=258); here they are None."""
    sequence_logits: torch.Tensor | None
    structure_logits: torch.Tensor
    secondary_structure_logits: torch.Tensor | None = None
    sasa_logits: torch.Tensor | None = None
    function_logits: torch.Tensor | None = None
    residue_logits: torch.Tensor | None = None
    embeddings: torch.Tensor | None = None


class _Heads:
    """``net.output_heads`` as the sampler inspects it (model.py:375)."""
    sequence_head = None


class CustomizedESM3(nn.Module):
    def __init__(self, d_model=1536, n_heads=24, v_heads=256, n_layers=48, pretrained=True,
                 n_structure_heads=4101, n_sequence_heads=0, *args, device=None,
                 time_conditioning=True, **kwargs):
        super().__init__()
        if n_sequence_heads:
            raise NotImplementedError("sequence head is off on the ddpm path (mdlm.yaml:58)")
        # `pretrained=True` in the reference downloads ESM3-open weights before the checkpoint
        # overwrites them (net.py:357-360); here weights only ever come from load_state_dict.
        self.dims = Dims(d_model=d_model, n_heads=n_heads, v_heads=v_heads, n_layers=n_layers,
                         n_structure_heads=n_structure_heads, time_conditioning=time_conditioning)
        self.d_model = d_model
        self.engine = Engine(self.dims, device=device)
        self.output_heads = _Heads()
        self._want_embeddings = False

    # -- nn.Module surface the loaders use ---------------------------------------------------
    @property
    def device(self):
        return self.engine.device

    def load_state_dict(self, state_dict, strict=True, prefix="net."):
        sd = {(k if k.startswith(("net.", "sigma_embedder.")) else prefix + k): v
              for k, v in state_dict.items()}
        self.engine.load_state_dict(sd, strict=strict)
        return self

    def to(self, *a, **k):
        return self        # weights are already resident on the engine's device

    def forward(self, structure_tokens, labels=None, mask=None, sequence_tokens=None, *,
                encoder_embeddings=None, ss8_tokens=None, sasa_tokens=None, function_tokens=None,
                residue_annotation_tokens=None, average_plddt=None, per_res_plddt=None,
                structure_coords=None, chain_id=None, sequence_id=None, auxiliary_embeddings=None):
        if labels is not None:
            raise NotImplementedError("training branch (net.py:471-481) is out of the ddpm path")
        for name, v in (("ss8_tokens", ss8_tokens), ("sasa_tokens", sasa_tokens),
                        ("function_tokens", function_tokens),
                        ("residue_annotation_tokens", residue_annotation_tokens),
                        ("average_plddt", average_plddt), ("per_res_plddt", per_res_plddt),
                        ("sequence_id", sequence_id)):
            if v is not None:
                raise NotImplementedError(f"{name}: only the default track values of the ddpm path "
                                          "(net.py:410-436) are implemented")
        if structure_tokens is None:
            raise ValueError("At least one of the inputs must be non-None")
        if sequence_tokens is None:
            sequence_tokens = torch.full_like(structure_tokens, 32)     # sequence mask id, net.py:411
        if structure_coords is not None:
            # net.py:437-441: [..., :3, :] of an atom3 / atom14 / atom37 array -> backbone frames; block 0's
            # geometric attention is live for this call (the ddpm path never passes coordinates: exact zero)
            B, T = structure_tokens.shape
            self.engine.set_structure_coords(structure_coords.expand(B, T, *structure_coords.shape[2:]))
        try:
            logits, emb = self.engine.forward(sequence_tokens, structure_tokens, aux=auxiliary_embeddings,
                                              want_embeddings=self._want_embeddings)
        finally:
            if structure_coords is not None:
                self.engine.set_structure_coords(None)
        return ESMOutput(sequence_logits=None, structure_logits=logits, embeddings=emb)


class TimestepEmbedder(nn.Module):
    """Sinusoidal timestep features -> 2-layer MLP (reference net.py:486-522).  Kept as a torch
    module (3 M parameters) so the reference's own ``_model_wrapper`` can call it; the fused loop
    evaluates the same MLP inside the library once per step (esmdiff_time_embed)."""

    def __init__(self, hidden_size, frequency_embedding_size=256):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(frequency_embedding_size, hidden_size, bias=True), nn.SiLU(),
                                 nn.Linear(hidden_size, hidden_size, bias=True))
        self.frequency_embedding_size = frequency_embedding_size

    @staticmethod
    def timestep_embedding(t, dim, max_period=10000):
        half = dim // 2
        k = torch.arange(start=0, end=half, dtype=t.dtype)
        freqs = torch.exp(-math.log(max_period) * k / half).to(device=t.device, dtype=t.dtype)
        args = t[:, None] * freqs[None]
        emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
        if dim % 2:
            emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
        return emb

    def forward(self, t):
        return self.mlp(self.timestep_embedding(t, self.frequency_embedding_size))
