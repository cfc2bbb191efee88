"""Synthetic weights and inputs for benchmarks (BASELINE.json: "random-init ESM3-open dims").

There is no network for checkpoints, so the benchmark model is the ESM3-open architecture with
the modules' default initialisers: ``nn.Linear`` U(+-1/sqrt(fan_in)) for weight and bias,
``nn.Embedding`` N(0,1) with zeroed padding rows, LayerNorm weight 1 / bias 0, geometric-attention
scales 0.  Generated on the target device (1.4 B parameters take seconds there instead of
minutes on the host) under the state-dict key names of a DeepSpeed ``['module']`` dict
(SURVEY.md 8b), i.e. exactly what ``load_state_dict_from_lightning_ckpt`` would feed the model.
"""
from __future__ import annotations

import math

import torch

from .engine import Dims


def random_state_dict(dims: Dims | None = None, device="cuda", seed: int = 0, full: bool = False) -> dict:
    """``full=False`` leaves out tensors the ddpm path never reads (function/residue embeddings,
    block-0 geometric attention), which the library would drop anyway."""
    d = dims or Dims()
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    D, F, V = d.d_model, d.ffn_hidden, d.n_structure_heads

    def lin(out_f, in_f):
        k = 1.0 / math.sqrt(in_f)
        return (torch.rand(out_f, in_f, device=dev, generator=g) * 2 - 1) * k

    def lin_b(out_f, in_f):
        k = 1.0 / math.sqrt(in_f)
        return (torch.rand(out_f, device=dev, generator=g) * 2 - 1) * k

    def emb(n, w):
        return torch.randn(n, w, device=dev, generator=g)

    ones = lambda n: torch.ones(n, device=dev)
    zeros = lambda n: torch.zeros(n, device=dev)
    sd = {
        "net.encoder.sequence_embed.weight": emb(d.seq_vocab, D),
        "net.encoder.plddt_projection.weight": lin(D, 16),
        "net.encoder.plddt_projection.bias": lin_b(D, 16),
        "net.encoder.structure_per_res_plddt_projection.weight": lin(D, 16),
        "net.encoder.structure_per_res_plddt_projection.bias": lin_b(D, 16),
        "net.encoder.structure_tokens_embed.weight": emb(d.struct_vocab, D),
        "net.encoder.ss8_embed.weight": emb(11, D),
        "net.encoder.sasa_embed.weight": emb(19, D),
    }
    if full:
        for i in range(8):
            w = emb(260, D // 8)
            w[0] = 0
            sd[f"net.encoder.function_embed.{i}.weight"] = w
        w = emb(1478, D)
        w[0] = 0
        sd["net.encoder.residue_embed.weight"] = w
    for l in range(d.n_layers):
        p = f"net.transformer.blocks.{l}."
        sd[p + "attn.layernorm_qkv.0.weight"] = ones(D)
        sd[p + "attn.layernorm_qkv.0.bias"] = zeros(D)
        sd[p + "attn.layernorm_qkv.1.weight"] = lin(3 * D, D)
        sd[p + "attn.out_proj.weight"] = lin(D, D)
        sd[p + "attn.q_ln.weight"] = ones(D)
        sd[p + "attn.k_ln.weight"] = ones(D)
        if l == 0 and full:
            sd[p + "geom_attn.s_norm.weight"] = ones(D)
            sd[p + "geom_attn.proj.weight"] = lin(15 * d.v_heads, D)
            sd[p + "geom_attn.out_proj.weight"] = lin(D, 3 * d.v_heads)
            sd[p + "geom_attn.distance_scale_per_head"] = zeros(d.v_heads)
            sd[p + "geom_attn.rotation_scale_per_head"] = zeros(d.v_heads)
        sd[p + "ffn.0.weight"] = ones(D)
        sd[p + "ffn.0.bias"] = zeros(D)
        sd[p + "ffn.1.weight"] = lin(2 * F, D)
        sd[p + "ffn.3.weight"] = lin(D, F)
    sd["net.transformer.norm.weight"] = ones(D)
    h = "net.output_heads.structure_head."
    sd[h + "0.weight"], sd[h + "0.bias"] = lin(D, D), lin_b(D, D)
    sd[h + "2.weight"], sd[h + "2.bias"] = ones(D), zeros(D)
    sd[h + "3.weight"], sd[h + "3.bias"] = lin(V, D), lin_b(V, D)
    sd["sigma_embedder.mlp.0.weight"] = lin(D, d.time_freq_dim)
    sd["sigma_embedder.mlp.0.bias"] = lin_b(D, d.time_freq_dim)
    sd["sigma_embedder.mlp.2.weight"] = lin(D, D)
    sd["sigma_embedder.mlp.2.bias"] = lin_b(D, D)
    return sd


def random_decoder_state_dict(dims=None, device="cuda", seed: int = 0, full: bool = False) -> dict:
    """Default-initialiser weights of esm's ``StructureTokenDecoder`` (ESM3_structure_decoder_v0
    dims) under its state-dict key names.  ``full`` adds the pairwise (pTM / PAE) head the path
    drops."""
    from .engine import DecoderDims
    d = dims or DecoderDims()
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    D, F = d.d_model, d.ffn_hidden

    def lin(out_f, in_f):
        return (torch.rand(out_f, in_f, device=dev, generator=g) * 2 - 1) / math.sqrt(in_f)

    def lin_b(out_f, in_f):
        return (torch.rand(out_f, device=dev, generator=g) * 2 - 1) / math.sqrt(in_f)

    sd = {"embed.weight": torch.randn(d.struct_vocab, D, device=dev, generator=g)}
    for l in range(d.n_layers):
        p = f"decoder_stack.blocks.{l}."
        sd[p + "attn.layernorm_qkv.0.weight"] = torch.ones(D, device=dev)
        sd[p + "attn.layernorm_qkv.0.bias"] = torch.zeros(D, device=dev)
        sd[p + "attn.layernorm_qkv.1.weight"] = lin(3 * D, D)
        sd[p + "attn.out_proj.weight"] = lin(D, D)
        sd[p + "attn.q_ln.weight"] = torch.ones(D, device=dev)
        sd[p + "attn.k_ln.weight"] = torch.ones(D, device=dev)
        sd[p + "ffn.0.weight"] = torch.ones(D, device=dev)
        sd[p + "ffn.0.bias"] = torch.zeros(D, device=dev)
        sd[p + "ffn.1.weight"] = lin(2 * F, D)
        sd[p + "ffn.3.weight"] = lin(D, F)
    sd["decoder_stack.norm.weight"] = torch.ones(D, device=dev)
    a = "affine_output_projection."
    sd[a + "ffn1.weight"], sd[a + "ffn1.bias"] = lin(D, D), lin_b(D, D)
    sd[a + "norm.weight"], sd[a + "norm.bias"] = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    sd[a + "proj.weight"], sd[a + "proj.bias"] = lin(d.n_affine_out, D), lin_b(d.n_affine_out, D)
    if d.plddt_bins:
        h = "plddt_head."
        sd[h + "0.weight"], sd[h + "0.bias"] = lin(D, D), lin_b(D, D)
        sd[h + "2.weight"], sd[h + "2.bias"] = torch.ones(D, device=dev), torch.zeros(D, device=dev)
        sd[h + "3.weight"], sd[h + "3.bias"] = lin(d.plddt_bins, D), lin_b(d.plddt_bins, D)
    if full:
        h = "pairwise_classification_head."
        sd[h + "linear1.weight"], sd[h + "linear1.bias"] = lin(128, D), lin_b(128, D)
    return sd


def random_encoder_state_dict(dims=None, device="cuda", seed: int = 0, full: bool = False) -> dict:
    """Default-initialiser weights of esm's ``StructureTokenEncoder`` (ESM3_structure_encoder_v0 dims) under its
    state-dict key names.  ``full`` adds the EMA bookkeeping buffers of the codebook that the path drops."""
    from .encoder import EncoderDims
    d = dims or EncoderDims()
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    D, F, H = d.d_model, d.ffn_hidden, d.v_heads

    def lin(out_f, in_f):
        return (torch.rand(out_f, in_f, device=dev, generator=g) * 2 - 1) / math.sqrt(in_f)

    sd = {}
    for l in range(d.n_layers):
        p = f"transformer.blocks.{l}."
        sd[p + "geom_attn.s_norm.weight"] = torch.ones(D, device=dev)
        sd[p + "geom_attn.proj.weight"] = lin(15 * H, D)
        sd[p + "geom_attn.out_proj.weight"] = lin(D, 3 * H)
        sd[p + "geom_attn.distance_scale_per_head"] = torch.zeros(H, device=dev)
        sd[p + "geom_attn.rotation_scale_per_head"] = torch.zeros(H, device=dev)
        sd[p + "ffn.0.weight"] = torch.ones(D, device=dev)
        sd[p + "ffn.0.bias"] = torch.zeros(D, device=dev)
        sd[p + "ffn.1.weight"] = lin(2 * F, D)
        sd[p + "ffn.3.weight"] = lin(D, F)
    sd["transformer.norm.weight"] = torch.ones(D, device=dev)
    sd["pre_vq_proj.weight"] = lin(d.d_out, D)
    sd["pre_vq_proj.bias"] = (torch.rand(d.d_out, device=dev, generator=g) * 2 - 1) / math.sqrt(D)
    sd["codebook.embeddings"] = torch.randn(d.n_codes, d.d_out, device=dev, generator=g)
    sd["relative_positional_embedding.embedding.weight"] = 0.02 * torch.randn(2 * d.rel_bins + 2, D, device=dev, generator=g)
    if full:
        sd["codebook.cluster_size"] = torch.zeros(d.n_codes, device=dev)
        sd["codebook.embeddings_avg"] = sd["codebook.embeddings"].clone()
        sd["codebook.data_initialized"] = torch.zeros(1, device=dev)
    return sd
