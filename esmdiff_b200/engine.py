"""Thin object over the C ABI: owns an ``esmdiff_ctx`` and moves torch tensors' pointers across it.

PyTorch is plumbing here (device memory, streams); all arithmetic on the path happens inside
``libesmdiff_b200.so``.  No CPU fallback: constructing an :class:`Engine` without a B200 raises.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import torch

from . import _lib
from ._lib import Cfg, EsmdiffError

MASK = 4096           # C.STRUCTURE_MASK_TOKEN (model.py:381)


@dataclass
class Dims:
    """CustomizedESM3.__init__ arguments (net.py:323-334) + mdlm.yaml:26-58."""
    d_model: int = 1536
    n_heads: int = 24
    v_heads: int = 256
    n_layers: int = 48
    n_structure_heads: int = 4101
    seq_vocab: int = 64
    struct_vocab: int = 4101
    time_freq_dim: int = 256
    time_conditioning: bool = True

    @property
    def ffn_hidden(self) -> int:
        return int(((8.0 / 3.0 * self.d_model) + 255) // 256 * 256)


def _ptr(t: torch.Tensor | None):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@dataclass
class DecoderDims:
    """esm ``StructureTokenDecoder(d_model=1280, n_heads=20, n_layers=30)`` (ESM3_structure_decoder_v0):
    the VQ-VAE decoder behind ``ESM3.decode`` (reference call site slm/sample_esmdiff.py:56-61)."""
    d_model: int = 1280
    n_heads: int = 20
    n_layers: int = 30
    n_affine_out: int = 9 + 7 * 2          # Dim6RotStructureHead.proj: trans 3, x 3, y 3, 7 torsion sin/cos pairs
    plddt_bins: int = 50                   # C.VQVAE_PLDDT_BINS; 0 = no pLDDT head
    struct_vocab: int = 4101               # 4096 codes + 5 special tokens

    @property
    def ffn_hidden(self) -> int:
        return int(((8.0 / 3.0 * self.d_model) + 255) // 256 * 256)


class Engine:
    def __init__(self, dims: Dims | DecoderDims | None = None, device: int | None = None):
        self.dims = dims or Dims()
        self.L = _lib.lib()
        if not torch.cuda.is_available():
            raise EsmdiffError("esmdiff_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        d = self.dims
        if isinstance(d, DecoderDims):
            cfg = Cfg(d.d_model, d.n_heads, d.n_layers, d.ffn_hidden, d.n_affine_out, 0, d.struct_vocab, 0, 0,
                      1, d.plddt_bins)
        else:
            cfg = Cfg(d.d_model, d.n_heads, d.n_layers, d.ffn_hidden, d.n_structure_heads, d.seq_vocab,
                      d.struct_vocab, d.time_freq_dim, int(d.time_conditioning), 0, 0, d.v_heads)
        h = C.c_void_p()
        rc = self.L.esmdiff_create(C.byref(cfg), self.device_index, C.byref(h))
        if rc != 0:
            raise EsmdiffError(self.L.esmdiff_last_error(None).decode())
        self.h = h
        self.finalized = False

    # -- plumbing ---------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != 0:
            msg = self.L.esmdiff_last_error(self.h).decode()
            if msg.startswith("IndexError"):
                raise IndexError(msg)
            raise EsmdiffError(msg)

    def close(self):
        if getattr(self, "h", None):
            self.L.esmdiff_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_time_conditioning(self, on: bool):
        """MaskedDiffusionLanguageModeling(time_conditioning=...) (model.py:333): False zeroes sigma."""
        self._check(self.L.esmdiff_set_time_conditioning(self.h, int(bool(on))))
        self.dims.time_conditioning = bool(on)

    def synchronize(self):
        self._check(self.L.esmdiff_synchronize(self.h, _stream()))

    @property
    def launch_count(self) -> int:
        return int(self.L.esmdiff_launch_count(self.h))

    PROF_KINDS = {"gemm_store_bf16": 0, "gemm_resid_f32": 1, "gemm_swiglu": 2, "gemm_bias_gelu": 3,
                  "gemm_bias": 4, "attention": 5, "sampler": 6, "layernorm": 7, "qk_norm_rope": 8,
                  "embed": 9}

    def profile(self, on: bool):
        self._check(self.L.esmdiff_profile_enable(self.h, int(on)))

    def profile_read(self) -> dict:
        """{kind: (ms, work, launches)} measured with CUDA events around each launch."""
        out = {}
        for name, k in self.PROF_KINDS.items():
            ms, work, n = C.c_double(), C.c_double(), C.c_int64()
            self._check(self.L.esmdiff_profile_read(self.h, k, C.byref(ms), C.byref(work), C.byref(n)))
            out[name] = (ms.value, work.value, n.value)
        return out

    # -- weights ----------------------------------------------------------------------------
    def set_weight(self, key: str, tensor: torch.Tensor):
        t = tensor.detach()
        if t.dtype not in (torch.float32, torch.bfloat16):
            t = t.float()
        t = t.contiguous()
        shape = (C.c_int64 * t.dim())(*t.shape)
        dtype = 0 if t.dtype == torch.float32 else 1
        self._check(self.L.esmdiff_set_weight(self.h, key.encode(), _ptr(t), int(t.is_cuda), dtype,
                                              shape, t.dim()))
        self.finalized = False

    def load_state_dict(self, sd: dict, strict: bool = True):
        """Keys as in the DeepSpeed ``['module']`` dict (checkpoint_utils.py:62-64)."""
        for k, v in sd.items():
            if not torch.is_tensor(v):
                continue
            try:
                self.set_weight(k, v)
            except EsmdiffError as e:
                if strict:
                    raise RuntimeError(f"Error(s) in loading state_dict: {e}") from e
        self.finalize()

    def finalize(self):
        rc = self.L.esmdiff_finalize_weights(self.h)
        if rc != 0:
            raise RuntimeError("Error(s) in loading state_dict: " + self.L.esmdiff_last_error(self.h).decode())
        self.finalized = True

    # -- the path ---------------------------------------------------------------------------
    def time_embed(self, sigma: float) -> torch.Tensor:
        out = torch.empty(self.dims.d_model, dtype=torch.float32, device=self.device)
        self._check(self.L.esmdiff_time_embed(self.h, float(sigma), _ptr(out), _stream()))
        return out

    def forward(self, sequence_tokens, structure_tokens, aux=None, want_embeddings=False,
                logits_out=None):
        """aux: None, (d,), or (B,T,d) fp32.  Returns (logits (B,T,V) fp32, embeddings|None)."""
        B, T = structure_tokens.shape
        seq = sequence_tokens.to(self.device, torch.int64).expand(B, T).contiguous()
        xt = structure_tokens.to(self.device, torch.int64).contiguous()
        V, D = self.dims.n_structure_heads, self.dims.d_model
        logits = logits_out if logits_out is not None else torch.empty(
            B, T, V, dtype=torch.float32, device=self.device)
        emb = torch.empty(B, T, D, dtype=torch.float32, device=self.device) if want_embeddings else None
        stride = 0
        if aux is not None:
            aux = aux.to(self.device, torch.float32)
            if aux.dim() == 1:
                aux, stride = aux.contiguous(), 0
            else:
                aux, stride = aux.expand(B, T, D).contiguous(), D
        self._check(self.L.esmdiff_forward(self.h, _ptr(seq), _ptr(xt), B, T, _ptr(aux), stride,
                                           _ptr(logits), _ptr(emb), _stream()))
        return logits, emb

    def set_structure_coords(self, coords: torch.Tensor | None):
        """``structure_coords`` of ``CustomizedESM3.forward`` (net.py:385, 433-441) for the NEXT forwards and sampling
        loops: (B, T, >=3, 3) N, CA, C per token position (NaN / inf = unknown; BOS / EOS rows too), or None = the
        ddpm path's default (no frames: block 0's geometric attention is exactly 0 and skipped)."""
        if coords is None:
            self._check(self.L.esmdiff_set_structure_coords(self.h, None, 0, 0, _stream()))
            return
        assert coords.dim() == 4 and coords.size(-1) == 3 and coords.size(-2) >= 3, "need (B, T, >=3, 3): N, CA, C"
        B, T = coords.shape[:2]
        c = coords[..., :3, :].to(self.device, torch.float32).contiguous()
        self._check(self.L.esmdiff_set_structure_coords(self.h, _ptr(c), B, T, _stream()))
        torch.cuda.current_stream().synchronize()      # `c` is a temporary: the frames are built before it goes away

    def forward_sigma(self, sequence_tokens, structure_tokens, sigma: float, logits_out=None):
        B, T = structure_tokens.shape
        seq = sequence_tokens.to(self.device, torch.int64).expand(B, T).contiguous()
        xt = structure_tokens.to(self.device, torch.int64).contiguous()
        logits = logits_out if logits_out is not None else torch.empty(
            B, T, self.dims.n_structure_heads, dtype=torch.float32, device=self.device)
        self._check(self.L.esmdiff_forward_sigma(self.h, _ptr(seq), _ptr(xt), B, T, float(sigma),
                                                 _ptr(logits), _stream()))
        return logits

    def logits_parameterization(self, logits, xt, out=None):
        B, T, _ = logits.shape
        out = logits if out is None else out
        self._check(self.L.esmdiff_logits_parameterization(self.h, _ptr(logits), _ptr(xt), B, T,
                                                           _ptr(out), _stream()))
        return out

    def sample_step(self, x, logits, u, mc_t: float, mc_s: float, seed: int = 0, step: int = 0):
        """In place on ``x`` (int64 (B,T) on device).  ``u`` None -> library Philox stream."""
        B, T = x.shape
        assert x.is_contiguous() and logits.is_contiguous() and (u is None or u.is_contiguous())
        self._check(self.L.esmdiff_sample_step(self.h, _ptr(x), _ptr(logits), _ptr(u), float(mc_t),
                                               float(mc_s), B, T, int(seed), int(step), _stream()))
        return x

    def denoise_argmax(self, x, logits):
        B, T = x.shape
        self._check(self.L.esmdiff_denoise_argmax(self.h, _ptr(x), _ptr(logits), B, T, _stream()))
        return x

    def schedule(self, steps: int, eps: float = 1e-5, noise_eps: float = 1e-3):
        s = (C.c_float * (steps + 1))()
        a = (C.c_float * steps)()
        b = (C.c_float * steps)()
        rc = self.L.esmdiff_schedule(steps, eps, noise_eps, s, a, b)
        assert rc == 0
        return list(s), list(a), list(b)

    def ddpm_sample(self, sequence_tokens, prior, steps, sigma, mc_t, mc_s, seed=0,
                    noise_removal=True):
        """Device-resident fused loop with the library's Philox uniforms."""
        B, T = sequence_tokens.shape
        seq = sequence_tokens.to(self.device, torch.int64).contiguous()
        pr = prior.to(self.device, torch.int64).contiguous() if prior is not None else None
        out = torch.empty(B, T, dtype=torch.int64, device=self.device)
        fa = lambda v: (C.c_float * len(v))(*[float(z) for z in v])
        self._check(self.L.esmdiff_ddpm_sample(self.h, _ptr(seq), _ptr(pr), B, T, int(steps), fa(sigma),
                                               fa(mc_t), fa(mc_s), int(seed), int(noise_removal),
                                               _ptr(out), _stream()))
        return out

    def gibbs_step(self, x, logits, noise, temperature: float, top_p: float, k: int, seed: int = 0, step: int = 0):
        """One entropy-ordered unmasking step in place on ``x`` (int64 (B,T) on device).  ``noise`` None ->
        library Philox stream, else Exp(1) draws (B,T,V) fp32."""
        B, T = x.shape
        assert x.is_contiguous() and logits.is_contiguous() and (noise is None or noise.is_contiguous())
        self._check(self.L.esmdiff_gibbs_step(self.h, _ptr(x), _ptr(logits), _ptr(noise), B, T, float(temperature),
                                              float(top_p), int(k), int(seed), int(step), _stream()))
        return x

    def gibbs_sample(self, sequence_tokens, prior, k_per_step, temperature: float, top_p: float, seed: int = 0):
        """Device-resident loop: ``len(k_per_step)`` x {forward without time conditioning, gibbs step}."""
        B, T = prior.shape
        seq = sequence_tokens.to(self.device, torch.int64).expand(B, T).contiguous()
        pr = prior.to(self.device, torch.int64).contiguous()
        out = torch.empty(B, T, dtype=torch.int64, device=self.device)
        ks = (C.c_int * len(k_per_step))(*[int(k) for k in k_per_step])
        self._check(self.L.esmdiff_gibbs_sample(self.h, _ptr(seq), _ptr(pr), B, T, len(k_per_step), ks,
                                                float(temperature), float(top_p), int(seed), _ptr(out), _stream()))
        return out

    def ddpm_sample_host(self, seq_host: torch.Tensor, prior_host, steps, eps=1e-5, seed=0,
                         noise_removal=True) -> torch.Tensor:
        """End to end through host buffers (H2D + loop + D2H inside the call)."""
        B, T = seq_host.shape
        assert not seq_host.is_cuda and seq_host.dtype == torch.int64 and seq_host.is_contiguous()
        out = torch.empty(B, T, dtype=torch.int64, pin_memory=True)
        self._check(self.L.esmdiff_ddpm_sample_host(self.h, _ptr(seq_host), _ptr(prior_host), B, T,
                                                    int(steps), float(eps), int(seed),
                                                    int(noise_removal), _ptr(out)))
        return out

    def decode_structure(self, structure_tokens, want_affine=False):
        """Batched VQ-VAE structure decode (engine built from :class:`DecoderDims`).  int64 (B,T)
        tokens INCLUDING BOS/EOS -> (bb (B,T,3,3) N/CA/C, o (B,T,3), plddt (B,T) | None, affine | None)."""
        B, T = structure_tokens.shape
        tok = structure_tokens.to(self.device, torch.int64).contiguous()
        bb = torch.empty(B, T, 3, 3, dtype=torch.float32, device=self.device)
        o = torch.empty(B, T, 3, dtype=torch.float32, device=self.device)
        pl = torch.empty(B, T, dtype=torch.float32, device=self.device) if self.dims.plddt_bins else None
        aff = torch.empty(B, T, self.dims.n_affine_out, dtype=torch.float32, device=self.device) if want_affine else None
        self._check(self.L.esmdiff_decode_structure(self.h, _ptr(tok), B, T, _ptr(bb), _ptr(o), _ptr(pl), _ptr(aff),
                                                    _stream()))
        return bb, o, pl, aff

    # -- single kernels (tests, roofline timing) ----------------------------------------------
    def op_gemm(self, epilogue: int, a, w, out, bias=None, scale=1.0, n=None):
        M, K = a.shape
        N = w.shape[0] if n is None else n
        self._check(self.L.esmdiff_op_gemm(self.h, epilogue, _ptr(a), _ptr(w), M, N, K, _ptr(out),
                                           out.stride(0), _ptr(bias), float(scale), _stream()))
        return out

    def op_gemm_ln(self, epilogue: int, a, w, out, bias=None, scale=1.0, stats_in=None, colsum=None,
                   stats_out=None, xb_out=None):
        """LayerNorm-folded epilogues 5/6/7 (gemm.cuh)."""
        M, K = a.shape
        N = w.shape[0]
        self._check(self.L.esmdiff_op_gemm_ln(self.h, epilogue, _ptr(a), _ptr(w), M, N, K, _ptr(out),
                                              out.stride(0), _ptr(bias), float(scale), _ptr(stats_in),
                                              _ptr(colsum), _ptr(stats_out), _ptr(xb_out), _stream()))
        return out

    @property
    def stats_span(self) -> int:
        """Columns per partial LayerNorm statistic the last residual+statistics GEMM (epilogue 6) wrote."""
        return int(self.L.esmdiff_op_stats_span(self.h))

    def op_fold_layernorm(self, w, gamma, beta=None, swiglu_hidden=0, center_rows=0, center_block=1):
        """center_rows > 0: column means of each block of ``center_block`` rows are removed from the
        first ``center_rows`` rows first (q_ln / k_ln centring folded into the QKV weight)."""
        rows, cols = w.shape
        dst = torch.empty(rows, cols, dtype=torch.bfloat16, device=w.device)
        colsum = torch.empty(rows, dtype=torch.float32, device=w.device)
        bias = torch.empty(rows, dtype=torch.float32, device=w.device)
        if center_rows:
            assert not swiglu_hidden
            self._check(self.L.esmdiff_op_fold_layernorm_centered(
                self.h, _ptr(w), _ptr(gamma), _ptr(beta), _ptr(dst), _ptr(colsum), _ptr(bias), rows, cols,
                int(center_rows), int(center_block), _stream()))
        else:
            self._check(self.L.esmdiff_op_fold_layernorm(self.h, _ptr(w), _ptr(gamma), _ptr(beta), _ptr(dst),
                                                         _ptr(colsum), _ptr(bias), rows, cols,
                                                         int(swiglu_hidden), _stream()))
        return dst, colsum, bias

    def op_gemm_qkv_rope(self, a, w, bias, stats_in, colsum, qk_gamma, T, n_rope):
        """QKV projection with the pre-LN, q_ln / k_ln and RoPE folded in (gemm.cuh epilogue 8).
        Returns (qkv bf16 [M, N], qk_sumsq fp32 [M, n_rope / 128])."""
        M, K = a.shape
        N = w.shape[0]
        out = torch.empty(M, N, dtype=torch.bfloat16, device=a.device)
        sumsq = torch.zeros(M, n_rope // 128, dtype=torch.float32, device=a.device)
        self._check(self.L.esmdiff_op_gemm_qkv_rope(self.h, _ptr(a), _ptr(w), M, N, K, _ptr(out), out.stride(0),
                                                    _ptr(bias), _ptr(stats_in), _ptr(colsum), _ptr(qk_gamma),
                                                    _ptr(sumsq), int(T), int(n_rope), _stream()))
        return out, sumsq

    def op_layernorm(self, x, w, b=None):
        M, D = x.shape
        y = torch.empty(M, D, dtype=torch.bfloat16, device=x.device)
        self._check(self.L.esmdiff_op_layernorm(self.h, _ptr(x), _ptr(w), _ptr(b), _ptr(y), M, D, _stream()))
        return y

    def op_qk_norm_rope(self, qkv, q_w, k_w, B, T):
        D = qkv.shape[-1] // 3
        self._check(self.L.esmdiff_op_qk_norm_rope(self.h, _ptr(qkv), _ptr(q_w), _ptr(k_w), B, T, D, _stream()))
        return qkv

    def op_attention(self, qkv, B, T, H, qk_sumsq=None):
        """qk_sumsq None: q, k already normalised + rotated.  Otherwise the un-normalised q', k' and
        the partial sums of squares of the fused QKV epilogue ([B*T, 2*H*64/128])."""
        out = torch.empty(B * T, H * 64, dtype=torch.bfloat16, device=qkv.device)
        if qk_sumsq is None:
            self._check(self.L.esmdiff_op_attention(self.h, _ptr(qkv), _ptr(out), B, T, H, _stream()))
        else:
            assert qk_sumsq.is_contiguous() and qk_sumsq.shape == (B * T, 2 * H * 64 // 128)
            self._check(self.L.esmdiff_op_attention_ln(self.h, _ptr(qkv), _ptr(qk_sumsq), _ptr(out), B, T, H,
                                                       _stream()))
        return out

    def op_convert_bf16(self, src, swiglu_hidden=0):
        rows, cols = src.shape
        dst = torch.empty(rows, cols, dtype=torch.bfloat16, device=src.device)
        self._check(self.L.esmdiff_op_convert_bf16(self.h, _ptr(src), _ptr(dst), rows, cols,
                                                   int(swiglu_hidden), _stream()))
        return dst


def residue_scale(n_layers: int) -> float:
    return math.sqrt(n_layers / 36)
