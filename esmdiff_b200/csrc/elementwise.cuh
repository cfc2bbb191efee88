// HBM-bound row kernels of the ESM3 forward.
// One warp per token row, 128-bit loads/stores, fp32 statistics.
#pragma once
#include "ptx.cuh"

namespace esmdiff {
namespace ew {

constexpr int ROWS_PER_BLOCK = 8;     // 8 warps = 256 threads

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over the last dim (eps 1e-5), fp32 in -> bf16 out (the next GEMM's A operand).
// nn.LayerNorm semantics: biased variance, y = (x - mean) * rstd * w (+ b).
// D = 128 * VEC4 * 4 ... generic: D % 128 == 0, D <= 128 * MAXV.
// ---------------------------------------------------------------------------------------------
template <int MAXV>
__global__ void __launch_bounds__(ROWS_PER_BLOCK * 32)
layernorm_f32_to_bf16_kernel(const float* __restrict__ x, const float* __restrict__ w,
                             const float* __restrict__ b, __nv_bfloat16* __restrict__ y, int M,
                             int D, float eps) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const int nv = D / 128;                  // float4 per lane
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * D);
    float4 v[MAXV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        if (i < nv) {
            v[i] = xr[i * 32 + lane];
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    const float mean = warp_sum(s) / D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        if (i < nv) {
            float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            q += (a * a + bb * bb) + (c * c + d * d);
        }
    const float rstd = rsqrtf(warp_sum(q) / D + eps);
    uint2* yr = reinterpret_cast<uint2*>(y + static_cast<long long>(row) * D);
    const float4* w4 = reinterpret_cast<const float4*>(w);
    const float4* b4 = reinterpret_cast<const float4*>(b);
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        if (i < nv) {
            const float4 ww = w4[i * 32 + lane];
            float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
            if (b != nullptr) bb = b4[i * 32 + lane];
            const float o0 = (v[i].x - mean) * rstd * ww.x + bb.x;
            const float o1 = (v[i].y - mean) * rstd * ww.y + bb.y;
            const float o2 = (v[i].z - mean) * rstd * ww.z + bb.z;
            const float o3 = (v[i].w - mean) * rstd * ww.w + bb.w;
            yr[i * 32 + lane] = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
        }
}

// ---------------------------------------------------------------------------------------------
// q_ln / k_ln (LayerNorm over the FULL width D, weight only) followed by rotary embedding per
// 64-wide head (non-interleaved rotate-half), in place on the q and k thirds of qkv [M, 3D] bf16.
// esm MultiHeadAttention: q_ln/k_ln then _apply_rotary.  Position = row % T.
// cos/sin tables: fp32 [T, 32].   D % 256 == 0 (lane owns 8 contiguous elements per 256 chunk;
// the rotate-half partner of column o is o +- 32, i.e. lane +- 4).
// ---------------------------------------------------------------------------------------------
template <int MAXC>
__global__ void __launch_bounds__(ROWS_PER_BLOCK * 32, 4)
qk_layernorm_rope_kernel(__nv_bfloat16* __restrict__ qkv, const float* __restrict__ q_w,
                         const float* __restrict__ k_w, const float* __restrict__ cos_t,
                         const float* __restrict__ sin_t, int M, int D, int T, float eps) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const int t = row % T;
    const int nc = D / 256;
    const int o = (lane * 8) & 63;               // offset inside the head
    const bool upper = o >= 32;
    const int fi = o & 31;                       // rotary frequency index of element 0
    float cs[8], sn[8];
    {
        const float4* c4 = reinterpret_cast<const float4*>(cos_t + t * 32 + fi);
        const float4* s4 = reinterpret_cast<const float4*>(sin_t + t * 32 + fi);
        float4 a = c4[0], bq = c4[1], c = s4[0], d = s4[1];
        cs[0] = a.x; cs[1] = a.y; cs[2] = a.z; cs[3] = a.w;
        cs[4] = bq.x; cs[5] = bq.y; cs[6] = bq.z; cs[7] = bq.w;
        sn[0] = c.x; sn[1] = c.y; sn[2] = c.z; sn[3] = c.w;
        sn[4] = d.x; sn[5] = d.y; sn[6] = d.z; sn[7] = d.w;
        if (!upper) {                            // lower half: x1*c - x2*s ; upper half: x2*c + x1*s
#pragma unroll
            for (int e = 0; e < 8; ++e) sn[e] = -sn[e];
        }
    }
    // q, then k.  The values stay PACKED (bf16 pairs) in registers and are unpacked again in each of
    // the three passes (one shift or mask per element): keeping fp32 copies of both rows cost 119
    // registers = 16 warps per SM and 3.4 TB/s; this form fits four blocks (32 warps) per SM, and
    // occupancy, not bytes in flight per warp, is what this kernel was short of.
    auto lo = [](uint32_t w) { return __uint_as_float(w << 16); };
    auto hi = [](uint32_t w) { return __uint_as_float(w & 0xffff0000u); };
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
        __nv_bfloat16* base = qkv + static_cast<long long>(row) * 3 * D + which * D;
        const float* w = which == 0 ? q_w : k_w;
        uint4 rawq[MAXC];
#pragma unroll
        for (int i = 0; i < MAXC; ++i)
            if (i < nc) rawq[i] = *reinterpret_cast<const uint4*>(base + i * 256 + lane * 8);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < MAXC; ++i)
            if (i < nc) {
                const uint32_t r4[4] = {rawq[i].x, rawq[i].y, rawq[i].z, rawq[i].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) s += lo(r4[e]) + hi(r4[e]);
            }
        const float mean = warp_sum(s) / D;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < MAXC; ++i)
            if (i < nc) {
                const uint32_t r4[4] = {rawq[i].x, rawq[i].y, rawq[i].z, rawq[i].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float d0 = lo(r4[e]) - mean, d1 = hi(r4[e]) - mean;
                    q = fmaf(d0, d0, q);
                    q = fmaf(d1, d1, q);
                }
            }
        const float rstd = rsqrtf(warp_sum(q) / D + eps);
        const float nmean = -mean;
#pragma unroll
        for (int i = 0; i < MAXC; ++i)
            if (i < nc) {
                const float4* w4 = reinterpret_cast<const float4*>(w + i * 256 + lane * 8);
                const float4 wa = w4[0], wb = w4[1];
                const float ww[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
                const uint32_t r4[4] = {rawq[i].x, rawq[i].y, rawq[i].z, rawq[i].w};
                float n[8], out[8];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float a0 = rstd * ww[2 * e], a1 = rstd * ww[2 * e + 1];
                    n[2 * e] = fmaf(lo(r4[e]), a0, nmean * a0);          // (x - mean) * rstd * w
                    n[2 * e + 1] = fmaf(hi(r4[e]), a1, nmean * a1);
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float partner = __shfl_xor_sync(0xffffffffu, n[e], 4);
                    // lower half: x1*c - x2*s ; upper half: x2*c + x1*s   (sn carries the sign)
                    out[e] = fmaf(partner, sn[e], n[e] * cs[e]);
                }
                *reinterpret_cast<uint4*>(base + i * 256 + lane * 8) =
                    make_uint4(pack_bf16x2(out[0], out[1]), pack_bf16x2(out[2], out[3]),
                               pack_bf16x2(out[4], out[5]), pack_bf16x2(out[6], out[7]));
            }
    }
}

// Rotary tables as esm RotaryEmbedding builds them: freqs = outer(t, inv_freq) in fp32,
// inv_freq[i] = 1 / base^(2i/64) passed in from the host; cos/sin in fp32.
__global__ void rope_table_kernel(const float* __restrict__ inv_freq, float* __restrict__ cos_t,
                                  float* __restrict__ sin_t, int T) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T * 32) return;
    const float f = static_cast<float>(i / 32) * inv_freq[i % 32];
    cos_t[i] = cosf(f);
    sin_t[i] = sinf(f);
}

// The same table in the layout the QKV GEMM epilogue reads (gemm.cuh EPI_QKV_ROPE_LN): one
// 256-byte row per token position, cos(t f_i) for i < 32 then sin(t f_i).
__global__ void rope_rows_kernel(const float* __restrict__ inv_freq, float* __restrict__ rows, int T) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T * 32) return;
    const float f = static_cast<float>(i / 32) * inv_freq[i % 32];
    rows[(i / 32) * 64 + (i % 32)] = cosf(f);
    rows[(i / 32) * 64 + 32 + (i % 32)] = sinf(f);
}

// ---------------------------------------------------------------------------------------------
// Input embedding of the ddpm path (esm EncodeInputs with six tracks at their defaults, then
// CustomizedESM3.forward's "+ auxiliary_embeddings", net.py:445-466):
//   x[m,:] = seq_embed[seq[m]] + const_vec + struct_embed[force(xt[m], seq[m])] + aux[m or 0]
// const_vec = plddt_projection(rbf(1)) + per_res_plddt_projection(rbf(0)) + ss8[0] + sasa[0]
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ long long force_structure_id(long long st, long long sq) {
    if (st == -1) st = 4096;
    if (sq == 0) st = 4098;        // BOS
    if (sq == 1) st = 4099;        // PAD
    if (sq == 2) st = 4097;        // EOS
    if (sq == 31) st = 4100;       // CHAINBREAK
    return st;
}

__global__ void __launch_bounds__(256)
embed_kernel(const long long* __restrict__ seq, const long long* __restrict__ xt,
             const float* __restrict__ seq_embed, const float* __restrict__ struct_embed,
             const float* __restrict__ const_vec, const float* __restrict__ aux,
             long long aux_row_stride, float* __restrict__ x, int M, int D, int seq_vocab,
             int struct_vocab, int* __restrict__ err, __nv_bfloat16* __restrict__ xb,
             float2* __restrict__ stats) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    long long sq = seq[row];
    long long st = force_structure_id(xt[row], sq);
    if (sq < 0 || sq >= seq_vocab || st < 0 || st >= struct_vocab) {
        if (lane == 0) atomicExch(err, 1);     // torch would raise IndexError; reported by the host
        sq = 0;
        st = 0;
    }
    const float4* a = reinterpret_cast<const float4*>(seq_embed + sq * D);
    const float4* b = reinterpret_cast<const float4*>(struct_embed + st * D);
    const float4* c = reinterpret_cast<const float4*>(const_vec);
    const float4* d = aux ? reinterpret_cast<const float4*>(aux + row * aux_row_stride) : nullptr;
    float4* o = reinterpret_cast<float4*>(x + static_cast<long long>(row) * D);
    for (int i = lane; i < D / 4; i += 32) {
        const float4 va = a[i], vb = b[i], vc = c[i];
        float4 r;
        // reference order: ((seq + plddt + per_res) + structure) + ss8 + sasa  [+ aux]
        r.x = (va.x + vc.x) + vb.x; r.y = (va.y + vc.y) + vb.y;
        r.z = (va.z + vc.z) + vb.z; r.w = (va.w + vc.w) + vb.w;
        if (d) { const float4 vd = d[i]; r.x += vd.x; r.y += vd.y; r.z += vd.z; r.w += vd.w; }
        o[i] = r;
        if (xb != nullptr) {
            // LayerNorm folded through the first QKV GEMM (gemm.cuh): bf16 copy of the row and the
            // (mean, M2) of each 128-column span (= one trip of this loop across the warp)
            reinterpret_cast<uint2*>(xb + static_cast<long long>(row) * D)[i] =
                make_uint2(pack_bf16x2(r.x, r.y), pack_bf16x2(r.z, r.w));
            const float mean = warp_sum((r.x + r.y) + (r.z + r.w)) * (1.0f / 128.0f);
            const float a0 = r.x - mean, a1 = r.y - mean, a2 = r.z - mean, a3 = r.w - mean;
            const float m2 = warp_sum((a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3));
            if (lane == 0) stats[static_cast<long long>(row) * (D / 128) + i / 32] = make_float2(mean, m2);
        }
    }
}

// const_vec[d] for the default tracks.  rbf(v)[k] = exp(-((v - k/15) * 16)^2), k < 16.
__global__ void default_tracks_kernel(const float* __restrict__ plddt_w, const float* __restrict__ plddt_b,
                                      const float* __restrict__ res_w, const float* __restrict__ res_b,
                                      const float* __restrict__ ss8, const float* __restrict__ sasa,
                                      float* __restrict__ out, int D) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    float a = 0.f, b = 0.f;
    for (int k = 0; k < 16; ++k) {
        const float c = static_cast<float>(k) / 15.0f;
        const float z1 = (1.0f - c) / 0.0625f, z0 = (0.0f - c) / 0.0625f;
        a += plddt_w[d * 16 + k] * expf(-(z1 * z1));
        b += res_w[d * 16 + k] * expf(-(z0 * z0));
    }
    out[d] = ((a + plddt_b[d]) + (b + res_b[d])) + ss8[d] + sasa[d];
}

// ---------------------------------------------------------------------------------------------
// TimestepEmbedder (net.py:486-522): cond = W2 silu(W0 [cos(s f_k), sin(s f_k)] + b0) + b2,
// f_k = exp(-ln(1e4) k / 128), k < 128.  One sigma for the whole batch (all rows identical on
// this path), so it is computed once per step instead of B times.
// Two launches: hidden, then output.  One warp per output feature, coalesced weight rows.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
time_embed_hidden_kernel(float sigma, const float* __restrict__ w0, const float* __restrict__ b0,
                         float* __restrict__ hidden, int D, int F) {
    pdl_launch_dependents();
    pdl_wait();
    const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (j >= D) return;
    const int half = F / 2;
    float s = 0.f;
    for (int k = lane; k < F; k += 32) {
        const int kk = k < half ? k : k - half;
        const float freq = expf(-9.210340371976184f * static_cast<float>(kk) / static_cast<float>(half));
        const float arg = sigma * freq;
        const float feat = k < half ? cosf(arg) : sinf(arg);
        s += w0[j * F + k] * feat;
    }
    s = warp_sum(s);
    if (lane == 0) {
        const float z = s + b0[j];
        hidden[j] = z / (1.0f + expf(-z));
    }
}
__global__ void __launch_bounds__(256)
time_embed_out_kernel(const float* __restrict__ hidden, const float* __restrict__ w2,
                      const float* __restrict__ b2, float* __restrict__ cond, int D) {
    pdl_launch_dependents();
    pdl_wait();
    const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (j >= D) return;
    float s = 0.f;
    for (int k = lane; k < D; k += 32) s += w2[static_cast<long long>(j) * D + k] * hidden[k];
    s = warp_sum(s);
    if (lane == 0) cond[j] = s + b2[j];
}

// Column means over each block of `block_rows` rows: out[b][k] = mean_{r in block b} W[r][k].
// For the q and k thirds of the QKV weight: q_ln / k_ln subtract the mean over the OUTPUT features
// of a row, which is linear in the input -- removing these column means from the weight rows makes
// the GEMM produce q - mean(q) directly (gemm.cuh EPI_QKV_ROPE_LN).
__global__ void column_mean_kernel(const float* __restrict__ src, float* __restrict__ out, long long block_rows,
                                   long long cols) {
    const long long k = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (k >= cols) return;
    const float* base = src + static_cast<long long>(blockIdx.y) * block_rows * cols + k;
    float s = 0.f, comp = 0.f;                       // Kahan: 1536 terms of alternating sign
    for (long long r = 0; r < block_rows; ++r) {
        const float y = base[r * cols] - comp;
        const float t = s + y;
        comp = (t - s) - y;
        s = t;
    }
    out[static_cast<long long>(blockIdx.y) * cols + k] = s / static_cast<float>(block_rows);
}

// LayerNorm folded into the following Linear (gemm.cuh, *_LN epilogues), one warp per output row n:
//   w[n, k]    = W[src(n), k] - colmean[n / center_block][k]   for n < center_rows (else W[src(n), k])
//                for those rows src(n) also interleaves the rotary partners of every 64-wide head:
//                position p of a head holds feature (p >> 1) + 32 (p & 1), i.e. (d, d + 32) adjacent,
//                so that an epilogue thread finds both halves of a rotation in consecutive registers.
//                q and k get the SAME permutation, which leaves every q.k dot product unchanged.
//   dst[n, k]  = bf16(w[n, k] * gamma[k])
//   colsum[n]  = sum_k float(dst[n, k])          (of the ROUNDED weights: it multiplies the row mean)
//   bias[n]    = sum_k beta[k] * w[n, k]         (fp32; beta may be null)
// src(n) applies the SwiGLU gate/up interleave of convert_rows_bf16_kernel when swiglu_hidden > 0.
__global__ void __launch_bounds__(256)
fold_layernorm_weight_kernel(const float* __restrict__ src, const float* __restrict__ gamma,
                             const float* __restrict__ beta, __nv_bfloat16* __restrict__ dst,
                             float* __restrict__ colsum, float* __restrict__ bias, long long rows,
                             long long cols, int swiglu_hidden, const float* __restrict__ colmean = nullptr,
                             long long center_rows = 0, long long center_block = 1) {
    const long long r = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    long long sr = r;
    if (swiglu_hidden > 0) {
        const long long blk = r / 256, within = r % 256;
        sr = within < 128 ? blk * 128 + within : swiglu_hidden + blk * 128 + (within - 128);
    }
    const float* cm = (colmean != nullptr && r < center_rows) ? colmean + (r / center_block) * cols : nullptr;
    if (r < center_rows) {
        const long long within = r % 64;
        sr = r - within + (within >> 1) + 32 * (within & 1);
    }
    float cs = 0.f, bs = 0.f;
    for (long long k = lane; k < cols; k += 32) {
        float w = src[sr * cols + k];
        if (cm != nullptr) w -= cm[k];
        const __nv_bfloat16 wf = __float2bfloat16_rn(w * gamma[k]);
        dst[r * cols + k] = wf;
        cs += __bfloat162float(wf);
        if (beta != nullptr) bs += beta[k] * w;
    }
    cs = warp_sum(cs);
    bs = warp_sum(bs);
    if (lane == 0) {
        colsum[r] = cs;
        bias[r] = bs;
    }
}

// dst[p] = src[rotary-partner interleave of p] per 64-wide head (see fold_layernorm_weight_kernel).
__global__ void interleave_rotary_pairs_kernel(const float* __restrict__ src, float* __restrict__ dst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int within = i % 64;
    dst[i] = src[i - within + (within >> 1) + 32 * (within & 1)];
}

// fp32 -> bf16 weight conversion with an optional row permutation (SwiGLU gate/up interleave).
__global__ void convert_rows_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                         long long rows, long long cols, int swiglu_hidden) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const long long r = i / cols, c = i % cols;
    long long sr = r;
    if (swiglu_hidden > 0) {
        // dst rows: per 256 block -> [128 gate rows | 128 up rows] of the same hidden indices
        const long long blk = r / 256, within = r % 256;
        sr = within < 128 ? blk * 128 + within : swiglu_hidden + blk * 128 + (within - 128);
    }
    dst[i] = __float2bfloat16_rn(src[sr * cols + c]);
}

}  // namespace ew
}  // namespace esmdiff
