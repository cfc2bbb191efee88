// Fused MDLM sampling step (SURVEY.md 2.2 k14 + k15): one pass over the logits row instead of
// the reference's ~13 element-wise passes.
//   logits_parameterization  reference slm/models/model.py:527-533
//   _ddpm_update tail        reference slm/models/model.py:602-607
//   _sample_categorical      reference slm/models/model.py:24-28
// One CTA per token row (V = 4101 fp32, 16 KiB): HBM-bound, rows that are already unmasked are
// skipped without touching their logits (their update is the identity, model.py:606-607).
#pragma once
#include "ptx.cuh"

namespace esmdiff {
namespace sampler {

constexpr int THREADS = 256;
constexpr int MAX_PER_THREAD = 17;           // V <= 256 * 17 = 4352
constexpr int GROUPS = 5;                    // a thread owns columns 4 (t + 256 k) .. + 3, k < 5: 5120 >= 4352
constexpr float NEG_INF = -1000000.0f;

struct RowStats {
    float lse;
};

__device__ __forceinline__ float block_max(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int i = 1; i < THREADS / 32; ++i) r = fmaxf(r, red[i]);
    return r;
}
__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < THREADS / 32; ++i) r += red[i];
    return r;
}

// Philox4x32-10 (Salmon et al. 2011), counter-based: the library's own uniform stream when the
// caller does not pass uniforms.  counter = (element/4, row, step, 0), key = seed.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}
// The uniform of (row, col) is word (col & 3) of the block with counter (col >> 2, row, step, 0): a
// thread owns four consecutive columns, so one Philox block serves four elements (the first
// version had consecutive THREADS on consecutive columns and spent 4 blocks per 4 elements: ncu
// r1j, 77 % of the issue slots, 630 GB/s).
__device__ __forceinline__ void philox_uniform4(unsigned long long seed, uint32_t step, uint32_t row,
                                                uint32_t col4, float* uu) {
    const uint4 r = philox4x32_10(make_uint4(col4, row, step, 0u),
                                  make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
    uu[0] = static_cast<float>(r.x >> 8) * (1.0f / 16777216.0f);    // [0, 1)
    uu[1] = static_cast<float>(r.y >> 8) * (1.0f / 16777216.0f);
    uu[2] = static_cast<float>(r.z >> 8) * (1.0f / 16777216.0f);
    uu[3] = static_cast<float>(r.w >> 8) * (1.0f / 16777216.0f);
}

// MODE 0: ddpm update (race argmax with uniforms)   MODE 1: noise removal (argmax of log p)
// MODE 2: write log p(x0) to out_logp (logits_parameterization only; all rows)
template <int MODE>
__global__ void __launch_bounds__(THREADS)
sample_rows_kernel(const float* __restrict__ logits, long long ld, const float* __restrict__ u,
                   long long* __restrict__ x, float* __restrict__ out_logp, int M, int V,
                   int mask_index, float mc_t, float mc_s, unsigned long long seed, uint32_t step,
                   uint32_t row_offset) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float red[THREADS / 32];
    __shared__ int red_i[THREADS / 32];
    const int row = blockIdx.x;
    const long long xt = x[row];
    const bool masked = xt == mask_index;
    if (!masked) {
        if constexpr (MODE == 2) {
            float* o = out_logp + row * ld;
            for (int i = threadIdx.x; i < V; i += THREADS) o[i] = (i == xt) ? 0.0f : NEG_INF;
        }
        return;     // MODE 0/1: copy_flag * x -> x keeps its value
    }
    const float* lr = logits + row * ld;
    float v[GROUPS * 4];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < GROUPS; ++k) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = 4 * (threadIdx.x + k * THREADS) + e;
            float a = -INFINITY;
            if (i < V) {
                a = lr[i];
                if (i == mask_index) a += NEG_INF;
            }
            v[4 * k + e] = a;
            mx = fmaxf(mx, a);
        }
    }
    mx = block_max(mx, red);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < GROUPS * 4; ++k) s += expf(v[k] - mx);          // exp(-inf) = 0 for padding
    s = block_sum(s, red);
    const float lse = logf(s) + mx;

    if constexpr (MODE == 2) {
        float* o = out_logp + row * ld;
#pragma unroll
        for (int k = 0; k < GROUPS; ++k)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int i = 4 * (threadIdx.x + k * THREADS) + e;
                if (i < V) o[i] = v[4 * k + e] - lse;
            }
        return;
    }

    float best = -INFINITY;
    int best_i = 0x7fffffff;
    const float dmc = mc_t - mc_s;
#pragma unroll
    for (int k = 0; k < GROUPS; ++k) {
        const int i0 = 4 * (threadIdx.x + k * THREADS);
        if (i0 < V) {
            float uu[4] = {0.f, 0.f, 0.f, 0.f};
            if constexpr (MODE == 0) {
                if (u == nullptr) philox_uniform4(seed, step, row + row_offset, i0 >> 2, uu);    // counter = row of the whole batch
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int i = i0 + e;
                if (i < V) {
                    float score;
                    if constexpr (MODE == 0) {
                        float q = expf(v[4 * k + e] - lse) * dmc;
                        if (i == mask_index) q = mc_s;
                        const float ur = u ? u[row * ld + i] : uu[e];
                        const float g = 1e-10f - logf(ur + 1e-10f);
                        score = q / g;
                    } else {
                        score = v[4 * k + e] - lse;
                    }
                    if (score > best) { best = score; best_i = i; }      // ascending i: first max wins
                }
            }
        }
    }
    // (max score, min index) reduction
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) { red[w] = best; red_i[w] = best_i; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < THREADS / 32; ++i)
            if (red[i] > best || (red[i] == best && red_i[i] < best_i)) { best = red[i]; best_i = red_i[i]; }
        // no comparison succeeds when every score is NaN (NaN logits): torch.argmax then returns the
        // first NaN's index, 0 -- never leave an out-of-range id behind for the next embedding lookup
        x[row] = best_i == 0x7fffffff ? 0 : best_i;
    }
}

}  // namespace sampler
}  // namespace esmdiff
