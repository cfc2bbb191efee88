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
constexpr int MAX_PER_THREAD = 17;           // ceil(4101 / 256); V <= 4352
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
__device__ __forceinline__ float philox_uniform(unsigned long long seed, uint32_t step, uint32_t row,
                                                uint32_t col) {
    const uint4 r = philox4x32_10(make_uint4(col >> 2, row, step, 0u),
                                  make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
    const uint32_t w = (col & 3) == 0 ? r.x : (col & 3) == 1 ? r.y : (col & 3) == 2 ? r.z : r.w;
    return static_cast<float>(w >> 8) * (1.0f / 16777216.0f);       // [0, 1)
}

// MODE 0: ddpm update (race argmax with uniforms)   MODE 1: noise removal (argmax of log p)
// MODE 2: write log p(x0) to out_logp (logits_parameterization only; all rows)
template <int MODE>
__global__ void __launch_bounds__(THREADS)
sample_rows_kernel(const float* __restrict__ logits, long long ld, const float* __restrict__ u,
                   long long* __restrict__ x, float* __restrict__ out_logp, int M, int V,
                   int mask_index, float mc_t, float mc_s, unsigned long long seed, uint32_t step) {
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
    float v[MAX_PER_THREAD];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < MAX_PER_THREAD; ++k) {
        const int i = threadIdx.x + k * THREADS;
        float a = -INFINITY;
        if (i < V) {
            a = lr[i];
            if (i == mask_index) a += NEG_INF;
        }
        v[k] = a;
        mx = fmaxf(mx, a);
    }
    mx = block_max(mx, red);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < MAX_PER_THREAD; ++k) s += expf(v[k] - mx);      // exp(-inf) = 0 for padding
    s = block_sum(s, red);
    const float lse = logf(s) + mx;

    if constexpr (MODE == 2) {
        float* o = out_logp + row * ld;
#pragma unroll
        for (int k = 0; k < MAX_PER_THREAD; ++k) {
            const int i = threadIdx.x + k * THREADS;
            if (i < V) o[i] = v[k] - lse;
        }
        return;
    }

    float best = -INFINITY;
    int best_i = 0x7fffffff;
    const float dmc = mc_t - mc_s;
#pragma unroll
    for (int k = 0; k < MAX_PER_THREAD; ++k) {
        const int i = threadIdx.x + k * THREADS;
        if (i < V) {
            float score;
            if constexpr (MODE == 0) {
                float q = expf(v[k] - lse) * dmc;
                if (i == mask_index) q = mc_s;
                const float uu = u ? u[row * ld + i] : philox_uniform(seed, step, row, i);
                const float g = 1e-10f - logf(uu + 1e-10f);
                score = q / g;
            } else {
                score = v[k] - lse;
            }
            if (score > best) { best = score; best_i = i; }      // ascending i: first max wins
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
        x[row] = best_i;
    }
}

}  // namespace sampler
}  // namespace esmdiff
