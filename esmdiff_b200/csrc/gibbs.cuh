// Iterative (entropy-ordered) unmasking of the structure track: the sampling half of `--mode gibbs`
// (reference slm/sample_esmdiff.py:66-130 -> esm.utils.generation.iterative_sampling_raw, esm==3.0.4,
// not vendored in the reference: restated in oracle/gibbs_ref.py, PARITY UNPINNED).
// Per decoding step, after the same forward as the ddpm path (no time conditioning):
//   gibbs_rows_kernel    one CTA per still-masked token row -- esm.utils.sampling.sample_logits +
//                        _compute_track_metadata in one pass over the 16 KiB logits row instead of a
//                        sort, a cumsum, two softmaxes, a multinomial and an entropy over [M, 4101]:
//                          top-p filter on the raw logits (nucleus found by bisection on the
//                          probability threshold: no sort), special ids excluded, candidate id =
//                          argmax softmax(l / temperature) / Exp(1)  (what torch.multinomial(p, 1) is),
//                          entropy of the filtered distribution at temperature 1
//   gibbs_commit_kernel  one CTA per sample -- _get_iterative_sampling_mask_for_prompt_and_step:
//                        the k masked positions of lowest entropy take their candidate id
#pragma once
#include "sampler.cuh"

namespace esmdiff {
namespace gibbs {

using sampler::THREADS;
using sampler::GROUPS;
using sampler::block_max;
using sampler::block_sum;

// noise == null: Exp(1) = -log(1 - u) from the library's Philox stream (u in [0, 1)); otherwise the
// caller's exponentials ([M][V], e.g. torch's `empty_like(p).exponential_()`: bit-identical draws to
// torch.multinomial under the same generator state).
__global__ void __launch_bounds__(THREADS)
gibbs_rows_kernel(const float* __restrict__ logits, long long ld, const float* __restrict__ noise,
                  const long long* __restrict__ x, int* __restrict__ cand, float* __restrict__ entropy, int V,
                  int n_valid, int mask_index, float inv_temperature, float top_p, unsigned long long seed,
                  uint32_t step, uint32_t row_offset) {
    __shared__ float red[THREADS / 32];
    __shared__ int red_i[THREADS / 32];
    const int row = blockIdx.x;
    if (x[row] != mask_index) {
        if (threadIdx.x == 0) {
            cand[row] = -1;
            entropy[row] = INFINITY;
        }
        return;
    }
    const float* lr = logits + row * ld;
    float v[GROUPS * 4];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < GROUPS; ++k)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = 4 * (threadIdx.x + k * THREADS) + e;
            const float a = i < V ? lr[i] : -INFINITY;
            v[4 * k + e] = a;
            mx = fmaxf(mx, a);
        }
    mx = block_max(mx, red);
    // probabilities over the WHOLE vocabulary (top_p_logits runs before the special ids are masked)
    float pr[GROUPS * 4];
    float z = 0.f;
#pragma unroll
    for (int k = 0; k < GROUPS * 4; ++k) {
        pr[k] = expf(v[k] - mx);                       // exp(-inf) = 0 for the padding
        z += pr[k];
    }
    z = block_sum(z, red);
    // nucleus: token i stays iff the mass of the tokens at least as likely as i is <= top_p (the sorted
    // cumulative sum of the reference), the most likely token always stays.  Smallest kept
    // probability by bisection on the bit pattern (monotone for non-negative floats).
    float thr = 0.f;
    if (top_p < 1.0f) {
        const float budget = top_p * z;
        uint32_t lo = 0u, hi = __float_as_uint(1.0f);  // pr <= 1; mass(>= hi) <= budget or hi is the maximum itself
#pragma unroll 1
        for (int it = 0; it < 31 && lo + 1 < hi; ++it) {
            const uint32_t mid = lo + ((hi - lo) >> 1);
            const float t = __uint_as_float(mid);
            float m = 0.f;
#pragma unroll
            for (int k = 0; k < GROUPS * 4; ++k) m += pr[k] >= t ? pr[k] : 0.f;
            m = block_sum(m, red);
            if (m <= budget) hi = mid; else lo = mid;
        }
        thr = __uint_as_float(hi);
    }
    // filtered distribution at temperature 1: entropy; at the sampling temperature: the race
    float zk = 0.f;
#pragma unroll
    for (int k = 0; k < GROUPS; ++k)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = 4 * (threadIdx.x + k * THREADS) + e;
            const bool keep = i < n_valid && (pr[4 * k + e] >= thr || v[4 * k + e] == mx);
            if (!keep) pr[4 * k + e] = 0.f;
            zk += pr[4 * k + e];
        }
    zk = block_sum(zk, red);
    if (zk == 0.f) {
        // every token of the nucleus is a special id: the reference's softmax over its finfo.min / -inf
        // row is then uniform over the valid ids
#pragma unroll
        for (int k = 0; k < GROUPS; ++k)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int i = 4 * (threadIdx.x + k * THREADS) + e;
                pr[4 * k + e] = i < n_valid ? 1.f : 0.f;
                if (i < n_valid) v[4 * k + e] = mx;
            }
        zk = static_cast<float>(n_valid);
    }
    const float log_zk = logf(zk);
    float h = 0.f;
    float best = -INFINITY;
    int best_i = 0x7fffffff;
#pragma unroll
    for (int k = 0; k < GROUPS; ++k) {
        const int i0 = 4 * (threadIdx.x + k * THREADS);
        if (i0 < n_valid) {
            float uu[4] = {0.f, 0.f, 0.f, 0.f};
            if (noise == nullptr) sampler::philox_uniform4(seed, step, row + row_offset, i0 >> 2, uu);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int i = i0 + e;
                const float p = pr[4 * k + e];
                if (i < n_valid && p > 0.f) {
                    const float lp = (v[4 * k + e] - mx) - log_zk;          // log of the filtered probability
                    h -= (p / zk) * lp;
                    const float q = noise ? noise[row * ld + i] : -log1pf(-uu[e]);
                    // softmax(l / temperature) up to its normalisation (the same for the whole row)
                    const float score = expf((v[4 * k + e] - mx) * inv_temperature) / q;
                    if (score > best) { best = score; best_i = i; }          // ascending i: first max wins
                }
            }
        }
    }
    h = block_sum(h, red);
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
        cand[row] = best_i == 0x7fffffff ? 0 : best_i;
        entropy[row] = h;
    }
}

// The k masked positions of lowest entropy of every sample take their candidate (ties: lower position
// first).  Rank by counting: T^2 compares per sample, T <= a few thousand.
__global__ void __launch_bounds__(256)
gibbs_commit_kernel(long long* __restrict__ x, const int* __restrict__ cand, const float* __restrict__ entropy, int T,
                    int k) {
    extern __shared__ float sh_e[];                    // [T]
    const long long base = static_cast<long long>(blockIdx.x) * T;
    for (int i = threadIdx.x; i < T; i += blockDim.x) sh_e[i] = cand[base + i] >= 0 ? entropy[base + i] : INFINITY;
    __syncthreads();
    for (int i = threadIdx.x; i < T; i += blockDim.x) {
        const int c = cand[base + i];
        if (c < 0) continue;
        const float e = sh_e[i];
        int rank = 0;
        for (int j = 0; j < T; ++j) {
            const float f = sh_e[j];
            rank += (f < e || (f == e && j < i)) ? 1 : 0;
        }
        if (rank < k) x[base + i] = c;
    }
}

}  // namespace gibbs
}  // namespace esmdiff
