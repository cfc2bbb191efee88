// VQ-VAE structure ENCODER kernels (the inpainting front end).
// Reference call sites: slm/models/utils.py:105-146 (protseq_to_data: mask_ids -> coordinates[idx] = inf,
// model.encode(ESMProtein(sequence, coordinates)).structure), slm/sample_esmdiff.py:166-175, 197-209.
// The arithmetic is esm==3.0.4's StructureTokenEncoder (esm/models/vqvae.py; not vendored, restated in
// oracle/vqvae_enc_ref.py, parity unpinned): 16 nearest neighbours by CA distance, relative-position
// embedding, two geometric-attention + SwiGLU blocks over every neighbourhood at d = 1024, the query node's
// row through the final LayerNorm and pre_vq_proj, nearest of 4096 codes.
//
// The result is an INDEX (a code per residue), so the whole encoder computes in fp32 on the CUDA cores: bf16
// tensor-core operands would flip the argmin of close codes.  It runs once per target (L = 256: 4096
// neighbourhood rows, 0.18 TFLOP) against 26 forwards of 72 TFLOP for the samples of that target.
#pragma once
#include <cuda_runtime.h>
#include <float.h>

namespace esmdiff {
namespace enc {

// knn_graph: for residue i of sample b the E nearest residues in the order of
//   d(i, j) = |CA_i - CA_j|                      when both have a frame
//           = 100 |i - j| + 1e6                  otherwise (sequence distance sorts after every structural one)
// ascending, ties -> lower j.  One block per (i, b); round e picks the smallest (d, j) above the previous pick.
__global__ void __launch_bounds__(128)
knn_kernel(const float* __restrict__ coords, const unsigned char* __restrict__ mask, int* __restrict__ edges,
           int L, int E) {
    const int i = blockIdx.x, b = blockIdx.y;
    const float* cb = coords + static_cast<long long>(b) * L * 9;
    const unsigned char* mb = mask + static_cast<long long>(b) * L;
    const bool mi = mb[i];
    const float xi = mi ? cb[i * 9 + 3] : 0.f, yi = mi ? cb[i * 9 + 4] : 0.f, zi = mi ? cb[i * 9 + 5] : 0.f;
    __shared__ float sd[4];
    __shared__ int sj[4];
    float last_d = -1.f;
    int last_j = -1;
    for (int e = 0; e < E; ++e) {
        float best_d = INFINITY;
        int best_j = 0x7fffffff;
        for (int j = threadIdx.x; j < L; j += blockDim.x) {
            float d;
            if (mi && mb[j]) {
                const float dx = xi - cb[j * 9 + 3], dy = yi - cb[j * 9 + 4], dz = zi - cb[j * 9 + 5];
                d = sqrtf(dx * dx + dy * dy + dz * dz);
            } else {
                d = fabsf(static_cast<float>(i - j)) * 1e2f + 1e6f;
            }
            const bool after = d > last_d || (d == last_d && j > last_j);
            if (after && (d < best_d || (d == best_d && j < best_j))) { best_d = d; best_j = j; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, best_d, off);
            const int oj = __shfl_xor_sync(0xffffffffu, best_j, off);
            if (od < best_d || (od == best_d && oj < best_j)) { best_d = od; best_j = oj; }
        }
        if ((threadIdx.x & 31) == 0) { sd[threadIdx.x >> 5] = best_d; sj[threadIdx.x >> 5] = best_j; }
        __syncthreads();
        best_d = sd[0]; best_j = sj[0];
        for (int w = 1; w < (blockDim.x >> 5); ++w)
            if (sd[w] < best_d || (sd[w] == best_d && sj[w] < best_j)) { best_d = sd[w]; best_j = sj[w]; }
        __syncthreads();
        last_d = best_d; last_j = best_j;
        if (threadIdx.x == 0) edges[(static_cast<long long>(b) * L + i) * E + e] = best_j;
    }
}

// Row m = (b, i, e) of the neighbourhood stream: z[m] = table[clamp(res[nbr] - res[i], -bins, bins) + bins + 1]
// (RelativePositionEmbedding), frame_idx[m] = b L + nbr.  residue_index NULL -> the positions themselves.
__global__ void __launch_bounds__(256)
relpos_gather_kernel(const int* __restrict__ edges, const long long* __restrict__ residue_index,
                     const float* __restrict__ table, float* __restrict__ z, int* __restrict__ frame_idx,
                     long long M, int L, int E, int D, int bins) {
    const long long m = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (m >= M) return;
    const long long bi = m / E;                      // b L + i
    const long long b = bi / L;
    const int nbr = edges[m], self = edges[bi * E];
    const long long rn = residue_index ? residue_index[b * L + nbr] : nbr;
    const long long rs = residue_index ? residue_index[b * L + self] : self;
    long long diff = rn - rs;
    diff = diff < -bins ? -bins : (diff > bins ? bins : diff);
    const float4* src = reinterpret_cast<const float4*>(table + (diff + bins + 1) * D);
    float4* dst = reinterpret_cast<float4*>(z + m * D);
    for (int k = lane; k < D / 4; k += 32) dst[k] = src[k];
    if (lane == 0) frame_idx[m] = static_cast<int>(b * L + nbr);
}

// y[m] = LayerNorm(x[m * in_stride]) * w (+ b), eps 1e-5, fp32 -> fp32; one warp per row, D % 4 == 0.
// zero_mask (may be NULL): rows with zero_mask[m] == 0 are written as zeros (z.masked_fill(~affine_mask, 0)).
__global__ void __launch_bounds__(256)
layernorm_f32_kernel(const float* __restrict__ x, long long in_stride, const float* __restrict__ w,
                     const float* __restrict__ bias, float* __restrict__ y, long long M, int D,
                     const unsigned char* __restrict__ zero_mask) {
    const long long m = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (m >= M) return;
    const float4* src = reinterpret_cast<const float4*>(x + m * in_stride);
    float4* dst = reinterpret_cast<float4*>(y + m * D);
    if (zero_mask && !zero_mask[m]) {
        for (int k = lane; k < D / 4; k += 32) dst[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    float s = 0.f;
    for (int k = lane; k < D / 4; k += 32) { const float4 v = src[k]; s += (v.x + v.y) + (v.z + v.w); }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    const float mean = s / static_cast<float>(D);
    float q = 0.f;
    for (int k = lane; k < D / 4; k += 32) {
        const float4 v = src[k];
        const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
        q += (a * a + b * b) + (c * c + d * d);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) q += __shfl_xor_sync(0xffffffffu, q, off);
    const float rstd = rsqrtf(q / static_cast<float>(D) + 1e-5f);
    for (int k = lane; k < D / 4; k += 32) {
        const float4 v = src[k];
        const float4 g = reinterpret_cast<const float4*>(w)[k];
        float4 o = make_float4((v.x - mean) * rstd * g.x, (v.y - mean) * rstd * g.y, (v.z - mean) * rstd * g.z,
                               (v.w - mean) * rstd * g.w);
        if (bias) {
            const float4 bb = reinterpret_cast<const float4*>(bias)[k];
            o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
        }
        dst[k] = o;
    }
}

// fp32 SGEMM  C[M][N] = A[M][K] W[N][K]^T, K % 16 == 0, both operands K-major (nn.Linear layout).
//   128 x 128 tile, BK = 16, 256 threads, 8 x 8 per thread as 2 x 2 blocks of 4 x 4 (float4 shared-memory
//   reads without bank conflicts), next tile's global loads in flight during the FMAs of the current one.
//   EPI 0: C = acc (+ bias[n]);   EPI 1: C += acc * scale (residual in place).
template <int EPI>
__global__ void __launch_bounds__(256)
sgemm_tn_kernel(const float* __restrict__ A, const float* __restrict__ W, float* C, const float* __restrict__ bias,
                int M, int N, int K, long long lda, long long ldc, float scale) {
    __shared__ __align__(16) float As[2][16][128 + 4];
    __shared__ __align__(16) float Ws[2][16][128 + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * 128, n0 = blockIdx.x * 128;
    // loader: thread -> (row = tid / 4 [+64], k4 = tid % 4): one float4 along K per operand row
    const int lr = tid >> 2, lk = (tid & 3) * 4;
    const int tx = tid & 15, ty = tid >> 4;          // compute: rows ty*4 (+64), cols tx*4 (+64)
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    float4 ra[2], rw[2];
    auto gload = [&](int k0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int ar = m0 + lr + h * 64, wr = n0 + lr + h * 64;
            ra[h] = ar < M ? *reinterpret_cast<const float4*>(A + static_cast<long long>(ar) * lda + k0 + lk)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
            rw[h] = wr < N ? *reinterpret_cast<const float4*>(W + static_cast<long long>(wr) * K + k0 + lk)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = lr + h * 64;
            As[buf][lk][r] = ra[h].x; As[buf][lk + 1][r] = ra[h].y; As[buf][lk + 2][r] = ra[h].z; As[buf][lk + 3][r] = ra[h].w;
            Ws[buf][lk][r] = rw[h].x; Ws[buf][lk + 1][r] = rw[h].y; Ws[buf][lk + 2][r] = rw[h].z; Ws[buf][lk + 3][r] = rw[h].w;
        }
    };
    gload(0);
    sstore(0);
    __syncthreads();
    const int nk = K / 16;
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) gload((kt + 1) * 16);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4 + 64]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4 + 64]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            sstore(buf ^ 1);
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ty * 4 + (i & 3) + (i >> 2) * 64;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + tx * 4 + (j & 3) + (j >> 2) * 64;
            if (n >= N) continue;
            float* c = C + static_cast<long long>(m) * ldc + n;
            if (EPI == 0) *c = acc[i][j] + (bias ? bias[n] : 0.f);
            else *c = *c + acc[i][j] * scale;
        }
    }
}

// h[m][f] = silu(u[m][f]) * u[m][F + f]   (esm SwiGLU: chunk(2) of the W1 output, gate first)
__global__ void __launch_bounds__(256)
swiglu_kernel(const float* __restrict__ u, float* __restrict__ h, long long M, int F) {
    const long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (g >= M * F) return;
    const long long m = g / F;
    const int f = static_cast<int>(g - m * F);
    const float a = u[m * 2 * F + f], b = u[m * 2 * F + F + f];
    h[g] = a / (1.0f + expf(-a)) * b;
}

// e2[n] = sum_k e[n][k]^2 (once per weight load); one warp per code
__global__ void __launch_bounds__(256)
rowsumsq_kernel(const float* __restrict__ e, float* __restrict__ e2, int N, int K) {
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (n >= N) return;
    float s = 0.f;
    for (int k = lane; k < K; k += 32) { const float v = e[static_cast<long long>(n) * K + k]; s += v * v; }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) e2[n] = s;
}

// EMACodebook: code[m] = argmin_n (|z_m|^2 + |e_n|^2) - 2 z_m . e_n, first minimum on ties (torch.argmin).
// dots [M][N] = z e^T from the SGEMM; one warp per row.
__global__ void __launch_bounds__(256)
codebook_argmin_kernel(const float* __restrict__ z, const float* __restrict__ dots, const float* __restrict__ e2,
                       long long* __restrict__ codes, long long M, int N, int K) {
    const long long m = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (m >= M) return;
    float z2 = 0.f;
    for (int k = lane; k < K; k += 32) { const float v = z[m * K + k]; z2 += v * v; }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) z2 += __shfl_xor_sync(0xffffffffu, z2, off);
    float best = INFINITY;
    int bi = 0x7fffffff;
    for (int n = lane; n < N; n += 32) {
        const float d = (z2 + e2[n]) - 2.0f * dots[m * N + n];
        if (d < best) { best = d; bi = n; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) codes[m] = bi == 0x7fffffff ? 0 : bi;
}

}  // namespace enc
}  // namespace esmdiff
