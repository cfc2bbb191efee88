// Backbone frames and geometric attention.
// Reference call sites: slm/models/net.py:433-441 (structure_coords -> build_affine3d_from_coordinates),
// :337-345 (TransformerStack(..., v_heads, mask_and_zero_frameless=True): block 0 carries geom_attn), :468;
// slm/models/utils.py:136-137 (model.encode -> the VQ-VAE structure encoder, two geometric blocks per
// 16-residue neighbourhood).  The arithmetic is esm==3.0.4's (esm/utils/structure/affine3d.py,
// esm/layers/geom_attention.py GeometricReasoningOriginalImpl) -- not vendored, restated in
// oracle/geom_ref.py, parity unpinned.
//
// All of it is fp32 CUDA-core work on 3-vectors: per (sample, head) an S x S attention whose scores are
//   w_r[h] (R_i q_r) . (R_j k_r) / sqrt3  -  w_d[h] |(R_i q_d + t_i) - (R_j k_d + t_j)| / sqrt3
// -- neither a GEMM (K = 3) nor large (S = 16 in the encoder, S = T in block 0).  It is bound by the reads of
// the rotated key / value vectors (36 B per key and head), so one thread = one head keeps QPT queries in
// registers and walks the keys once for all of them: consecutive threads read consecutive heads (coalesced
// 12-byte vectors), and a key row is read S / QPT times instead of S times.
#pragma once
#include <cuda_bf16.h>
#include <float.h>

#include "ptx.cuh"

namespace esmdiff {
namespace geom {

// esm.utils.structure.affine3d._graham_schmidt (eps inside the square roots), R row-major with columns [e0, e1, e2]
__device__ __forceinline__ void gram_schmidt(const float* xa, const float* xy, float eps, float* R) {
    const float d0 = sqrtf(xa[0] * xa[0] + xa[1] * xa[1] + xa[2] * xa[2] + eps);
    const float e0[3] = {xa[0] / d0, xa[1] / d0, xa[2] / d0};
    const float dot = e0[0] * xy[0] + e0[1] * xy[1] + e0[2] * xy[2];
    float e1[3] = {xy[0] - e0[0] * dot, xy[1] - e0[1] * dot, xy[2] - e0[2] * dot};
    const float d1 = sqrtf(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2] + eps);
    e1[0] /= d1; e1[1] /= d1; e1[2] /= d1;
    const float e2[3] = {e0[1] * e1[2] - e0[2] * e1[1], e0[2] * e1[0] - e0[0] * e1[2], e0[0] * e1[1] - e0[1] * e1[0]};
#pragma unroll
    for (int i = 0; i < 3; ++i) { R[3 * i] = e0[i]; R[3 * i + 1] = e1[i]; R[3 * i + 2] = e2[i]; }
}

// build_affine3d_from_coordinates: coords [B][L][3][3] (N, CA, C; NaN / inf = unknown) ->
//   rot [B*L][9] (row-major R, columns e0 e1 e2), trans [B*L][3] (= CA), mask [B*L] (1 = has a frame).
// Residues without a frame take the frame of the average backbone of the valid ones of their sample
// (identity rotation, zero translation when there is none).  One block per sample.
__global__ void __launch_bounds__(256)
frames_kernel(const float* __restrict__ coords, float* __restrict__ rot, float* __restrict__ trans,
              unsigned char* __restrict__ mask, int L) {
    const int b = blockIdx.x;
    const float* cb = coords + static_cast<long long>(b) * L * 9;
    __shared__ float red[9][8];
    __shared__ int cnt[8];
    __shared__ float avgR[9], avgT[3];
    float s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int n = 0;
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        bool ok = true;
        float v[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            v[k] = cb[i * 9 + k];
            ok = ok && isfinite(v[k]) && v[k] < 1e6f;
        }
        float* R = rot + (static_cast<long long>(b) * L + i) * 9;
        float* t = trans + (static_cast<long long>(b) * L + i) * 3;
        mask[static_cast<long long>(b) * L + i] = ok ? 1 : 0;
        if (ok) {
            ++n;
#pragma unroll
            for (int k = 0; k < 9; ++k) s[k] += v[k];
            // Affine3D.from_graham_schmidt(C, CA, N): x axis = CA - C, xy plane = N - CA, origin CA
            const float xa[3] = {v[3] - v[6], v[4] - v[7], v[5] - v[8]};
            const float xy[3] = {v[0] - v[3], v[1] - v[4], v[2] - v[5]};
            float Rr[9];
            gram_schmidt(xa, xy, 1e-12f, Rr);
#pragma unroll
            for (int k = 0; k < 9; ++k) R[k] = Rr[k];
            t[0] = v[3]; t[1] = v[4]; t[2] = v[5];
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], off);
        if (lane == 0) red[k][warp] = s[k];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) n += __shfl_xor_sync(0xffffffffu, n, off);
    if (lane == 0) cnt[warp] = n;
    __syncthreads();
    if (threadIdx.x == 0) {
        float a[9];
        int tot = 0;
        const int nw = (blockDim.x + 31) >> 5;
        for (int w = 0; w < nw; ++w) tot += cnt[w];
        for (int k = 0; k < 9; ++k) {
            float acc = 0.f;
            for (int w = 0; w < nw; ++w) acc += red[k][w];
            a[k] = acc / (static_cast<float>(tot) + 1e-8f);
        }
        if (tot > 0) {
            const float xa[3] = {a[3] - a[6], a[4] - a[7], a[5] - a[8]};
            const float xy[3] = {a[0] - a[3], a[1] - a[4], a[2] - a[5]};
            float Rr[9];
            gram_schmidt(xa, xy, 1e-12f, Rr);
            for (int k = 0; k < 9; ++k) avgR[k] = Rr[k];
        } else {
            for (int k = 0; k < 9; ++k) avgR[k] = (k % 4 == 0) ? 1.f : 0.f;
        }
        avgT[0] = a[3]; avgT[1] = a[4]; avgT[2] = a[5];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        if (mask[static_cast<long long>(b) * L + i]) continue;
        float* R = rot + (static_cast<long long>(b) * L + i) * 9;
        float* t = trans + (static_cast<long long>(b) * L + i) * 3;
#pragma unroll
        for (int k = 0; k < 9; ++k) R[k] = avgR[k];
        t[0] = avgT[0]; t[1] = avgT[1]; t[2] = avgT[2];
    }
}

__device__ __forceinline__ float ld_f(const float* p) { return *p; }
__device__ __forceinline__ float ld_f(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st_f(float* p, float v) { *p = v; }
__device__ __forceinline__ void st_f(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// proj output p [M][15 H] = [q_rot | k_rot | value | q_dist | k_dist] x (H heads x 3) in the residues' local
// frames -> out fp32, same layout, in the global frame: R v for the first 9 H values, R v + t for the last 6 H.
// frame_idx[m] (or m itself when NULL) selects the frame of row m.  out may alias p when TIn = float.
template <typename TIn>
__global__ void __launch_bounds__(256)
rotate_kernel(const TIn* p, float* out, const float* __restrict__ rot, const float* __restrict__ trans,
              const int* __restrict__ frame_idx, long long M, int H) {
    const long long nvec = 5ll * H;
    const long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (g >= M * nvec) return;
    const long long m = g / nvec;
    const int v = static_cast<int>(g - m * nvec);
    const long long f = frame_idx ? frame_idx[m] : m;
    const float* R = rot + f * 9;
    const TIn* src = p + m * nvec * 3 + v * 3;
    const float x = ld_f(src), y = ld_f(src + 1), z = ld_f(src + 2);
    float o0 = R[0] * x + R[1] * y + R[2] * z;
    float o1 = R[3] * x + R[4] * y + R[5] * z;
    float o2 = R[6] * x + R[7] * y + R[8] * z;
    if (v >= 3 * H) {
        const float* t = trans + f * 3;
        o0 += t[0]; o1 += t[1]; o2 += t[2];
    }
    float* dst = out + m * nvec * 3 + v * 3;
    dst[0] = o0; dst[1] = o1; dst[2] = o2;
}

// Geometric attention over groups of S consecutive rows of the rotated projection r [M][15 H] (M = G S).
//   grid (ceil(S / QPT), G), block = H threads (thread = head), QPT queries per thread.
//   key j of a group is masked when its row has no frame: bias = finfo.min instead of the +1 of the
//   same-sequence mask (esm adds the FLOAT of that mask; a no-op under softmax) -- a group whose keys are ALL
//   frameless averages its values uniformly, exactly as torch's softmax of an all-finfo.min row does.
//   out [M][ldo] (first 3 H columns): softmax-weighted values rotated back into the query's frame (R_i^T),
//   zeroed for frameless queries when zero_frameless (TransformerStack(mask_and_zero_frameless=True)).
template <int QPT, typename TOut>
__global__ void __launch_bounds__(256, 2)
attention_kernel(const float* __restrict__ r, const float* __restrict__ rot, const unsigned char* __restrict__ mask,
                 const int* __restrict__ frame_idx, const float* __restrict__ w_rot, const float* __restrict__ w_dist,
                 TOut* __restrict__ out, int ldo, int S, int H, int zero_frameless) {
    const int h = threadIdx.x;
    const long long row0 = static_cast<long long>(blockIdx.y) * S;
    const int q0 = blockIdx.x * QPT;
    const long long ld = 15ll * H;
    // scores in the log2 domain: softplus of the per-head scales (F.softplus, threshold 20), 1/sqrt3 and log2(e)
    // folded into two per-head factors; the +1 of esm's float same-sequence mask becomes + log2(e)
    const float LOG2E = 1.4426950408889634f;
    const float wr_raw = w_rot[h], wd_raw = w_dist[h];
    const float wr = (wr_raw > 20.f ? wr_raw : log1pf(expf(wr_raw))) * (0.57735026918962576f * LOG2E);
    const float wd = (wd_raw > 20.f ? wd_raw : log1pf(expf(wd_raw))) * (0.57735026918962576f * LOG2E);
    float qr[QPT][3], qd[QPT][3], acc[QPT][3], mx[QPT], l[QPT];
#pragma unroll
    for (int q = 0; q < QPT; ++q) {
        const int i = min(q0 + q, S - 1);
        const float* base = r + (row0 + i) * ld;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            qr[q][c] = base[h * 3 + c];
            qd[q][c] = base[9 * H + h * 3 + c];
            acc[q][c] = 0.f;
        }
        mx[q] = -INFINITY;
        l[q] = 0.f;
    }
    // the key / value vectors of step j + 1 are fetched (L2) while step j is computed
    float kv[9];
    bool keyed_next;
    auto fetch = [&](int j) {
        const float* base = r + (row0 + j) * ld;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            kv[c] = base[3 * H + h * 3 + c];
            kv[3 + c] = base[6 * H + h * 3 + c];
            kv[6 + c] = base[12 * H + h * 3 + c];
        }
        keyed_next = mask[frame_idx ? frame_idx[row0 + j] : row0 + j] != 0;
    };
    fetch(0);
    for (int j = 0; j < S; ++j) {
        const float k0 = kv[0], k1 = kv[1], k2 = kv[2], v0 = kv[3], v1 = kv[4], v2 = kv[5], d0 = kv[6], d1 = kv[7], d2 = kv[8];
        const bool keyed = keyed_next;
        if (j + 1 < S) fetch(j + 1);
#pragma unroll
        for (int q = 0; q < QPT; ++q) {
            const float rt = qr[q][0] * k0 + qr[q][1] * k1 + qr[q][2] * k2;
            const float e0 = qd[q][0] - d0, e1 = qd[q][1] - d1, e2 = qd[q][2] - d2;
            float dt;
            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(dt) : "f"(e0 * e0 + e1 * e1 + e2 * e2));
            // a frameless key: score + finfo.min == finfo.min in fp32, whatever the score
            const float w = keyed ? fmaf(rt, wr, fmaf(-dt, wd, LOG2E)) : -FLT_MAX;
            if (w > mx[q]) {                                 // raise the running maximum (rare after the first keys)
                const float corr = fast_exp2(mx[q] - w);     // first key: exp2(-inf) = 0
                l[q] *= corr;
                acc[q][0] *= corr; acc[q][1] *= corr; acc[q][2] *= corr;
                mx[q] = w;
            }
            const float pe = fast_exp2(w - mx[q]);
            l[q] += pe;
            acc[q][0] = fmaf(pe, v0, acc[q][0]);
            acc[q][1] = fmaf(pe, v1, acc[q][1]);
            acc[q][2] = fmaf(pe, v2, acc[q][2]);
        }
    }
#pragma unroll
    for (int q = 0; q < QPT; ++q) {
        const int i = q0 + q;
        if (i >= S) break;
        const long long fi = frame_idx ? frame_idx[row0 + i] : row0 + i;
        const float* R = rot + fi * 9;
        const float inv = 1.0f / l[q];
        const float a0 = acc[q][0] * inv, a1 = acc[q][1] * inv, a2 = acc[q][2] * inv;
        const bool zero = zero_frameless && !mask[fi];
        TOut* o = out + (row0 + i) * ldo + h * 3;
        // R^T a: component c = sum_j R[j][c] a[j]
        st_f(o, zero ? 0.f : R[0] * a0 + R[3] * a1 + R[6] * a2);
        st_f(o + 1, zero ? 0.f : R[1] * a0 + R[4] * a1 + R[7] * a2);
        st_f(o + 2, zero ? 0.f : R[2] * a0 + R[5] * a1 + R[8] * a2);
    }
}

}  // namespace geom
}  // namespace esmdiff
