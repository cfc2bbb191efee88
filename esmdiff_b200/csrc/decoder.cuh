// Tail of the VQ-VAE structure-token decoder (reference call sites
// slm/sample_esmdiff.py:41-61, 225-231 -> esm ESM3.decode -> StructureTokenDecoder.decode ->
// Dim6RotStructureHead, ProteinChain.infer_oxygen -- esm==3.0.4, not vendored: restated from the
// published package, parity unpinned).
// The decoder trunk (token embedding, 30 pre-LN blocks at d = 1280 with q/k-LayerNorm + RoPE and a
// SwiGLU FFN, final LayerNorm) and its two regression heads run on the same tcgen05 GEMM /
// attention kernels as the sampling network; these row kernels turn the heads' outputs into
// coordinates.
#pragma once
#include "ptx.cuh"

namespace esmdiff {
namespace dec {

// Single-table embedding (StructureTokenDecoder.embed) + the bf16 copy and per-128-column partial
// statistics the first LayerNorm-folded GEMM needs (same contract as ew::embed_kernel).
__global__ void __launch_bounds__(256)
embed_tokens_kernel(const long long* __restrict__ tok, const float* __restrict__ table, float* __restrict__ x,
                    __nv_bfloat16* __restrict__ xb, float2* __restrict__ stats, int M, int D, int vocab,
                    int* __restrict__ err) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    long long t = tok[row];
    if (t < 0 || t >= vocab) {
        if (lane == 0) atomicExch(err, 1);
        t = 0;
    }
    const float4* a = reinterpret_cast<const float4*>(table + t * D);
    float4* o = reinterpret_cast<float4*>(x + static_cast<long long>(row) * D);
    for (int i = lane; i < D / 4; i += 32) {
        const float4 r = a[i];
        o[i] = r;
        if (xb != nullptr) {
            reinterpret_cast<uint2*>(xb + static_cast<long long>(row) * D)[i] =
                make_uint2(pack_bf16x2(r.x, r.y), pack_bf16x2(r.z, r.w));
            float s = (r.x + r.y) + (r.z + r.w);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            const float mean = s * (1.0f / 128.0f);
            const float a0 = r.x - mean, a1 = r.y - mean, a2 = r.z - mean, a3 = r.w - mean;
            float m2 = (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) m2 += __shfl_xor_sync(0xffffffffu, m2, off);
            if (lane == 0) stats[static_cast<long long>(row) * (D / 128) + i / 32] = make_float2(mean, m2);
        }
    }
}

// Gram-Schmidt frame from (x_axis, xy_plane) as esm.utils.structure.affine3d._graham_schmidt
// (eps inside the square roots), columns [e0, e1, e2].
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

// Dim6RotStructureHead.forward after its projection (esm/layers/structure_proj.py): per token the
// 23 outputs split [trans 3 | x 3 | y 3 | angles 14 (unused: predict_torsion_angles=False)];
//   trans *= 10;  x /= |x| + 1e-5;  y /= |y| + 1e-5
//   frame = from_graham_schmidt(neg_x_axis = x + trans, origin = trans, xy_plane = y + trans)
//         = rotation _graham_schmidt(-x, y, 1e-12) with translation trans (composed with the identity)
//   bb[a] = R * BB_COORDINATES[a] + trans  for a in (N, CA, C)
// One thread per token; bb_out [M][3][3] fp32.
__global__ void backbone_frames_kernel(const float* __restrict__ proj, int ld, float* __restrict__ bb_out, int M) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const float* p = proj + static_cast<long long>(m) * ld;
    const float t[3] = {p[0] * 10.0f, p[1] * 10.0f, p[2] * 10.0f};
    float x[3] = {p[3], p[4], p[5]}, y[3] = {p[6], p[7], p[8]};
    const float nx = sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]) + 1e-5f;
    const float ny = sqrtf(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]) + 1e-5f;
    float xa[3], xy[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        x[i] /= nx; y[i] /= ny;
        xa[i] = t[i] - (x[i] + t[i]);          // origin - neg_x_axis, in the reference's order of operations
        xy[i] = (y[i] + t[i]) - t[i];          // xy_plane - origin
    }
    float R[9];
    gram_schmidt(xa, xy, 1e-12f, R);
    const float local[3][3] = {{0.5256f, 1.3612f, 0.0f}, {0.0f, 0.0f, 0.0f}, {-1.5251f, 0.0f, 0.0f}};   // esm BB_COORDINATES
    float* o = bb_out + static_cast<long long>(m) * 9;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int i = 0; i < 3; ++i)
            o[3 * a + i] = R[3 * i] * local[a][0] + R[3 * i + 1] * local[a][1] + R[3 * i + 2] * local[a][2] + t[i];
}

// ProteinChain.infer_oxygen (esm/utils/structure/protein_chain.py): O of residue i from the frame
// from_graham_schmidt(CA_i, C_i, N_{i+1}) applied to the fixed vector (0.6240, -1.0613, 0.0103);
// the last residue of a chain has no successor -> NaN (the atom is then left out of the PDB).
// bb [B][T][3][3] includes the BOS/EOS positions; residues are positions 1 .. T-2.  o_out [B][T][3].
__global__ void infer_oxygen_kernel(const float* __restrict__ bb, float* __restrict__ o_out, int B, int T) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= B * T) return;
    const int t = m % T;
    float* o = o_out + static_cast<long long>(m) * 3;
    if (t < 1 || t >= T - 2) {                 // BOS, EOS, or the last residue
        o[0] = o[1] = o[2] = nanf("");
        return;
    }
    const float* ca = bb + static_cast<long long>(m) * 9 + 3;
    const float* cc = ca + 3;
    const float* nn = bb + static_cast<long long>(m + 1) * 9;
    float xa[3], xy[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { xa[i] = cc[i] - ca[i]; xy[i] = nn[i] - cc[i]; }
    float R[9];
    gram_schmidt(xa, xy, 1e-10f, R);           // Affine3D.from_graham_schmidt passes its own eps = 1e-10
    const float v[3] = {0.6240f, -1.0613f, 0.0103f};
#pragma unroll
    for (int i = 0; i < 3; ++i) o[i] = R[3 * i] * v[0] + R[3 * i + 1] * v[1] + R[3 * i + 2] * v[2] + cc[i];
}

// pLDDT = CategoricalMixture(logits, bins).mean() (esm/utils/misc or structure heads): softmax over
// the bins times the bin centres (k + 0.5) / bins.  One warp per token.
__global__ void __launch_bounds__(256)
plddt_mean_kernel(const float* __restrict__ logits, int ld, int bins, float* __restrict__ out, int M) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const float* l = logits + static_cast<long long>(row) * ld;
    float mx = -INFINITY;
    for (int k = lane; k < bins; k += 32) mx = fmaxf(mx, l[k]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    float s = 0.f, w = 0.f;
    for (int k = lane; k < bins; k += 32) {
        const float e = expf(l[k] - mx);
        s += e;
        w += e * ((static_cast<float>(k) + 0.5f) / static_cast<float>(bins));
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, off);
        w += __shfl_xor_sync(0xffffffffu, w, off);
    }
    if (lane == 0) out[row] = w / s;
}

}  // namespace dec
}  // namespace esmdiff
