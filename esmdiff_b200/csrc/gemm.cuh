// Persistent warp-specialised tcgen05 GEMM for sm_100a on CTA PAIRS:  C[M,N] = A[M,K] * W[N,K]^T
//   A, W : bf16, K-major (activations row-major, nn.Linear weights as stored)
//   One cluster of two CTAs (one TPC) owns a 256x256 output tile: tcgen05.mma.cta_group::2 with
//   M = 256 (128 rows per CTA) and N = 256; each CTA TMA-loads its 128 rows of A and ITS HALF of
//   the W tile, so per output element only half the shared-memory fill traffic of a single-CTA
//   128x256 tile is needed (shared-memory bandwidth, TMA writes + MMA operand reads, is what
//   bounds the single-CTA form at ~2/3 of the tensor peak).  fp32 accumulators live in TMEM: two
//   stages of 256 columns, so the epilogue of tile i overlaps the main loop of tile i+1.
//   Shared-memory ring: 6 stages x (16 KiB A + 16 KiB W-half) per CTA, SWIZZLE_128B.
// Replaces the cuBLAS sgemm calls behind nn.Linear on the reference path (SURVEY.md 2.2 k3, k7,
// k8, k10, k13) with the element-wise tails fused into the epilogue:
//   EPI_STORE_BF16      out = acc                      TMA store            (QKV projection)
//   EPI_RESID_F32       x  += acc / scale              TMA load of x, add, TMA store (out_proj, FFN W2)
//   EPI_SWIGLU_BF16     out = silu(gate) * up          TMA store            (FFN W1; gate/up rows
//                                                                            interleaved per 128 offline)
//   EPI_BIAS_GELU_F32   out = gelu(acc + bias)         coalesced st.global  (RegressionHead Linear+GELU)
//   EPI_BIAS_F32        out = acc + bias  (ragged N = 4101, unaligned rows) (RegressionHead output)
// LayerNorm folded through the GEMM (the pre-LN of both residual branches):
//   LN(x) W^T = rstd * (x (gamma.W)^T - mean * colsum(gamma.W)) + beta W^T
// so the A operand is the bf16 copy of the RAW residual stream, gamma is folded into W offline
// (ew::fold_layernorm_weight_kernel), and the per-row mean/rstd enter in the epilogue:
//   EPI_RESID_F32_LN    EPI_RESID_F32 + bf16 copy of the new x + per-row partial statistics
//                       (mean, M2 over each 128-column span) for the NEXT LayerNorm
//   EPI_STORE_BF16_LN   out = rstd * (acc - mean * c[n]) + b[n]            (QKV projection)
//   EPI_SWIGLU_BF16_LN  the same on gate and up, then silu(gate) * up      (FFN W1)
//   EPI_QKV_ROPE_LN     EPI_STORE_BF16_LN for the whole QKV projection with q_ln / k_ln and the
//                       rotary embedding folded in (SURVEY.md 2.2 k4 + k5; esm MultiHeadAttention:
//                       q_ln, k_ln = LayerNorm over the FULL width, weight only, then rotate-half
//                       RoPE per 64-wide head).  The q and k rows of W are CENTRED offline (column
//                       means over the 1536 q / k output features removed: q - mean(q) is linear in
//                       the input), so the epilogue holds y = q - mean(q) in fp32.  It writes
//                       rope(gamma * y) as bf16 -- with the two halves (d, d + 32) of every rotation
//                       stored ADJACENT inside their head, a permutation applied to q and k alike
//                       (folded into the weight rows) that no q.k dot product can see -- and the
//                       per-row sum of y^2 over its 128 columns;
//                       the 1/sqrt(mean y^2 + eps) factor is a per-row scalar that commutes with
//                       the rotation and is applied inside the attention kernel (query rows: in
//                       the softmax scale; key rows: on the shared-memory K tile).  Statistics
//                       come from the fp32 accumulators -- q and k are rounded to bf16 once.
//                       This removes the stand-alone q/k-LayerNorm + RoPE kernel (5.6 % of a
//                       step, 12.3 KB of HBM traffic per token row and block).
// This removes the stand-alone LayerNorm kernel of every block (4 B/element re-read of x and a
// launch) at the cost of a 2 B/element extra store in the residual epilogue.
#pragma once
#include "ptx.cuh"

namespace esmdiff {
namespace gemm {

constexpr int BM = 256;                   // rows per CTA pair
constexpr int BM_CTA = 128;               // rows per CTA (= TMEM lanes)
constexpr int BN_MAX = 256;               // tile width is a template parameter: 256, or 192 for the
                                          // N = 1536 GEMMs (6 -> 8 column tiles: 512 tiles over 74
                                          // CTA pairs = 6.92 waves instead of 5.19 -> 6)
constexpr int BK = 64;                    // 64 bf16 = one 128-byte swizzle row
constexpr int MAX_STAGES = 8;
constexpr int A_BYTES = BM_CTA * BK * 2;  // 16 KiB
constexpr int TMEM_COLS = 512;            // two fp32 accumulator stages, 256 columns apart
constexpr int ACC_STRIDE = 256;
constexpr int EPI_WARPS = 8;              // warp w < 8: TMEM lane quarter w % 4, column half w / 4
constexpr int BOX_BYTES = 4096;           // one 32-row x 128-byte staging box
constexpr int THREADS = 384;              // w0-7 epilogue, w8 TMA, w9 MMA (leader CTA), w10 TMEM alloc, w11 idle
// The single-thread roles sit in the HIGHEST warp slots on purpose: the warp scheduler of an SM
// sub-partition prefers the highest warp id among eligible warps (B300_MICROARCH.md, measured), so
// with the issuers in warps 0/1 two compute-heavy epilogue warps on the same sub-partition outrank
// the one thread that feeds the tensor pipe (ncu r1n: out_proj tensor pipe 56 % under the residual
// epilogue's ~1600 instructions per tile and warp, 94 % under the light SwiGLU epilogue).
constexpr int WARP_TMA = 8, WARP_MMA = 9, WARP_ALLOC = 10;

enum Epilogue {
    EPI_STORE_BF16 = 0,
    EPI_RESID_F32 = 1,
    EPI_SWIGLU_BF16 = 2,
    EPI_BIAS_GELU_F32 = 3,
    EPI_BIAS_F32 = 4,
    EPI_STORE_BF16_LN = 5,
    EPI_RESID_F32_LN = 6,
    EPI_SWIGLU_BF16_LN = 7,
    EPI_QKV_ROPE_LN = 8,
};
constexpr int LN_SPAN = 128;              // columns per partial LayerNorm statistic (= epilogue warp width)

// The residual epilogue double-buffers its x boxes (TMA load -> add -> TMA store), paid for with
// one pipeline stage.
template <int EPI, int BN> struct Cfg {
    static_assert(BN == 256 || BN == 192, "tile widths instantiated");
    static_assert(BN == 256 || EPI == EPI_RESID_F32 || EPI == EPI_RESID_F32_LN,
                  "BN = 192 (N = 1536 as 8 column tiles: finer wave quantisation) is instantiated for the residual epilogues");
    static constexpr int BN_CTA = BN / 2;                 // W rows each CTA loads
    static constexpr int B_BYTES = BN_CTA * BK * 2;       // 16 / 12 KiB
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BOXES = (EPI == EPI_RESID_F32 || EPI == EPI_RESID_F32_LN) ? 0 : 1;    // the residual epilogue stages nothing
    // EPI_QKV_ROPE_LN: per epilogue warp, the three per-column vectors (colsum, bias, q/k gamma) of its
    // 128 columns are fetched once per tile while the main loop runs (ncu r2c: as direct global loads
    // right before their use they cost the epilogue an exposed L2 round trip per 16 columns -- 7.2k of
    // 13.8k stall samples -- and the tensor pipe 15 points).  They are held as one float4 per lane and
    // broadcast with shuffles; the first form staged them in shared memory (-DESMDIFF_VEC_SMEM): 12 KiB
    // that cost the TMA ring its sixth stage, and 96 broadcast LDS.128 per thread and tile next to a main
    // loop that already saturates the shared-memory port -- 232.4 vs 226.3 us at M = 25 800
    // (profiles/r4e_qkv_vectors_by_shuffle.txt)
#ifndef ESMDIFF_VEC_SMEM                                     // the three vectors live in registers, one float4 per lane (below)
    static constexpr int VEC_BYTES = 0;
#else
    static constexpr int VEC_BYTES = EPI == EPI_QKV_ROPE_LN ? 3 * 128 * 4 : 0;
#endif
    static constexpr int STG_BYTES = EPI_WARPS * (BOXES * BOX_BYTES + VEC_BYTES);
    static constexpr int STAGES_FIT = (227 * 1024 - 1024 - 512 - STG_BYTES) / STAGE_BYTES;
#ifdef ESMDIFF_STAGES_CAP                                   // experiment builds: cap the TMA ring depth (tools/kbench_qkv_variants.py)
    static constexpr int STAGES_MAX_ = ESMDIFF_STAGES_CAP < MAX_STAGES ? ESMDIFF_STAGES_CAP : MAX_STAGES;
#else
    static constexpr int STAGES_MAX_ = MAX_STAGES;
#endif
    static constexpr int STAGES = STAGES_FIT < STAGES_MAX_ ? STAGES_FIT : STAGES_MAX_;
    static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + STG_BYTES + 512;
    static_assert(SMEM_BYTES <= 227 * 1024 && STAGES >= 4, "shared memory budget");
};

struct Params {
    int M, N, K;          // N = rows of W actually present (TMA zero-fills beyond)
    int m_tiles, n_tiles; // tiles of 256 x BN
    void* out;            // direct-store epilogues (3, 4): fp32
    long long ldo;        // elements between output rows
    const float* bias;    // [N] or null
    float scale;          // EPI_RESID_F32: divisor of the branch output
    // LayerNorm folded through the GEMM
    const float2* stats_in;    // *_LN consumers: [M][K / 128] (mean, M2) partials of the A rows
    const float* colsum;       // *_LN consumers: c[n] = sum_k bf16(gamma_k W_nk); bias = b[n] = sum_k beta_k W_nk
    float2* stats_out;         // EPI_RESID_F32_LN: [M][N / 128] partials of the updated rows
    __nv_bfloat16* xb_out;     // EPI_RESID_F32_LN: bf16 copy of the updated rows, [M][N]
    float ln_eps;
    int stats_span;            // *_LN consumers: columns per partial statistic of the A rows (128, or 96 when the
                               // producing residual GEMM ran 192-wide tiles)
    // EPI_QKV_ROPE_LN
    const float* rope;         // [>= T][64]: cos(t f_i) i < 32 | sin(t f_i), token position t = row % T
    const float* qk_gamma;     // [n_rope]: q_ln.weight | k_ln.weight
    float* qk_sumsq;           // [M][n_rope / 128]: sum over each 128-column span of (q - mean q)^2, (k - mean k)^2
    int T;                     // tokens per sample
    int n_rope;                // columns [0, n_rope) are q | k (multiple of 256); the rest (v) is stored as is
    int run;                   // EPI_QKV_ROPE_LN tile schedule (TileSchedule): 0 strided, R > 0 runs of R column tiles, < 0 contiguous ranges
};

__device__ __forceinline__ float gelu_erf(float x) {
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
// SwiGLU gate for a bf16 result: ex2/rcp approximations (2 ulp of fp32) are far below the bf16
// rounding of the output (2^-9)
__device__ __forceinline__ float silu_fast(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

// One 128-byte line (32 fp32) of a residual-stream row <-> registers, as four 256-bit accesses.
#ifndef RESID_DEBUG_SKIP
#define RESID_DEBUG_SKIP 0      // experiments only: 1 = no x loads, 2 = no x/xb stores, 3 = neither
#endif
__device__ __forceinline__ void ld_row128(float* r, const float* g) {
    if (RESID_DEBUG_SKIP & 1) return;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        asm volatile("ld.global.L1::no_allocate.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=f"(r[8 * i]), "=f"(r[8 * i + 1]), "=f"(r[8 * i + 2]), "=f"(r[8 * i + 3]),
                       "=f"(r[8 * i + 4]), "=f"(r[8 * i + 5]), "=f"(r[8 * i + 6]), "=f"(r[8 * i + 7])
                     : "l"(g + 8 * i)
                     : "memory");
}
__device__ __forceinline__ void st_row128(float* g, const float* r) {
    if (RESID_DEBUG_SKIP & 2) return;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        asm volatile("st.global.v8.f32 [%8], {%0, %1, %2, %3, %4, %5, %6, %7};"
                     ::"f"(r[8 * i]), "f"(r[8 * i + 1]), "f"(r[8 * i + 2]), "f"(r[8 * i + 3]),
                       "f"(r[8 * i + 4]), "f"(r[8 * i + 5]), "f"(r[8 * i + 6]), "f"(r[8 * i + 7]), "l"(g + 8 * i)
                     : "memory");
}

// 4 consecutive fp32 / bf16 of one row.
__device__ __forceinline__ void ld_quad(float* r, const float* g) {
    if (RESID_DEBUG_SKIP & 1) return;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]) : "l"(g) : "memory");
}
__device__ __forceinline__ void st_quad(float* g, const float* r) {
    if (RESID_DEBUG_SKIP & 2) return;
    asm volatile("st.global.v4.f32 [%4], {%0, %1, %2, %3};" ::"f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "l"(g) : "memory");
}
__device__ __forceinline__ void st_quad_bf16(__nv_bfloat16* g, const float* r) {
    if (RESID_DEBUG_SKIP & 2) return;
    asm volatile("st.global.v2.b32 [%2], {%0, %1};" ::"r"(pack_bf16x2(r[0], r[1])), "r"(pack_bf16x2(r[2], r[3])), "l"(g) : "memory");
}
// 8 x 8 transpose of 4-float blocks inside every group of 8 lanes: in, lane k holds block i in
// a[4 i .. 4 i + 3] (its own row, column quad i); out, lane i holds block k (row k of the group,
// column quad i) in the same slots.  Three butterfly stages of 16 shuffles, registers only (the
// same transpose through 4 KiB of swizzled shared memory per warp -- 16 LDS/STS.128 instead of
// ~190 instructions -- measured slower in situ: the main loop competes for shared-memory bandwidth).
__device__ __forceinline__ void transpose_quads8(float* a, int lane) {
#pragma unroll
    for (int b = 4; b >= 1; b >>= 1) {
        const bool up = lane & b;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j & b) continue;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float send = up ? a[4 * j + e] : a[4 * (j | b) + e];
                const float recv = __shfl_xor_sync(0xffffffffu, send, b);
                if (up) a[4 * j + e] = recv;
                else a[4 * (j | b) + e] = recv;
            }
        }
    }
}

// 32 bf16 of one row (64 bytes) as two 256-bit stores.
__device__ __forceinline__ void st_row64(__nv_bfloat16* g, const uint32_t* r) {
    if (RESID_DEBUG_SKIP & 2) return;
#pragma unroll
    for (int i = 0; i < 2; ++i)
        asm volatile("st.global.v8.b32 [%8], {%0, %1, %2, %3, %4, %5, %6, %7};"
                     ::"r"(r[8 * i]), "r"(r[8 * i + 1]), "r"(r[8 * i + 2]), "r"(r[8 * i + 3]),
                       "r"(r[8 * i + 4]), "r"(r[8 * i + 5]), "r"(r[8 * i + 6]), "r"(r[8 * i + 7]), "l"(g + 16 * i)
                     : "memory");
}
// N consecutive floats of a per-column vector, the same address in every lane (one L1 transaction
// per 16 bytes, broadcast).
template <int N>
__device__ __forceinline__ void ld_uniform(float* r, const float* g) {
#pragma unroll
    for (int i = 0; i < N / 4; ++i) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
        r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
    }
}
// The same from shared memory (one broadcast read per 16 bytes).
template <int N>
__device__ __forceinline__ void ld_uniform_smem(float* r, const float* s) {
#ifdef ESMDIFF_EXP_NOVEC                                    // timing experiment: no shared-memory broadcasts (results wrong)
#pragma unroll
    for (int i = 0; i < N; ++i) r[i] = 1.0f + 0.001f * i;
    return;
#endif
#pragma unroll
    for (int i = 0; i < N / 4; ++i) {
        const float4 v = reinterpret_cast<const float4*>(s)[i];
        r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
    }
}
// Row statistics from the partials the producing epilogue left: equal-sized spans, so
// mean = avg(mean_i), M2 = sum M2_i + span * sum (mean_i - mean)^2 (Chan et al.).  Returns
// rstd and -rstd * mean.
__device__ __forceinline__ void row_mean_rstd(const float2* st, int nspan, bool ok, float eps, float& rstd,
                                              float& nrm, int span = LN_SPAN) {
    rstd = 0.f;
    nrm = 0.f;
    if (!ok) return;
    float4 t[8];
    float ms = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (2 * i < nspan) {
            t[i] = __ldg(reinterpret_cast<const float4*>(st) + i);
            ms += t[i].x + t[i].z;
        }
    const float mean = ms / static_cast<float>(nspan);
    float m2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (2 * i < nspan) {
            const float d0 = t[i].x - mean, d1 = t[i].z - mean;
            m2 += (t[i].y + t[i].w) + static_cast<float>(span) * (d0 * d0 + d1 * d1);
        }
    rstd = rsqrtf(m2 / static_cast<float>(nspan * span) + eps);
    nrm = -rstd * mean;
}

// One 16-byte chunk of a 128-byte staging row under the 128B swizzle TMA expects.
__device__ __forceinline__ void st_swz16(uint8_t* box, int row, int chunk, uint32_t a, uint32_t b, uint32_t c,
                                         uint32_t d) {
    *reinterpret_cast<uint4*>(box + row * 128 + ((chunk ^ (row & 7)) << 4)) = make_uint4(a, b, c, d);
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}

// Tile schedule of one CTA pair over the (row tile, column tile) list, column tile fastest.
//   run == 0 : strided -- tile = cluster, cluster + n_clusters, ...  (every epilogue but EPI_QKV_ROPE_LN)
//   run == R : units of R consecutive column tiles of one row tile, units strided over the clusters.
//              Consecutive tiles of a unit share their token rows, so the rotary table row each
//              EPI_QKV_ROPE_LN epilogue thread needs (64 fp32, one row per thread = 32 L1 wavefronts per
//              load instruction) and the row statistics stay in registers for R tiles.  R must divide
//              n_tiles; the cost is coarser load balance (units of R tiles per cluster).
//   run < 0  : one CONTIGUOUS range of the list per cluster: maximal reuse and perfect balance, but all
//              74 clusters then sit on different A row tiles for the whole kernel (58 MB live in L2
//              beside 240 MB of streamed output; needs the evict-last hint on A, see the TMA producer).
struct TileSchedule {
    int count;                              // tiles of this cluster
    int mode, a, b, c, n_tiles;
    __device__ TileSchedule(int run, int cluster_id, int num_clusters, int num_tiles, int n_tiles_) {
        n_tiles = n_tiles_;
        if (run < 0) {
            mode = 2;
            a = static_cast<int>(static_cast<long long>(cluster_id) * num_tiles / num_clusters);
            count = static_cast<int>(static_cast<long long>(cluster_id + 1) * num_tiles / num_clusters) - a;
            b = c = 0;
        } else if (run == 0) {
            mode = 0;
            a = cluster_id;
            b = num_clusters;
            c = 0;
            count = cluster_id < num_tiles ? (num_tiles - cluster_id + num_clusters - 1) / num_clusters : 0;
        } else {
            mode = 1;
            a = cluster_id;
            b = num_clusters;
            c = run;
            const int units = num_tiles / run;
            count = cluster_id < units ? (units - cluster_id + num_clusters - 1) / num_clusters * run : 0;
        }
    }
    __device__ __forceinline__ int tile_at(int i) const {
        if (mode == 2) return a + i;
        if (mode == 0) return a + i * b;
        const int u = a + (i / c) * b, per_row = n_tiles / c;
        return (u / per_row) * n_tiles + (u % per_row) * c + i % c;
    }
};

template <int EPI, int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA,     // A   [M, K] bf16, box 128 x 64
                    const __grid_constant__ CUtensorMap tmB,     // W   [N, K] bf16, box 128 x 64
                    const __grid_constant__ CUtensorMap tmC,     // out: bf16 box 32 x 64 / fp32 box 32 x 32
                    const Params p) {
    using C = Cfg<EPI, BN>;
    constexpr int STAGES = C::STAGES;
    constexpr int BOXES = C::BOXES;
    constexpr int BN_CTA = C::BN_CTA;
    constexpr int STAGE_BYTES = C::STAGE_BYTES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* stg_all = smem + STAGES * STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + C::STG_BYTES);
    uint64_t* full = bars;                 // [STAGES]  TMA (both CTAs) -> MMA; used in the leader only
    uint64_t* empty = bars + MAX_STAGES;   // [STAGES]  MMA -> TMA, multicast to both CTAs
    uint64_t* tfull = bars + 2 * MAX_STAGES;   // [2]   MMA -> epilogue, multicast to both CTAs
    uint64_t* tempty = tfull + 2;          // [2]       epilogue warps of both CTAs -> MMA (leader's copy)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();             // 0 = leader
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;
    const int num_tiles = p.m_tiles * p.n_tiles;
    const int kblocks = p.K / BK;
    const TileSchedule sched(EPI == EPI_QKV_ROPE_LN ? p.run : 0, cluster_id, num_clusters, num_tiles, p.n_tiles);

    if (warp == WARP_TMA && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if constexpr (EPI == EPI_STORE_BF16 || EPI == EPI_SWIGLU_BF16 || EPI == EPI_STORE_BF16_LN || EPI == EPI_SWIGLU_BF16_LN || EPI == EPI_QKV_ROPE_LN) tma_prefetch_desc(&tmC);
    }
    if (warp == WARP_MMA && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], 2 * EPI_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == WARP_ALLOC) {
        tmem_alloc_pair(tmem_slot, TMEM_COLS);
        tmem_relinquish_pair();
    }
    tcgen05_fence_before();
    cluster_sync_all();                    // peer barriers initialised before any remote arrive / TMA
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();               // the next kernel may set itself up while this one runs

    if (warp == WARP_TMA) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            pdl_wait();                    // A (and any other input) is complete and visible
            int stage = 0;
            uint32_t phase = 0;
            for (int it = 0; it < sched.count; ++it) {
                const int tile = sched.tile_at(it);
                const int m0 = (tile / p.n_tiles) * BM + rank * BM_CTA;
                const int n0 = (tile % p.n_tiles) * BN + rank * BN_CTA;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    const uint32_t leader_full = mapa_shared(smem_u32(&full[stage]), 0);
                    if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * STAGE_BYTES);
                    if constexpr (EPI == EPI_QKV_ROPE_LN) {
                        // contiguous tile ranges: a CTA pair re-reads ITS A row tile for ~18 consecutive column
                        // tiles while the chip streams 240 MB of q|k|v through L2 -- without a hint the A tiles
                        // are evicted in between and come back from DRAM (ncu r2e: 793 MB of DRAM reads against
                        // 96 MB for the strided schedule, tensor pipe 79 %): keep A, let the output go first
                        tma_load_2d_pair_hint(sa, &tmA, leader_full, kb * BK, m0, L2_EVICT_LAST);
                    } else {
                        tma_load_2d_pair(sa, &tmA, leader_full, kb * BK, m0);
                    }
                    tma_load_2d_pair(sa + A_BYTES, &tmB, leader_full, kb * BK, n0);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == WARP_MMA) {
        // ===================== MMA issuer (leader CTA) =====================
        // The whole warp walks the schedule so that control flow, stage counters and descriptors
        // are warp-uniform (uniform registers, no per-MMA re-broadcast); one elected lane issues
        // the tcgen05 instructions.  The next stage's barrier is probed before this stage's MMAs
        // are issued, so its ~90-cycle try_wait latency is off the issue path.
        if (rank == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, 0);
            const uint64_t desc0 = umma_desc_sw128(smem_u32(smem), 16, 1024);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int it = 0; it < sched.count; ++it) {
                const int tile = sched.tile_at(it);
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + acc * ACC_STRIDE;
                bool ready = mbar_try_wait(&full[stage], phase);
                for (int kb = 0; kb < kblocks; ++kb) {
                    if (!ready) mbar_wait(&full[stage], phase);
                    tcgen05_fence_after();
                    const int next = stage + 1 == STAGES ? 0 : stage + 1;
                    const uint32_t next_phase = next == 0 ? phase ^ 1 : phase;
                    ready = kb + 1 < kblocks ? mbar_try_wait(&full[next], next_phase) : false;
                    const uint64_t adesc = desc0 + static_cast<uint64_t>(stage * (STAGE_BYTES >> 4));
                    const uint64_t bdesc = adesc + (A_BYTES >> 4);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            // +32 bytes (16 bf16) along K inside the 128B swizzle row = +2 encoded
                            umma_bf16_ss_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                        umma_commit_pair(&empty[stage]);  // frees the slot in both CTAs when the MMAs retire
                    }
                    __syncwarp();
                    stage = next;
                    phase = next_phase;
                }
                if (elect_one()) umma_commit_pair(&tfull[acc]);   // accumulators ready in both CTAs
                __syncwarp();
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp < EPI_WARPS) {
        // ===================== epilogue warps (both CTAs) =====================
        const int q = warp & 3;                       // TMEM lanes [32 q, 32 q + 32)
        const int half = warp >> 2;                   // accumulator columns [128 half, 128 half + 128)
        uint8_t* box = stg_all + warp * BOXES * BOX_BYTES;      // 1024-byte aligned: the 128B swizzle is a function of the address
        float* vec = reinterpret_cast<float*>(stg_all + EPI_WARPS * BOXES * BOX_BYTES + warp * C::VEC_BYTES);   // [3][128]: colsum | bias | gamma
        int acc = 0;
        uint32_t acc_phase = 0;
        constexpr bool RESID = EPI == EPI_RESID_F32 || EPI == EPI_RESID_F32_LN;
        float xres[RESID ? 64 : 1];                   // residual epilogue: two 32-float register sets of x
        pdl_wait();                        // statistics / residual stream reads, all global writes
        constexpr bool ROPE = EPI == EPI_QKV_ROPE_LN;
        float rcos[ROPE ? 32 : 1], rsin[ROPE ? 32 : 1];   // rotary table row of this thread's token position
        int rope_mt = -1;                             // row tile the table row was loaded for
        int stat_mt = -1;                             // row tile rstd_c / nrm_c belong to
        float rstd_c = 0.f, nrm_c = 0.f;
        for (int it = 0; it < sched.count; ++it) {
            const int tile = sched.tile_at(it);
            const int nb = tile % p.n_tiles;
            const int n0 = nb * BN + half * (BN / 2);
            const int row_base = (tile / p.n_tiles) * BM + rank * BM_CTA + q * 32;
            const uint32_t t_row = tmem_base + acc * ACC_STRIDE + half * (BN / 2) + (static_cast<uint32_t>(q * 32) << 16);

            if constexpr (EPI == EPI_STORE_BF16) {
                mbar_wait(&tfull[acc], acc_phase);
                tcgen05_fence_after();
#pragma unroll 1
                for (int c = 0; c < (BN / 2) / 64; ++c) {
                    uint32_t v[32], w[32];
                    tmem_ld_32x32b_x32(t_row + c * 64, v);
                    tmem_ld_32x32b_x32(t_row + c * 64 + 32, w);
                    tmem_ld_wait();
                    if (lane == 0) bulk_wait_group_read<0>();     // the previous store has read the box
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        st_swz16(box, lane, j,
                                 pack_bf16x2(__uint_as_float(v[8 * j]), __uint_as_float(v[8 * j + 1])),
                                 pack_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3])),
                                 pack_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5])),
                                 pack_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])));
                        st_swz16(box, lane, 4 + j,
                                 pack_bf16x2(__uint_as_float(w[8 * j]), __uint_as_float(w[8 * j + 1])),
                                 pack_bf16x2(__uint_as_float(w[8 * j + 2]), __uint_as_float(w[8 * j + 3])),
                                 pack_bf16x2(__uint_as_float(w[8 * j + 4]), __uint_as_float(w[8 * j + 5])),
                                 pack_bf16x2(__uint_as_float(w[8 * j + 6]), __uint_as_float(w[8 * j + 7])));
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmC, box, n0 + c * 64, row_base);
                        bulk_commit_group();
                    }
                }
            } else if constexpr (EPI == EPI_SWIGLU_BF16) {
                // tile columns [0,128) = gate rows, [128,256) = up rows of the same hidden units;
                // this warp produces output columns [64 half, 64 half + 64) of the tile's 128
                mbar_wait(&tfull[acc], acc_phase);
                tcgen05_fence_after();
                const uint32_t t_gate = tmem_base + acc * ACC_STRIDE + half * 64 + (static_cast<uint32_t>(q * 32) << 16);
                if (lane == 0) bulk_wait_group_read<0>();
                __syncwarp();
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t g[32], u[32];
                    tmem_ld_32x32b_x32(t_gate + h * 32, g);
                    tmem_ld_32x32b_x32(t_gate + BN / 2 + h * 32, u);
                    tmem_ld_wait();
                    float y[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) y[j] = silu_fast(__uint_as_float(g[j])) * __uint_as_float(u[j]);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        st_swz16(box, lane, h * 4 + j, pack_bf16x2(y[8 * j], y[8 * j + 1]),
                                 pack_bf16x2(y[8 * j + 2], y[8 * j + 3]), pack_bf16x2(y[8 * j + 4], y[8 * j + 5]),
                                 pack_bf16x2(y[8 * j + 6], y[8 * j + 7]));
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&tmC, box, nb * (BN / 2) + half * 64, row_base);
                    bulk_commit_group();
                }
            } else if constexpr (EPI == EPI_STORE_BF16_LN || EPI == EPI_QKV_ROPE_LN) {
                // out = rstd * acc + (b[n] - rstd * mean * c[n]); the row statistics are fetched
                // while the main loop of this tile is still running
                float rstd, nrm;
                bool rot = false;
#ifndef ESMDIFF_VEC_SMEM
                float4 vc4 = make_float4(0.f, 0.f, 0.f, 0.f), vb4 = vc4, vg4 = vc4;
                // 16 consecutive entries of a vector held as one float4 per lane: lanes base .. base + 3
                auto bcast16 = [&](float* r, const float4& v, int base) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        r[4 * j] = __shfl_sync(0xffffffffu, v.x, base + j);
                        r[4 * j + 1] = __shfl_sync(0xffffffffu, v.y, base + j);
                        r[4 * j + 2] = __shfl_sync(0xffffffffu, v.z, base + j);
                        r[4 * j + 3] = __shfl_sync(0xffffffffu, v.w, base + j);
                    }
                };
#endif
                if constexpr (ROPE) {
                    // contiguous tile ranges: consecutive tiles share their rows, so the row statistics
                    // and the rotary table row are fetched once per row tile
                    rot = n0 < p.n_rope;                           // warp-uniform: a q or k column tile
                    const int mt = tile / p.n_tiles;
                    if (mt != stat_mt) {
                        row_mean_rstd(p.stats_in + static_cast<long long>(row_base + lane) * (p.K / p.stats_span),
                                      p.K / p.stats_span, row_base + lane < p.M, p.ln_eps, rstd_c, nrm_c, p.stats_span);
                        stat_mt = mt;
                    }
                    rstd = rstd_c;
                    nrm = nrm_c;
                    // this warp's 128 columns of colsum / bias / gamma -> shared memory (latency hidden
                    // behind the wait for the accumulators)
#ifndef ESMDIFF_VEC_SMEM
                    vc4 = __ldg(reinterpret_cast<const float4*>(p.colsum + n0) + lane);
                    vb4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + lane);
                    vg4 = make_float4(1.f, 1.f, 1.f, 1.f);
                    if (rot) vg4 = __ldg(reinterpret_cast<const float4*>(p.qk_gamma + n0) + lane);
#else
                    __syncwarp();
                    {
                        const float4 c4 = __ldg(reinterpret_cast<const float4*>(p.colsum + n0) + lane);
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + lane);
                        float4 g4 = make_float4(1.f, 1.f, 1.f, 1.f);
                        if (rot) g4 = __ldg(reinterpret_cast<const float4*>(p.qk_gamma + n0) + lane);
                        reinterpret_cast<float4*>(vec)[lane] = c4;
                        reinterpret_cast<float4*>(vec + 128)[lane] = b4;
                        reinterpret_cast<float4*>(vec + 256)[lane] = g4;
                    }
                    __syncwarp();
#endif
#ifdef ESMDIFF_EXP_NOTABLE                                  // timing experiment: no rotary table loads (results wrong)
                    if (rot && rope_mt < 0) {
#else
                    if (rot && mt != rope_mt) {
#endif
                        const int row = row_base + lane < p.M ? row_base + lane : p.M - 1;
#ifndef ESMDIFF_ROPE_TABLE
                        // the rotary factors of this thread's token position from MUFU sin / cos (the special-function
                        // unit is idle in a GEMM) instead of a 256-byte table row per thread and tile through L1, which
                        // competes with the operand traffic of the main loop: 242.7 -> 230.4 us at M = 25 800
                        // (profiles/r4c_qkv_rope_on_the_fly.txt; -DESMDIFF_ROPE_TABLE builds the table form).
                        // sin.approx / cos.approx reduce the argument in fp32: <= ~1e-4 absolute at position 1026
                        // -- the reference's own fp32 angle t * inv_freq is only good to 6e-5 there, and both are
                        // 40x below the bf16 rounding of the rotated q', k'.
                        const float INV_FREQ[32] = {1.000000000e+00f, 7.498942018e-01f, 5.623413324e-01f, 4.216965139e-01f, 3.162277639e-01f, 2.371373922e-01f, 1.778279394e-01f, 1.333521456e-01f, 1.000000015e-01f, 7.498941571e-02f, 5.623412877e-02f, 4.216964915e-02f, 3.162277862e-02f, 2.371373586e-02f, 1.778279431e-02f, 1.333521493e-02f, 9.999999776e-03f, 7.498942316e-03f, 5.623413250e-03f, 4.216964822e-03f, 3.162277862e-03f, 2.371373819e-03f, 1.778279431e-03f, 1.333521446e-03f, 1.000000047e-03f, 7.498941850e-04f, 5.623413017e-04f, 4.216965463e-04f, 3.162277862e-04f, 2.371373848e-04f, 1.778279402e-04f, 1.333521504e-04f};
                        const float fp = static_cast<float>(row % p.T);
#pragma unroll
                        for (int i = 0; i < 32; ++i) __sincosf(fp * INV_FREQ[i], &rsin[i], &rcos[i]);
#else
                        const float4* rp = reinterpret_cast<const float4*>(p.rope + static_cast<long long>(row % p.T) * 64);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 c4 = __ldg(rp + i), s4 = __ldg(rp + 8 + i);
                            rcos[4 * i] = c4.x; rcos[4 * i + 1] = c4.y; rcos[4 * i + 2] = c4.z; rcos[4 * i + 3] = c4.w;
                            rsin[4 * i] = s4.x; rsin[4 * i + 1] = s4.y; rsin[4 * i + 2] = s4.z; rsin[4 * i + 3] = s4.w;
                        }
#endif
                        rope_mt = mt;
                    }
                } else {
                    row_mean_rstd(p.stats_in + static_cast<long long>(row_base + lane) * (p.K / p.stats_span),
                                  p.K / p.stats_span, row_base + lane < p.M, p.ln_eps, rstd, nrm, p.stats_span);
                }
                mbar_wait(&tfull[acc], acc_phase);
                tcgen05_fence_after();
                if (!rot) {
#pragma unroll 1
                    for (int c = 0; c < (BN / 2) / 64; ++c) {
                        if (lane == 0) bulk_wait_group_read<0>();     // the previous store has read the box
                        __syncwarp();
                        if constexpr (ROPE) {
                            // v column tiles of the QKV projection: 16 columns at a time (this kernel keeps
                            // 64 registers of rotary table per thread)
#pragma unroll 1
                            for (int hh = 0; hh < 4; ++hh) {
                                uint32_t v[16];
                                float cs[16], bs[16];
                                tmem_ld_32x32b_x16(t_row + c * 64 + hh * 16, v);
#ifndef ESMDIFF_VEC_SMEM
                                bcast16(cs, vc4, c * 16 + hh * 4);
                                bcast16(bs, vb4, c * 16 + hh * 4);
#else
                                ld_uniform_smem<16>(cs, vec + c * 64 + hh * 16);
                                ld_uniform_smem<16>(bs, vec + 128 + c * 64 + hh * 16);
#endif
                                tmem_ld_wait();
                                float y[16];
#pragma unroll
                                for (int j = 0; j < 16; ++j) y[j] = fmaf(__uint_as_float(v[j]), rstd, fmaf(nrm, cs[j], bs[j]));
#pragma unroll
                                for (int j = 0; j < 2; ++j)
                                    st_swz16(box, lane, hh * 2 + j, pack_bf16x2(y[8 * j], y[8 * j + 1]),
                                             pack_bf16x2(y[8 * j + 2], y[8 * j + 3]), pack_bf16x2(y[8 * j + 4], y[8 * j + 5]),
                                             pack_bf16x2(y[8 * j + 6], y[8 * j + 7]));
                            }
                        } else {
#pragma unroll 1
                            for (int hh = 0; hh < 2; ++hh) {
                                uint32_t v[32];
                                float cs[32], bs[32];
                                tmem_ld_32x32b_x32(t_row + c * 64 + hh * 32, v);
                                ld_uniform<32>(cs, p.colsum + n0 + c * 64 + hh * 32);
                                ld_uniform<32>(bs, p.bias + n0 + c * 64 + hh * 32);
                                tmem_ld_wait();
                                float y[32];
#pragma unroll
                                for (int j = 0; j < 32; ++j) y[j] = fmaf(__uint_as_float(v[j]), rstd, fmaf(nrm, cs[j], bs[j]));
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    st_swz16(box, lane, hh * 4 + j, pack_bf16x2(y[8 * j], y[8 * j + 1]),
                                             pack_bf16x2(y[8 * j + 2], y[8 * j + 3]), pack_bf16x2(y[8 * j + 4], y[8 * j + 5]),
                                             pack_bf16x2(y[8 * j + 6], y[8 * j + 7]));
                            }
                        }
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            if constexpr (ROPE) tma_store_2d_hint(&tmC, box, n0 + c * 64, row_base, L2_EVICT_FIRST);
                            else tma_store_2d(&tmC, box, n0 + c * 64, row_base);
                            bulk_commit_group();
                        }
                    }
                } else if constexpr (ROPE) {
                    // the q / k rows of W are stored with the two halves of every rotation adjacent
                    // (feature d at position 2 d', d + 32 at 2 d' + 1 of its head), so consecutive
                    // accumulator columns (2 e, 2 e + 1) of a 32-column chunk are one rotary pair with
                    // frequency index 16 hh + e
                    float sq = 0.f;
#pragma unroll 1
                    for (int c = 0; c < (BN / 2) / 64; ++c) {
                        if (lane == 0) bulk_wait_group_read<0>();
                        __syncwarp();
#pragma unroll
                        for (int hq = 0; hq < 4; ++hq) {                   // 16 columns = 8 rotary pairs at a time
                            uint32_t v[16];
                            float cs[16], bs[16], gs[16];
                            const int vo = c * 64 + hq * 16;
                            tmem_ld_32x32b_x16(t_row + vo, v);
#ifndef ESMDIFF_VEC_SMEM
                            bcast16(cs, vc4, vo >> 2);
                            bcast16(bs, vb4, vo >> 2);
                            bcast16(gs, vg4, vo >> 2);
#else
                            ld_uniform_smem<16>(cs, vec + vo);
                            ld_uniform_smem<16>(bs, vec + 128 + vo);
                            ld_uniform_smem<16>(gs, vec + 256 + vo);
#endif
                            tmem_ld_wait();
                            uint32_t o[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const float ya = fmaf(__uint_as_float(v[2 * e]), rstd, fmaf(nrm, cs[2 * e], bs[2 * e]));
                                const float yb = fmaf(__uint_as_float(v[2 * e + 1]), rstd, fmaf(nrm, cs[2 * e + 1], bs[2 * e + 1]));
                                sq = fmaf(ya, ya, sq);
                                sq = fmaf(yb, yb, sq);
                                const float za = ya * gs[2 * e], zb = yb * gs[2 * e + 1];
                                const float cc = rcos[hq * 8 + e], ss = rsin[hq * 8 + e];
                                o[e] = pack_bf16x2(fmaf(za, cc, -(zb * ss)),      // x1 cos - x2 sin
                                                   fmaf(zb, cc, za * ss));        // x2 cos + x1 sin
                            }
                            st_swz16(box, lane, hq * 2, o[0], o[1], o[2], o[3]);
                            st_swz16(box, lane, hq * 2 + 1, o[4], o[5], o[6], o[7]);
                        }
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            if constexpr (ROPE) tma_store_2d_hint(&tmC, box, n0 + c * 64, row_base, L2_EVICT_FIRST);
                            else tma_store_2d(&tmC, box, n0 + c * 64, row_base);
                            bulk_commit_group();
                        }
                    }
                    if (row_base + lane < p.M)
                        p.qk_sumsq[static_cast<long long>(row_base + lane) * (p.n_rope / LN_SPAN) + n0 / LN_SPAN] = sq;
                }
            } else if constexpr (EPI == EPI_SWIGLU_BF16_LN) {
                float rstd, nrm;
                row_mean_rstd(p.stats_in + static_cast<long long>(row_base + lane) * (p.K / p.stats_span),
                              p.K / p.stats_span, row_base + lane < p.M, p.ln_eps, rstd, nrm, p.stats_span);
                mbar_wait(&tfull[acc], acc_phase);
                tcgen05_fence_after();
                const uint32_t t_gate = tmem_base + acc * ACC_STRIDE + half * 64 + (static_cast<uint32_t>(q * 32) << 16);
                const float* cg = p.colsum + nb * BN + half * 64;       // gate columns of this warp; up = +BN/2
                const float* bg = p.bias + nb * BN + half * 64;
                if (lane == 0) bulk_wait_group_read<0>();
                __syncwarp();
#pragma unroll
                for (int h = 0; h < 4; ++h) {                           // 16 output columns at a time
                    uint32_t g[16], u[16];
                    float c0[16], b0[16], c1[16], b1[16];
                    tmem_ld_32x32b_x16(t_gate + h * 16, g);
                    tmem_ld_32x32b_x16(t_gate + BN / 2 + h * 16, u);
                    ld_uniform<16>(c0, cg + h * 16);
                    ld_uniform<16>(b0, bg + h * 16);
                    ld_uniform<16>(c1, cg + BN / 2 + h * 16);
                    ld_uniform<16>(b1, bg + BN / 2 + h * 16);
                    tmem_ld_wait();
                    float y[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float yg = fmaf(__uint_as_float(g[j]), rstd, fmaf(nrm, c0[j], b0[j]));
                        const float yu = fmaf(__uint_as_float(u[j]), rstd, fmaf(nrm, c1[j], b1[j]));
                        y[j] = silu_fast(yg) * yu;
                    }
#pragma unroll
                    for (int j = 0; j < 2; ++j)
                        st_swz16(box, lane, h * 2 + j, pack_bf16x2(y[8 * j], y[8 * j + 1]),
                                 pack_bf16x2(y[8 * j + 2], y[8 * j + 3]), pack_bf16x2(y[8 * j + 4], y[8 * j + 5]),
                                 pack_bf16x2(y[8 * j + 6], y[8 * j + 7]));
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&tmC, box, nb * (BN / 2) + half * 64, row_base);
                    bulk_commit_group();
                }
            } else if constexpr (RESID) {
                // residual stream update x = x + acc / scale: x moves global <-> registers directly
                // (a TMA-staged form, 512 KiB of shared-memory traffic per tile on top of a main loop
                // already near the shared-memory bandwidth limit, held out_proj to 0.74-0.93 PFLOP/s).
                // Access pattern: a 32x32b TMEM load puts a token ROW in each thread, and the first
                // version read/wrote x that way (a 32-byte sector per lane = 32 L1 wavefronts per
                // instruction, 10 240 per tile).  tools/l2test.py with loads or stores compiled out
                // shows each alone is free but together they exceed the 12.3k-clk main loop of a
                // K = 1536 tile (out_proj 123-145 us vs 80 us) independent of L2 residency: the LSU
                // pipe, at ~2 clk per wavefront.  So the accumulator chunk is transposed in 8 x 8
                // blocks of 4 floats across each group of 8 lanes (48 shuffles): afterwards lane i of
                // a group holds columns [4 i, 4 i + 4) of all 8 rows of the group, and every 128-bit
                // access has 8 lanes on one 128-byte line: 4x fewer wavefronts, same instruction
                // count.  Loads run two chunks ahead in two register sets; the first two chunks of
                // the NEXT tile are requested while this tile drains.  acc / scale is evaluated as
                // acc * (1 / scale): <= 1 ulp(fp32) from the reference's division.
                constexpr int NCH = (BN / 2) / 32;          // 4 chunks (BN = 256) or 3 (BN = 192) of 32 columns per warp
                constexpr int SPAN = BN / 2;                // columns per partial LayerNorm statistic this warp leaves
                float* xg = reinterpret_cast<float*>(p.out);
                const float inv = 1.0f / p.scale;
                const int grp = lane >> 3, qd = lane & 7;                  // rows 8 grp .. 8 grp + 7, column quad qd
                // first of my 8 rows and my 4 columns in chunk c of tile t
                auto quad_ptr = [&](int t, int c) {
                    const long long row = static_cast<long long>(t / p.n_tiles) * BM + rank * BM_CTA + q * 32 + grp * 8;
                    return xg + row * p.ldo + (t % p.n_tiles) * BN + half * (BN / 2) + c * 32 + qd * 4;
                };
                auto rows_left = [&](int t) { return p.M - ((t / p.n_tiles) * BM + static_cast<int>(rank) * BM_CTA + q * 32 + grp * 8); };
                auto ld_chunk = [&](float* xr, int t, int c) {
                    const float* src = quad_ptr(t, c);
                    const int nr = rows_left(t);
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (k < nr) ld_quad(xr + 4 * k, src + static_cast<long long>(k) * p.ldo);
                };
                const int nr = rows_left(tile);
                // x loads run two chunks ahead in two register sets.  With an even chunk count the sets
                // line up across tiles and the first two chunks of the NEXT tile are requested while this
                // one drains; with three chunks (BN = 192) they would swap roles every tile, so each
                // tile requests its own first two chunks before it waits for the accumulator instead.
                constexpr bool CROSS = NCH % 2 == 0;
                if (!CROSS || it == 0) {
                    ld_chunk(xres, tile, 0);
                    ld_chunk(xres + 32, tile, 1);
                }
                const int nt = it + 1 < sched.count ? sched.tile_at(it + 1) : num_tiles;
                mbar_wait(&tfull[acc], acc_phase);
                tcgen05_fence_after();
                // EPI_RESID_F32_LN: (sum, sum of squares) of my 4 columns of each of the 8 rows, over
                // the warp's chunks.  Plain sums over a SPAN-column span: M2 = q - s^2/SPAN loses
                // ~1e-7 (1 + (span mean / span std)^2) relative, harmless for a residual stream whose
                // span means are of the order of its spread; the spans are then combined exactly
                // (Chan) by the consumer.
                float st_s[8], st_q[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) { st_s[k] = 0.f; st_q[k] = 0.f; }
                auto do_chunk = [&](const int c, float* xr) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(t_row + c * 32, v);
                    tmem_ld_wait();
                    float* a = reinterpret_cast<float*>(v);
                    transpose_quads8(a, lane);
                    float* dst = quad_ptr(tile, c);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) xr[4 * k + e] = fmaf(a[4 * k + e], inv, xr[4 * k + e]);
                        if (k < nr) st_quad(dst + static_cast<long long>(k) * p.ldo, xr + 4 * k);
                    }
                    if constexpr (EPI == EPI_RESID_F32_LN) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                st_s[k] += xr[4 * k + e];
                                st_q[k] = fmaf(xr[4 * k + e], xr[4 * k + e], st_q[k]);
                            }
                            if (k < nr) st_quad_bf16(p.xb_out + (dst - xg) + static_cast<long long>(k) * p.ldo, xr + 4 * k);
                        }
                    }
                    if (c + 2 < NCH) ld_chunk(xr, tile, c + 2);
                    else if (CROSS && nt < num_tiles) ld_chunk(xr, nt, c + 2 - NCH);
                };
                if constexpr (NCH == 4) {
#pragma unroll 1
                    for (int c2 = 0; c2 < NCH; c2 += 2) {
                        do_chunk(c2, xres);
                        do_chunk(c2 + 1, xres + 32);
                    }
                } else {
                    do_chunk(0, xres);
                    do_chunk(1, xres + 32);
                    do_chunk(2, xres);
                }
                if constexpr (EPI == EPI_RESID_F32_LN) {
                    // add the 8 column quads of each row: 8 items over 8 lanes -> lane qd ends with row qd
#pragma unroll
                    for (int b = 4; b >= 1; b >>= 1) {
                        const bool up = lane & b;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (j & b) continue;
                            const float ss = up ? st_s[j] : st_s[j | b], sq = up ? st_q[j] : st_q[j | b];
                            const float rs = __shfl_xor_sync(0xffffffffu, ss, b), rq = __shfl_xor_sync(0xffffffffu, sq, b);
                            if (up) { st_s[j | b] += rs; st_q[j | b] += rq; }
                            else { st_s[j] += rs; st_q[j] += rq; }
                        }
                    }
                    // lane qd now holds the totals of row qd of its group in slot qd
                    float my_s = st_s[0], my_q = st_q[0];
#pragma unroll
                    for (int k = 1; k < 8; ++k)
                        if (qd == k) { my_s = st_s[k]; my_q = st_q[k]; }
                    if (qd < nr) {
                        const long long row = static_cast<long long>(row_base + grp * 8 + qd);
                        const float sm = my_s * (1.0f / SPAN);
                        p.stats_out[row * (p.N / SPAN) + nb * 2 + half] = make_float2(sm, my_q - my_s * sm);
                    }
                }
            } else {
                // bias (+ GELU), fp32 out with arbitrary row stride: transpose through shared
                // memory (XOR-swizzled 32x32 words) so that a warp writes 128-byte row segments
                mbar_wait(&tfull[acc], acc_phase);
                tcgen05_fence_after();
                float* tb = reinterpret_cast<float*>(box);
                float* out = reinterpret_cast<float*>(p.out);
#pragma unroll 1
                for (int c = 0; c < (BN / 2) / 32; ++c) {
                    if (n0 + c * 32 >= p.N) break;               // warp-uniform
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(t_row + c * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) tb[lane * 32 + (j ^ lane)] = __uint_as_float(v[j]);
                    __syncwarp();
                    const int col = n0 + c * 32 + lane;
                    if (col < p.N) {
                        const float b = p.bias[col];
#pragma unroll 8
                        for (int r = 0; r < 32; ++r) {
                            const int row = row_base + r;
                            if (row < p.M) {
                                float y = tb[r * 32 + (lane ^ r)] + b;
                                if constexpr (EPI == EPI_BIAS_GELU_F32) y = gelu_erf(y);
                                out[static_cast<long long>(row) * p.ldo + col] = y;
                            }
                        }
                    }
                    __syncwarp();
                }
            }
            // all TMEM reads of this accumulator stage are complete (tcgen05.wait::ld above)
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_relaxed(mapa_shared(smem_u32(&tempty[acc]), 0));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if constexpr (EPI == EPI_STORE_BF16 || EPI == EPI_SWIGLU_BF16 || EPI == EPI_STORE_BF16_LN || EPI == EPI_SWIGLU_BF16_LN || EPI == EPI_QKV_ROPE_LN) {
            if (lane == 0) bulk_wait_group<0>();                  // stores complete before exit
        }
    }

    tcgen05_fence_before();
    cluster_sync_all();     // the leader's MMAs read the peer's shared memory: nobody leaves early
    if (warp == WARP_ALLOC) {
        tcgen05_fence_after();
        tmem_dealloc_pair(tmem_base, TMEM_COLS);
    }
}

}  // namespace gemm
}  // namespace esmdiff
