// Non-causal multi-head attention, d_head = 64, with the K and V of one (sample, head) RESIDENT in
// shared memory: the variant for the sequence lengths the sampling path is quoted on
// (T = L + 2 <= 766; T = 258 keeps two CTAs per SM).  Replaces F.scaled_dot_product_attention on
// the reference path (esm MultiHeadAttention.forward with seq_id None).
//
// One CTA = one (sample b, head h); it walks ALL query tiles of 128 rows, so K/V are read from
// L2/HBM once per (b, h) instead of once per query tile.  Per step (query tile qt, kv tile j):
//   S warp    : S = Q_qt K_j^T     UMMA 128 x ncols x 16 (x4), fp32 into one of THREE TMEM S buffers,
//               issued up to three steps ahead of the softmax (a buffer is free once the P V that
//               read P from it has retired)
//   P V warp  : O (+)= P V_j       UMMA with A = P FROM TMEM (bf16 pairs written over the S buffer
//                                  by the softmax threads), B = V_j as MN-major smem operand
//               (two issuing warps: the issue of these small MMAs plus a commit cost one warp ~1100 clk
//               per step, more than the softmax of the step -- see the P V issuer below)
//   4 softmax warps (thread = query row = TMEM lane): row max of the tile; the running max is only
//               raised when the tile max exceeds it by more than 2^8 (then O is rescaled in TMEM and
//               the row sum in its register -- rare after the first tile), p = exp2(s*scale - m),
//               packed to bf16 and stored back to TMEM.  No shared-memory round trip for P, no
//               per-tile read of O.
//   The last kv tile is only as wide as needed (multiple of 16 columns): T = 258 costs 4 x 64 + 16
//   columns, not 5 x 64.  Warps whose 32 query rows are all >= T skip the softmax (the 2-row tail
//   tile of T = 258 keeps one warp busy, not four).
// Input  qkv : bf16 [M = B*T, 3*D]  (q | k | v, each D = H*64).  Two forms of q, k:
//   qk_rstd == null : q, k already LayerNormed + RoPE'd (stand-alone ew::qk_layernorm_rope_kernel)
//   qk_rstd != null : q' = rope(gamma_q (q - mean q)), k' likewise, NOT yet divided by their row
//                      standard deviation; qk_rstd holds 1/std per row (qk_rstd_kernel below, from the
//                      partial sums of squares the QKV GEMM epilogue left, gemm.cuh EPI_QKV_ROPE_LN).  The
//                      missing factors are per-row scalars, both applied to the fp32 scores: rstd_q[i]
//                      goes into the softmax scale of query row i (thread = row), rstd_k[j] multiplies
//                      score column j (one packed multiply per two scores, the factors read as
//                      shared-memory broadcasts).  q and k are therefore rounded to bf16 exactly once.
//                      (First version: K rows rescaled in shared memory before the first S MMA -- a
//                      second rounding of k, a proxy fence per tile and the whole pass on the critical
//                      path of every CTA: +15 % kernel time; this form: see DESIGN.md.)
// Output ctx : bf16 [M, D]
//
// Loads (round 2, from a per-CTA clock trace, tools/attn_timeline.py): the CTAs of a wave start together,
// ask for their 96 KB at the same moment (28 MB chip-wide = 4.4 us of HBM time with every tensor pipe
// idle), compute together with HBM idle, and end together.  So (1) the TMA warp issues its loads BEFORE
// the CTA-wide set-up barrier, one tile per lane in one go (a single thread needed ~140 clk per load);
// (2) once its own tiles have landed, every CTA prefetches the tiles of the CTA that will take its place
// (block index + resident CTAs) into L2 (cp.async.bulk.prefetch.tensor): the next wave's loads then hit
// L2 and HBM works while the tensor pipes do; (3) the 1/std tables come from a tiny kernel run once per
// layer instead of 24 KB of partial sums re-reduced by each of the 24 head CTAs of a sample with the
// TMA loads held back behind them.
#pragma once
#include "ptx.cuh"

namespace esmdiff {
namespace attn2 {

#ifndef ATTN_DBG
#define ATTN_DBG 0      // development builds only (tools/attn_timeline.py): 32 = per-CTA clock trace, 4 = no softmax math
#endif
#if ATTN_DBG & 32
constexpr int TRACE_SLOTS = 64;
static __device__ long long g_attn_trace[4096 * TRACE_SLOTS];
__device__ __forceinline__ void trace_ev(int slot) {
    if (blockIdx.x < 4096 && slot < TRACE_SLOTS) g_attn_trace[blockIdx.x * TRACE_SLOTS + slot] = clock64();
}
#define TRACE(slot) trace_ev(slot)
#else
#define TRACE(slot)
#endif
constexpr int BQ = 128;
constexpr int BKV = 64;
constexpr int DH = 64;
constexpr int MAX_KV_TILES = 12;            // T <= 768
constexpr int Q_BYTES = BQ * DH * 2;        // 16 KiB
constexpr int KV_TILE_BYTES = BKV * DH * 2; // 8 KiB
constexpr int BAR_BYTES = 512;
constexpr int THREADS = 256;                // warps 0-3 softmax, 4 TMA + S issuer, 5 P V issuer, 6-7 leftover query rows (CUDA cores)
constexpr int MAX_LEFT = 2;                 // T mod 128 <= 2 (BOS/EOS around L = 128 k residues): no tensor tile for them
constexpr int NSBUF = 3;                    // S / P buffers in TMEM
constexpr int TMEM_COLS = 256;              // S0 [0,64) S1 [64,128) S2 [128,192) O [192,256)
constexpr int COL_O = 192;
constexpr float RESCALE_LOG2 = 8.0f;        // lazy rescale threshold: p <= 2^8

struct Params {
    int B, T, H;
    int nq;                     // query tiles of 128 rows on the tensor path: ceil(T / 128), or floor when n_left > 0
    int n_left;                 // 0..MAX_LEFT trailing query rows done by warps 6-7 on CUDA cores
    int q_splits;               // CTAs per (sample, head): each takes a contiguous range of the query tiles (the last one
                                // also the trailing rows) and loads K/V for itself.  2 when the grid would otherwise be
                                // just over a multiple of the resident CTA count (13 samples: 312 CTAs on 296 slots = two
                                // rounds for 1.05 rounds of work; 624 half-length CTAs = three rounds of 0.55)
    int nkv;                    // ceil(T / 64)
    int tail_cols;              // width of the last kv tile: multiple of 16 in [16, 64]
    const __nv_bfloat16* qkv;   // [B*T, 3*H*64] (the leftover warps read their query rows directly)
    __nv_bfloat16* ctx;         // [B*T, H*64]
    float scale_log2;           // (1/sqrt(64)) * log2(e)
    const float* qk_rstd;       // q_ln / k_ln 1/std per token row: q rows at [0, M), k rows at [rstd_ld, rstd_ld + M); or null
    long long rstd_ld;
    int prefetch_stride;        // L2-prefetch the tiles of CTA blockIdx.x + prefetch_stride (0 = off)
};

// 1/std of every q row and k row from the QKV epilogue's per-128-column partial sums of squares
// ([M][2 * nspan]: q spans then k spans).  One thread per (operand, row).
__global__ void __launch_bounds__(256)
qk_rstd_kernel(const float* __restrict__ sumsq, float* __restrict__ rstd, long long M, long long ld, int nspan, float eps) {
    pdl_launch_dependents();
    pdl_wait();
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= 2 * M) return;
    const bool isk = idx >= M;
    const long long row = isk ? idx - M : idx;
    const float2* src = reinterpret_cast<const float2*>(sumsq + row * 2 * nspan + (isk ? nspan : 0));
    float sum = 0.f;
    for (int i = 0; i < (nspan >> 1); ++i) {
        const float2 v = __ldg(src + i);
        sum += v.x + v.y;
    }
    rstd[(isk ? ld : 0) + row] = rsqrtf(sum * (1.0f / static_cast<float>(nspan * 128)) + eps);
}

__host__ __device__ inline int kv_bytes(int nkv, int tail_cols) {
    return (nkv - 1) * KV_TILE_BYTES + tail_cols * 128;
}
__host__ __device__ inline int left_bytes(int nkv) { return MAX_LEFT * nkv * BKV * 4; }   // fp32 p per leftover row
__host__ __device__ inline int rstd_bytes(int nkv) { return 2 * nkv * BKV * 4; }          // 1/std of every k row, then q row
__host__ inline int smem_bytes(int nkv, int tail_cols) {
    return 1024 + 2 * Q_BYTES + 2 * kv_bytes(nkv, tail_cols) + BAR_BYTES + left_bytes(nkv) + rstd_bytes(nkv);
}
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// One trailing query row on CUDA cores (one warp), K and V read from the swizzled shared-memory
// tiles the tensor path uses.  T = L + 2 tokens puts exactly two rows (the last residue and EOS)
// past the last full 128-row tile whenever L is a multiple of 128 -- every configuration the path
// is quoted on -- and a tensor tile for them costs a third of the kernel at T = 258 (5 of 15
// MMA/softmax steps with 2 of 128 rows live).  Scores: lane = key (k = lane + 32 m), q in
// registers; P V: lane = two output dims, p broadcast from shared memory.  fp32 throughout.
__device__ __forceinline__ void leftover_row(const __nv_bfloat16* __restrict__ qrow, __nv_bfloat16* __restrict__ orow,
                                             const uint8_t* sK, const uint8_t* sV, float* pf, int T, float sc,
                                             uint64_t* k_ready, uint64_t* v_full, int nkv, int lane,
                                             const float* rstd_k) {
    // q: every lane needs all 64 dims -> 8 x 16-byte loads of the same 128-byte row
    float q[DH];
    {
        const uint4* src = reinterpret_cast<const uint4*>(qrow);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const uint4 v = __ldg(src + c);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                q[8 * c + 2 * e] = __uint_as_float(w[e] << 16);
                q[8 * c + 2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
            }
        }
    }
    for (int j = 0; j < nkv; ++j) mbar_wait(&k_ready[j], 0);
    const int nk = (T + 31) >> 5;                      // keys per lane
    float mx = -INFINITY;
    for (int m = 0; m < nk; ++m) {
        const int k = m * 32 + lane;
        const int kk = k < T ? k : T - 1;              // clamp: rows past T may not be loaded
        const uint8_t* rowp = sK + (kk >> 6) * KV_TILE_BYTES + (kk & 63) * 128;
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const uint4 v = *reinterpret_cast<const uint4*>(rowp + ((c ^ (kk & 7)) << 4));
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                a0 = fmaf(q[8 * c + 2 * e], __uint_as_float(w[e] << 16), a0);
                a1 = fmaf(q[8 * c + 2 * e + 1], __uint_as_float(w[e] & 0xffff0000u), a1);
            }
        }
        float sv = k < T ? (a0 + a1) * sc : -INFINITY;
        if (rstd_k != nullptr && k < T) sv *= rstd_k[k];
        pf[k] = sv;                                    // pf holds nkv * 64 >= nk * 32 floats
        mx = fmaxf(mx, sv);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float l = 0.f;
    for (int m = 0; m < nk; ++m) {
        const int k = m * 32 + lane;
        const float pv = fast_exp2(pf[k] - mx);        // exp2(-inf) = 0 for the padding keys
        pf[k] = pv;
        l += pv;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    __syncwarp();
    for (int j = 0; j < nkv; ++j) mbar_wait(&v_full[j], 0);
    // out[2 lane, 2 lane + 1] = sum_k p[k] V[k][2 lane, 2 lane + 1]
    int off[8];                                        // byte offset of my dim pair inside row r, r & 7 = i
#pragma unroll
    for (int i = 0; i < 8; ++i) off[i] = i * 128 + ((((lane >> 2) ^ i) << 4) | ((lane & 3) << 2));
    float o0 = 0.f, o1 = 0.f;
    const int k8 = T >> 3;
    for (int g = 0; g < k8; ++g) {                     // 8 keys per trip: rows 8 g .. 8 g + 7 of one tile
        const uint8_t* base = sV + (g >> 3) * KV_TILE_BYTES + (g & 7) * 1024;
        const float4 pa = *reinterpret_cast<const float4*>(pf + 8 * g);
        const float4 pb = *reinterpret_cast<const float4*>(pf + 8 * g + 4);
        const float pp[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t v = *reinterpret_cast<const uint32_t*>(base + off[i]);
            o0 = fmaf(pp[i], __uint_as_float(v << 16), o0);
            o1 = fmaf(pp[i], __uint_as_float(v & 0xffff0000u), o1);
        }
    }
    for (int k = k8 * 8; k < T; ++k) {
        const uint32_t v = *reinterpret_cast<const uint32_t*>(sV + (k >> 6) * KV_TILE_BYTES + (k & 63) * 128 +
                                                             ((((lane >> 2) ^ (k & 7)) << 4) | ((lane & 3) << 2)));
        o0 = fmaf(pf[k], __uint_as_float(v << 16), o0);
        o1 = fmaf(pf[k], __uint_as_float(v & 0xffff0000u), o1);
    }
    const float inv = 1.0f / l;
    reinterpret_cast<uint32_t*>(orow)[lane] = pack_bf16x2(o0 * inv, o1 * inv);
}

__global__ void __launch_bounds__(THREADS, 2)
attention_resident_kernel(const __grid_constant__ CUtensorMap tmQ,      // box [128 rows][64 cols]
                          const __grid_constant__ CUtensorMap tmKV,     // box [ 64 rows][64 cols]
                          const __grid_constant__ CUtensorMap tmKVt,    // box [tail rows][64 cols]
                          const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const int kvb = kv_bytes(p.nkv, p.tail_cols);
    uint8_t* sQ = smem;                                   // two query-tile buffers
    uint8_t* sK = sQ + 2 * Q_BYTES;
    uint8_t* sV = sK + kvb;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kvb);
    uint64_t* q_full = bars;                              // [2]
    uint64_t* s_full = bars + 2;                          // [3]  S issuer -> softmax
    uint64_t* p_full = bars + 5;                          // [3]  softmax warps (4 arrivals) -> P V issuer
    uint64_t* pv_done = bars + 8;                         // [3]  P V of the step that used this buffer retired -> S issuer, rescale, epilogue
    uint64_t* o_free = bars + 11;                         // 1    completes once per query tile
    uint64_t* k_full = bars + 12;                         // [MAX_KV_TILES], single use
    uint64_t* v_full = k_full + MAX_KV_TILES;             // [MAX_KV_TILES], single use
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(v_full + MAX_KV_TILES);
    float* left_p = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + BAR_BYTES);   // [MAX_LEFT][nkv * 64]
    float* rstd_k = left_p + MAX_LEFT * p.nkv * BKV;                                          // [nkv * 64]
    float* rstd_q = rstd_k + p.nkv * BKV;                                                     // [nkv * 64] (>= T)

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int split = blockIdx.x % p.q_splits;
    const int h = (blockIdx.x / p.q_splits) % p.H;
    const int b = blockIdx.x / (p.q_splits * p.H);
    const int D = p.H * DH;
    const int row0 = b * p.T;
    const int qt0 = split * p.nq / p.q_splits;                            // first query tile of this CTA
    const int nq = (split + 1) * p.nq / p.q_splits - qt0, nkv = p.nkv;   // its query tiles
    const int n_left = split == p.q_splits - 1 ? p.n_left : 0;           // trailing rows go with the last range
    const bool fused_ln = p.qk_rstd != nullptr;
    uint64_t* k_ready = k_full;                           // K tiles are consumed as the TMA loads leave them

#if ATTN_DBG & 32
    if (threadIdx.x == 0 && blockIdx.x < 4096) {
        g_attn_trace[blockIdx.x * TRACE_SLOTS + 0] = static_cast<long long>(global_timer_ns());
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        g_attn_trace[blockIdx.x * TRACE_SLOTS + 1] = smid;
        TRACE(2);
    }
#endif
    if (warp == 4) {
        if (lane == 0) {
            tma_prefetch_desc(&tmQ);
            tma_prefetch_desc(&tmKV);
            tma_prefetch_desc(&tmKVt);
            for (int s = 0; s < 2; ++s) mbar_init(&q_full[s], 1);
            for (int s = 0; s < NSBUF; ++s) {
                mbar_init(&s_full[s], 1);
                mbar_init(&p_full[s], 4);              // one arrival per softmax warp
                mbar_init(&pv_done[s], 1);
            }
            mbar_init(o_free, 4);
            for (int j = 0; j < nkv; ++j) {
                mbar_init(&k_full[j], 1);
                mbar_init(&v_full[j], 1);
            }
            fence_barrier_init();
        }
        __syncwarp();
        // All tile loads of this CTA, one per lane, before the CTA-wide barrier: lane 0 = Q tile 0, lanes 1..nkv =
        // K tiles, then Q tile 1 (if any), then the V tiles.  The same lanes later prefetch the same tiles of the
        // CTA that will replace this one into L2.
        pdl_wait();
        const int has_q1 = nq > 1 ? 1 : 0;
        if (lane < 2 * nkv + 1 + has_q1) {
            const bool is_q = lane == 0 || (has_q1 && lane == nkv + 1);
            const bool is_k = lane >= 1 && lane <= nkv;
            const int j = is_k ? lane - 1 : lane - nkv - 1 - has_q1;       // kv tile (K or V lanes)
            const int qt = lane == 0 ? 0 : 1;
            const bool last = !is_q && j == nkv - 1;
            uint64_t* bar = is_q ? &q_full[qt] : is_k ? &k_full[j] : &v_full[j];
            uint8_t* dst = is_q ? sQ + qt * Q_BYTES : (is_k ? sK : sV) + j * KV_TILE_BYTES;
            const CUtensorMap* tm = is_q ? &tmQ : last ? &tmKVt : &tmKV;
            const int col = is_q ? h * DH : (is_k ? D : 2 * D) + h * DH;
            const int rowoff = is_q ? (qt0 + qt) * BQ : j * BKV;
            mbar_arrive_expect_tx(bar, is_q ? Q_BYTES : last ? p.tail_cols * 128 : KV_TILE_BYTES);
            tma_load_2d(dst, tm, bar, col, row0 + rowoff);
        }
        if (lane == 0) TRACE(3);
    }
    if (warp == 5) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    // 1/std of this sample's k rows and q rows (softmax + trailing-row warps): the global loads are issued here,
    // before the set-up barrier, and land in shared memory after it
    constexpr int RSTD_PER_THREAD = MAX_KV_TILES * BKV / 192;            // 4
    const int rtid = warp < 4 ? threadIdx.x : threadIdx.x - 64;         // 0..191 over warps 0-3, 6-7
    float rk[RSTD_PER_THREAD], rq[RSTD_PER_THREAD];
    if (warp != 4) pdl_wait();
    if (fused_ln && warp != 4 && warp != 5) {
#pragma unroll
        for (int u = 0; u < RSTD_PER_THREAD; ++u) {
            const int idx = rtid + u * 192;
            const bool in = idx < p.T;
            rk[u] = in ? __ldg(p.qk_rstd + p.rstd_ld + row0 + idx) : 0.f;       // 0: padding keys of the last tile
            rq[u] = in ? __ldg(p.qk_rstd + row0 + idx) : 1.f;
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();
    if (fused_ln && warp != 4 && warp != 5) {
#pragma unroll
        for (int u = 0; u < RSTD_PER_THREAD; ++u) {
            const int idx = rtid + u * 192;
            if (idx < nkv * BKV) {
                rstd_k[idx] = rk[u];
                rstd_q[idx] = rq[u];
            }
        }
        named_bar_sync(2, 192);
    }

    if (warp == 4) {
        // ===================== S = Q K^T issuer =====================
        // The whole warp walks the schedule (warp-uniform control flow: counters and descriptors stay in uniform
        // registers); one elected lane issues the tcgen05 instructions.  S runs up to three steps ahead of the
        // softmax: buffer s_i % 3 is free once the P V that read P from it (step s_i - 3) has retired.
        const uint32_t idesc_s_full = umma_idesc_bf16(BQ, BKV, 0);
        const uint32_t idesc_s_tail = umma_idesc_bf16(BQ, p.tail_cols, 0);
        const uint64_t desc_q0 = umma_desc_sw128(smem_u32(sQ), 16, 1024);
        const uint64_t desc_k0 = umma_desc_sw128(smem_u32(sK), 16, 1024);
        int buf = 0, ph = 0, s_i = 0;
        for (int s_qt = 0; s_qt < nq; ++s_qt) {
            for (int s_j = 0; s_j < nkv; ++s_j, ++s_i) {
                if (s_j == 0) mbar_wait(&q_full[s_qt & 1], (s_qt >> 1) & 1);
                if (s_qt == 0) mbar_wait(&k_ready[s_j], 0);
                if (lane == 0 && s_i < 8) TRACE(8 + s_i);               // operands in
                if (s_i >= NSBUF) mbar_wait(&pv_done[buf], ph ^ 1);
                if (lane == 0 && s_i < 8) TRACE(16 + s_i);              // buffer free
                tcgen05_fence_after();
                const uint64_t qdesc = desc_q0 + static_cast<uint64_t>((s_qt & 1) * (Q_BYTES >> 4));
                const uint64_t kdesc = desc_k0 + static_cast<uint64_t>(s_j * (KV_TILE_BYTES >> 4));
                const uint32_t idesc = s_j == nkv - 1 ? idesc_s_tail : idesc_s_full;
                const uint32_t ts = tmem_base + buf * BKV;
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < DH / 16; ++k)
                        umma_bf16_ss(ts, qdesc + 2 * k, kdesc + 2 * k, idesc, k != 0 ? 1u : 0u);
                    umma_commit(&s_full[buf]);
                }
                __syncwarp();
                if (lane == 0 && s_i < 8) TRACE(24 + s_i);              // S issued
                if (++buf == NSBUF) { buf = 0; ph ^= 1; }
            }
        }
    } else if (warp == 5) {
        // ===================== O += P V issuer (+ the loads of query tiles 2, 3, ...) =====================
        // Two issuing warps because the ISSUE of these small MMAs, not their execution, bounded the step:
        // tools/mma_bench.cu: ~45 clk per M128 N64 K16 instruction and ~190 clk per tcgen05.commit in the issuing
        // thread against 32 clk of tensor-pipe time; one warp issuing P V_i, S_{i+2} and a commit needed ~1100 clk
        // per step (clock trace, tools/attn_timeline.py) with the softmax warps waiting behind it.
        constexpr uint32_t idesc_pv = umma_idesc_bf16(BQ, DH, 1);       // P V : V is MN-major
        const uint64_t desc_v0 = umma_desc_sw128(smem_u32(sV), 16, 1024);
        const uint32_t tmem_o = tmem_base + COL_O;
        int buf = 0, ph = 0;
        for (int qt = 0; qt < nq; ++qt) {
            for (int j = 0; j < nkv; ++j) {
                if (qt == 0) mbar_wait(&v_full[j], 0);
                mbar_wait(&p_full[buf], ph);
                if (lane == 0 && qt * nkv + j < 8) TRACE(32 + qt * nkv + j);   // P ready
                if (j == 0 && qt > 0) mbar_wait(o_free, (qt - 1) & 1);
                tcgen05_fence_after();
                // V tile [kv rows][64 d] is an MN-major B operand: 128-byte rows along N = d,
                // 8-row (k) groups 1024 B apart; one UMMA K-step (16 kv rows) = 2048 B.
                const uint64_t vdesc = desc_v0 + static_cast<uint64_t>(j * (KV_TILE_BYTES >> 4));
                const uint32_t tp = tmem_base + buf * BKV;              // P: bf16 pairs, 8 columns per K-step
                const bool last = j == nkv - 1;
                if (elect_one()) {
                    if (!last) {
#pragma unroll
                        for (int k = 0; k < BKV / 16; ++k)
                            umma_bf16_ts(tmem_o, tp + 8 * k, vdesc + 128 * k, idesc_pv, (j | k) != 0 ? 1u : 0u);
                    } else {
                        const int ksteps = p.tail_cols >> 4;
                        for (int k = 0; k < ksteps; ++k)
                            umma_bf16_ts(tmem_o, tp + 8 * k, vdesc + 128 * k, idesc_pv, (j | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&pv_done[buf]);
                    // all S of query tile qt have been consumed by the softmax (p_full of its last step): its
                    // buffer takes query tile qt + 2
                    if (last && qt + 2 < nq) {
                        uint64_t* bar = &q_full[qt & 1];
                        mbar_arrive_expect_tx(bar, Q_BYTES);
                        tma_load_2d(sQ + (qt & 1) * Q_BYTES, &tmQ, bar, h * DH, row0 + (qt0 + qt + 2) * BQ);
                    }
                }
                __syncwarp();
                if (lane == 0 && qt * nkv + j < 8) TRACE(40 + qt * nkv + j);   // P V issued
                if (++buf == NSBUF) { buf = 0; ph ^= 1; }
            }
        }
    } else if (warp >= 6) {
        // ===================== trailing query rows past the last full tile =====================
        const int lw = warp - 6;
        if (lw < n_left) {
            const int t = p.nq * BQ + lw;
            float sc = p.scale_log2;
            if (fused_ln) sc *= rstd_q[t];
            leftover_row(p.qkv + static_cast<long long>(row0 + t) * 3 * D + h * DH,
                         p.ctx + static_cast<long long>(row0 + t) * D + h * DH, sK, sV, left_p + lw * nkv * BKV, p.T,
                         sc, k_ready, v_full, nkv, lane, fused_ln ? rstd_k : nullptr);
        }
        // L2 prefetch of the tiles of the CTA that will take this one's place, once my own have landed
        const int next = blockIdx.x + p.prefetch_stride;
        const int has_q1 = nq > 1 ? 1 : 0;
        if (lw == 0 && p.prefetch_stride > 0 && next < static_cast<int>(gridDim.x) && lane < 2 * nkv + 1 + has_q1) {
            mbar_wait(&v_full[nkv - 1], 0);
            const int nsplit = next % p.q_splits;
            const int nh = (next / p.q_splits) % p.H;
            const int nrow0 = (next / (p.q_splits * p.H)) * p.T;
            const int nqt0 = nsplit * p.nq / p.q_splits;
            const bool is_q = lane == 0 || (has_q1 && lane == nkv + 1);
            const bool is_k = lane >= 1 && lane <= nkv;
            const int j = is_k ? lane - 1 : lane - nkv - 1 - has_q1;
            const int qt = lane == 0 ? 0 : 1;
            const bool last = !is_q && j == nkv - 1;
            const CUtensorMap* tm = is_q ? &tmQ : last ? &tmKVt : &tmKV;
            const int col = is_q ? nh * DH : (is_k ? D : 2 * D) + nh * DH;
            const int rowoff = is_q ? (nqt0 + qt) * BQ : j * BKV;
            tma_prefetch_l2_2d(tm, col, nrow0 + rowoff);
        }
    } else {
        // ===================== softmax / output warps: thread = query row =====================
        const int r = threadIdx.x;                                   // 0..127 == TMEM lane
        const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
        const uint32_t t_o = tmem_base + lane_addr + COL_O;
        float sc = p.scale_log2;              // per query row once q_ln's 1/std is folded in (set per query tile)
        float thresh = RESCALE_LOG2 / sc;
        // out[row] = O / l for query tile qt (after its last PV has retired), then free O
        auto epilogue = [&](int qt, float l, int ebuf, int eph) {
            mbar_wait(&pv_done[ebuf], eph);                          // the last P V of query tile qt
            tcgen05_fence_after();
            if ((qt0 + qt) * BQ + warp * 32 < p.T) {
                const float inv = 1.0f / l;
                const int t = (qt0 + qt) * BQ + r;
                uint4* dst = reinterpret_cast<uint4*>(p.ctx + static_cast<long long>(row0 + t) * D + h * DH);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t o[16];
                    tmem_ld_32x32b_x16(t_o + c * 16, o);
                    tmem_ld_wait();
                    if (t < p.T) {
#pragma unroll
                        for (int g = 0; g < 2; ++g)
                            dst[c * 2 + g] = make_uint4(
                                pack_bf16x2(__uint_as_float(o[8 * g]) * inv, __uint_as_float(o[8 * g + 1]) * inv),
                                pack_bf16x2(__uint_as_float(o[8 * g + 2]) * inv, __uint_as_float(o[8 * g + 3]) * inv),
                                pack_bf16x2(__uint_as_float(o[8 * g + 4]) * inv, __uint_as_float(o[8 * g + 5]) * inv),
                                pack_bf16x2(__uint_as_float(o[8 * g + 6]) * inv, __uint_as_float(o[8 * g + 7]) * inv));
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(o_free);
        };

        float m_run = 0.f;                    // running max of the raw scores (set at j == 0)
        float l_run = 0.f, l_prev = 0.f;      // running row sum (relative to m_run); previous tile's final sum
        // Software pipeline over the steps: the TMEM loads of step i + 1's scores are issued into the score
        // registers as soon as step i's exponentials have left them -- before step i's P stores are waited for,
        // fenced and signalled -- so the barrier round trip (~350 clk from an arrive to the next successful wait,
        // clock trace) and the TMEM load latency of every step hide behind the tail of the previous one.  ONE
        // tcgen05.ld site in the loop: the registers it writes must not be touched before the wait::ld at the top
        // of the next trip.
        uint32_t s[64];
        const int nsteps = nq * nkv;
        int buf = 0, ph = 0;                  // S / P buffer of the step being computed and the parity of its use count
        int lbuf = 0, lph = 0;                // ... of the step being loaded (one ahead)
        int qt = 0, j = -1;                   // the step being computed; (0, -1) = none yet
#pragma unroll 1
        for (int step = -1; step < nsteps; ++step) {
            const bool active = (qt0 + qt) * BQ + warp * 32 < p.T;   // warp-uniform
            const uint32_t t_s = tmem_base + lane_addr + buf * BKV;
            const int pbuf = buf == 0 ? NSBUF - 1 : buf - 1, pph = buf == 0 ? ph ^ 1 : ph;   // the previous step's
            if (step >= 0) {
                if (j == 0) {                 // new query tile: keep the finished tile's row sum for its epilogue
                    l_prev = l_run;
                    l_run = 0.f;
                    if (fused_ln) {
                        const int t = (qt0 + qt) * BQ + r;
                        sc = p.scale_log2 * (t < p.T ? rstd_q[t] : 1.0f);
                        thresh = RESCALE_LOG2 / sc;
                    }
                }
                if (active && !(ATTN_DBG & 4)) {
                    const bool last = j == nkv - 1;
                    const int nch = (last ? p.tail_cols : BKV) >> 4;     // 16-column chunks
                    tmem_ld_wait();
                    if (fused_ln) {
                        // k_ln's 1/std: one factor per score column (= key row), shared by all query rows
                        const float4* rk4 = reinterpret_cast<const float4*>(rstd_k + j * BKV);
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if (c < nch) {
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float4 rk = rk4[c * 4 + e];
                                    float a0, a1, a2, a3;
                                    fmul2(a0, a1, __uint_as_float(s[c * 16 + 4 * e]), __uint_as_float(s[c * 16 + 4 * e + 1]), rk.x, rk.y);
                                    fmul2(a2, a3, __uint_as_float(s[c * 16 + 4 * e + 2]), __uint_as_float(s[c * 16 + 4 * e + 3]), rk.z, rk.w);
                                    s[c * 16 + 4 * e] = __float_as_uint(a0);
                                    s[c * 16 + 4 * e + 1] = __float_as_uint(a1);
                                    s[c * 16 + 4 * e + 2] = __float_as_uint(a2);
                                    s[c * 16 + 4 * e + 3] = __float_as_uint(a3);
                                }
                            }
                    }
                    float mx = -INFINITY;
                    if (last) {
                        const int valid = p.T - j * BKV;             // >= 1 valid kv columns in this tile
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if (c < nch) {
#pragma unroll
                                for (int e = 0; e < 16; ++e) {
                                    const float a = (c * 16 + e < valid) ? __uint_as_float(s[c * 16 + e]) : -INFINITY;
                                    s[c * 16 + e] = __float_as_uint(a);
                                    mx = fmaxf(mx, a);
                                }
                            }
                    } else {
                        // four independent FMNMX3 chains of 8 (a single 64-deep chain is 64 x 4 clk of latency)
                        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                        for (int e = 0; e < 64; e += 8)
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                m4[k] = fmax3(m4[k], __uint_as_float(s[e + 2 * k]), __uint_as_float(s[e + 2 * k + 1]));
                        mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
                    }
                    if (j == 0) {
                        m_run = mx;
                    } else {
                        const bool need = mx > m_run + thresh;
                        if (__any_sync(0xffffffffu, need)) {
                            // raise the running max: rescale O and l in TMEM once PV_{i-1} has retired
                            const float m_new = need ? mx : m_run;
                            const float f = fast_exp2((m_run - m_new) * sc);
                            mbar_wait(&pv_done[pbuf], pph);          // P V of the previous step has retired
                            tcgen05_fence_after();
                            l_run *= f;
#pragma unroll 1
                            for (int c = 0; c < 4; ++c) {
                                uint32_t o[16];
                                tmem_ld_32x32b_x16(t_o + c * 16, o);
                                tmem_ld_wait();
#pragma unroll
                                for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * f);
                                tmem_st_32x32b_x16(t_o + c * 16, o);
                            }
                            tmem_st_wait();
                            m_run = m_new;
                        }
                    }
                    const float nm = -m_run * sc;
                    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (c < nch) {
                            uint32_t pk[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                float a0, a1;
                                ffma2(a0, a1, __uint_as_float(s[c * 16 + 2 * e]), __uint_as_float(s[c * 16 + 2 * e + 1]), sc, sc, nm, nm);
                                const float p0 = fast_exp2(a0), p1 = fast_exp2(a1);
                                fadd2(rs0, rs1, rs0, rs1, p0, p1);
                                pk[e] = pack_bf16x2(p0, p1);
                            }
                            tmem_st_32x32b_x8(t_s + c * 8, pk);      // P over the S buffer: 2 bf16 per column
                        }
                    l_run += rs0 + rs1;
                }
            }
            // ---- the next step's scores: wait for its S and start the TMEM loads ----
            if (step + 1 < nsteps) {
                int lqt = qt, lj = j + 1;
                if (lj == nkv) { lj = 0; ++lqt; }
                mbar_wait(&s_full[lbuf], lph);
                if (threadIdx.x == 0 && step + 1 < 8) TRACE(48 + step + 1);   // S seen by softmax warp 0
                tcgen05_fence_after();
                if ((qt0 + lqt) * BQ + warp * 32 < p.T && !(ATTN_DBG & 4)) {
                    const int nch = (lj == nkv - 1 ? p.tail_cols : BKV) >> 4;
                    const uint32_t t_l = tmem_base + lane_addr + lbuf * BKV;
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (c < nch) tmem_ld_32x32b_x16(t_l + c * 16, s + c * 16);
                }
                if (++lbuf == NSBUF) { lbuf = 0; lph ^= 1; }
            }
            if (step >= 0) {
                if (active && !(ATTN_DBG & 4)) tmem_st_wait();
                // one arrival per warp: 128 per-thread arrivals are 128 serialised shared-memory atomics
                // on the critical path of every step
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[buf]);
                if (threadIdx.x == 0 && step < 8) TRACE(56 + step);   // softmax warp 0 done
                if (j == 0 && qt > 0) epilogue(qt - 1, l_prev, pbuf, pph);   // deferred: overlaps this tile's first MMAs
                if (++buf == NSBUF) { buf = 0; ph ^= 1; }
            }
            if (++j == nkv && step + 1 < nsteps) { j = 0; ++qt; }
        }
        epilogue(nq - 1, l_run, buf == 0 ? NSBUF - 1 : buf - 1, buf == 0 ? ph ^ 1 : ph);
        if (threadIdx.x == 0) TRACE(4);
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 5) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
        if (lane == 0) TRACE(5);
    }
}

}  // namespace attn2
}  // namespace esmdiff
