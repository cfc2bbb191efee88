// Non-causal multi-head attention, d_head = 64, with the K and V of one (sample, head) RESIDENT in
// shared memory: the variant for the sequence lengths the sampling path is quoted on
// (T = L + 2 <= 766; T = 258 keeps two CTAs per SM).  Replaces F.scaled_dot_product_attention on
// the reference path (esm MultiHeadAttention.forward with seq_id None).
//
// One CTA = one (sample b, head h); it walks ALL query tiles of 128 rows, so K/V are read from
// L2/HBM once per (b, h) instead of once per query tile.  Per step (query tile qt, kv tile j):
//   MMA warp  : S = Q_qt K_j^T     UMMA 128 x ncols x 16 (x4), fp32 into one of two TMEM S buffers,
//               issued two steps ahead of the softmax
//               O (+)= P V_j       UMMA with A = P FROM TMEM (bf16 pairs written over the S buffer
//                                  by the softmax threads), B = V_j as MN-major smem operand
//   4 softmax warps (thread = query row = TMEM lane): row max of the tile; the running max is only
//               raised when the tile max exceeds it by more than 2^8 (then O is rescaled in TMEM and
//               the row sum in its register -- rare after the first tile), p = exp2(s*scale - m),
//               packed to bf16 and stored back to TMEM.  No shared-memory round trip for P, no
//               per-tile read of O.
//   The last kv tile is only as wide as needed (multiple of 16 columns): T = 258 costs 4 x 64 + 16
//   columns, not 5 x 64.  Warps whose 32 query rows are all >= T skip the softmax (the 2-row tail
//   tile of T = 258 keeps one warp busy, not four).
// Input  qkv : bf16 [M = B*T, 3*D]  (q | k | v, each D = H*64).  Two forms of q, k:
//   qk_sumsq == null : q, k already LayerNormed + RoPE'd (stand-alone ew::qk_layernorm_rope_kernel)
//   qk_sumsq != null : q' = rope(gamma_q (q - mean q)), k' likewise, NOT yet divided by their row
//                      standard deviation, plus the per-row partial sums of (q - mean q)^2 and
//                      (k - mean k)^2 the QKV GEMM epilogue left (gemm.cuh EPI_QKV_ROPE_LN).  The
//                      missing factors are per-row scalars, both applied to the fp32 scores: rstd_q[i]
//                      goes into the softmax scale of query row i (thread = row), rstd_k[j] multiplies
//                      score column j (one packed multiply per two scores, the factors read as
//                      shared-memory broadcasts).  q and k are therefore rounded to bf16 exactly once.
//                      (First version: K rows rescaled in shared memory before the first S MMA -- a
//                      second rounding of k, a proxy fence per tile and the whole pass on the critical
//                      path of every CTA: +15 % kernel time; this form: see DESIGN.md.)
// Output ctx : bf16 [M, D]
#pragma once
#include "ptx.cuh"

namespace esmdiff {
namespace attn2 {

constexpr int BQ = 128;
constexpr int BKV = 64;
constexpr int DH = 64;
constexpr int MAX_KV_TILES = 12;            // T <= 768
constexpr int Q_BYTES = BQ * DH * 2;        // 16 KiB
constexpr int KV_TILE_BYTES = BKV * DH * 2; // 8 KiB
constexpr int BAR_BYTES = 512;
constexpr int THREADS = 256;                // warps 0-3 softmax, 4 TMA, 5 MMA, 6-7 leftover query rows (CUDA cores)
constexpr int MAX_LEFT = 2;                 // T mod 128 <= 2 (BOS/EOS around L = 128 k residues): no tensor tile for them
constexpr int TMEM_COLS = 256;              // S0 [0,64) S1 [64,128) O [128,192)
constexpr int COL_O = 128;
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
    const float* qk_sumsq;      // [B*T][2 * nspan]: q spans then k spans (gemm.cuh EPI_QKV_ROPE_LN), or null
    int nspan;                  // D / 128 partial sums per row and operand (<= 12)
    float ln_eps;               // q_ln / k_ln epsilon
};

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
    uint64_t* q_empty = bars + 2;                         // [2]
    uint64_t* s_full = bars + 4;                          // [2]  MMA -> softmax
    uint64_t* p_full = bars + 6;                          // [2]  softmax warps (4 arrivals) -> MMA
    uint64_t* pv_done = bars + 8;                         // 1    P V of step nsteps-2 retired (no S follows it)
    uint64_t* o_free = bars + 9;                          // 1    completes once per query tile
    uint64_t* o_full = bars + 10;                         // 1    last PV of a query tile retired
    uint64_t* k_full = bars + 11;                         // [MAX_KV_TILES], single use
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
    const int nsteps = nq * nkv;
    const bool fused_ln = p.qk_sumsq != nullptr;
    uint64_t* k_ready = k_full;                           // K tiles are consumed as the TMA loads leave them

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmKV);
        tma_prefetch_desc(&tmKVt);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&q_full[s], 1);
            mbar_init(&q_empty[s], 1);
            mbar_init(&s_full[s], 1);
            mbar_init(&p_full[s], 4);              // one arrival per softmax warp
        }
        mbar_init(pv_done, 1);
        mbar_init(o_free, 4);
        mbar_init(o_full, 1);
        for (int j = 0; j < nkv; ++j) {
            mbar_init(&k_full[j], 1);
            mbar_init(&v_full[j], 1);
        }
        fence_barrier_init();
    }
    if (warp == 5) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();
    if (warp != 5) pdl_wait();             // every role but the MMA issuer touches global memory

    if (fused_ln) {
        // 1/std of every k row and q row of this (sample, head)'s sample from the QKV epilogue's partial
        // sums of squares, into shared memory, by the six warps that have nothing to do until the first
        // tiles land.  The TMA warp holds its loads back until the first batch of these small loads is in
        // flight (named barrier 1): issued behind 82 KB of tile traffic per CTA they came back ~5k clk
        // late, and the first S MMA needs the k factors (ncu r2c: +8k clk per CTA).
        if (warp != 4 && warp != 5) {
            const int tid = warp < 4 ? threadIdx.x : threadIdx.x - 64;       // 0..191
            const int n = 2 * p.T;                                             // k rows, then q rows
            const int half = p.nspan >> 1;
            const float inv_n = 1.0f / static_cast<float>(p.nspan * 128);
            bool released = false;
#pragma unroll 1
            for (int base = 0; base < n; base += 4 * 192) {
                float2 raw[4][6];
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int idx = base + b * 192 + tid;
                    const bool isq = idx >= p.T;
                    const int row = isq ? idx - p.T : idx;
                    const float2* src = reinterpret_cast<const float2*>(
                        p.qk_sumsq + static_cast<long long>(row0 + row) * 2 * p.nspan + (isq ? 0 : p.nspan));
#pragma unroll
                    for (int i = 0; i < 6; ++i)
                        raw[b][i] = (idx < n && i < half) ? __ldg(src + i) : make_float2(0.f, 0.f);
                }
                if (!released) {
                    named_bar_arrive(1, 224);
                    released = true;
                }
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int idx = base + b * 192 + tid;
                    float sum = 0.f;
#pragma unroll
                    for (int i = 0; i < 6; ++i) sum += raw[b][i].x + raw[b][i].y;
                    if (idx < n) (idx >= p.T ? rstd_q[idx - p.T] : rstd_k[idx]) = rsqrtf(sum * inv_n + p.ln_eps);
                }
            }
            for (int idx = p.T + tid; idx < p.nkv * BKV; idx += 192) rstd_k[idx] = 0.f;      // padding keys of the last tile
            named_bar_sync(2, 192);
        } else if (warp == 4) {
            named_bar_sync(1, 224);
        }
    }

    if (warp == 4) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            auto load_q = [&](int qt) {
                uint64_t* bar = &q_full[qt & 1];
                mbar_arrive_expect_tx(bar, Q_BYTES);
                tma_load_2d(sQ + (qt & 1) * Q_BYTES, &tmQ, bar, h * DH, row0 + (qt0 + qt) * BQ);
            };
            auto load_kv = [&](uint8_t* dst, uint64_t* bar, int col, int j) {
                const bool last = j == nkv - 1;
                mbar_arrive_expect_tx(bar, last ? p.tail_cols * 128 : KV_TILE_BYTES);
                tma_load_2d(dst + j * KV_TILE_BYTES, last ? &tmKVt : &tmKV, bar, col, row0 + j * BKV);
            };
            load_q(0);
            for (int j = 0; j < nkv; ++j) load_kv(sK, &k_full[j], D + h * DH, j);
            if (nq > 1) load_q(1);
            for (int j = 0; j < nkv; ++j) load_kv(sV, &v_full[j], 2 * D + h * DH, j);
            for (int qt = 2; qt < nq; ++qt) {
                mbar_wait(&q_empty[qt & 1], ((qt >> 1) - 1) & 1);
                load_q(qt);
            }
        }
    } else if (warp == 5) {
        // ===================== MMA issuer =====================
        // The whole warp walks the schedule (warp-uniform control flow: counters and descriptors
        // stay in uniform registers); one elected lane issues the tcgen05 instructions.  The
        // first cut did this from a single thread with per-step integer divisions and took
        // ~2.5k cycles per step -- the issuing thread, not MUFU or the tensor pipe, was the
        // bottleneck (ncu r1c: softmax warps 33 % stalled on s_full).
        constexpr uint32_t idesc_pv = umma_idesc_bf16(BQ, DH, 1);       // P V : V is MN-major
        const uint32_t idesc_s_full = umma_idesc_bf16(BQ, BKV, 0);
        const uint32_t idesc_s_tail = umma_idesc_bf16(BQ, p.tail_cols, 0);
        const uint64_t desc_q0 = umma_desc_sw128(smem_u32(sQ), 16, 1024);
        const uint64_t desc_k0 = umma_desc_sw128(smem_u32(sK), 16, 1024);
        const uint64_t desc_v0 = umma_desc_sw128(smem_u32(sV), 16, 1024);
        const uint32_t tmem_o = tmem_base + COL_O;
        int s_i = 0, s_qt = 0, s_j = 0;                                 // next S = Q K^T to issue
        auto issue_s = [&]() {
            if (s_j == 0) mbar_wait(&q_full[s_qt & 1], (s_qt >> 1) & 1);
            if (s_qt == 0) mbar_wait(&k_ready[s_j], 0);
            tcgen05_fence_after();
            const uint64_t qdesc = desc_q0 + static_cast<uint64_t>((s_qt & 1) * (Q_BYTES >> 4));
            const uint64_t kdesc = desc_k0 + static_cast<uint64_t>(s_j * (KV_TILE_BYTES >> 4));
            const uint32_t idesc = s_j == nkv - 1 ? idesc_s_tail : idesc_s_full;
            const uint32_t ts = tmem_base + (s_i & 1) * BKV;
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < DH / 16; ++k)
                    umma_bf16_ss(ts, qdesc + 2 * k, kdesc + 2 * k, idesc, k != 0 ? 1u : 0u);
                umma_commit(&s_full[s_i & 1]);
                if (s_j == nkv - 1) umma_commit(&q_empty[s_qt & 1]);
            }
            __syncwarp();
            ++s_i;
            if (++s_j == nkv) { s_j = 0; ++s_qt; }
        };
        issue_s();
        if (nsteps > 1) issue_s();
        int i = 0;
        for (int qt = 0; qt < nq; ++qt) {
            for (int j = 0; j < nkv; ++j, ++i) {
                if (qt == 0) mbar_wait(&v_full[j], 0);
                mbar_wait(&p_full[i & 1], (i >> 1) & 1);
                if (j == 0 && qt > 0) mbar_wait(o_free, (qt - 1) & 1);
                tcgen05_fence_after();
                // V tile [kv rows][64 d] is an MN-major B operand: 128-byte rows along N = d,
                // 8-row (k) groups 1024 B apart; one UMMA K-step (16 kv rows) = 2048 B.
                const uint64_t vdesc = desc_v0 + static_cast<uint64_t>(j * (KV_TILE_BYTES >> 4));
                const uint32_t tp = tmem_base + (i & 1) * BKV;          // P: bf16 pairs, 8 columns per K-step
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
                    // every tcgen05.commit stalls this thread ~200 clk (tools/mma_bench.cu): the "P V of
                    // step i retired" signal the rare rescale path needs rides on the commit of S_{i+2},
                    // issued right behind it; only the P Vs that no S follows commit their own
                    if (last) umma_commit(o_full);
                    else if (s_i >= nsteps) umma_commit(pv_done);
                }
                __syncwarp();
                if (s_i < nsteps) issue_s();              // overwrites P_i's buffer: ordered after PV_i
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
    } else {
        // ===================== softmax / output warps: thread = query row =====================
        const int r = threadIdx.x;                                   // 0..127 == TMEM lane
        const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
        const uint32_t t_o = tmem_base + lane_addr + COL_O;
        float sc = p.scale_log2;              // per query row once q_ln's 1/std is folded in (set per query tile)
        float thresh = RESCALE_LOG2 / sc;
        // out[row] = O / l for query tile qt (after its last PV has retired), then free O
        auto epilogue = [&](int qt, float l) {
            mbar_wait(o_full, qt & 1);
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
#pragma unroll 1
        for (int qt = 0; qt < nq; ++qt) {
            const bool active = (qt0 + qt) * BQ + warp * 32 < p.T;   // warp-uniform
            for (int j = 0; j < nkv; ++j) {
                const int i = qt * nkv + j;
                const uint32_t t_s = tmem_base + lane_addr + (i & 1) * BKV;
                if (j == 0) {                 // new query tile: keep the finished tile's row sum for its epilogue
                    l_prev = l_run;
                    l_run = 0.f;
                    if (fused_ln) {
                        const int t = (qt0 + qt) * BQ + r;
                        sc = p.scale_log2 * (t < p.T ? rstd_q[t] : 1.0f);
                        thresh = RESCALE_LOG2 / sc;
                    }
                }
                mbar_wait(&s_full[i & 1], (i >> 1) & 1);
                tcgen05_fence_after();
                if (active) {
                    const bool last = j == nkv - 1;
                    const int nch = (last ? p.tail_cols : BKV) >> 4;     // 16-column chunks
                    uint32_t s[64];
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (c < nch) tmem_ld_32x32b_x16(t_s + c * 16, s + c * 16);
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
                            // P V of step i-1 retired: implied by the commit of S_{i+1} (issued after it);
                            // the very last step has no S_{i+1} and waits for the single pv_done commit
                            if (i + 1 < nsteps) mbar_wait(&s_full[(i + 1) & 1], ((i + 1) >> 1) & 1);
                            else mbar_wait(pv_done, 0);
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
                    tmem_st_wait();
                }
                // one arrival per warp: 128 per-thread arrivals are 128 serialised shared-memory atomics
                // on the critical path of every step
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[i & 1]);
                if (j == 0 && qt > 0) epilogue(qt - 1, l_prev);   // deferred: overlaps this tile's first MMAs
            }
        }
        epilogue(nq - 1, l_run);
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 5) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace attn2
}  // namespace esmdiff
