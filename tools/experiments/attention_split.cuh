// Resident-K/V attention with the softmax of every 128 x 64 score tile split over TWO threads per
// query row (attention_resident.cuh is the one-thread-per-row form; same inputs, same outputs).
//
// Why: at d_head = 64 the kernel is bound by the softmax warps (one ex2 per 256 tensor FLOP; MUFU
// 16 / clk / SM), and attention_resident runs only 2 x 4 of them per SM -- two per scheduler, each a
// dependent chain tcgen05.ld -> max -> ex2 -> pack -> tcgen05.st.  ncu: XU pipe 31 %, tensor pipe
// 15 %, issue 44 %.  Here eight softmax warps per CTA (16 per SM, four per scheduler) interleave
// twice as many independent chains:
//   warp w, w + 4 (w = 0..3) own the same 32 TMEM lanes (query rows); warp w takes score columns
//   [0, 32) of every kv tile, warp w + 4 columns [32, 64).
// The two halves never synchronise inside the kv loop: each keeps ITS OWN running maximum and row
// sum and accumulates into ITS OWN output accumulator in TMEM
//   O_a += P[:, 0:32] V[0:32, :]      O_b += P[:, 32:64] V[32:64, :]      (2 UMMA K-steps each)
// and the two partial softmaxes are merged once per query tile in the epilogue:
//   m = max(m_a, m_b),  out = (O_a 2^(m_a - m) + O_b 2^(m_b - m)) / (l_a 2^(m_a - m) + l_b 2^(m_b - m)).
// The MMA count per step is unchanged (4 + 4); TMEM: S0 | S1 | O_a | O_b = 256 columns, two CTAs
// per SM as before.  P of half b is written over the first 16 columns of ITS OWN half of the S
// buffer (half a may still be reading columns 16..31).
// Warps: 0-7 softmax, 8 MMA issuer (+ TMEM allocation), 9 TMA producer, then the <= 2 trailing query
// rows of T = 128 k + 2 on CUDA cores (sequentially; attention_resident spends two warps on them).
#pragma once
#include "attention_resident.cuh"

namespace esmdiff {
namespace attn3 {

using attn2::BKV;
using attn2::BQ;
using attn2::DH;
using attn2::KV_TILE_BYTES;
using attn2::MAX_KV_TILES;
using attn2::MAX_LEFT;
using attn2::Params;
using attn2::Q_BYTES;
using attn2::RESCALE_LOG2;
using attn2::named_bar_arrive;
using attn2::named_bar_sync;

#ifndef ATTN_DBG
#define ATTN_DBG 0      // knock-out experiments (tools/gpu_attn_dbg.sh): 1 no ex2, 2 no S load, 4 no softmax math, 16 no trailing rows
#endif
#if ATTN_DBG & 32
// per-CTA event trace (SM clock; slot 0 = globaltimer at entry), read back by esmdiff_dbg_read_trace
constexpr int TRACE_SLOTS = 48;
static __device__ long long g_attn_trace[4096 * TRACE_SLOTS];
__device__ __forceinline__ void trace_ev(int slot) {
    if (blockIdx.x < 4096 && slot < TRACE_SLOTS) g_attn_trace[blockIdx.x * TRACE_SLOTS + slot] = clock64();
}
#define TRACE(slot) trace_ev(slot)
#else
#define TRACE(slot)
#endif
constexpr int THREADS = 320;
constexpr int SOFTMAX_THREADS = 256;
constexpr int BAR_BYTES = 512;
constexpr int TMEM_COLS = 256;              // S0 [0,64) S1 [64,128) O_a [128,192) O_b [192,256)
constexpr int COL_O = 128;
constexpr int HALF = 32;                    // score columns per softmax thread and kv tile
constexpr int XCH_BYTES = 2 * 2 * BQ * 8;   // (m, l) of both halves, double buffered over query tiles
constexpr int LEFTQ_BYTES = DH * 4;         // the trailing row's q in fp32

__host__ inline int smem_bytes(int nkv, int tail_cols) {
    return 1024 + 2 * Q_BYTES + 2 * attn2::kv_bytes(nkv, tail_cols) + BAR_BYTES + nkv * BKV * 4 /* one row of p */ +
           attn2::rstd_bytes(nkv) + XCH_BYTES + LEFTQ_BYTES;
}

// One trailing query row on CUDA cores (one warp); attention_resident.cuh's leftover_row with q read
// from shared memory (broadcast) instead of 64 registers: this kernel lives under a 96-register cap.
__device__ __forceinline__ void leftover_row(const __nv_bfloat16* __restrict__ qrow, __nv_bfloat16* __restrict__ orow,
                                             const uint8_t* sK, const uint8_t* sV, float* pf, float* qf, int T,
                                             float sc, int lane, const float* rstd_k) {
    {
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(qrow) + lane);
        qf[2 * lane] = __uint_as_float(w << 16);
        qf[2 * lane + 1] = __uint_as_float(w & 0xffff0000u);
    }
    __syncwarp();
    const float4* q4 = reinterpret_cast<const float4*>(qf);
    const int nk = (T + 31) >> 5;                      // keys per lane
    float mx = -INFINITY;
#pragma unroll 1
    for (int m = 0; m < nk; ++m) {
        const int k = m * 32 + lane;
        const int kk = k < T ? k : T - 1;              // clamp: rows past T may not be loaded
        const uint8_t* rowp = sK + (kk >> 6) * KV_TILE_BYTES + (kk & 63) * 128;
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const uint4 v = *reinterpret_cast<const uint4*>(rowp + ((c ^ (kk & 7)) << 4));
            const float4 qa = q4[2 * c], qb = q4[2 * c + 1];
            a0 = fmaf(qa.x, __uint_as_float(v.x << 16), a0);
            a1 = fmaf(qa.y, __uint_as_float(v.x & 0xffff0000u), a1);
            a0 = fmaf(qa.z, __uint_as_float(v.y << 16), a0);
            a1 = fmaf(qa.w, __uint_as_float(v.y & 0xffff0000u), a1);
            a0 = fmaf(qb.x, __uint_as_float(v.z << 16), a0);
            a1 = fmaf(qb.y, __uint_as_float(v.z & 0xffff0000u), a1);
            a0 = fmaf(qb.z, __uint_as_float(v.w << 16), a0);
            a1 = fmaf(qb.w, __uint_as_float(v.w & 0xffff0000u), a1);
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
    // out[2 lane, 2 lane + 1] = sum_k p[k] V[k][2 lane, 2 lane + 1]
    const int cl = lane >> 2, wl = (lane & 3) << 2;    // 16-byte chunk and byte inside it of my dim pair
    float o0 = 0.f, o1 = 0.f;
    const int k8 = T >> 3;
#pragma unroll 1
    for (int g = 0; g < k8; ++g) {                     // 8 keys per trip: rows 8 g .. 8 g + 7 of one tile
        const uint8_t* base = sV + (g >> 3) * KV_TILE_BYTES + (g & 7) * 1024;
        const float4 pa = *reinterpret_cast<const float4*>(pf + 8 * g);
        const float4 pb = *reinterpret_cast<const float4*>(pf + 8 * g + 4);
        const float pp[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t v = *reinterpret_cast<const uint32_t*>(base + i * 128 + (((cl ^ i) << 4) | wl));
            o0 = fmaf(pp[i], __uint_as_float(v << 16), o0);
            o1 = fmaf(pp[i], __uint_as_float(v & 0xffff0000u), o1);
        }
    }
    for (int k = k8 * 8; k < T; ++k) {
        const uint32_t v = *reinterpret_cast<const uint32_t*>(sV + (k >> 6) * KV_TILE_BYTES + (k & 63) * 128 +
                                                             (((cl ^ (k & 7)) << 4) | wl));
        o0 = fmaf(pf[k], __uint_as_float(v << 16), o0);
        o1 = fmaf(pf[k], __uint_as_float(v & 0xffff0000u), o1);
    }
    const float inv = 1.0f / l;
    reinterpret_cast<uint32_t*>(orow)[lane] = pack_bf16x2(o0 * inv, o1 * inv);
    __syncwarp();                                      // pf / qf are reused by the next row
}

__global__ void __launch_bounds__(THREADS, 2)
attention_split_kernel(const __grid_constant__ CUtensorMap tmQ,      // box [128 rows][64 cols]
                       const __grid_constant__ CUtensorMap tmKV,     // box [ 64 rows][64 cols]
                       const __grid_constant__ CUtensorMap tmKVt,    // box [tail rows][64 cols]
                       const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const int kvb = attn2::kv_bytes(p.nkv, p.tail_cols);
    uint8_t* sQ = smem;                                   // two query-tile buffers
    uint8_t* sK = sQ + 2 * Q_BYTES;
    uint8_t* sV = sK + kvb;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kvb);
    uint64_t* q_full = bars;                              // [2]
    uint64_t* q_empty = bars + 2;                         // [2]
    uint64_t* s_full = bars + 4;                          // [2]  MMA -> softmax
    uint64_t* p_full = bars + 6;                          // [2]  softmax warps (8 arrivals) -> MMA
    uint64_t* pv_done = bars + 8;                         // 1    P V of step nsteps-2 retired (no S follows it)
    uint64_t* o_free = bars + 9;                          // 1    completes once per query tile
    uint64_t* o_full = bars + 10;                         // 1    last PV of a query tile retired
    uint64_t* k_full = bars + 11;                         // [MAX_KV_TILES], single use
    uint64_t* v_full = k_full + MAX_KV_TILES;             // [MAX_KV_TILES], single use
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(v_full + MAX_KV_TILES);
    float* left_p = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + BAR_BYTES);   // [nkv * 64]
    float* rstd_k = left_p + p.nkv * BKV;                                                     // [nkv * 64]
    float* rstd_q = rstd_k + p.nkv * BKV;                                                     // [nkv * 64] (>= T)
    float2* xch = reinterpret_cast<float2*>(rstd_q + p.nkv * BKV);                            // [2 bufs][2 halves][128]
    float* left_q = reinterpret_cast<float*>(xch + 2 * 2 * BQ);                               // [64]

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
    const bool fused_ln = p.qk_rstd != nullptr;

#if ATTN_DBG & 32
    if (threadIdx.x == 0 && blockIdx.x < 4096) {
        g_attn_trace[blockIdx.x * TRACE_SLOTS + 0] = static_cast<long long>(global_timer_ns());
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        g_attn_trace[blockIdx.x * TRACE_SLOTS + 1] = smid;
        TRACE(2);
    }
#endif
    const int has_q1 = nq > 1 ? 1 : 0;
    const int nloads = 2 * nkv + 1 + has_q1;
    if (warp == 9) {
        if (lane == 0) {
            tma_prefetch_desc(&tmQ);
            tma_prefetch_desc(&tmKV);
            tma_prefetch_desc(&tmKVt);
            for (int s = 0; s < 2; ++s) {
                mbar_init(&q_full[s], 1);
                mbar_init(&q_empty[s], 1);
                mbar_init(&s_full[s], 1);
                mbar_init(&p_full[s], 8);              // one arrival per softmax warp
            }
            mbar_init(pv_done, 1);
            mbar_init(o_free, 8);
            mbar_init(o_full, 1);
            for (int j = 0; j < nkv; ++j) {
                mbar_init(&k_full[j], 1);
                mbar_init(&v_full[j], 1);
            }
            fence_barrier_init();
        }
        __syncwarp();
        // all tile loads of this CTA, one per lane, before the CTA-wide barrier (attention_resident.cuh)
        pdl_wait();
        if (lane < nloads) {
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
        if (lane == 0) TRACE(6);
    }
    if (warp == 8) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) TRACE(3);
    pdl_launch_dependents();
    if (warp < 8) pdl_wait();              // the softmax warps read the 1/std tables and write the output
    if (threadIdx.x == 0) TRACE(4);

    if (fused_ln && warp != 8) {
        // 1/std of this sample's k rows and q rows into shared memory
        const int tid = warp < 8 ? threadIdx.x : threadIdx.x - 32;       // 0..287
        for (int idx = tid; idx < nkv * BKV; idx += SOFTMAX_THREADS + 32) {
            const bool in = idx < p.T;
            rstd_k[idx] = in ? __ldg(p.qk_rstd + p.rstd_ld + row0 + idx) : 0.f;   // 0: padding keys of the last tile
            rstd_q[idx] = in ? __ldg(p.qk_rstd + row0 + idx) : 1.f;
        }
        named_bar_sync(2, SOFTMAX_THREADS + 32);
        if (threadIdx.x == 0) TRACE(5);
    }

    if (warp == 9) {
        // ===================== L2 prefetch for the next CTA, the trailing query rows, later query tiles =====================
        const int next = blockIdx.x + p.prefetch_stride;
        if (p.prefetch_stride > 0 && next < static_cast<int>(gridDim.x) && lane < nloads) {
            mbar_wait(&v_full[nkv - 1], 0);                  // my own tiles are (as good as) in
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
        __syncwarp();
        if (n_left > 0 && !(ATTN_DBG & 16)) {
            for (int j = 0; j < nkv; ++j) mbar_wait(&k_full[j], 0);
            for (int j = 0; j < nkv; ++j) mbar_wait(&v_full[j], 0);
            for (int lw = 0; lw < n_left; ++lw) {
                const int t = p.nq * BQ + lw;
                float sc = p.scale_log2;
                if (fused_ln) sc *= rstd_q[t];
                leftover_row(p.qkv + static_cast<long long>(row0 + t) * 3 * D + h * DH,
                             p.ctx + static_cast<long long>(row0 + t) * D + h * DH, sK, sV, left_p, left_q, p.T, sc, lane,
                             fused_ln ? rstd_k : nullptr);
            }
        }
        if (lane == 0) {
            for (int qt = 2; qt < nq; ++qt) {
                mbar_wait(&q_empty[qt & 1], ((qt >> 1) - 1) & 1);
                uint64_t* bar = &q_full[qt & 1];
                mbar_arrive_expect_tx(bar, Q_BYTES);
                tma_load_2d(sQ + (qt & 1) * Q_BYTES, &tmQ, bar, h * DH, row0 + (qt0 + qt) * BQ);
            }
        }
    } else if (warp == 8) {
        // ===================== MMA issuer =====================
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
            if (s_qt == 0) mbar_wait(&k_full[s_j], 0);
            if (lane == 0 && s_qt == 0 && s_j == 0) TRACE(7);
            if (lane == 0 && s_qt == 0 && s_j == nkv - 1) TRACE(8);
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
                if (lane == 0 && qt == 0 && j == 0) TRACE(9);
                if (lane == 0 && qt == 0 && j == nkv - 1) TRACE(10);
                mbar_wait(&p_full[i & 1], (i >> 1) & 1);
                if (lane == 0 && i < 8) TRACE(16 + i);
                if (j == 0 && qt > 0) mbar_wait(o_free, (qt - 1) & 1);
                tcgen05_fence_after();
                // V tile [kv rows][64 d] is an MN-major B operand: 128-byte rows along N = d,
                // 8-row (k) groups 1024 B apart; one UMMA K-step (16 kv rows) = 2048 B.
                const uint64_t vdesc = desc_v0 + static_cast<uint64_t>(j * (KV_TILE_BYTES >> 4));
                const uint32_t tp = tmem_base + (i & 1) * BKV;          // P: bf16 pairs, 8 columns per K-step
                const bool last = j == nkv - 1;
                const uint32_t acc = j != 0 ? 1u : 0u;
                if (elect_one()) {
                    // K-steps 0, 1: keys [0, 32) of the tile, P at the start of half a's columns, into O_a;
                    // K-steps 2, 3: keys [32, 64), P at the start of half b's columns, into O_b
                    if (!last) {
                        umma_bf16_ts(tmem_o, tp, vdesc, idesc_pv, acc);
                        umma_bf16_ts(tmem_o, tp + 8, vdesc + 128, idesc_pv, 1u);
                        umma_bf16_ts(tmem_o + DH, tp + HALF, vdesc + 256, idesc_pv, acc);
                        umma_bf16_ts(tmem_o + DH, tp + HALF + 8, vdesc + 384, idesc_pv, 1u);
                    } else {
                        const int ksteps = p.tail_cols >> 4;
                        for (int k = 0; k < ksteps; ++k)
                            umma_bf16_ts(tmem_o + (k >> 1) * DH, tp + (k >> 1) * HALF + (k & 1) * 8, vdesc + 128 * k,
                                         idesc_pv, (k & 1) ? 1u : acc);
                    }
                    // every tcgen05.commit stalls this thread ~200 clk (tools/mma_bench.cu): the "P V of
                    // step i retired" signal the rare rescale path needs rides on the commit of S_{i+2},
                    // issued right behind it; only the P Vs that no S follows commit their own
                    if (last) umma_commit(o_full);
                    else if (s_i >= nsteps) umma_commit(pv_done);
                }
                __syncwarp();
                if (s_i < nsteps) issue_s();              // overwrites P_i's buffer: ordered after PV_i
                if (lane == 0 && i < 8) TRACE(24 + i);
            }
        }
    } else {
        // ===================== softmax / output warps: two threads per query row =====================
        const int quad = warp & 3, half = warp >> 2;
        const int r = quad * 32 + lane;                              // query row of the tile == TMEM lane
        const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
        const uint32_t t_o = tmem_base + lane_addr + COL_O + half * DH;      // my half's accumulator
        const int col0 = half * HALF;
        float sc = p.scale_log2;              // per query row once q_ln's 1/std is folded in (set per query tile)
        float thresh = RESCALE_LOG2 / sc;
        // out[row][32 half .. 32 half + 32) for query tile qt from both halves' accumulators, then free O.
        // m2 = my half's running maximum in log2 units (-inf when the half saw no column), l = its row sum
        auto epilogue = [&](int qt, float m2, float l) {
            float2* x = xch + (qt & 1) * 2 * BQ;
            x[half * BQ + r] = make_float2(m2, l);
            named_bar_sync(3 + quad, 64);                            // the two warps that share these rows
            const float2 o = x[(half ^ 1) * BQ + r];
            mbar_wait(o_full, qt & 1);
            tcgen05_fence_after();
            if ((qt0 + qt) * BQ + quad * 32 < p.T) {
                const float m = fmaxf(m2, o.x);
                const float f_me = l > 0.f ? fast_exp2(m2 - m) : 0.f;
                const float f_ot = o.y > 0.f ? fast_exp2(o.x - m) : 0.f;
                const float inv = 1.0f / (l * f_me + o.y * f_ot);
                const float fa = (half == 0 ? f_me : f_ot) * inv, fb = (half == 0 ? f_ot : f_me) * inv;
                const bool has_a = half == 0 ? l > 0.f : o.y > 0.f, has_b = half == 0 ? o.y > 0.f : l > 0.f;
                const int t = (qt0 + qt) * BQ + r;
                uint4* dst = reinterpret_cast<uint4*>(p.ctx + static_cast<long long>(row0 + t) * D + h * DH + col0);
                const uint32_t t_oa = tmem_base + lane_addr + COL_O + col0, t_ob = t_oa + DH;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t oa[16], ob[16];
                    tmem_ld_32x32b_x16(t_oa + c * 16, oa);
                    tmem_ld_32x32b_x16(t_ob + c * 16, ob);
                    tmem_ld_wait();
                    float v[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        v[e] = (has_a ? __uint_as_float(oa[e]) * fa : 0.f) + (has_b ? __uint_as_float(ob[e]) * fb : 0.f);
                    if (t < p.T) {
#pragma unroll
                        for (int g = 0; g < 2; ++g)
                            dst[c * 2 + g] = make_uint4(pack_bf16x2(v[8 * g], v[8 * g + 1]), pack_bf16x2(v[8 * g + 2], v[8 * g + 3]),
                                                        pack_bf16x2(v[8 * g + 4], v[8 * g + 5]), pack_bf16x2(v[8 * g + 6], v[8 * g + 7]));
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(o_free);
        };

        float m_run = 0.f;                    // running max of my half's raw scores
        float l_run = 0.f;                    // running row sum (relative to m_run)
        float m2_prev = -INFINITY, l_prev = 0.f;   // the finished query tile's values, for its deferred epilogue
        bool seen = false;                    // my half has had a column in this query tile
#pragma unroll 1
        for (int qt = 0; qt < nq; ++qt) {
            const bool active = (qt0 + qt) * BQ + quad * 32 < p.T;   // warp-uniform
            for (int j = 0; j < nkv; ++j) {
                const int i = qt * nkv + j;
                const uint32_t t_s = tmem_base + lane_addr + (i & 1) * BKV + col0;
                if (j == 0) {                 // new query tile: keep the finished tile's state for its epilogue
                    m2_prev = seen ? m_run * sc : -INFINITY;
                    l_prev = seen ? l_run : 0.f;
                    l_run = 0.f;
                    seen = false;
                    if (fused_ln) {
                        const int t = (qt0 + qt) * BQ + r;
                        sc = p.scale_log2 * (t < p.T ? rstd_q[t] : 1.0f);
                        thresh = RESCALE_LOG2 / sc;
                    }
                }
                const bool last = j == nkv - 1;
                const int mycols = last ? min(max(p.tail_cols - col0, 0), HALF) : HALF;   // 0, 16 or 32
                mbar_wait(&s_full[i & 1], (i >> 1) & 1);
                if (threadIdx.x == 0 && i == 0) TRACE(11);
                if (threadIdx.x == 0 && i < 8) TRACE(32 + i);
                tcgen05_fence_after();
                if (active && mycols > 0 && !(ATTN_DBG & 4)) {
                    const int nch = mycols >> 4;                         // 16-column chunks
                    uint32_t s[HALF];
#if ATTN_DBG & 2
#pragma unroll
                    for (int e = 0; e < HALF; ++e) s[e] = __float_as_uint(0.01f * static_cast<float>((r + e * 7 + i) & 63));
#else
                    tmem_ld_32x32b_x16(t_s, s);
                    if (nch > 1) tmem_ld_32x32b_x16(t_s + 16, s + 16);
                    tmem_ld_wait();
#endif
                    if (fused_ln) {
                        // k_ln's 1/std: one factor per score column (= key row), shared by all query rows
                        const float4* rk4 = reinterpret_cast<const float4*>(rstd_k + j * BKV + col0);
#pragma unroll
                        for (int c = 0; c < 2; ++c)
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
                        const int valid = p.T - j * BKV - col0;      // >= 1 valid kv columns in my part of this tile
#pragma unroll
                        for (int c = 0; c < 2; ++c)
                            if (c < nch) {
#pragma unroll
                                for (int e = 0; e < 16; ++e) {
                                    const float a = (c * 16 + e < valid) ? __uint_as_float(s[c * 16 + e]) : -INFINITY;
                                    s[c * 16 + e] = __float_as_uint(a);
                                    mx = fmaxf(mx, a);
                                }
                            }
                    } else {
                        // four independent FMNMX3 chains of 4
                        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                        for (int e = 0; e < HALF; e += 8)
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                m4[k] = fmax3(m4[k], __uint_as_float(s[e + 2 * k]), __uint_as_float(s[e + 2 * k + 1]));
                        mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
                    }
                    if (!seen) {
                        m_run = mx;
                        seen = true;
                    } else {
                        const bool need = mx > m_run + thresh;
                        if (__any_sync(0xffffffffu, need)) {
                            // raise the running max: rescale my O half and l once PV_{i-1} has retired
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
                    for (int c = 0; c < 2; ++c)
                        if (c < nch) {
                            uint32_t pk[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                float a0, a1;
                                ffma2(a0, a1, __uint_as_float(s[c * 16 + 2 * e]), __uint_as_float(s[c * 16 + 2 * e + 1]), sc, sc, nm, nm);
#if ATTN_DBG & 1
                                const float p0 = a0, p1 = a1;
#else
                                const float p0 = fast_exp2(a0), p1 = fast_exp2(a1);
#endif
                                fadd2(rs0, rs1, rs0, rs1, p0, p1);
                                pk[e] = pack_bf16x2(p0, p1);
                            }
                            tmem_st_32x32b_x8(t_s + c * 8, pk);      // P over the start of my columns: 2 bf16 per column
                        }
                    l_run += rs0 + rs1;
                    tmem_st_wait();
                }
                // one arrival per warp
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[i & 1]);
                if (threadIdx.x == 0 && i < 8) TRACE(40 + i);
                if (j == 0 && qt > 0) epilogue(qt - 1, m2_prev, l_prev);   // deferred: overlaps this tile's first MMAs
            }
        }
        epilogue(nq - 1, seen ? m_run * sc : -INFINITY, seen ? l_run : 0.f);
        if (threadIdx.x == 0) TRACE(12);
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 8) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
        if (lane == 0) TRACE(13);
    }
}

}  // namespace attn3
}  // namespace esmdiff
