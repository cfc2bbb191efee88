// EXPERIMENT, not part of the library (see DESIGN.md "What bounds attention here"): correct but
// slower than attention_resident.cuh (214 vs 156 us at B = 100, T = 258).  Built only by
// tools/attn_trace.cu.
// Non-causal multi-head attention, d_head = 64, for the sequence lengths the sampling path is quoted
// on (T = L + 2 <= 766).  Replaces F.scaled_dot_product_attention on the reference path
// (esm MultiHeadAttention.forward with seq_id None).
//
// What bounds this kernel on sm_100 (measured: profiles/r1f_ncu_full_attention.txt,
// tools/attn_trace.cu, tools/mma_bench.cu):
//   * MUFU: 16 ex2/clk/SM -> 41 us at B = 100, T = 258 if the unit never idles;
//   * the tensor pipe at d_head = 64: a 128 x 64 x 16 UMMA with A in shared memory costs ~47 clk
//     (operand reads, not math); with A in TMEM it is math bound (32 clk) -> 28 us;
//   * the MMA-issuing thread: every tcgen05.commit stalls it ~190 clk, every small UMMA ~40 clk.
// The earlier resident-K/V kernel (attention_resident.cuh: 8 UMMAs + 2 commits per 64-key step,
// 4 softmax warps per CTA) sat at 155 us: issue-thread bound with MUFU 31 % busy.  This version:
//   * Q lives in TMEM (the softmax threads load their query rows from global memory straight into
//     TMEM, two tiles ahead), so S = Q K^T is a TS-mode UMMA like P V and no shared memory is
//     spent on queries;
//   * TWO issuing warps, one for S = Q K^T and one for P V, one commit per step each;
//   * EIGHT softmax warps per CTA: two per TMEM lane quarter, each thread owns one query row and
//     HALF of the 64 key columns of a step; the two halves of a row agree on the running max
//     through shared memory (named barrier per quarter).  With two CTAs per SM that is 4 softmax
//     warps per scheduler, enough to keep MUFU fed;
//   * S is double buffered (step i+2 is computed while i and i+1 are in softmax), P overwrites S
//     in place as bf16 pairs and feeds P V from TMEM; the running max is only raised when a step
//     max exceeds it by 2^8 (lazy rescale of O in TMEM);
//   * optionally ATTN_POLY_PER_8 of every 8 exponentials run on the FMA pipe (Cody-Waite +
//     degree-3 minimax polynomial, 7.5e-5 relative error, below the bf16 rounding of P).
// One CTA = one (sample b, head h), walking its query tiles of 128 rows; K and V of the head are
// resident in shared memory (TMA, 128B swizzle).  The last kv tile is only as wide as needed
// (multiple of 16 columns).
// Input  qkv : bf16 [M = B*T, 3*D]  (q | k | v, each D = H*64; q,k already LayerNormed + RoPE'd)
// Output ctx : bf16 [M, D]
#pragma once
#include "../../esmdiff_b200/csrc/ptx.cuh"

namespace esmdiff {
namespace attn4 {

constexpr int BQ = 128;
constexpr int BKV = 64;
constexpr int DH = 64;
constexpr int MAX_KV_TILES = 12;            // T <= 768
constexpr int KV_TILE_BYTES = BKV * DH * 2; // 8 KiB
constexpr int XCHG_BYTES = 3 * 2 * BQ * 4;  // row max (two slots, by step parity) / row sum exchange between the column halves
constexpr int BAR_BYTES = 512;
constexpr int THREADS = 320;                // warps 0-7 softmax (quarter = w & 3, column half = w >> 2), 8 TMA + S issue, 9 P V issue
constexpr int TMEM_COLS = 256;
constexpr int COL_S = 0;                    // S/P buffers [0,64) and [64,128)
constexpr int COL_O = 128;                  // O [128,192)
constexpr int COL_Q = 192;                  // Q buffers [192,224) and [224,256): 64 bf16 = 32 columns per row
constexpr float RESCALE_LOG2 = 8.0f;        // lazy rescale threshold: p <= 2^8
#ifndef ATTN_POLY_PER_8
#define ATTN_POLY_PER_8 1                   // of every 8 exponentials, this many run on the FMA pipe
#endif

struct Params {
    int B, T, H;
    int nq;                     // ceil(T / 128)
    int nkv;                    // ceil(T / 64)
    int tail_cols;              // width of the last kv tile: multiple of 16 in [16, 64]
    const __nv_bfloat16* qkv;   // [B*T, 3*H*64]
    __nv_bfloat16* ctx;         // [B*T, H*64]
    float scale_log2;           // (1/sqrt(64)) * log2(e)
#ifdef ATTN_TRACE
    long long* trace;           // tools/attn_trace.cu: [gridDim.x][64] SM-clock stamps
#endif
};

#ifdef ATTN_TRACE
#define ATTN_STAMP(slot) do { if ((slot) < 64) p.trace[blockIdx.x * 64 + (slot)] = clock64(); } while (0)
#else
#define ATTN_STAMP(slot) do { } while (0)
#endif

__host__ __device__ inline int kv_bytes(int nkv, int tail_cols) {
    return (nkv - 1) * KV_TILE_BYTES + tail_cols * 128;
}
__host__ inline int smem_bytes(int nkv, int tail_cols) {
    return 1024 + 2 * kv_bytes(nkv, tail_cols) + XCHG_BYTES + BAR_BYTES;
}

// 2^x on the FMA pipe for x <= ~8 (x is clamped at -125): n = round(x), f = x - n in [-0.5, 0.5],
// 2^f by a degree-3 minimax polynomial (7.5e-5 relative), 2^n by adding n to the exponent field.
__device__ __forceinline__ float exp2_poly(float x) {
    x = fmaxf(x, -125.0f);
    const float t = x + 12582912.0f;                    // 1.5 * 2^23: n sits in the low mantissa bits
    const float f = x - (t - 12582912.0f);
    float r = fmaf(0.05517162f, f, 0.24261113f);
    r = fmaf(r, f, 0.69326097f);
    r = fmaf(r, f, 0.99992806f);
    return __uint_as_float(__float_as_uint(r) + (__float_as_uint(t) << 23));
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(THREADS, 2)
attention_qtmem_kernel(const __grid_constant__ CUtensorMap tmKV,     // box [ 64 rows][64 cols]
                       const __grid_constant__ CUtensorMap tmKVt,    // box [tail rows][64 cols]
                       const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const int kvb = kv_bytes(p.nkv, p.tail_cols);
    uint8_t* sK = smem;
    uint8_t* sV = sK + kvb;
    float* xchg = reinterpret_cast<float*>(sV + kvb);     // [3 slots][2 halves][128 rows]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kvb + XCHG_BYTES);
    uint64_t* q_full = bars;                              // [2]  softmax threads (256) -> MMA: Q tile in TMEM
    uint64_t* s_full = bars + 2;                          // [2]  MMA -> softmax
    uint64_t* p_full = bars + 4;                          // [2]  softmax threads (256) -> MMA
    uint64_t* pv_done = bars + 6;                         // [2]  P V of a step retired (buffer free; tile's last: O complete)
    uint64_t* k_full = bars + 8;                          // [MAX_KV_TILES], single use
    uint64_t* v_full = k_full + MAX_KV_TILES;             // [MAX_KV_TILES], single use
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(v_full + MAX_KV_TILES);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int h = blockIdx.x % p.H;
    const int b = blockIdx.x / p.H;
    const int D = p.H * DH;
    const int row0 = b * p.T;
    const int nq = p.nq, nkv = p.nkv;
    const int nsteps = nq * nkv;
    if (threadIdx.x == 0) ATTN_STAMP(0);                  // CTA start

    if (warp == 8 && lane == 0) {
        tma_prefetch_desc(&tmKV);
        tma_prefetch_desc(&tmKVt);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&q_full[s], 256);
            mbar_init(&s_full[s], 1);
            mbar_init(&p_full[s], 256);
        }
        mbar_init(&pv_done[0], 1);
        mbar_init(&pv_done[1], 1);
        for (int j = 0; j < nkv; ++j) {
            mbar_init(&k_full[j], 1);
            mbar_init(&v_full[j], 1);
        }
        fence_barrier_init();
    }
    if (warp == 9) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) ATTN_STAMP(1);                  // setup done

    if (warp == 8) {
        // ===================== TMA producer + S = Q K^T issuer =====================
        // Two issuing warps (this one for S, warp 9 for P V): a tcgen05.commit stalls its issuing
        // thread for ~200+ clk and every small UMMA for ~40-100, so one warp issuing 8 UMMAs and a
        // commit per step was the bottleneck of the whole kernel (1000 clk per 64-key step with the
        // softmax switched off, tools/attn_trace.cu).  S of step s reuses the buffer of step s-2,
        // whose P the P V of step s-2 reads: ordered through pv_done (another thread issues it).
        if (lane == 0) {
            auto load_kv = [&](uint8_t* dst, uint64_t* bar, int col, int j) {
                const bool last = j == nkv - 1;
                mbar_arrive_expect_tx(bar, last ? p.tail_cols * 128 : KV_TILE_BYTES);
                tma_load_2d(dst + j * KV_TILE_BYTES, last ? &tmKVt : &tmKV, bar, col, row0 + j * BKV);
            };
            for (int j = 0; j < nkv; ++j) load_kv(sK, &k_full[j], D + h * DH, j);
            for (int j = 0; j < nkv; ++j) load_kv(sV, &v_full[j], 2 * D + h * DH, j);
        }
        __syncwarp();
        const uint32_t idesc_s_full = umma_idesc_bf16(BQ, BKV, 0);
        const uint32_t idesc_s_tail = umma_idesc_bf16(BQ, p.tail_cols, 0);
        const uint64_t desc_k0 = umma_desc_sw128(smem_u32(sK), 16, 1024);
        int s_i = 0;
        for (int qt = 0; qt < nq; ++qt) {
            for (int j = 0; j < nkv; ++j, ++s_i) {
                if (j == 0) mbar_wait(&q_full[qt & 1], (qt >> 1) & 1);
                if (qt == 0) mbar_wait(&k_full[j], 0);
                if (s_i >= 2) mbar_wait(&pv_done[s_i & 1], ((s_i - 2) >> 1) & 1);
                if (lane == 0 && s_i < 15) ATTN_STAMP(2 + 2 * s_i);
                tcgen05_fence_after();
                const uint64_t kdesc = desc_k0 + static_cast<uint64_t>(j * (KV_TILE_BYTES >> 4));
                const uint32_t idesc = j == nkv - 1 ? idesc_s_tail : idesc_s_full;
                const uint32_t ts = tmem_base + COL_S + (s_i & 1) * BKV;
                const uint32_t tq = tmem_base + COL_Q + (qt & 1) * 32;      // Q: bf16 pairs, 8 columns per K-step
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < DH / 16; ++k)
                        umma_bf16_ts(ts, tq + 8 * k, kdesc + 2 * k, idesc, k != 0 ? 1u : 0u);
                    umma_commit(&s_full[s_i & 1]);
                }
                __syncwarp();
            }
        }
    } else if (warp == 9) {
        // ===================== O (+)= P V issuer =====================
        constexpr uint32_t idesc_pv = umma_idesc_bf16(BQ, DH, 1);       // P V : V is MN-major
        const uint64_t desc_v0 = umma_desc_sw128(smem_u32(sV), 16, 1024);
        const uint32_t tmem_o = tmem_base + COL_O;
        int i = 0;
        for (int qt = 0; qt < nq; ++qt) {
            for (int j = 0; j < nkv; ++j, ++i) {
                if (qt == 0) mbar_wait(&v_full[j], 0);
                mbar_wait(&p_full[i & 1], (i >> 1) & 1);
                if (lane == 0 && i < 15) ATTN_STAMP(3 + 2 * i);
                tcgen05_fence_after();
                // V tile [kv rows][64 d] is an MN-major B operand: 128-byte rows along N = d,
                // 8-row (k) groups 1024 B apart; one UMMA K-step (16 kv rows) = 2048 B.
                const uint64_t vdesc = desc_v0 + static_cast<uint64_t>(j * (KV_TILE_BYTES >> 4));
                const uint32_t tp = tmem_base + COL_S + (i & 1) * BKV;  // P: bf16 pairs, 8 columns per K-step
                const int ksteps = j == nkv - 1 ? p.tail_cols >> 4 : BKV / 16;
                if (elect_one()) {
                    for (int k = 0; k < ksteps; ++k)
                        umma_bf16_ts(tmem_o, tp + 8 * k, vdesc + 128 * k, idesc_pv, (j | k) != 0 ? 1u : 0u);
                    umma_commit(&pv_done[i & 1]);   // frees the S/P buffer; the tile's last one also means "O complete"
                }
                __syncwarp();
            }
        }
    } else {
        // ============ softmax / output warps: thread = one query row x half of the key columns ============
        const int q4 = warp & 3;                                     // TMEM lane quarter
        const int hf = warp >> 2;                                    // column half
        const int r = q4 * 32 + lane;                                // row in the tile == TMEM lane
        const uint32_t lane_addr = static_cast<uint32_t>(q4 * 32) << 16;
        const uint32_t t_o = tmem_base + lane_addr + COL_O + hf * 32;
        const float sc = p.scale_log2;
        const float thresh = RESCALE_LOG2 / sc;
        // slot s: written by this thread at [s][hf][r], read by the partner thread of the other half.
        // Step maxima alternate between slots 0/1 by step parity (a thread can only be one barrier
        // ahead of its partner, so a slot is never rewritten before it was read); slot 2 = row sums.
        float* my_x = xchg + hf * BQ + r;
        const float* other_x = xchg + (hf ^ 1) * BQ + r;

        // this thread's 32 of the 64 query dims of row t (dims [32 hf, 32 hf + 32)) -> 16 TMEM columns
        auto load_q = [&](int qt, uint4 (&qv)[4]) {
            const int t = qt * BQ + r;
            if (t < p.T) {
                const uint4* src = reinterpret_cast<const uint4*>(
                    p.qkv + static_cast<long long>(row0 + t) * 3 * D + h * DH + hf * 32);
#pragma unroll
                for (int e = 0; e < 4; ++e) qv[e] = __ldg(src + e);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) qv[e] = make_uint4(0u, 0u, 0u, 0u);
            }
        };
        auto store_q = [&](int qt, const uint4 (&qv)[4]) {
            tmem_st_32x32b_x16(tmem_base + lane_addr + COL_Q + (qt & 1) * 32 + hf * 16,
                               reinterpret_cast<const uint32_t*>(qv));
            tmem_st_wait();
            tcgen05_fence_before();
            mbar_arrive(&q_full[qt & 1]);
        };
        uint4 qv[4];
        load_q(0, qv);
        store_q(0, qv);
        if (nq > 1) {
            load_q(1, qv);
            store_q(1, qv);
        }

        int i = 0;
        for (int qt = 0; qt < nq; ++qt) {
            const bool active = qt * BQ + q4 * 32 < p.T;             // warp-uniform, same in both halves
            // Q of tile qt+1 goes into the buffer tile qt-1 used (all its S have been seen complete)
            const bool prefetch = qt >= 1 && qt + 1 < nq;
            if (prefetch) load_q(qt + 1, qv);
            float m_run = 0.f, l_run = 0.f;
            for (int j = 0; j < nkv; ++j, ++i) {
                const uint32_t t_s = tmem_base + lane_addr + COL_S + (i & 1) * BKV;
                mbar_wait(&s_full[i & 1], (i >> 1) & 1);
                if (threadIdx.x == 0 && i < 14) ATTN_STAMP(32 + 2 * i);
                tcgen05_fence_after();
#ifdef ATTN_SKIP_SOFTMAX
                if (false) {
#else
                if (active) {
#endif
                    const bool last = j == nkv - 1;
                    const int nch = (last ? p.tail_cols : BKV) >> 4;     // 16-column chunks in this step
                    const int valid = (last ? p.T - j * BKV : BKV) - hf * 32;   // valid columns of this half
                    const bool c0 = 2 * hf < nch, c1 = 2 * hf + 1 < nch;     // which of my two chunks exist
                    uint32_t s[32];
                    if (c0) tmem_ld_32x32b_x16(t_s + hf * 32, s);
                    if (c1) tmem_ld_32x32b_x16(t_s + hf * 32 + 16, s + 16);
                    if (!c0) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) s[e] = 0u;
                    }
                    if (!c1) {
#pragma unroll
                        for (int e = 16; e < 32; ++e) s[e] = 0u;
                    }
                    tmem_ld_wait();
                    float mx0 = -INFINITY, mx1 = -INFINITY;
                    if (!last) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            mx0 = fmaxf(mx0, __uint_as_float(s[2 * e]));
                            mx1 = fmaxf(mx1, __uint_as_float(s[2 * e + 1]));
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 32; ++e)
                            if (e < valid) mx0 = fmaxf(mx0, __uint_as_float(s[e]));
                    }
                    // the two halves of the row agree on the step max (also orders: both halves have
                    // read S before either writes P over it)
                    my_x[(i & 1) * 2 * BQ] = fmaxf(mx0, mx1);
                    named_bar_sync(1 + q4, 64);
                    const float mx = fmaxf(fmaxf(mx0, mx1), other_x[(i & 1) * 2 * BQ]);
                    if (j == 0) {
                        m_run = mx;
                    } else {
                        const bool need = mx > m_run + thresh;
                        if (__any_sync(0xffffffffu, need)) {
                            // raise the running max: rescale my half of O and my partial row sum once P V of
                            // the previous step has retired
                            const float m_new = need ? mx : m_run;
                            const float f = fast_exp2((m_run - m_new) * sc);
                            mbar_wait(&pv_done[(i - 1) & 1], ((i - 1) >> 1) & 1);
                            tcgen05_fence_after();
                            l_run *= f;
#pragma unroll 1
                            for (int c = 0; c < 2; ++c) {
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
                    // p = exp2(s*scale - m): MUFU, a share on the FMA pipe
                    const float nm = -m_run * sc;
                    float rs0 = 0.f, rs1 = 0.f;
                    uint32_t pk[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const float x0 = fmaf(__uint_as_float(s[2 * e]), sc, nm);
                        const float x1 = fmaf(__uint_as_float(s[2 * e + 1]), sc, nm);
                        float p0 = ((2 * e) & 7) < ATTN_POLY_PER_8 ? exp2_poly(x0) : fast_exp2(x0);
                        float p1 = ((2 * e + 1) & 7) < ATTN_POLY_PER_8 ? exp2_poly(x1) : fast_exp2(x1);
                        if (last) {
                            p0 = 2 * e < valid ? p0 : 0.f;
                            p1 = 2 * e + 1 < valid ? p1 : 0.f;
                        }
                        rs0 += p0;
                        rs1 += p1;
                        pk[e] = pack_bf16x2(p0, p1);
                    }
                    l_run += rs0 + rs1;
                    // S columns [32 hf, 32 hf + 32) -> P columns [16 hf, 16 hf + 16) of the same buffer
                    if (c0) tmem_st_32x32b_x8(t_s + hf * 16, pk);
                    if (c1) tmem_st_32x32b_x8(t_s + hf * 16 + 8, pk + 8);
                    tmem_st_wait();
                }
                tcgen05_fence_before();
                mbar_arrive(&p_full[i & 1]);
                if (threadIdx.x == 0 && i < 14) ATTN_STAMP(33 + 2 * i);
                if (j == 0 && prefetch) store_q(qt + 1, qv);
            }
            // ---- out[row] = O / l once the tile's last P V has retired ----
            if (active) {
                my_x[2 * 2 * BQ] = l_run;
                named_bar_sync(1 + q4, 64);
                l_run += other_x[2 * 2 * BQ];
            }
            mbar_wait(&pv_done[(i - 1) & 1], ((i - 1) >> 1) & 1);
            if (threadIdx.x == 0) ATTN_STAMP(60 + (qt & 1) * 2);
            tcgen05_fence_after();
            if (active) {
                const float inv = 1.0f / l_run;
                const int t = qt * BQ + r;
                uint4* dst = reinterpret_cast<uint4*>(p.ctx + static_cast<long long>(row0 + t) * D + h * DH + hf * 32);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
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
            // the next tile's first P V (accumulate = 0 into O) is only issued after this thread's
            // next p_full arrive, i.e. after the reads above
            tcgen05_fence_before();
            if (threadIdx.x == 0) ATTN_STAMP(61 + (qt & 1) * 2);
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 9) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace attn4
}  // namespace esmdiff
