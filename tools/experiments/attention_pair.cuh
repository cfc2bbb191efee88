// EXPERIMENT, not part of the library: numerically identical to attention_resident.cuh on the first
// run, but slower (B = 100, T = 258: 203 us against 138 us; the cross-CTA arrive/commit round trip per
// 64-key step costs more than the halved UMMA issue saves).  To try it again: include it from
// esmdiff_b200.cu and launch attn5::attention_pair_kernel with 2*B*H CTAs (see git history of
// launch_attention at the commit that added this file).
// Non-causal multi-head attention, d_head = 64, on CTA PAIRS: the two CTAs of a cluster (one TPC)
// take two consecutive 128-row query tiles of the SAME (sample, head) and share every UMMA:
//   S  = Q K^T   tcgen05.mma.cta_group::2, M = 256 (128 query rows per CTA), N = 64 keys per step,
//                each CTA supplying 32 of the 64 key rows of the tile from its own shared memory
//   O += P V     cta_group::2 with A = P from each CTA's TMEM, N = 64 head dims, each CTA supplying
//                32 of them
// Why: at d_head = 64 the single-CTA kernel (attention_resident.cuh) is bound by the thread that
// ISSUES the tensor instructions -- every tcgen05.commit stalls it ~190 clk and every 128x64x16
// UMMA 40-100 clk (tools/mma_bench.cu), ~1000 clk per 64-key step with the softmax switched off
// (tools/attn_trace.cu) against ~500 clk of MUFU work.  One instruction stream for two query tiles
// halves that cost per tile; everything else (K/V resident in shared memory, S/P/O in TMEM, P fed
// to P V from TMEM, lazy rescale, thread = query row, trailing T mod 128 <= 2 rows on CUDA cores)
// is as in attention_resident.cuh.
// Operand placement: the instruction carries ONE shared-memory descriptor that both CTAs apply to
// their own memory, and takes the first N/2 rows (K, K-major) or columns (V, MN-major) from each.
// Both CTAs therefore keep all keys of a tile, CTA 1 with the two 32-row halves swapped, and CTA 1
// loads V shifted by 32 head dims (its columns 0-31 hold dims 32-63).  The leader CTA, whose copy
// is in natural order, also runs the trailing query rows on CUDA cores.
// Used when there are at least two full query tiles (T >= 256); input/output as the other kernels:
//   qkv : bf16 [M = B*T, 3*D]  (q | k | v, each D = H*64; q,k already LayerNormed + RoPE'd)
//   ctx : bf16 [M, D]
#pragma once
#include "../../esmdiff_b200/csrc/attention_resident.cuh"

namespace esmdiff {
namespace attn5 {

using attn2::BKV;
using attn2::BQ;
using attn2::DH;
using attn2::KV_TILE_BYTES;
using attn2::MAX_KV_TILES;
using attn2::MAX_LEFT;
using attn2::Q_BYTES;
constexpr int BAR_BYTES = 512;
constexpr int THREADS = 256;                // warps 0-3 softmax, 4 TMA, 5 MMA (leader), 6-7 trailing rows (leader)
constexpr int TMEM_COLS = 256;              // S0 [0,64) S1 [64,128) O [128,192)
constexpr int COL_O = 128;
constexpr float RESCALE_LOG2 = 8.0f;

struct Params {
    int B, T, H;
    int nq;                     // full 128-row query tiles handled on the tensor path (>= 2)
    int n_left;                 // 0..MAX_LEFT trailing query rows (leader CTA, CUDA cores)
    int nkv;                    // ceil(T / 64)
    int tail_cols;              // width of the last kv tile: multiple of 16 in [16, 64]
    const __nv_bfloat16* qkv;
    __nv_bfloat16* ctx;
    float scale_log2;
};

__host__ inline int smem_bytes(int nkv, int tail_cols) {
    return 1024 + 2 * Q_BYTES + 2 * attn2::kv_bytes(nkv, tail_cols) + BAR_BYTES + attn2::left_bytes(nkv);
}

// D[tmem of both CTAs] (+)= A[tmem of each CTA] * B[smem halves of both CTAs]
__device__ __forceinline__ void umma_bf16_ts_pair(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                                  uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 2)
attention_pair_kernel(const __grid_constant__ CUtensorMap tmQ,      // box [128 rows][64 cols]
                      const __grid_constant__ CUtensorMap tmK32,    // box [ 32 rows][64 cols]
                      const __grid_constant__ CUtensorMap tmKt,     // box [tail/2 rows][64 cols]
                      const __grid_constant__ CUtensorMap tmV,      // box [ 64 rows][64 cols]
                      const __grid_constant__ CUtensorMap tmVt,     // box [tail rows][64 cols]
                      const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const int kvb = attn2::kv_bytes(p.nkv, p.tail_cols);
    uint8_t* sQ = smem;                                   // two query-tile buffers (tile pair tp -> buffer tp & 1)
    uint8_t* sK = sQ + 2 * Q_BYTES;
    uint8_t* sV = sK + kvb;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kvb);
    uint64_t* q_full = bars;                              // [2]  TMA of both CTAs -> MMA           (leader's copy)
    uint64_t* q_empty = bars + 2;                         // [2]  MMA -> TMA of both CTAs            (multicast)
    uint64_t* s_full = bars + 4;                          // [2]  MMA -> softmax of both CTAs        (multicast)
    uint64_t* p_full = bars + 6;                          // [2]  softmax warps of both CTAs -> MMA  (leader's copy, 8 arrivals)
    uint64_t* pv_done = bars + 8;                         // 1    P V of step nsteps-2 retired       (multicast)
    uint64_t* o_free = bars + 9;                          // 1    softmax warps of both CTAs -> MMA  (leader's copy, 8 arrivals)
    uint64_t* o_full = bars + 10;                         // 1    last P V of a tile pair retired    (multicast)
    uint64_t* k_full = bars + 11;                         // [MAX_KV_TILES]  TMA of both CTAs        (leader's copy)
    uint64_t* v_full = k_full + MAX_KV_TILES;             // [MAX_KV_TILES]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(v_full + MAX_KV_TILES);
    float* left_p = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + BAR_BYTES);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();              // 0 = leader
    const int item = blockIdx.x >> 1;
    const int h = item % p.H;
    const int b = item / p.H;
    const int D = p.H * DH;
    const int row0 = b * p.T;
    const int nq = p.nq, nkv = p.nkv;
    const int npairs = (nq + 1) >> 1;                     // tile pairs; the last one may have an idle CTA 1
    const int nsteps = npairs * nkv;
    const int half_tail = p.tail_cols >> 1;

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK32);
        tma_prefetch_desc(&tmKt);
        tma_prefetch_desc(&tmV);
        tma_prefetch_desc(&tmVt);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&q_full[s], 1);
            mbar_init(&q_empty[s], 1);
            mbar_init(&s_full[s], 1);
            mbar_init(&p_full[s], 8);
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
    if (warp == 5) {
        tmem_alloc_pair(tmem_slot, TMEM_COLS);
        tmem_relinquish_pair();
    }
    tcgen05_fence_before();
    cluster_sync_all();                    // peer barriers initialised before any remote arrive / TMA
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        // ===================== TMA producer (both CTAs, completion on the leader's barriers) =====================
        if (lane == 0) {
            auto load_q = [&](int tp) {
                const int qt = min(2 * tp + static_cast<int>(rank), nq - 1);     // an idle CTA 1 reloads the last tile
                const uint32_t bar = mapa_shared(smem_u32(&q_full[tp & 1]), 0);
                if (rank == 0) mbar_arrive_expect_tx(&q_full[tp & 1], 2 * Q_BYTES);
                tma_load_2d_pair(sQ + (tp & 1) * Q_BYTES, &tmQ, bar, h * DH, row0 + qt * BQ);
            };
            // K tile j: the half this CTA supplies to the UMMA first (keys 32 r .. of the tile), then the other
            auto load_k = [&](int j) {
                const bool last = j == nkv - 1;
                const int hr = last ? half_tail : BKV / 2;                       // rows per half
                const uint32_t bar = mapa_shared(smem_u32(&k_full[j]), 0);
                if (rank == 0) mbar_arrive_expect_tx(&k_full[j], 2 * 2 * hr * 128);
                const CUtensorMap* tm = last ? &tmKt : &tmK32;
                uint8_t* dst = sK + j * KV_TILE_BYTES;
                tma_load_2d_pair(dst, tm, bar, D + h * DH, row0 + j * BKV + static_cast<int>(rank) * hr);
                tma_load_2d_pair(dst + hr * 128, tm, bar, D + h * DH, row0 + j * BKV + static_cast<int>(1 - rank) * hr);
            };
            // V tile j: all key rows, head dims [32 r, 32 r + 64) (the UMMA takes the first 32 columns per CTA)
            auto load_v = [&](int j) {
                const bool last = j == nkv - 1;
                const int bytes = last ? p.tail_cols * 128 : KV_TILE_BYTES;
                const uint32_t bar = mapa_shared(smem_u32(&v_full[j]), 0);
                if (rank == 0) mbar_arrive_expect_tx(&v_full[j], 2 * bytes);
                tma_load_2d_pair(sV + j * KV_TILE_BYTES, last ? &tmVt : &tmV, bar, 2 * D + h * DH + static_cast<int>(rank) * 32,
                                 row0 + j * BKV);
            };
            load_q(0);
            for (int j = 0; j < nkv; ++j) load_k(j);
            if (npairs > 1) load_q(1);
            for (int j = 0; j < nkv; ++j) load_v(j);
            for (int tp = 2; tp < npairs; ++tp) {
                mbar_wait(&q_empty[tp & 1], ((tp >> 1) - 1) & 1);
                load_q(tp);
            }
        }
    } else if (warp == 5) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (rank == 0) {
            constexpr uint32_t idesc_pv = umma_idesc_bf16(2 * BQ, DH, 1);       // P V : V is MN-major
            const uint32_t idesc_s_full = umma_idesc_bf16(2 * BQ, BKV, 0);
            const uint32_t idesc_s_tail = umma_idesc_bf16(2 * BQ, p.tail_cols, 0);
            const uint64_t desc_q0 = umma_desc_sw128(smem_u32(sQ), 16, 1024);
            const uint64_t desc_k0 = umma_desc_sw128(smem_u32(sK), 16, 1024);
            const uint64_t desc_v0 = umma_desc_sw128(smem_u32(sV), 16, 1024);
            const uint32_t tmem_o = tmem_base + COL_O;
            int s_i = 0, s_tp = 0, s_j = 0;                                 // next S = Q K^T to issue
            auto issue_s = [&]() {
                if (s_j == 0) mbar_wait(&q_full[s_tp & 1], (s_tp >> 1) & 1);
                if (s_tp == 0) mbar_wait(&k_full[s_j], 0);
                tcgen05_fence_after();
                const uint64_t qdesc = desc_q0 + static_cast<uint64_t>((s_tp & 1) * (Q_BYTES >> 4));
                const uint64_t kdesc = desc_k0 + static_cast<uint64_t>(s_j * (KV_TILE_BYTES >> 4));
                const uint32_t idesc = s_j == nkv - 1 ? idesc_s_tail : idesc_s_full;
                const uint32_t ts = tmem_base + (s_i & 1) * BKV;
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < DH / 16; ++k)
                        umma_bf16_ss_pair(ts, qdesc + 2 * k, kdesc + 2 * k, idesc, k != 0 ? 1u : 0u);
                    umma_commit_pair(&s_full[s_i & 1]);
                    if (s_j == nkv - 1) umma_commit_pair(&q_empty[s_tp & 1]);
                }
                __syncwarp();
                ++s_i;
                if (++s_j == nkv) { s_j = 0; ++s_tp; }
            };
            issue_s();
            if (nsteps > 1) issue_s();
            int i = 0;
            for (int tp = 0; tp < npairs; ++tp) {
                for (int j = 0; j < nkv; ++j, ++i) {
                    if (tp == 0) mbar_wait(&v_full[j], 0);
                    mbar_wait(&p_full[i & 1], (i >> 1) & 1);
                    if (j == 0 && tp > 0) mbar_wait(o_free, (tp - 1) & 1);
                    tcgen05_fence_after();
                    const uint64_t vdesc = desc_v0 + static_cast<uint64_t>(j * (KV_TILE_BYTES >> 4));
                    const uint32_t tpm = tmem_base + (i & 1) * BKV;         // P: bf16 pairs, 8 columns per K-step
                    const bool last = j == nkv - 1;
                    const int ksteps = last ? p.tail_cols >> 4 : BKV / 16;
                    if (elect_one()) {
                        for (int k = 0; k < ksteps; ++k)
                            umma_bf16_ts_pair(tmem_o, tpm + 8 * k, vdesc + 128 * k, idesc_pv, (j | k) != 0 ? 1u : 0u);
                        if (last) umma_commit_pair(o_full);
                        else if (s_i >= nsteps) umma_commit_pair(pv_done);
                    }
                    __syncwarp();
                    if (s_i < nsteps) issue_s();              // overwrites P_i's buffer: ordered after PV_i
                }
            }
        }
    } else if (warp >= 6) {
        // ===================== trailing query rows (leader: its K/V copies are in natural order) =====================
        const int lw = warp - 6;
        if (rank == 0 && lw < p.n_left) {
            const int t = nq * BQ + lw;
            attn2::leftover_row(p.qkv + static_cast<long long>(row0 + t) * 3 * D + h * DH,
                                p.ctx + static_cast<long long>(row0 + t) * D + h * DH, sK, sV, left_p + lw * nkv * BKV, p.T,
                                p.scale_log2, k_full, v_full, nkv, lane);
        }
    } else {
        // ===================== softmax / output warps: thread = query row of this CTA's tile =====================
        const int r = threadIdx.x;                                   // 0..127 == TMEM lane
        const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
        const uint32_t t_o = tmem_base + lane_addr + COL_O;
        const float sc = p.scale_log2;
        const float thresh = RESCALE_LOG2 / sc;
        const uint32_t p_full_leader[2] = {mapa_shared(smem_u32(&p_full[0]), 0), mapa_shared(smem_u32(&p_full[1]), 0)};
        const uint32_t o_free_leader = mapa_shared(smem_u32(o_free), 0);

        auto epilogue = [&](int tp, float l) {
            const int qt = 2 * tp + static_cast<int>(rank);
            mbar_wait(o_full, tp & 1);
            tcgen05_fence_after();
            if (qt < nq && qt * BQ + warp * 32 < p.T) {
                const float inv = 1.0f / l;
                const int t = qt * BQ + r;
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
            if (lane == 0) mbar_arrive_cluster_relaxed(o_free_leader);
        };

        float m_run = 0.f;
        float l_run = 0.f, l_prev = 0.f;
        for (int tp = 0; tp < npairs; ++tp) {
            const int qt = 2 * tp + static_cast<int>(rank);
            const bool active = qt < nq && qt * BQ + warp * 32 < p.T;    // warp-uniform
            for (int j = 0; j < nkv; ++j) {
                const int i = tp * nkv + j;
                const uint32_t t_s = tmem_base + lane_addr + (i & 1) * BKV;
                if (j == 0) {
                    l_prev = l_run;
                    l_run = 0.f;
                }
                mbar_wait(&s_full[i & 1], (i >> 1) & 1);
                tcgen05_fence_after();
                if (active) {
                    const bool last = j == nkv - 1;
                    const int nch = (last ? p.tail_cols : BKV) >> 4;
                    uint32_t s[64];
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (c < nch) tmem_ld_32x32b_x16(t_s + c * 16, s + c * 16);
                    tmem_ld_wait();
                    float mx = -INFINITY;
                    if (last) {
                        const int valid = p.T - j * BKV;
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
#pragma unroll
                        for (int e = 0; e < 64; ++e) mx = fmaxf(mx, __uint_as_float(s[e]));
                    }
                    if (j == 0) {
                        m_run = mx;
                    } else {
                        const bool need = mx > m_run + thresh;
                        if (__any_sync(0xffffffffu, need)) {
                            const float m_new = need ? mx : m_run;
                            const float f = fast_exp2((m_run - m_new) * sc);
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
                                const float p0 = fast_exp2(fmaf(__uint_as_float(s[c * 16 + 2 * e]), sc, nm));
                                const float p1 = fast_exp2(fmaf(__uint_as_float(s[c * 16 + 2 * e + 1]), sc, nm));
                                rs0 += p0;
                                rs1 += p1;
                                pk[e] = pack_bf16x2(p0, p1);
                            }
                            tmem_st_32x32b_x8(t_s + c * 8, pk);
                        }
                    l_run += rs0 + rs1;
                    tmem_st_wait();
                }
                // (inactive warps -- rows past T, or the idle CTA 1 of an odd tile count -- leave their P rows as
                // they are: every output row depends on its own P row only, and theirs are never stored)
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster_relaxed(p_full_leader[i & 1]);
                if (j == 0 && tp > 0) epilogue(tp - 1, l_prev);
            }
        }
        epilogue(npairs - 1, l_run);
    }

    tcgen05_fence_before();
    cluster_sync_all();     // the leader's MMAs read the peer's shared and tensor memory: nobody leaves early
    if (warp == 5) {
        tcgen05_fence_after();
        tmem_dealloc_pair(tmem_base, TMEM_COLS);
    }
}

}  // namespace attn5
}  // namespace esmdiff
