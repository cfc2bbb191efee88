// Non-causal, unmasked multi-head attention for d_head = 64 on sm_100a tensor cores (tcgen05).
// Replaces F.scaled_dot_product_attention on the reference path (SURVEY.md 2.2 k6;
// esm MultiHeadAttention.forward with seq_id None).  softmax(q k^T / 8) v per (sample, head).
//
// One CTA = one (sample b, head h, 128-row query tile).  KV is walked in tiles of 64:
//   TMA warp : Q once, then K_j / V_j into two 3-slot rings (SWIZZLE_128B tiles of [rows][64]).
//   MMA warp : S_j = Q K_j^T  (UMMA 128x64x16 x4, fp32 in TMEM, two S buffers)
//              O_j = P_j V_j  (UMMA 128x64x16 x4, V as MN-major B operand, fresh accumulator)
//   4 softmax warps (thread = query row = TMEM lane): online softmax in the log2 domain, P_j
//              written as bf16 into a 128B-swizzled K-major smem tile, running output kept in
//              registers and rescaled there (no TMEM read-modify-write).
// Two CTAs fit per SM (about 81 KiB smem, 256 TMEM columns each), so one CTA's softmax overlaps
// the other's MMAs.
// Input  qkv : bf16 [M = B*T, 3*D]  (q | k | v, each D = H*64).  With qk_sumsq == null q, k are
//   already LayerNormed + RoPE'd; otherwise they are the un-normalised q', k' of gemm.cuh's
//   EPI_QKV_ROPE_LN and the per-row factors 1/std are applied here in fp32: rstd_q[i] in the scale
//   of query row i, rstd_k[j] per score column from a shared-memory table (K tiles are streamed
//   through a ring here, so they are not rescaled in place as attention_resident.cuh does).
// Output ctx : bf16 [M, D]
#pragma once
#include "ptx.cuh"

namespace esmdiff {
namespace attn {

constexpr int BQ = 128;
constexpr int BKV = 64;
constexpr int DH = 64;
constexpr int KV_SLOTS = 3;
constexpr int Q_BYTES = BQ * DH * 2;        // 16 KiB
constexpr int KV_BYTES = BKV * DH * 2;      // 8 KiB
constexpr int P_BYTES = BQ * BKV * 2;       // 16 KiB
constexpr int SMEM_BYTES = 1024 + Q_BYTES + 2 * KV_SLOTS * KV_BYTES + P_BYTES + 256;
constexpr int TMEM_COLS = 256;              // S0 [0,64) S1 [64,128) O [128,192)
constexpr int THREADS = 192;                // warps 0-3 softmax, 4 TMA, 5 MMA

struct Params {
    int B, T, H;
    int q_tiles;                // ceil(T / BQ)
    __nv_bfloat16* ctx;         // [B*T, H*64]
    float scale_log2;           // (1/sqrt(64)) * log2(e)
    const float* qk_sumsq;      // [B*T][2 * nspan] (q spans then k spans) or null
    int nspan;                  // D / 128
    float ln_eps;
};
__host__ inline int smem_bytes(int T) { return SMEM_BYTES + ((T + BKV - 1) / BKV) * BKV * 4; }

__global__ void __launch_bounds__(THREADS, 2)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV_q,    // box [128 rows][64 cols]
                     const __grid_constant__ CUtensorMap tmQKV_kv,   // box [ 64 rows][64 cols]
                     const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + Q_BYTES;
    uint8_t* sV = sK + KV_SLOTS * KV_BYTES;
    uint8_t* sP = sV + KV_SLOTS * KV_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + P_BYTES);
    uint64_t* q_full = bars;                  // 1
    uint64_t* k_full = bars + 1;              // [3]
    uint64_t* k_empty = bars + 4;             // [3]
    uint64_t* v_full = bars + 7;              // [3]
    uint64_t* v_empty = bars + 10;            // [3]
    uint64_t* s_full = bars + 13;             // [2]
    uint64_t* s_empty = bars + 15;            // [2]
    uint64_t* p_full = bars + 17;             // 1
    uint64_t* o_full = bars + 18;             // 1
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);
    float* rk_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [nkv * 64] rstd_k per key

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    int idx = blockIdx.x;
    const int qt = idx % p.q_tiles; idx /= p.q_tiles;
    const int h = idx % p.H;
    const int b = idx / p.H;
    const int D = p.H * DH;
    const int row0 = b * p.T;                   // first token row of this sample
    const int q0 = qt * BQ;                     // first query position of this tile
    const int nkv = (p.T + BKV - 1) / BKV;
    const bool fused_ln = p.qk_sumsq != nullptr;
    pdl_launch_dependents();
    pdl_wait();

    if (fused_ln) {
        for (int t = threadIdx.x; t < nkv * BKV; t += THREADS) {
            float rk = 0.f;
            if (t < p.T) {
                const float* part = p.qk_sumsq + static_cast<long long>(row0 + t) * 2 * p.nspan + p.nspan;
                float ss = 0.f;
                for (int i = 0; i < p.nspan; ++i) ss += __ldg(part + i);
                rk = rsqrtf(ss / static_cast<float>(p.nspan * 128) + p.ln_eps);
            }
            rk_s[t] = rk;
        }
    }
    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmQKV_q);
        tma_prefetch_desc(&tmQKV_kv);
        mbar_init(q_full, 1);
        for (int s = 0; s < KV_SLOTS; ++s) {
            mbar_init(&k_full[s], 1);
            mbar_init(&k_empty[s], 1);
            mbar_init(&v_full[s], 1);
            mbar_init(&v_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&s_full[s], 1);
            mbar_init(&s_empty[s], 128);
        }
        mbar_init(p_full, 128);
        mbar_init(o_full, 1);
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
    const uint32_t tmem_o = tmem_base + 128;

    if (warp == 4) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, Q_BYTES);
            tma_load_2d(sQ, &tmQKV_q, q_full, h * DH, row0 + q0);
            for (int j = 0; j < nkv; ++j) {
                const int slot = j % KV_SLOTS;
                const uint32_t ph = (j / KV_SLOTS) & 1;
                mbar_wait(&k_empty[slot], ph ^ 1);
                mbar_arrive_expect_tx(&k_full[slot], KV_BYTES);
                tma_load_2d(sK + slot * KV_BYTES, &tmQKV_kv, &k_full[slot], D + h * DH,
                            row0 + j * BKV);
                mbar_wait(&v_empty[slot], ph ^ 1);
                mbar_arrive_expect_tx(&v_full[slot], KV_BYTES);
                tma_load_2d(sV + slot * KV_BYTES, &tmQKV_kv, &v_full[slot], 2 * D + h * DH,
                            row0 + j * BKV);
            }
        }
    } else if (warp == 5) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc_s = umma_idesc_bf16(BQ, BKV, 0);   // Q K^T : both K-major
            constexpr uint32_t idesc_o = umma_idesc_bf16(BQ, DH, 1);    // P V   : V is MN-major
            const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ), 16, 1024);
            const uint64_t pdesc = umma_desc_sw128(smem_u32(sP), 16, 1024);
            auto issue_s = [&](int j) {
                const int slot = j % KV_SLOTS;
                mbar_wait(&k_full[slot], (j / KV_SLOTS) & 1);
                mbar_wait(&s_empty[j & 1], ((j >> 1) & 1) ^ 1);
                tcgen05_fence_after();
                const uint64_t kdesc = umma_desc_sw128(smem_u32(sK + slot * KV_BYTES), 16, 1024);
#pragma unroll
                for (int k = 0; k < DH / 16; ++k)
                    umma_bf16_ss(tmem_base + (j & 1) * BKV, qdesc + 2 * k, kdesc + 2 * k, idesc_s,
                                 k != 0 ? 1u : 0u);
                umma_commit(&s_full[j & 1]);
                umma_commit(&k_empty[slot]);
            };
            mbar_wait(q_full, 0);
            issue_s(0);
            for (int j = 0; j < nkv; ++j) {
                if (j + 1 < nkv) issue_s(j + 1);
                const int slot = j % KV_SLOTS;
                mbar_wait(&v_full[slot], (j / KV_SLOTS) & 1);
                mbar_wait(p_full, j & 1);
                tcgen05_fence_after();
                // V tile [64 kv rows][64 d] is an MN-major B operand: 128-byte rows along N=d,
                // 8-row (k) groups 1024 B apart; one UMMA K-step (16 kv rows) = 2048 B.
                const uint64_t vdesc = umma_desc_sw128(smem_u32(sV + slot * KV_BYTES), 16, 1024);
#pragma unroll
                for (int k = 0; k < BKV / 16; ++k)
                    umma_bf16_ss(tmem_o, pdesc + 2 * k, vdesc + 128 * k, idesc_o, k != 0 ? 1u : 0u);
                umma_commit(o_full);
                umma_commit(&v_empty[slot]);
            }
        }
    } else {
        // ===================== softmax / output warps: thread = query row =====================
        const int r = threadIdx.x;                                   // 0..127 == TMEM lane
        const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
        float m_run = -INFINITY;       // running max, log2 domain (already scaled)
        float l_run = 0.f;
        float acc[DH];
#pragma unroll
        for (int d = 0; d < DH; ++d) acc[d] = 0.f;
        uint8_t* p_row = sP + r * 128;
        const int sw = r & 7;
        float sc = p.scale_log2;
        if (fused_ln && q0 + r < p.T) {
            const float* part = p.qk_sumsq + static_cast<long long>(row0 + q0 + r) * 2 * p.nspan;
            float ss = 0.f;
            for (int i = 0; i < p.nspan; ++i) ss += __ldg(part + i);
            sc *= rsqrtf(ss / static_cast<float>(p.nspan * 128) + p.ln_eps);
        }

        // acc += O_{j}: both are relative to the running max m_run at the time of the call
        auto fold_o = [&](uint32_t parity) {
            mbar_wait(o_full, parity);
            tcgen05_fence_after();
            uint32_t o[32];
            tmem_ld_32x32b_x32(tmem_o + lane_addr, o);
            tmem_ld_wait();
#pragma unroll
            for (int d = 0; d < 32; ++d) acc[d] += __uint_as_float(o[d]);
            tmem_ld_32x32b_x32(tmem_o + lane_addr + 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int d = 0; d < 32; ++d) acc[d + 32] += __uint_as_float(o[d]);
        };

        for (int j = 0; j < nkv; ++j) {
            // O_{j-1} retires after S_j on the tensor pipe, so this wait also covers s_full.
            if (j > 0) fold_o((j - 1) & 1);
            mbar_wait(&s_full[j & 1], (j >> 1) & 1);
            tcgen05_fence_after();
            uint32_t s0[32], s1[32];
            tmem_ld_32x32b_x32(tmem_base + lane_addr + (j & 1) * BKV, s0);
            tmem_ld_32x32b_x32(tmem_base + lane_addr + (j & 1) * BKV + 32, s1);
            tmem_ld_wait();
            tcgen05_fence_before();
            mbar_arrive(&s_empty[j & 1]);

            const int kv_valid = p.T - j * BKV;        // >= 1; < 64 only in the last tile
            // Scores stay UNSCALED by the per-row factor sc (> 0, so the max commutes): the per-column
            // factor rstd_k takes the slot of the old scale multiply and sc moves into the exp2 argument
            // as an FMA -- folding q_ln / k_ln's 1/std in costs no arithmetic instruction here, only the
            // broadcast reads of the table.
            float mx = -INFINITY;
            if (fused_ln) {
                const float4* rk4 = reinterpret_cast<const float4*>(rk_s + j * BKV);
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    const float4 ra = rk4[c4], rb = rk4[c4 + 8];
                    const float fa[4] = {ra.x, ra.y, ra.z, ra.w}, fb[4] = {rb.x, rb.y, rb.z, rb.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        s0[4 * c4 + e] = __float_as_uint(__uint_as_float(s0[4 * c4 + e]) * fa[e]);
                        s1[4 * c4 + e] = __float_as_uint(__uint_as_float(s1[4 * c4 + e]) * fb[e]);
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const float a = (c < kv_valid) ? __uint_as_float(s0[c]) : -INFINITY;
                const float bq = (c + 32 < kv_valid) ? __uint_as_float(s1[c]) : -INFINITY;
                s0[c] = __float_as_uint(a);
                s1[c] = __float_as_uint(bq);
                mx = fmax3(mx, a, bq);
            }
            const float m_new = fmaxf(m_run, mx);
            const float alpha = fast_exp2((m_run - m_new) * sc);      // 0 on the first tile (m_run = -inf)
#pragma unroll
            for (int d = 0; d < DH; ++d) acc[d] *= alpha;
            const float nm = -m_new * sc;

            float rs = 0.f;
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
                float e[8];
                uint32_t w[4];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    e[i] = fast_exp2(fmaf(__uint_as_float(s0[c8 * 8 + i]), sc, nm));
                    rs += e[i];
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) w[i] = pack_bf16x2(e[2 * i], e[2 * i + 1]);
                *reinterpret_cast<uint4*>(p_row + ((c8 ^ sw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
            }
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
                float e[8];
                uint32_t w[4];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    e[i] = fast_exp2(fmaf(__uint_as_float(s1[c8 * 8 + i]), sc, nm));
                    rs += e[i];
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) w[i] = pack_bf16x2(e[2 * i], e[2 * i + 1]);
                *reinterpret_cast<uint4*>(p_row + (((c8 + 4) ^ sw) << 4)) =
                    make_uint4(w[0], w[1], w[2], w[3]);
            }
            l_run = l_run * alpha + rs;
            m_run = m_new;
            fence_proxy_async_smem();          // generic-proxy smem writes -> visible to UMMA
            tcgen05_fence_before();
            mbar_arrive(p_full);
        }
        // last partial product, normalise, store
        fold_o((nkv - 1) & 1);
        {
            const float inv = 1.0f / l_run;
            const int t = q0 + r;
            if (t < p.T) {
                uint4* dst = reinterpret_cast<uint4*>(
                    p.ctx + static_cast<long long>(row0 + t) * D + h * DH);
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) {
                    const float* y = acc + c8 * 8;
                    dst[c8] = make_uint4(pack_bf16x2(y[0] * inv, y[1] * inv),
                                         pack_bf16x2(y[2] * inv, y[3] * inv),
                                         pack_bf16x2(y[4] * inv, y[5] * inv),
                                         pack_bf16x2(y[6] * inv, y[7] * inv));
                }
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 5) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace attn
}  // namespace esmdiff
