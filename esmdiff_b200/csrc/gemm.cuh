// Persistent warp-specialised tcgen05 GEMM for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T
//   A, W : bf16, K-major (activations row-major, nn.Linear weights as stored)
//   accumulate fp32 in TMEM, two accumulator stages so the epilogue of tile i overlaps the
//   main loop of tile i+1; TMA (SWIZZLE_128B) feeds a 4-stage shared-memory ring.
// Replaces the cuBLAS sgemm calls behind nn.Linear on the reference path (SURVEY.md 2.2 k3,
// k7, k8, k10, k13) with the element-wise tails fused into the epilogue:
//   EPI_STORE_BF16      out = acc                               (QKV projection)
//   EPI_RESID_F32       x  += acc / scale                       (out_proj, FFN W2; blocks.py residual)
//   EPI_SWIGLU_BF16     out = silu(gate) * up                   (FFN W1 + SwiGLU; gate/up rows
//                                                                interleaved per 128 offline)
//   EPI_BIAS_GELU_F32   out = gelu(acc + bias)                  (RegressionHead Linear+GELU)
//   EPI_BIAS_F32        out = acc + bias   (ragged N, e.g. 4101) (RegressionHead output Linear)
#pragma once
#include "ptx.cuh"

namespace esmdiff {
namespace gemm {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;                    // 64 bf16 = one 128-byte swizzle row
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;      // 16 KiB
constexpr int B_BYTES = BN * BK * 2;      // 32 KiB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int TMEM_COLS = 2 * BN;         // two fp32 accumulator stages = all 512 columns
constexpr int EPI_WARPS = 4;
constexpr int STG_LD = 33;                // padded row of the per-warp transpose buffer
constexpr int STG_BYTES = EPI_WARPS * 32 * STG_LD * 4;
constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + STG_BYTES + 256;
constexpr int THREADS = 256;              // w0 TMA, w1 MMA, w2 TMEM alloc, w3 idle, w4-7 epilogue

enum Epilogue {
    EPI_STORE_BF16 = 0,
    EPI_RESID_F32 = 1,
    EPI_SWIGLU_BF16 = 2,
    EPI_BIAS_GELU_F32 = 3,
    EPI_BIAS_F32 = 4,
};

struct Params {
    int M, N, K;          // N = rows of W actually present (TMA zero-fills beyond)
    int m_tiles, n_tiles;
    void* out;            // bf16 or fp32, see Epilogue
    long long ldo;        // elements between output rows
    const float* bias;    // [N] or null
    float scale;          // EPI_RESID_F32: divisor of the branch output
};

__device__ __forceinline__ float gelu_erf(float x) {
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float silu(float x) { return x / (1.0f + expf(-x)); }

template <int EPI>
__global__ void __launch_bounds__(THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA,
                    const __grid_constant__ CUtensorMap tmB, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    float* stg_all = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + STG_BYTES);
    uint64_t* full = bars;                 // [STAGES]  TMA -> MMA
    uint64_t* empty = bars + STAGES;       // [STAGES]  MMA -> TMA
    uint64_t* tfull = bars + 2 * STAGES;   // [2]       MMA -> epilogue
    uint64_t* tempty = tfull + 2;          // [2]       epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = p.m_tiles * p.n_tiles;
    const int kblocks = p.K / BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], EPI_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m0 = (tile / p.n_tiles) * BM;
                const int n0 = (tile % p.n_tiles) * BN;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
                    tma_load_2d(sa, &tmA, &full[stage], kb * BK, m0);
                    tma_load_2d(sa + A_BYTES, &tmB, &full[stage], kb * BK, n0);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, 0);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                    const uint64_t adesc = umma_desc_sw128(sa, 16, 1024);
                    const uint64_t bdesc = umma_desc_sw128(sa + A_BYTES, 16, 1024);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        // +32 bytes (16 bf16) along K inside the 128B swizzle row = +2 encoded
                        umma_bf16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc,
                                     (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty[stage]);       // frees the smem slot when the MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull[acc]);             // accumulator ready for the epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue warps =====================
        const int ew = warp - 4;                      // == warp % 4 -> TMEM lanes [32 ew, 32 ew + 32)
        float* stg = stg_all + ew * 32 * STG_LD;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m0 = (tile / p.n_tiles) * BM;
            const int nb = tile % p.n_tiles;
            const int n0 = nb * BN;
            mbar_wait(&tfull[acc], acc_phase);
            tcgen05_fence_after();
            const uint32_t t_row = tmem_base + acc * BN + (static_cast<uint32_t>(ew * 32) << 16);
            const int row_base = m0 + ew * 32;
            constexpr int NCHUNK = (EPI == EPI_SWIGLU_BF16) ? (BN / 2) / 32 : BN / 32;
#pragma unroll 1
            for (int c = 0; c < NCHUNK; ++c) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(t_row + c * 32, v);
                if constexpr (EPI == EPI_SWIGLU_BF16) {
                    uint32_t w[32];
                    tmem_ld_32x32b_x32(t_row + BN / 2 + c * 32, w);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        stg[lane * STG_LD + j] = silu(__uint_as_float(v[j])) * __uint_as_float(w[j]);
                } else {
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) stg[lane * STG_LD + j] = __uint_as_float(v[j]);
                }
                __syncwarp();
                // transposed read-back: lane = column, coalesced row segments to global
                if constexpr (EPI == EPI_SWIGLU_BF16) {
                    const int col = nb * (BN / 2) + c * 32 + lane;
                    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
                    if (col < p.N / 2) {
#pragma unroll 8
                        for (int r = 0; r < 32; ++r) {
                            const int row = row_base + r;
                            if (row < p.M)
                                out[static_cast<long long>(row) * p.ldo + col] =
                                    __float2bfloat16_rn(stg[r * STG_LD + lane]);
                        }
                    }
                } else {
                    const int col = n0 + c * 32 + lane;
                    if (col < p.N) {
                        float b = 0.f;
                        if constexpr (EPI == EPI_BIAS_GELU_F32 || EPI == EPI_BIAS_F32)
                            b = p.bias[col];
                        if constexpr (EPI == EPI_STORE_BF16) {
                            __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
#pragma unroll 8
                            for (int r = 0; r < 32; ++r) {
                                const int row = row_base + r;
                                if (row < p.M)
                                    out[static_cast<long long>(row) * p.ldo + col] =
                                        __float2bfloat16_rn(stg[r * STG_LD + lane]);
                            }
                        } else if constexpr (EPI == EPI_RESID_F32) {
                            // read-modify-write of the fp32 residual stream: issue all 32 loads
                            // before the first store so they overlap (same pointer -> the
                            // compiler would otherwise serialise load/store pairs).
                            float* out = reinterpret_cast<float*>(p.out) +
                                         static_cast<long long>(row_base) * p.ldo + col;
                            float old[32];
#pragma unroll
                            for (int r = 0; r < 32; ++r)
                                old[r] = (row_base + r < p.M) ? __ldcg(out + static_cast<long long>(r) * p.ldo) : 0.f;
#pragma unroll
                            for (int r = 0; r < 32; ++r)
                                if (row_base + r < p.M)
                                    out[static_cast<long long>(r) * p.ldo] =
                                        old[r] + stg[r * STG_LD + lane] / p.scale;
                        } else {
                            float* out = reinterpret_cast<float*>(p.out);
#pragma unroll 8
                            for (int r = 0; r < 32; ++r) {
                                const int row = row_base + r;
                                if (row < p.M) {
                                    float y = stg[r * STG_LD + lane] + b;
                                    if constexpr (EPI == EPI_BIAS_GELU_F32) y = gelu_erf(y);
                                    out[static_cast<long long>(row) * p.ldo + col] = y;
                                }
                            }
                        }
                    }
                }
                __syncwarp();
            }
            tcgen05_fence_before();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace gemm
}  // namespace esmdiff
