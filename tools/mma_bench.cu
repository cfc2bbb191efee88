// Dev tool: issue cost of small tcgen05.mma instructions (M = 128, K = 16, bf16) on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mma_bench.bin tools/mma_bench.cu
// One warp per CTA issues `reps` groups of `per_commit` MMAs followed by one commit, then waits for
// the last commit.  Prints cycles per MMA for SS (A in smem) and TS (A in TMEM) operands.
#include <cstdio>
#include <cstdlib>
#include "../esmdiff_b200/csrc/ptx.cuh"
using namespace esmdiff;

template <int N, bool TS>
__global__ void __launch_bounds__(128) k(long long* out, int reps, int per_commit) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(&slot, 256); tmem_relinquish(); }
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tm = slot;
    if (warp == 1) {
        constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0);
        const uint64_t adesc = umma_desc_sw128(smem_u32(smem), 16, 1024);
        const uint64_t bdesc = umma_desc_sw128(smem_u32(smem + 16384), 16, 1024);
        long long t0 = clock64();
        uint32_t phase = 0;
        for (int r = 0; r < reps; ++r) {
            if (elect_one()) {
                for (int i = 0; i < per_commit; ++i) {
                    if (TS) umma_bf16_ts(tm + 128, tm + 8 * (i & 3), bdesc + 2 * (i & 3), idesc, 1u);
                    else umma_bf16_ss(tm + 128, adesc + 2 * (i & 3), bdesc + 2 * (i & 3), idesc, 1u);
                }
                umma_commit(&bar);
            }
            __syncwarp();
        }
        long long t1 = clock64();
        // wait for all commits
        for (int r = 0; r < reps; ++r) { mbar_wait(&bar, phase); phase ^= 1; }
        long long t2 = clock64();
        if (threadIdx.x == 32) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) { tcgen05_fence_after(); tmem_dealloc(tm, 256); }
}

template <int N, bool TS>
void run(int grid, int reps, int per_commit, int smem) {
    long long* d; cudaMalloc(&d, grid * 16);
    cudaFuncSetAttribute(k<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<N, TS><<<grid, 128, smem>>>(d, reps, per_commit);
    cudaDeviceSynchronize();
    k<N, TS><<<grid, 128, smem>>>(d, reps, per_commit);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("N=%3d %s grid=%3d smem=%3dK per_commit=%2d: issue %.1f clk/MMA, complete %.1f clk/MMA (%s)\n", N, TS ? "TS" : "SS",
           grid, smem >> 10, per_commit, (double)h[0] / (reps * per_commit), (double)h[1] / (reps * per_commit),
           cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    const int one = 120 * 1024, two = 100 * 1024;     // smem per CTA: 1 or 2 CTAs per SM
    for (int pc : {1, 2, 4, 16}) {
        run<64, false>(148, 256 / pc, pc, one);
        run<64, true>(148, 256 / pc, pc, one);
        run<32, false>(148, 256 / pc, pc, one);
        run<256, false>(148, 256 / pc, pc, one);
        run<64, false>(296, 256 / pc, pc, two);
        run<64, true>(296, 256 / pc, pc, two);
        run<32, false>(296, 256 / pc, pc, two);
    }
    run<16, false>(148, 64, 4, one);
    run<128, false>(148, 64, 4, one);
    run<128, true>(148, 64, 4, one);
    return 0;
}
