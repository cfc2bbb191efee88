// Dev tool: per-event SM-clock trace of attention_qtmem_kernel for a few CTAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DATTN_TRACE -o /tmp/attn_trace tools/attn_trace.cu -lcuda
//   /tmp/attn_trace [B=100] [T=258] [H=24]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "experiments/attention_qtmem.cuh"
using namespace esmdiff;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn enc;
static CUtensorMap tmap(void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    cuuint64_t gdim[2] = {cols, rows}, gstr[1] = {cols * 2};
    cuuint32_t box[2] = {64, box_rows}, es[2] = {1, 1};
    CUtensorMap tm;
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    return tm;
}
__global__ void fill(__nv_bfloat16* p, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = __float2bfloat16(((i * 2654435761u) % 2001) / 1000.0f - 1.0f);
}
int main(int argc, char** argv) {
    int B = argc > 1 ? atoi(argv[1]) : 100, T = argc > 2 ? atoi(argv[2]) : 258, H = argc > 3 ? atoi(argv[3]) : 24;
    void* fn; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    enc = (EncodeTiledFn)fn;
    const int D = H * 64; const long long M = (long long)B * T;
    __nv_bfloat16 *qkv, *out; long long* trace;
    cudaMalloc(&qkv, M * 3 * D * 2); cudaMalloc(&out, M * D * 2);
    const int grid = B * H;
    cudaMalloc(&trace, (size_t)grid * 64 * 8); cudaMemset(trace, 0, (size_t)grid * 64 * 8);
    fill<<<(unsigned)((M * 3 * D + 255) / 256), 256>>>(qkv, M * 3 * D);
    attn4::Params p;
    p.B = B; p.T = T; p.H = H; p.nq = (T + 127) / 128; p.nkv = (T + 63) / 64;
    p.tail_cols = ((T - (p.nkv - 1) * 64) + 15) / 16 * 16;
    p.qkv = qkv; p.ctx = out; p.scale_log2 = 0.125f * 1.4426950408889634f; p.trace = trace;
    CUtensorMap tkv = tmap(qkv, M, 3 * D, 64), tkt = tmap(qkv, M, 3 * D, p.tail_cols);
    const int smem = attn4::smem_bytes(p.nkv, p.tail_cols);
    cudaFuncSetAttribute(attn4::attention_qtmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int it = 0; it < 3; ++it) {
        cudaEventRecord(e0);
        attn4::attention_qtmem_kernel<<<grid, attn4::THREADS, smem>>>(tkv, tkt, p);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("run %d: %.1f us (%s)\n", it, ms * 1e3, cudaGetErrorString(cudaGetLastError()));
    }
    std::vector<long long> h((size_t)grid * 64);
    cudaMemcpy(h.data(), trace, h.size() * 8, cudaMemcpyDeviceToHost);
    const int steps = p.nkv * p.nq;
    int picks[] = {0, 1, grid / 2, grid / 2 + 1, grid - 1};
    for (int c : picks) {
        const long long* t = &h[(size_t)c * 64];
        printf("CTA %d: setup %lld |", c, t[1] - t[0]);
        for (int i = 0; i < steps && i < 14; ++i)
            printf(" [%d] Sissue %lld Sarr %lld Pdone %lld PVissue %lld |", i, t[2 + 2 * i] - t[0], t[32 + 2 * i] - t[0],
                   t[33 + 2 * i] - t[0], t[3 + 2 * i] - t[0]);
        printf("\n  Ofull0 %lld out0 %lld Ofull1 %lld out1 %lld\n", t[60] - t[0], t[61] - t[0], t[62] - t[0], t[63] - t[0]);
    }
    // CTA durations
    double sum = 0; long long mx = 0;
    for (int c = 0; c < grid; ++c) { long long d = h[(size_t)c * 64 + 61 + 2 * ((p.nq - 1) & 1)] - h[(size_t)c * 64]; sum += d; if (d > mx) mx = d; }
    printf("mean CTA (chain 0 end) %.0f clk, max %lld clk\n", sum / grid, mx);
    return 0;
}
