// Integer-pipe microbenchmarks: the roofline denominator of the fill kernels.
// MEASURED_PEAKS.json has no integer entry, so bench.py measures it on the box it runs on:
// independent dependency chains of IADD3 / VIMNMX / VIADDMNMX per thread, all SMs busy, CUDA events.
#include "../../include/bialign_b200.h"
#include "common.cuh"

namespace ba {

template <int KIND>
__global__ void __launch_bounds__(256) int_pipe_kernel(int* out, int iters, int seed) {
    constexpr int CH = 8;  // independent chains per thread (latency 4, one issue per 2 clk per SMSP)
    int v[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) v[c] = seed + threadIdx.x * 3 + c * 17 + blockIdx.x;
    const int k1 = seed | 1, k2 = seed * 3 + 5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                if (KIND == 0) v[c] = v[c] + v[(c + 1) % CH];          // dependent add ring (IADD3 / IMAD), not foldable
                else if (KIND == 1) v[c] = max(v[c] ^ k1, k2);         // xor + max (LOP3 + VIMNMX), 2 ALU instr
                else v[c] = __viaddmax_s32(v[c], k1, k2 - it);         // fused add+max (VIADDMNMX)
            }
        }
    }
    int acc = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) acc ^= v[c];
    if (acc == 0x7fffffff) out[0] = acc;  // keep the chains alive
}

}  // namespace ba

extern "C" BA_API int ba_microbench_int(int device, int kind, double* instr_per_s, int* sm_count) {
    using namespace ba;
    if (!instr_per_s) return BA_ERR_INVALID_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return BA_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return BA_ERR_CUDA;
    int* d = nullptr;
    if (cudaMalloc(&d, 4) != cudaSuccess) return BA_ERR_OOM;
    const int grid = prop.multiProcessorCount * 8, iters = 4096;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a);
        if (kind == 0) int_pipe_kernel<0><<<grid, 256>>>(d, iters, rep + 1);
        else if (kind == 1) int_pipe_kernel<1><<<grid, 256>>>(d, iters, rep + 1);
        else int_pipe_kernel<2><<<grid, 256>>>(d, iters, rep + 1);
        cudaEventRecord(b);
        if (cudaEventSynchronize(b) != cudaSuccess) { cudaFree(d); return BA_ERR_CUDA; }
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d);
    const double instr = (double)grid * 256 * (double)iters * 16 * 8 * (kind == 1 ? 2 : 1);
    *instr_per_s = instr / (best * 1e-3);
    if (sm_count) *sm_count = prop.multiProcessorCount;
    return BA_OK;
}
