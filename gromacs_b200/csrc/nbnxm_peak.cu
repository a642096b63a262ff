/* FP32 FMA-pipe peak of the device, measured live: the roofline denominator bench.py reports the
 * force kernel against (MEASURED_PEAKS.json carries no FP32 figure). 16 independent FFMA chains per
 * thread, 8 CTAs of 256 threads per SM; same kernel as profiles/microbench/fp32_peak.cu mode 0. */
#include <cuda_runtime.h>

#include "../../include/nbnxm_b200.h"

namespace
{
__global__ void __launch_bounds__(256) ffma_kernel(float* out, int iters, float seed)
{
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = seed + i + threadIdx.x;
    const float b = seed * 0.999f, c = seed * 0.001f;
    for (int it = 0; it < iters; it++)
    {
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], b, c);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
} // namespace

extern "C" int nbnxm_b200_measure_fp32_peak(int device, double* tflops)
{
    if (!tflops || cudaSetDevice(device) != cudaSuccess) return 1;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return 1;
    const int blocks = prop.multiProcessorCount * 8, iters = 20000;
    float*    d      = nullptr;
    if (cudaMalloc(&d, size_t(blocks) * 256 * sizeof(float)) != cudaSuccess) return 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    ffma_kernel<<<blocks, 256>>>(d, iters, 1.0f);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++)
    {
        cudaEventRecord(e0);
        ffma_kernel<<<blocks, 256>>>(d, iters, 1.0f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    if (cudaGetLastError() != cudaSuccess) return 1;
    *tflops = 2.0 * 16.0 * iters * 256.0 * blocks / (best * 1e-3) * 1e-12;
    return 0;
}
