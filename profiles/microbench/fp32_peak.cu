// Microbenchmark: FP32 FMA-pipe peak on B200 (sm_100a) with scalar FFMA and packed FFMA2
// (fma.rn.f32x2), and how many issue slots the packed form leaves for ALU / MUFU work.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_peak fp32_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c)
{
    u64 d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

template<int MODE>
__global__ void __launch_bounds__(256) kern(float* out, int iters, float seed)
{
    float a[16];
    u64   p[8];
    unsigned int n[8];
    float m[4];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = seed + i + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; i++)
    {
        float2 t = make_float2(seed + i, seed - i + threadIdx.x);
        p[i]     = *reinterpret_cast<u64*>(&t);
        n[i]     = threadIdx.x * 7 + i;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) m[i] = seed + 1.5f + i;
    const float b = seed * 0.999f, c = seed * 0.001f;
    float2      bb = make_float2(b, b), cc = make_float2(c, c);
    const u64   b2 = *reinterpret_cast<u64*>(&bb), c2 = *reinterpret_cast<u64*>(&cc);
    for (int it = 0; it < iters; it++)
    {
        if (MODE == 0 || MODE == 2 || MODE == 4 || MODE == 6)
        {
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], b, c);
        }
        if (MODE == 1 || MODE == 3 || MODE == 5 || MODE == 7)
        {
#pragma unroll
            for (int i = 0; i < 8; i++) p[i] = fma2(p[i], b2, c2);
        }
        if (MODE == 2 || MODE == 3)
        {
#pragma unroll
            for (int i = 0; i < 4; i++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(n[i]) : "r"(n[i + 4]), "r"(it));
        }
        if (MODE == 6 || MODE == 7)
        {
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(n[i]) : "r"(n[(i + 4) & 7]), "r"(it));
        }
        if (MODE == 4 || MODE == 5)
        {
#pragma unroll
            for (int i = 0; i < 2; i++) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(m[i]));
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
#pragma unroll
    for (int i = 0; i < 8; i++)
    {
        float2 t = *reinterpret_cast<float2*>(&p[i]);
        s += t.x + t.y + n[i];
    }
#pragma unroll
    for (int i = 0; i < 4; i++) s += m[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template<int MODE>
void run(const char* name, float* d, int blocks, int iters, double fmas_per_iter)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    kern<MODE><<<blocks, 256>>>(d, iters, 1.0f);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++)
    {
        cudaEventRecord(e0);
        kern<MODE><<<blocks, 256>>>(d, iters, 1.0f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double fl = 2.0 * fmas_per_iter * iters * 256.0 * blocks;
    printf("%-28s %8.3f ms  %8.2f TFLOP/s (FMA flops only)\n", name, best, fl / best * 1e-9);
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%s SMs %d clock %d kHz -> nominal FP32 peak %.2f TFLOP/s\n", prop.name, prop.multiProcessorCount, clk,
           prop.multiProcessorCount * 128.0 * 2.0 * clk * 1e-9);
    const int blocks = prop.multiProcessorCount * 8, iters = 20000;
    float*    d;
    cudaMalloc(&d, blocks * 256 * sizeof(float));
    run<0>("FFMA x16", d, blocks, iters, 16);
    run<1>("FFMA2 x8", d, blocks, iters, 16);
    run<2>("FFMA x16 + 4 LOP3", d, blocks, iters, 16);
    run<3>("FFMA2 x8 + 4 LOP3", d, blocks, iters, 16);
    run<6>("FFMA x16 + 8 LOP3", d, blocks, iters, 16);
    run<7>("FFMA2 x8 + 8 LOP3", d, blocks, iters, 16);
    run<4>("FFMA x16 + 2 MUFU.RSQ", d, blocks, iters, 16);
    run<5>("FFMA2 x8 + 2 MUFU.RSQ", d, blocks, iters, 16);
    cudaFree(d);
    return 0;
}
