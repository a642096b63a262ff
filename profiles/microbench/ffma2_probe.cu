// Microbenchmark: what bounds packed FP32 (FFMA2) code on B200 beyond the pure-FMA peak of fp32_peak.cu:
//   A  FFMA2 with three DISTINCT register-pair sources per instruction (register-file read bandwidth)
//   B  dependent-chain latency of FFMA / FFMA2 / MUFU.RSQ / LDS
//   C  FFMA2 with a broadcast constant / immediate as one source
//   D  MUFU throughput
//   E  throughput of a dependent chain as a function of resident warps per scheduler
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c)
{
    u64 d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 pk(float lo, float hi)
{
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float sum2(u64 v)
{
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return a + b;
}

// MODE 0: FFMA  a[i] = fma(b[i], c[(i+1)&15], a[i])   (3 distinct regs)
// MODE 1: FFMA2 p[i] = fma2(q[i], r[(i+1)&7], p[i])   (3 distinct reg pairs)
// MODE 2: FFMA2 p[i] = fma2(p[i], q[i], bc(param))    (2 reg pairs + uniform broadcast)
// MODE 3: FFMA2 p[i] = fma2(p[i], q0, r0)             (shared sources, as fp32_peak.cu)
// MODE 4: FFMA2 p[i] = fma2(q[i], q[i], p[i])         (2 distinct pairs)
// MODE 5: MUFU.RSQ x8 independent
// MODE 6: 6 x FFMA2 (3 distinct) + 2 MUFU
template<int MODE>
__global__ void __launch_bounds__(128) tput(float* out, int iters, float seed, float param)
{
    float a[16], b[16], c[16];
    u64   p[8], q[8], r[8];
    float m[8];
#pragma unroll
    for (int i = 0; i < 16; i++)
    {
        a[i] = seed + i + threadIdx.x;
        b[i] = seed * 0.999f + i * 1e-6f;
        c[i] = seed * 0.001f + i * 1e-7f + threadIdx.x * 1e-9f;
    }
#pragma unroll
    for (int i = 0; i < 8; i++)
    {
        p[i] = pk(seed + i, seed - i + threadIdx.x);
        q[i] = pk(seed * 0.999f + i * 1e-6f, seed * 0.998f + i * 1e-6f);
        r[i] = pk(seed * 0.001f + i * 1e-7f, seed * 0.002f + threadIdx.x * 1e-9f);
        m[i] = seed + 1.5f + i;
    }
    for (int it = 0; it < iters; it++)
    {
        if (MODE == 0)
        {
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = fmaf(b[i], c[(i + 1) & 15], a[i]);
        }
        if (MODE == 1 || MODE == 6)
        {
#pragma unroll
            for (int i = 0; i < (MODE == 6 ? 6 : 8); i++) p[i] = fma2(q[i], r[(i + 1) & 7], p[i]);
        }
        if (MODE == 2)
        {
#pragma unroll
            for (int i = 0; i < 8; i++) p[i] = fma2(p[i], q[i], pk(param, param));
        }
        if (MODE == 3)
        {
#pragma unroll
            for (int i = 0; i < 8; i++) p[i] = fma2(p[i], q[0], r[0]);
        }
        if (MODE == 4)
        {
#pragma unroll
            for (int i = 0; i < 8; i++) p[i] = fma2(q[i], q[i], p[i]);
        }
        if (MODE == 5)
        {
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(m[i]));
        }
        if (MODE == 6)
        {
#pragma unroll
            for (int i = 0; i < 2; i++) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(m[i]));
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i] + b[i] + c[i];
#pragma unroll
    for (int i = 0; i < 8; i++) s += sum2(p[i]) + sum2(q[i]) + sum2(r[i]) + m[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// dependent chains, 16 links per iteration. KIND 0 FFMA, 1 FFMA2, 2 MUFU.RSQ, 3 LDS (pointer chase in shared), 4 FMUL2->FADD2 alternating
template<int KIND>
__global__ void __launch_bounds__(32) chain(float* out, int iters, float seed, long long* cycles)
{
    __shared__ int sm[64];
    sm[threadIdx.x]      = (threadIdx.x + 1) & 31;
    sm[threadIdx.x + 32] = 0;
    __syncwarp();
    float x = seed + threadIdx.x * 1e-3f, y = 0.999f, z = 1e-3f;
    u64   px = pk(x, x + 1.0f), py = pk(y, y), pz = pk(z, z);
    int   idx = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++)
    {
#pragma unroll
        for (int i = 0; i < 16; i++)
        {
            if (KIND == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(y), "f"(z));
            if (KIND == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(px) : "l"(py), "l"(pz));
            if (KIND == 2) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x));
            if (KIND == 3) idx = *(volatile int*)(sm + idx);
            if (KIND == 4)
            {
                if (i & 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(px) : "l"(pz));
                else asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(px) : "l"(py));
            }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
    out[blockIdx.x * 32 + threadIdx.x] = x + sum2(px) + idx;
}

// a pair-body-like dependent chain (LDS.128 -> 3 FADD2 -> 3 r2 -> MUFU -> 12 dependent FFMA2 -> 6 accumulate) per "body",
// one warp per CTA, WPS resident warps per scheduler: cycles per body per scheduler
__global__ void __launch_bounds__(32) bodychain(float* out, int iters, float seed)
{
    __shared__ float4 sm[64];
    sm[threadIdx.x]      = make_float4(seed + threadIdx.x, seed, seed * 2, 1.0f);
    sm[threadIdx.x + 32] = make_float4(seed - threadIdx.x, seed, seed * 3, 1.0f);
    __syncwarp();
    u64 xj = pk(seed + 0.1f, seed + 0.2f), yj = pk(seed, seed + 0.3f), zj = pk(seed, seed + 0.4f);
    u64 f0 = 0, f1 = 0, f2 = 0, g0 = 0, g1 = 0, g2 = 0;
    u64 k1 = pk(0.999f, 0.998f), k2 = pk(1e-3f, 2e-3f);
    for (int it = 0; it < iters; it++)
    {
#pragma unroll 1
        for (int ci = 0; ci < 8; ci++)
        {
            float4 xi;
            {
                const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(sm + ((ci * 8 + (threadIdx.x & 7) + it) & 63)));
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(xi.x), "=f"(xi.y), "=f"(xi.z), "=f"(xi.w) : "r"(a) : "memory");
            }
            u64    dx, dy, dz, r2;
            asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(dx) : "l"(pk(xi.x, xi.x)), "l"(xj));
            asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(dy) : "l"(pk(xi.y, xi.y)), "l"(yj));
            asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(dz) : "l"(pk(xi.z, xi.z)), "l"(zj));
            asm volatile("mul.rn.f32x2 %0, %1, %1;" : "=l"(r2) : "l"(dx));
            r2 = fma2(dy, dy, r2);
            r2 = fma2(dz, dz, r2);
            float a, b;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r2));
            asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a));
            asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(b));
            u64 t = pk(a, b);
#pragma unroll
            for (int k = 0; k < 12; k++) t = fma2(t, k1, k2);
            f0 = fma2(t, dx, f0);
            f1 = fma2(t, dy, f1);
            f2 = fma2(t, dz, f2);
            g0 = fma2(t, dx, g0);
            g1 = fma2(t, dy, g1);
            g2 = fma2(t, dz, g2);
        }
    }
    out[blockIdx.x * 32 + threadIdx.x] = sum2(f0) + sum2(f1) + sum2(f2) + sum2(g0) + sum2(g1) + sum2(g2);
}

static float timeit(void (*launch)(void*), void* arg)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch(arg);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++)
    {
        cudaEventRecord(e0);
        launch(arg);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

struct Ctx
{
    float*     d;
    long long* cyc;
    int        blocks, iters;
};

template<int MODE>
void runT(const char* name, Ctx& c, double instrPerIter, double clkGHz, int sms)
{
    auto  l  = [](void* a) { Ctx* c = (Ctx*)a; tput<MODE><<<c->blocks, 128>>>(c->d, c->iters, 1.0f, 0.5f); };
    float ms = timeit(l, &c);
    // cycles per warp-instruction per scheduler
    double warpsPerSched = c.blocks * 4.0 / (sms * 4.0);
    double cyc           = ms * 1e-3 * clkGHz * 1e9 / (c.iters * instrPerIter * warpsPerSched);
    printf("%-44s %8.3f ms  %6.3f cycles per warp-instruction per scheduler\n", name, ms, cyc);
}

template<int KIND>
void runC(const char* name, Ctx& c)
{
    chain<KIND><<<1, 32>>>(c.d, 2000, 1.0f, c.cyc);
    cudaDeviceSynchronize();
    chain<KIND><<<1, 32>>>(c.d, 2000, 1.0f, c.cyc);
    long long h;
    cudaMemcpy(&h, c.cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-44s %6.2f cycles per link (dependent chain, one warp)\n", name, (double)h / (2000.0 * 16.0));
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double ghz = clk * 1e-6;
    const int    sms = prop.multiProcessorCount;
    printf("%s SMs %d clock %d kHz\n", prop.name, sms, clk);
    Ctx c;
    c.blocks = sms * 8;
    c.iters  = 20000;
    cudaMalloc(&c.d, c.blocks * 128 * sizeof(float));
    cudaMalloc(&c.cyc, sizeof(long long));
    runT<0>("FFMA  x16, 3 distinct registers", c, 16, ghz, sms);
    runT<1>("FFMA2 x8, 3 distinct register pairs", c, 8, ghz, sms);
    runT<4>("FFMA2 x8, 2 distinct register pairs", c, 8, ghz, sms);
    runT<2>("FFMA2 x8, 2 pairs + uniform broadcast", c, 8, ghz, sms);
    runT<3>("FFMA2 x8, shared sources (reuse)", c, 8, ghz, sms);
    runT<5>("MUFU.RSQ x8", c, 8, ghz, sms);
    runT<6>("6 FFMA2 (3 distinct) + 2 MUFU.RSQ", c, 8, ghz, sms);
    runC<0>("FFMA latency", c);
    runC<1>("FFMA2 latency", c);
    runC<4>("FMUL2/FADD2 latency", c);
    runC<2>("MUFU.RSQ latency", c);
    runC<3>("LDS latency", c);
    // body chain: 27 packed FP32 + 2 MUFU + 1 LDS.128 per body (54 FMA-pipe cycles)
    for (int wps = 1; wps <= 8; wps++)
    {
        struct B { float* d; int blocks; } b{c.d, sms * 4 * wps};
        auto  l  = [](void* a) { B* b = (B*)a; bodychain<<<b->blocks, 32>>>(b->d, 2000, 1.0f); };
        float ms = timeit(l, &b);
        printf("body chain, %d warps/scheduler: %7.1f cycles per body per scheduler (FMA-pipe floor 54)\n", wps,
               ms * 1e-3 * ghz * 1e9 / (2000.0 * 8.0 * wps));
    }
    return 0;
}
