/* Coordinate / force buffer kernels around the nbnxm_b200 force kernel.
 *
 * x_to_nbat_x : replaces nbnxm_gpu_x_to_nbat_x_kernel
 *               (src/gromacs/nbnxm/cuda/nbnxm_gpu_buffer_ops_internal.cu:73-120)
 * f4_to_f3    : packs the 16-byte-aligned internal force accumulator into the float3 layout of
 *               NBAtomDataGpu::f (src/gromacs/nbnxm/gpu_types_common.h:178)
 * reduce_f     : replaces reduceKernel (src/gromacs/mdlib/gpuforcereduction_impl_internal.cu:61-118): the nbat-order
 *               forces gathered into the caller's atom-order rvec array, read straight from the float4 accumulator
 *               (no f4 -> f3 pass in between)
 * halo pack / unpack : the x-slab analogue of packSendBufKernel / unpackRecvBufKernel
 *               (src/gromacs/domdec/gpuhaloexchange_impl_gpu.cu:82-137)
 * All are HBM-bound streaming kernels: one thread per atom, 16-byte accesses on the nbat side.
 */
#include "nbnxm_device.cuh"

namespace nbb
{

__global__ void __launch_bounds__(256)
        x_to_nbat_x_kernel(float4* __restrict__ xq, const float* __restrict__ x, const int* __restrict__ atomIndex, int first, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
    {
        const int a = atomIndex[first + i];
        if (a >= 0)
        {
            /* q (w) is left untouched, like the reference (it is set at search steps) */
            float4 v = xq[first + i];
            v.x      = x[3 * a];
            v.y      = x[3 * a + 1];
            v.z      = x[3 * a + 2];
            xq[first + i] = v;
        }
    }
}

template<bool ACCUMULATE>
__global__ void __launch_bounds__(256) f4_to_f3_kernel(const float4* __restrict__ f4, float* __restrict__ f3, int first, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
    {
        const float4 v = f4[first + i];
        float*       o = f3 + 3 * size_t(first + i);
        if (ACCUMULATE)
        {
            /* forces other device code (GPU listed forces) left in f3 since the clear stay */
            o[0] += v.x;
            o[1] += v.y;
            o[2] += v.z;
        }
        else
        {
            o[0] = v.x;
            o[1] = v.y;
            o[2] = v.z;
        }
    }
}

__global__ void __launch_bounds__(256)
        pack_xq_kernel(const float4* __restrict__ xq, const int* __restrict__ index, int n, float sx, float sy, float sz, float4* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
    {
        float4 v = xq[index[i]];
        v.x += sx;
        v.y += sy;
        v.z += sz;
        out[i] = v;
    }
}

__global__ void __launch_bounds__(256) copy4_kernel(const float4* __restrict__ in, float4* __restrict__ out, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

__global__ void __launch_bounds__(256)
        unpack_add_f_kernel(float4* __restrict__ f4, const int* __restrict__ index, int n, const float4* __restrict__ in)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
    {
        const float4 v = in[i];
        float4*      d = f4 + index[i];
        /* each destination atom appears once in the index list, plain read-modify-write */
        float4 o = *d;
        o.x += v.x;
        o.y += v.y;
        o.z += v.z;
        *d = o;
    }
}

template<bool ADD_RVEC, bool ACCUMULATE>
__global__ void __launch_bounds__(256) reduce_f_kernel(const float4* __restrict__ f4, const float* __restrict__ rvecToAdd,
                                                       float* __restrict__ fTotal, const int* __restrict__ cell, int atomStart, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
    {
        const int    a = atomStart + i;
        const float4 v = f4[cell[a]];
        float        x = v.x, y = v.y, z = v.z;
        if (ACCUMULATE)
        {
            x += fTotal[3 * a];
            y += fTotal[3 * a + 1];
            z += fTotal[3 * a + 2];
        }
        if (ADD_RVEC)
        {
            x += rvecToAdd[3 * a];
            y += rvecToAdd[3 * a + 1];
            z += rvecToAdd[3 * a + 2];
        }
        fTotal[3 * a]     = x;
        fTotal[3 * a + 1] = y;
        fTotal[3 * a + 2] = z;
    }
}

static inline int nblk(int n) { return (n + 255) / 256; }

void launch_reduce_f(const float4* f4, const float* rvecToAdd, float* fTotal, const int* cell, int atomStart, int n, bool accumulate,
                     cudaStream_t s)
{
    if (n <= 0) return;
    if (rvecToAdd != nullptr)
    {
        if (accumulate) reduce_f_kernel<true, true><<<nblk(n), 256, 0, s>>>(f4, rvecToAdd, fTotal, cell, atomStart, n);
        else reduce_f_kernel<true, false><<<nblk(n), 256, 0, s>>>(f4, rvecToAdd, fTotal, cell, atomStart, n);
    }
    else
    {
        if (accumulate) reduce_f_kernel<false, true><<<nblk(n), 256, 0, s>>>(f4, rvecToAdd, fTotal, cell, atomStart, n);
        else reduce_f_kernel<false, false><<<nblk(n), 256, 0, s>>>(f4, rvecToAdd, fTotal, cell, atomStart, n);
    }
}

void launch_x_to_nbat_x(float4* xq, const float* x, const int* atomIndex, int first, int n, cudaStream_t s)
{
    if (n > 0) x_to_nbat_x_kernel<<<nblk(n), 256, 0, s>>>(xq, x, atomIndex, first, n);
}
void launch_f4_to_f3(const float4* f4, float* f3, int first, int n, cudaStream_t s, bool accumulate)
{
    if (n <= 0) return;
    if (accumulate)
        f4_to_f3_kernel<true><<<nblk(n), 256, 0, s>>>(f4, f3, first, n);
    else
        f4_to_f3_kernel<false><<<nblk(n), 256, 0, s>>>(f4, f3, first, n);
}
void launch_pack_xq(const float4* xq, const int* index, int n, const float* shift, float4* out, cudaStream_t s)
{
    if (n > 0) pack_xq_kernel<<<nblk(n), 256, 0, s>>>(xq, index, n, shift[0], shift[1], shift[2], out);
}
void launch_copy4(const float4* in, float4* out, int n, cudaStream_t s)
{
    if (n > 0) copy4_kernel<<<nblk(n), 256, 0, s>>>(in, out, n);
}
void launch_unpack_add_f(float4* f4, const int* index, int n, const float4* in, cudaStream_t s)
{
    if (n > 0) unpack_add_f_kernel<<<nblk(n), 256, 0, s>>>(f4, index, n, in);
}

} // namespace nbb
