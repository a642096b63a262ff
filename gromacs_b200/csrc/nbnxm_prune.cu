/* Pair-list pruning and sci sorting kernels of the nbnxm_b200 path.
 *
 * prune kernel  : replaces nbnxm_kernel_prune_cuda<haveFreshList>
 *                 (src/gromacs/nbnxm/cuda/nbnxm_cuda_kernel_pruneonly.cuh:109-346)
 * histogram scan: replaces cub::DeviceScan::ExclusiveSum (cuda/nbnxm_cuda_data_mgmt.cu:135-165)
 * bucket sort   : replaces nbnxmKernelBucketSciSort (cuda/nbnxm_cuda_kernel_sci_sort.cuh:67-90)
 *
 * Mask semantics are the reference's, bit for bit: for every (cjPacked, half w, j-cluster jm,
 * i-cluster ci) a bit survives the outer prune iff any of its 8 x 4 atom pairs has
 * r2 < rlistOuter^2 and is set in the inner mask iff any pair has r2 < rlistInner^2, with
 * xi = float(x_i + shift) and r2 = fma(dz,dz, fma(dy,dy, dx*dx)).
 *
 * Work decomposition: one warp (= one 32-thread CTA) per sci entry; lane = il + 8*jl holds the 8 shifted
 * i-atoms (il of every i-cluster) in registers and tests j-atoms jl (half 0) and jl+4 (half 1), so both
 * halves are pruned by the same warp with two warp votes per cluster pair.
 */
#include <cstdlib>

#include "nbnxm_device.cuh"

namespace nbb
{

/* One warp = one 32-thread CTA = one sci entry, like the force kernel: list masks and trip counts are
 * CTA-uniform.  Only the i-cluster loop is unrolled (the shifted i-atoms live in registers); the j-cluster loop
 * is not, which keeps the kernel inside the instruction cache. */
template<bool FRESH>
__global__ void __launch_bounds__(32) nbnxm_prune_kernel(const AtomDataDev ad, const ParamsDev p, const PairlistDev pl, const int numParts)
{
    constexpr unsigned c_full = 0xffffffffu;
    const int          lane   = threadIdx.x;
    const int          il     = lane & 7;
    const int          jl     = lane >> 3;
    const int          unit   = blockIdx.x;

    /* rolling part index lives on the device, one copy per work unit, so consecutive launches need
     * no host bookkeeping (same idea as pruneonly.cuh:123-131) */
    const int part = pl.rollingPart[unit];
    __syncwarp();
    if (lane == 0)
    {
        pl.rollingPart[unit] = (part + 1) % numParts;
    }
    const int numSciInPart = (pl.numSci - part + numParts - 1) / numParts;
    if (unit >= numSciInPart)
    {
        return;
    }
    const int              sciIdx = unit * numParts + part;
    const nbnxm_b200_sci_t s      = FRESH ? pl.sci[sciIdx] : pl.sciSorted[sciIdx];

    const float shx = ad.shiftVec[3 * s.shift], shy = ad.shiftVec[3 * s.shift + 1], shz = ad.shiftVec[3 * s.shift + 2];
    float       xi[c_superClusterSize], yi[c_superClusterSize], zi[c_superClusterSize];
#pragma unroll
    for (int ci = 0; ci < c_superClusterSize; ci++)
    {
        const float4 v = ad.xq[(s.sci * c_superClusterSize + ci) * c_clusterSize + il];
        xi[ci]         = __fadd_rn(v.x, shx);
        yi[ci]         = __fadd_rn(v.y, shy);
        zi[ci]         = __fadd_rn(v.z, shz);
    }
    const float rlistOuter2 = p.rlist_outer_sq;
    const float rlistInner2 = p.rlist_inner_sq;
    int         count       = 0;

    /* the 32 j-atoms of a cjPacked group are fetched by one coalesced 16-byte load per lane (lane L: atom
     * L & 7 of j-cluster L >> 3), one group ahead of use, and parked in shared memory */
    __shared__ float4 sm_xqj[32];
    const uint4*      cjGroups = reinterpret_cast<const uint4*>(pl.cjPacked);
    const uint2*      outerMasks = reinterpret_cast<const uint2*>(pl.imaskOuter);

    auto checkMasks = [&](const int jpx, uint4& cjv, unsigned& full0, unsigned& full1, unsigned& new0, unsigned& new1) {
        cjv             = cjGroups[2 * jpx];
        const uint4 mev = cjGroups[2 * jpx + 1];
        if (FRESH)
        {
            full0 = mev.x;
            full1 = mev.z;
            new0  = 0u;
            new1  = 0u;
        }
        else
        {
            const uint2 o = outerMasks[jpx];
            full0         = o.x;
            full1         = o.y;
            new0          = mev.x;
            new1          = mev.z;
        }
    };
    auto fetchAtom = [&](const uint4 cjv, const unsigned checkAny, float4& xj) {
        if (checkAny & (0xffu << (8 * jl)))
        {
            const int cj = static_cast<int>(jl == 0 ? cjv.x : (jl == 1 ? cjv.y : (jl == 2 ? cjv.z : cjv.w)));
            xj           = ad.xqJ[cj * c_clusterSize + il];
        }
    };

    /* Two-deep software pipeline over the groups of the entry: the masks and j-cluster indices of group jp + 2 are requested
     * while the j-atoms of group jp + 1 - whose address comes out of ITS masks, requested one iteration earlier - are, and group
     * jp is tested.  With the masks only one group ahead every iteration waited for them before it could ask for the atoms
     * (ncu, rolling pass at 12.3 M atoms: long_scoreboard 1.7 per issue). */
    uint4    cjNext = make_uint4(0u, 0u, 0u, 0u), cjAfter = make_uint4(0u, 0u, 0u, 0u);
    unsigned full0N = 0u, full1N = 0u, new0N = 0u, new1N = 0u;
    unsigned full0A = 0u, full1A = 0u, new0A = 0u, new1A = 0u;
    float4   xjNext = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    int      jp     = s.cj_packed_begin;
    if (jp < s.cj_packed_end)
    {
        checkMasks(jp, cjNext, full0N, full1N, new0N, new1N);
        if (jp + 1 < s.cj_packed_end) checkMasks(jp + 1, cjAfter, full0A, full1A, new0A, new1A);
        fetchAtom(cjNext, FRESH ? (full0N | full1N) : ((new0N ^ full0N) | (new1N ^ full1N)), xjNext);
    }
    for (; jp < s.cj_packed_end; jp++)
    {
        unsigned       full0 = full0N, full1 = full1N, new0 = new0N, new1 = new1N;
        const unsigned check0 = FRESH ? full0 : (new0 ^ full0);
        const unsigned check1 = FRESH ? full1 : (new1 ^ full1);
        __syncwarp();
        sm_xqj[lane] = xjNext;
        __syncwarp();
        /* rotate: group jp + 1 becomes "next" (its masks were requested an iteration ago), group jp + 2 is requested now */
        cjNext = cjAfter;
        full0N = full0A;
        full1N = full1A;
        new0N  = new0A;
        new1N  = new1A;
        if (jp + 2 < s.cj_packed_end) checkMasks(jp + 2, cjAfter, full0A, full1A, new0A, new1A);
        if (jp + 1 < s.cj_packed_end)
        {
            fetchAtom(cjNext, FRESH ? (full0N | full1N) : ((new0N ^ full0N) | (new1N ^ full1N)), xjNext);
        }
        if ((check0 | check1) == 0u)
        {
            continue;
        }
        /* one iteration per j-cluster with anything left to check; both halves in the same pass */
        unsigned      c0 = check0, c1 = check1;
        unsigned      clear0 = 0u, clear1 = 0u, set0 = 0u, set1 = 0u; /* outer bits to clear, inner bits to set */
        const float4* xjPtr = sm_xqj + jl;
#pragma unroll 1
        for (int shift = 0; (c0 | c1) != 0u; shift += 8, c0 >>= 8, c1 >>= 8, xjPtr += c_clusterSize)
        {
            const unsigned m0 = c0 & 0xffu, m1 = c1 & 0xffu;
            if ((m0 | m1) == 0u)
            {
                continue;
            }
            const float4 xj0 = xjPtr[0];
            const float4 xj1 = xjPtr[4];
            unsigned     out0 = 0u, out1 = 0u, in0 = 0u, in1 = 0u; /* cluster pairs with a pair inside rlistOuter / rlistInner */
#pragma unroll
            for (int ci = 0; ci < c_superClusterSize; ci++)
            {
                if ((m0 | m1) & (1u << ci))
                {
                    const float r20 = norm2_fma(__fsub_rn(xi[ci], xj0.x), __fsub_rn(yi[ci], xj0.y), __fsub_rn(zi[ci], xj0.z));
                    const float r21 = norm2_fma(__fsub_rn(xi[ci], xj1.x), __fsub_rn(yi[ci], xj1.y), __fsub_rn(zi[ci], xj1.z));
                    if (FRESH)
                    {
                        if (__any_sync(c_full, r20 < rlistOuter2)) out0 |= 1u << ci;
                        if (__any_sync(c_full, r21 < rlistOuter2)) out1 |= 1u << ci;
                    }
                    if (__any_sync(c_full, r20 < rlistInner2)) in0 |= 1u << ci;
                    if (__any_sync(c_full, r21 < rlistInner2)) in1 |= 1u << ci;
                }
            }
            if (FRESH)
            {
                clear0 |= (m0 & ~out0) << shift;
                clear1 |= (m1 & ~out1) << shift;
            }
            set0 |= (m0 & in0) << shift;
            set1 |= (m1 & in1) << shift;
        }
        full0 &= ~clear0;
        full1 &= ~clear1;
        new0 |= set0;
        new1 |= set1;
        if (lane == 0)
        {
            /* like the reference, a half whose check mask is empty is left untouched */
            if (FRESH)
            {
                if (check0) pl.imaskOuter[2 * jp] = full0;
                if (check1) pl.imaskOuter[2 * jp + 1] = full1;
            }
            if (check0) pl.cjPacked[jp].imei[0].imask = new0;
            if (check1) pl.cjPacked[jp].imei[1].imask = new1;
        }
        if (FRESH)
        {
            count += (check0 ? __popc(new0) : 0) + (check1 ? __popc(new1) : 0);
        }
    }
    if (FRESH && lane == 0)
    {
        const int index = max(c_sciHistogramSize - count - 1, 0);
        atomicAdd(pl.sciHistogram + index, 1);
        pl.sciCount[sciIdx] = index;
    }
}

/* exclusive prefix sum of the 8192-bin histogram, one CTA */
__global__ void __launch_bounds__(1024) nbnxm_sci_histogram_scan_kernel(const int* __restrict__ histogram, int* __restrict__ offset)
{
    constexpr int perThread = c_sciHistogramSize / 1024;
    __shared__ int warpSums[32];
    const int      t = threadIdx.x;
    int            v[perThread];
    int            sum = 0;
#pragma unroll
    for (int i = 0; i < perThread; i++)
    {
        v[i] = histogram[t * perThread + i];
        sum += v[i];
    }
    int incl = sum;
#pragma unroll
    for (int m = 1; m < 32; m <<= 1)
    {
        const int n = __shfl_up_sync(0xffffffffu, incl, m);
        if ((t & 31) >= m) incl += n;
    }
    if ((t & 31) == 31) warpSums[t >> 5] = incl;
    __syncthreads();
    if (t < 32)
    {
        int w = warpSums[t];
#pragma unroll
        for (int m = 1; m < 32; m <<= 1)
        {
            const int n = __shfl_up_sync(0xffffffffu, w, m);
            if (t >= m) w += n;
        }
        warpSums[t] = w;
    }
    __syncthreads();
    int base = incl - sum + ((t >> 5) > 0 ? warpSums[(t >> 5) - 1] : 0);
#pragma unroll
    for (int i = 0; i < perThread; i++)
    {
        offset[t * perThread + i] = base;
        base += v[i];
    }
}

__global__ void __launch_bounds__(256) nbnxm_sci_bucket_sort_kernel(const PairlistDev pl)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < pl.numSci)
    {
        const nbnxm_b200_sci_t s   = pl.sci[i];
        const int              pos = atomicAdd(pl.sciOffset + pl.sciCount[i], 1);
        pl.sciSorted[pos]          = s;
    }
}

/* Locality-preserving form of the sort for lists far larger than the GPU holds at once: the entries are sorted by
 * decreasing pair count within tiles of c_sciSortTile consecutive entries of the (spatially ordered) sci array, so CTAs
 * that run at the same time work on neighbouring super-clusters and share their j-atoms in L2, while every tile - the
 * last one in particular, which forms the tail of the launch - still runs its longest entries first.  Any permutation
 * of sci is a valid result (nbnxm_cuda_kernel_sci_sort.cuh:41-58); the reference notes that the global order stops
 * paying above ~400 k atoms (gpu_types_common.h:77-82).  One CTA per tile: (bucket << 13 | index) keys, bitonic network
 * in shared memory; ties keep the list order, so the result is deterministic. */
constexpr int c_sciSortTile = 4096;

template<int c_sciSortTile>
__global__ void __launch_bounds__(1024) nbnxm_sci_tile_sort_kernel(const PairlistDev pl)
{
    __shared__ unsigned int key[c_sciSortTile];
    const int               first = blockIdx.x * c_sciSortTile;
    const int               n     = min(c_sciSortTile, pl.numSci - first);
    for (int i = threadIdx.x; i < c_sciSortTile; i += 1024)
    {
        key[i] = i < n ? (static_cast<unsigned int>(pl.sciCount[first + i]) << 13) | static_cast<unsigned int>(i) : 0xffffffffu;
    }
    __syncthreads();
    for (int k = 2; k <= c_sciSortTile; k <<= 1)
    {
        for (int j = k >> 1; j > 0; j >>= 1)
        {
            for (int t = threadIdx.x; t < c_sciSortTile / 2; t += 1024)
            {
                const int          lo  = 2 * t - (t & (j - 1)); /* index with bit j clear */
                const int          hi  = lo + j;
                const bool         up  = (lo & k) == 0;
                const unsigned int a   = key[lo];
                const unsigned int b   = key[hi];
                if ((a > b) == up)
                {
                    key[lo] = b;
                    key[hi] = a;
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < n; i += 1024)
    {
        pl.sciSorted[first + i] = pl.sci[first + (key[i] & (c_sciSortTile - 1))];
    }
}

/* diagnostics: 32 atom pairs per set imask bit over the cjPacked ranges of all sci entries */
__global__ void __launch_bounds__(256) nbnxm_count_pairs_kernel(const PairlistDev pl)
{
    const int          i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long n = 0;
    if (i < pl.numSci)
    {
        const nbnxm_b200_sci_t s = pl.sci[i];
        for (int jp = s.cj_packed_begin; jp < s.cj_packed_end; jp++)
        {
            n += 32ull * (__popc(pl.cjPacked[jp].imei[0].imask) + __popc(pl.cjPacked[jp].imei[1].imask));
        }
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1)
    {
        n += __shfl_xor_sync(0xffffffffu, n, m);
    }
    if ((threadIdx.x & 31) == 0 && n != 0)
    {
        atomicAdd(pl.pairCount, n);
    }
}

void launch_count_pairs(const PairlistDev& pl, cudaStream_t stream)
{
    if (pl.numSci > 0 && pl.pairCount != nullptr)
    {
        nbnxm_count_pairs_kernel<<<(pl.numSci + 255) / 256, 256, 0, stream>>>(pl);
    }
}

void launch_prune(bool fresh, const AtomDataDev& ad, const ParamsDev& p, const PairlistDev& pl, int numParts, cudaStream_t stream)
{
    const int units = (pl.numSci + numParts - 1) / numParts;
    if (fresh)
    {
        nbnxm_prune_kernel<true><<<units, 32, 0, stream>>>(ad, p, pl, numParts);
    }
    else
    {
        nbnxm_prune_kernel<false><<<units, 32, 0, stream>>>(ad, p, pl, numParts);
    }
}

bool sci_sort_is_tiled(int numSci)
{
    /* NBNXM_B200_SCI_SORT=global / tiled for A/B runs; default: tiles once the list is many times what the GPU holds at once
     * (148 SMs x 20 one-warp CTAs) and the atom data no longer fits L2 */
    static const char* mode = getenv("NBNXM_B200_SCI_SORT");
    if (mode != nullptr && mode[0] == 'g') return false;
    if (mode != nullptr && mode[0] == 't') return true;
    return numSci > 65536;
}

int launch_sci_sort(const PairlistDev& pl, cudaStream_t stream)
{
    if (sci_sort_is_tiled(pl.numSci))
    {
        /* NBNXM_B200_SCI_SORT_TILE=8192 for A/B runs of the tile size */
        static const char* tile = getenv("NBNXM_B200_SCI_SORT_TILE");
        if (tile != nullptr && atoi(tile) == 8192)
        {
            nbnxm_sci_tile_sort_kernel<8192><<<(pl.numSci + 8191) / 8192, 1024, 0, stream>>>(pl);
        }
        else
        {
            nbnxm_sci_tile_sort_kernel<c_sciSortTile><<<(pl.numSci + c_sciSortTile - 1) / c_sciSortTile, 1024, 0, stream>>>(pl);
        }
        return 1;
    }
    nbnxm_sci_histogram_scan_kernel<<<1, 1024, 0, stream>>>(pl.sciHistogram, pl.sciOffset);
    nbnxm_sci_bucket_sort_kernel<<<(pl.numSci + 255) / 256, 256, 0, stream>>>(pl);
    return 2; /* kernels launched */
}

} // namespace nbb
