/* GPU pair-list construction for the GPU layout (include/nbnxm_b200_search.h, nbnxm_b200_gpu_search_*).
 *
 * Builds, on the device and from the coordinates already resident in the handle (`xq`, nbat order), what the
 * reference builds on the CPU and uploads every nstlist steps: nonbonded_verlet_t::constructPairlist
 * (src/gromacs/nbnxm/pairlist.cpp:4056) -> gpu_init_pairlist (nbnxm_gpu_data_mgmt.cpp:739).  The list never
 * visits the host (at 12.3 M atoms it is 0.65 GB of cjPacked plus 0.14 GB of exclusion masks).
 *
 * The search is a sequence of one-thread-per-item passes separated by prefix sums; the per-item bodies live in
 * gpusearch_bodies.h, the sequence in gpusearch_driver.h (both shared with the CPU emulation of the tests).  This
 * file provides the CUDA backend: the generic item kernel, a two-level exclusive scan, buffers and the C ABI.
 * All passes are integer / bounding-box work bound by memory latency and divergence, not by FP32 throughput.
 */
#include <cstdlib>
#include <vector>

#include "../../include/nbnxm_b200_search.h"
#include "gpusearch_driver.h"
#include "nbnxm_handle.cuh"

namespace nbb
{

template<typename F>
__global__ void __launch_bounds__(256) search_items_kernel(const F f, int n)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n)
    {
        f(i);
    }
}

/* warp kernel: one warp per item, lanes cooperate through votes inside F::device */
template<typename F>
__global__ void __launch_bounds__(256) search_warps_kernel(const F f, int n)
{
    const int warp = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (warp < n)
    {
        f.device(warp, threadIdx.x & 31);
    }
}

/* block kernel: stages of independent items separated by block barriers, scratch in dynamic shared memory */
template<typename F>
__global__ void __launch_bounds__(1024) search_block_kernel(const F f)
{
    extern __shared__ __align__(16) unsigned char scratch[];
    const int b  = blockIdx.x;
    const int ns = f.numStages(b);
    for (int s = 0; s < ns; s++)
    {
        const int ni = f.numItems(b, s);
        for (int t = threadIdx.x; t < ni; t += blockDim.x)
        {
            f(b, s, t, scratch);
        }
        __syncthreads();
    }
}

/* ---- exclusive prefix sum: tiles of 8192 ints per CTA (the block scan of nbnxm_sci_histogram_scan_kernel), tile
 * totals scanned by the same kernel, then added back ---- */
constexpr int c_scanTile = 8192;

__global__ void __launch_bounds__(1024) search_scan_tile_kernel(const int* __restrict__ in, int* __restrict__ out, int n, int* __restrict__ tileSums)
{
    constexpr int  perThread = c_scanTile / 1024;
    __shared__ int warpSums[32];
    const int      t    = threadIdx.x;
    const int      base = blockIdx.x * c_scanTile + t * perThread;
    int            v[perThread];
    int            sum = 0;
#pragma unroll
    for (int i = 0; i < perThread; i++)
    {
        v[i] = (base + i < n) ? in[base + i] : 0;
        sum += v[i];
    }
    int incl = sum;
#pragma unroll
    for (int m = 1; m < 32; m <<= 1)
    {
        const int up = __shfl_up_sync(0xffffffffu, incl, m);
        if ((t & 31) >= m) incl += up;
    }
    if ((t & 31) == 31) warpSums[t >> 5] = incl;
    __syncthreads();
    if (t < 32)
    {
        int w = warpSums[t];
#pragma unroll
        for (int m = 1; m < 32; m <<= 1)
        {
            const int up = __shfl_up_sync(0xffffffffu, w, m);
            if (t >= m) w += up;
        }
        warpSums[t] = w;
    }
    __syncthreads();
    int run = incl - sum + ((t >> 5) > 0 ? warpSums[(t >> 5) - 1] : 0);
#pragma unroll
    for (int i = 0; i < perThread; i++)
    {
        if (base + i < n) out[base + i] = run;
        run += v[i];
    }
    if (t == 0 && tileSums != nullptr)
    {
        tileSums[blockIdx.x] = warpSums[31];
    }
}

__global__ void __launch_bounds__(256) search_scan_add_kernel(int* __restrict__ out, int n, const int* __restrict__ tileOffsets)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n)
    {
        out[i] += tileOffsets[i / c_scanTile];
    }
}

struct CudaSearchBackend
{
    template<typename T>
    using Buf = DevBuf<T>;

    cudaStream_t st       = nullptr;
    long long    launches = 0;
    DevBuf<int>  tileSums, tileOffsets;
    int*         h_value = nullptr; /* pinned, 64 bytes: up to four values per read-back */

    int fail(const char* msg) { return nbb::fail("%s", msg); }

    template<typename T>
    int reserve(Buf<T>& b, size_t count)
    {
        if (count > b.alloc)
        {
            /* the buffer may still be read by kernels of the previous build */
            CU(cudaStreamSynchronize(st));
        }
        CU(b.reserve(count));
        return 0;
    }
    int zero(void* p, size_t bytes)
    {
        if (bytes > 0) CU(cudaMemsetAsync(p, 0, bytes, st));
        return 0;
    }
    int ones(void* p, size_t bytes)
    {
        if (bytes > 0) CU(cudaMemsetAsync(p, 0xff, bytes, st));
        return 0;
    }
    template<typename T>
    int upload(T* dst, const T* src, size_t count)
    {
        if (count > 0) CU(cudaMemcpyAsync(dst, src, sizeof(T) * count, cudaMemcpyHostToDevice, st));
        return 0;
    }
    int scan(const int* in, int* out, int n)
    {
        if (n <= 0) return 0;
        const int numTiles = (n + c_scanTile - 1) / c_scanTile;
        if (numTiles > c_scanTile) return nbb::fail("pair search: scan of %d items exceeds two levels", n);
        if (numTiles == 1)
        {
            search_scan_tile_kernel<<<1, 1024, 0, st>>>(in, out, n, nullptr);
            launches++;
        }
        else
        {
            if (reserve(tileSums, numTiles) || reserve(tileOffsets, numTiles)) return 1;
            search_scan_tile_kernel<<<numTiles, 1024, 0, st>>>(in, out, n, tileSums.p);
            search_scan_tile_kernel<<<1, 1024, 0, st>>>(tileSums.p, tileOffsets.p, numTiles, nullptr);
            search_scan_add_kernel<<<(n + 255) / 256, 256, 0, st>>>(out, n, tileOffsets.p);
            launches += 3;
        }
        CU(cudaGetLastError());
        return 0;
    }
    int readInt(const int* p, int* v)
    {
        CU(cudaMemcpyAsync(h_value, p, sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        *v = *reinterpret_cast<int*>(h_value);
        return 0;
    }
    /* several scalars with ONE stream synchronisation: two ints and one 64-bit count */
    int readInts2ULL(const int* p0, int* v0, const int* p1, int* v1, const unsigned long long* p2, unsigned long long* v2)
    {
        CU(cudaMemcpyAsync(h_value, p0, sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(h_value + 1, p1, sizeof(int), cudaMemcpyDeviceToHost, st));
        if (p2) CU(cudaMemcpyAsync(h_value + 2, p2, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        *v0 = h_value[0];
        *v1 = h_value[1];
        if (p2) *v2 = *reinterpret_cast<unsigned long long*>(h_value + 2);
        return 0;
    }
    int readULL(const unsigned long long* p, unsigned long long* v)
    {
        CU(cudaMemcpyAsync(h_value, p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        *v = *reinterpret_cast<unsigned long long*>(h_value);
        return 0;
    }
    template<typename F>
    int forEach(int n, F f)
    {
        if (n <= 0) return 0;
        search_items_kernel<F><<<(n + 255) / 256, 256, 0, st>>>(f, n);
        launches++;
        CU(cudaGetLastError());
        return 0;
    }
    template<typename F>
    int forEachWarp(int n, F f)
    {
        if (n <= 0) return 0;
        search_warps_kernel<F><<<(n + 7) / 8, 256, 0, st>>>(f, n);
        launches++;
        CU(cudaGetLastError());
        return 0;
    }
    template<typename F>
    int forEachBlock(int blocks, int threads, size_t scratchBytes, F f)
    {
        if (blocks <= 0) return 0;
        if (scratchBytes > 48 * 1024)
        {
            CU(cudaFuncSetAttribute(search_block_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(scratchBytes)));
        }
        search_block_kernel<F><<<blocks, threads, scratchBytes, st>>>(f);
        launches++;
        CU(cudaGetLastError());
        return 0;
    }
};

template<typename... B>
static void releaseAll(B&... b)
{
    (b.release(), ...);
}

} // namespace nbb

struct nbnxm_b200_gpu_search
{
    nbnxm_b200*                              nb = nullptr;
    int                                      device = 0;
    nbb::CudaSearchBackend                   be;
    nbs::SearchState<nbb::CudaSearchBackend> st;
    cudaEvent_t                              evStart = nullptr, evStop = nullptr;
    float                                    lastBuildMs = 0.0f, lastGridMs = 0.0f;
    int                                      numAtomsSet = 0;
    bool                                     haveExclSet = false;
};

using nbb::fail;

extern "C" {

int nbnxm_b200_gpu_search_create(nbnxm_b200_gpu_search_t** out, nbnxm_b200_t* nb)
{
    if (!out || !nb) return fail("nbnxm_b200_gpu_search_create: null argument");
    CU(cudaSetDevice(nb->device));
    void*       pinned = nullptr;
    cudaEvent_t ev[2]  = { nullptr, nullptr };
    if (cudaMallocHost(&pinned, 64) != cudaSuccess || cudaEventCreate(&ev[0]) != cudaSuccess || cudaEventCreate(&ev[1]) != cudaSuccess)
    {
        if (pinned) cudaFreeHost(pinned);
        if (ev[0]) cudaEventDestroy(ev[0]);
        return fail("nbnxm_b200_gpu_search_create: %s", cudaGetErrorString(cudaGetLastError()));
    }
    nbnxm_b200_gpu_search* s = new nbnxm_b200_gpu_search();
    s->nb                    = nb;
    s->device                = nb->device;
    s->be.st                 = nb->stream[0];
    s->be.h_value            = static_cast<int*>(pinned);
    s->evStart               = ev[0];
    s->evStop                = ev[1];
    /* the mask pass runs one warp per bin pair (2.0 ms against 3.1 ms for the 1.5 M-atom box, profiles/r02a_*);
     * NBNXM_B200_SEARCH_COOP=0 selects the one-thread-per-j-cluster form for A/B runs */
    const char* coop        = getenv("NBNXM_B200_SEARCH_COOP");
    s->st.cooperativeMasks = !(coop != nullptr && coop[0] == '0');
    /* column sort of the gridding: buckets + ranks; NBNXM_B200_SEARCH_BITONIC_SORT=1 selects the bitonic networks for A/B runs */
    const char* bitonic      = getenv("NBNXM_B200_SEARCH_BITONIC_SORT");
    s->st.bitonicColumnSort = (bitonic != nullptr && bitonic[0] == '1');
    /* column counts and scatter of the gridding: counters of a block of atoms in shared memory;
     * NBNXM_B200_SEARCH_GLOBAL_ATOMICS=1 selects one thread per atom with global atomics for A/B runs */
    const char* globalAtomics   = getenv("NBNXM_B200_SEARCH_GLOBAL_ATOMICS");
    s->st.globalColumnAtomics = (globalAtomics != nullptr && globalAtomics[0] == '1');
    *out = s;
    return 0;
}

int nbnxm_b200_gpu_search_free(nbnxm_b200_gpu_search_t* s)
{
    if (!s) return 0;
    /* the handle may already be gone (nbnxm_b200_free synchronises the device): nothing of it is touched here */
    cudaSetDevice(s->device);
    cudaDeviceSynchronize();
    auto& t = s->st;
    nbb::releaseAll(t.colFirstBin, t.atomIndex, t.slotOfAtom, t.clCount, t.exclIndex, t.exclAtoms, t.clBB, t.binBB,
                    t.entryNumBinPairs, t.entryBinPairOff, t.binPairJ, t.binPairEntry, t.binPairMask, t.entryNumJ,
                    t.entryGroups, t.entryCjOff, t.entryNumSci, t.entrySciOff, t.entryNonEmpty, t.entryCompactOff,
                    t.compactEntry, t.exclFlag, t.exclOff, t.numClusterPairs, t.sci, t.cjp, t.excl, s->be.tileSums, s->be.tileOffsets,
                    t.colOfAtom, t.colCount, t.colAtomStart, t.colBins, t.colFill, t.colAtoms, t.maxColCount, t.rankOfAtom, t.perturbedAtom, t.slotPert, t.clPert, t.fepInteracts, t.fepCount, t.fepOff, t.fepNonEmpty,
                    t.fepIOff, t.fepPairEntry, t.fepJjnr, t.fepIinr, t.fepShift, t.qAtom, t.ljCombPerType,
                    t.typeAtom);
    if (s->be.h_value) cudaFreeHost(s->be.h_value);
    if (s->evStart) cudaEventDestroy(s->evStart);
    if (s->evStop) cudaEventDestroy(s->evStop);
    delete s;
    return 0;
}

int nbnxm_b200_gpu_search_set_grid(nbnxm_b200_gpu_search_t* s, const float* box, int ncx, int ncy, const int* first_bin_of_column,
                                   const int* atom_index, int nbins, int natoms, const int* excl_index, const int* excl_atoms)
{
    if (!s || !box || !first_bin_of_column || !atom_index || ncx < 1 || ncy < 1 || nbins < 1 || natoms < 1)
    {
        return fail("nbnxm_b200_gpu_search_set_grid: bad argument");
    }
    if (nbins * nbs::c_binAtoms != s->nb->natoms)
    {
        return fail("nbnxm_b200_gpu_search_set_grid: the grid has %d nbat slots, the handle %d atoms (call nbnxm_b200_init_atomdata first)",
                    nbins * nbs::c_binAtoms, s->nb->natoms);
    }
    CU(cudaSetDevice(s->nb->device));
    /* the previous build may still read the grid arrays */
    CU(cudaStreamSynchronize(s->be.st));
    if (nbs::setGrid(s->be, s->st, box, ncx, ncy, first_bin_of_column, atom_index, nbins, natoms, excl_index, excl_atoms)) return 1;
    /* the host arrays are the caller's: finish the uploads before returning */
    CU(cudaStreamSynchronize(s->be.st));
    s->nb->launches += s->be.launches;
    s->be.launches = 0;
    return 0;
}

int nbnxm_b200_gpu_search_set_atoms(nbnxm_b200_gpu_search_t* s, int natoms, const float* q, const int* type, int ntypes,
                                    const float* lj_comb_per_type, const int* excl_index, const int* excl_atoms)
{
    if (!s || natoms < 1 || ntypes < 1) return fail("nbnxm_b200_gpu_search_set_atoms: bad argument");
    CU(cudaSetDevice(s->device));
    CU(cudaStreamSynchronize(s->be.st));
    if (nbs::setAtomProperties(s->be, s->st, natoms, q, type, ntypes, lj_comb_per_type)) return 1;
    if (nbs::setExclusions(s->be, s->st, natoms, excl_index, excl_atoms)) return 1;
    s->numAtomsSet  = natoms;
    s->haveExclSet  = excl_index != nullptr && excl_atoms != nullptr;
    CU(cudaStreamSynchronize(s->be.st));
    return 0;
}

int nbnxm_b200_gpu_search_put_atoms_on_grid(nbnxm_b200_gpu_search_t* s, const float* box, int nslabs, const float* d_x,
                                            void* x_ready_event, int* natoms_nbat, int* nbins_out, int* ncx_out, int* ncy_out)
{
    if (!s || !box || !d_x) return fail("nbnxm_b200_gpu_search_put_atoms_on_grid: null argument");
    if (s->numAtomsSet < 1) return fail("nbnxm_b200_gpu_search_put_atoms_on_grid: call nbnxm_b200_gpu_search_set_atoms first");
    nbnxm_b200* nb = s->nb;
    const bool  comb = (nb->params.vdw_type == NBNXM_B200_VDW_CUT_COMB_GEOM || nb->params.vdw_type == NBNXM_B200_VDW_CUT_COMB_LB);
    if (comb && !s->st.haveLjComb) return fail("nbnxm_b200_gpu_search_put_atoms_on_grid: this VdW flavor needs lj_comb_per_type");
    if (!comb && !s->st.haveType) return fail("nbnxm_b200_gpu_search_put_atoms_on_grid: this VdW flavor needs atom types");
    CU(cudaSetDevice(s->device));
    const int natoms = s->numAtomsSet;
    int       ncx = 1, ncy = 1, nbins = 0;
    if (nbnxm_b200_grid_dims(box, natoms, nslabs, &ncx, &ncy)) return fail("nbnxm_b200_gpu_search_put_atoms_on_grid: grid dimensions");
    if (x_ready_event) CU(cudaStreamWaitEvent(s->be.st, static_cast<cudaEvent_t>(x_ready_event), 0));
    CU(cudaEventRecord(s->evStart, s->be.st));
    if (nbs::putAtomsOnGrid(s->be, s->st, box, ncx, ncy, natoms, d_x, &nbins)) return 1;
    const int nslots = nbins * nbs::c_binAtoms;
    /* gpu_init_atomdata for the new grid, then atom data in nbat order, written on the device */
    if (nbnxm_b200_init_atomdata_device(nb, nslots, nslots)) return 1;
    if (nbs::fillAtomData(s->be, s->st, d_x, reinterpret_cast<nbs::XQ*>(nb->xq.p), comb ? nullptr : nb->atomType.p,
                          comb ? reinterpret_cast<float*>(nb->ljComb.p) : nullptr))
    {
        return 1;
    }
    /* the buffer ops of the following steps use the same order: x_to_nbat_x (slot -> atom), reduce_f (atom -> slot) */
    if (size_t(nslots) > nb->atomIndex.alloc || size_t(natoms) > nb->cell.alloc)
    {
        CU(cudaStreamSynchronize(nb->stream[0]));
        CU(cudaStreamSynchronize(nb->stream[1]));
    }
    CU(nb->atomIndex.reserve(nslots));
    CU(nb->cell.reserve(natoms));
    CU(cudaMemcpyAsync(nb->atomIndex.p, s->st.atomIndex.p, sizeof(int) * nslots, cudaMemcpyDeviceToDevice, s->be.st));
    CU(cudaMemcpyAsync(nb->cell.p, s->st.slotOfAtom.p, sizeof(int) * natoms, cudaMemcpyDeviceToDevice, s->be.st));
    nb->numCells = natoms;
    nb->xgrids.assign(1, nbnxm_b200::XGrid());
    nb->xgrids[0].first = 0;
    nb->xgrids[0].n     = nslots;
    CU(cudaEventRecord(s->evStop, s->be.st));
    CU(cudaStreamSynchronize(s->be.st));
    CU(cudaGetLastError());
    CU(cudaEventElapsedTime(&s->lastGridMs, s->evStart, s->evStop));
    nb->launches += s->be.launches;
    s->be.launches = 0;
    if (!s->haveExclSet)
    {
        s->st.g.exclIndex = nullptr;
        s->st.g.exclAtoms = nullptr;
    }
    if (natoms_nbat) *natoms_nbat = nslots;
    if (nbins_out) *nbins_out = nbins;
    if (ncx_out) *ncx_out = ncx;
    if (ncy_out) *ncy_out = ncy;
    return 0;
}

int nbnxm_b200_gpu_search_get_order(nbnxm_b200_gpu_search_t* s, int* atom_index, int* first_bin_of_column, float* grid_ms)
{
    if (!s || !s->st.haveGrid) return fail("nbnxm_b200_gpu_search_get_order: no grid");
    CU(cudaSetDevice(s->device));
    CU(cudaStreamSynchronize(s->be.st));
    const nbs::Grid& g = s->st.g;
    if (atom_index) CU(cudaMemcpy(atom_index, s->st.atomIndex.p, sizeof(int) * g.nbins * nbs::c_binAtoms, cudaMemcpyDeviceToHost));
    if (first_bin_of_column)
    {
        CU(cudaMemcpy(first_bin_of_column, s->st.colFirstBin.p, sizeof(int) * (g.ncx * g.ncy + 1), cudaMemcpyDeviceToHost));
    }
    if (grid_ms) *grid_ms = s->lastGridMs;
    return 0;
}

int nbnxm_b200_gpu_search_build(nbnxm_b200_gpu_search_t* s, int iloc, float rlist, int min_sci, int bin_begin, int bin_end,
                                int j_bin_lo, int j_bin_hi, int inter_zone, int required_tx)
{
    if (!s || iloc < 0 || iloc > 1) return fail("nbnxm_b200_gpu_search_build: bad argument");
    if (!s->st.haveGrid) return fail("nbnxm_b200_gpu_search_build: call nbnxm_b200_gpu_search_set_grid first");
    const nbs::Grid& g = s->st.g;
    if (bin_end < 0) bin_end = g.nbins;
    if (j_bin_hi < 0) j_bin_hi = g.nbins;
    if (bin_begin < 0 || bin_end > g.nbins || bin_begin > bin_end || j_bin_lo < 0 || j_bin_hi > g.nbins)
    {
        return fail("nbnxm_b200_gpu_search_build: bin range");
    }
    for (int d = 0; d < 3; d++)
    {
        if (2 * rlist >= g.box[d]) return fail("nbnxm_b200_gpu_search_build: rlist %g must be shorter than half the box (%g)", rlist, g.box[d]);
    }
    nbnxm_b200* nb = s->nb;
    CU(cudaSetDevice(nb->device));
    /* coordinates are copied on the local / non-local streams; the search runs on the local stream */
    if (nb->stream[1] != nb->stream[0]) CU(cudaStreamSynchronize(nb->stream[1]));
    CU(cudaEventRecord(s->evStart, s->be.st));
    if (nbs::buildPairlist(s->be, s->st, reinterpret_cast<const nbs::XQ*>(nb->xq.p), rlist, min_sci, bin_begin, bin_end, j_bin_lo,
                           j_bin_hi, inter_zone, required_tx))
    {
        return 1;
    }
    CU(cudaEventRecord(s->evStop, s->be.st));
    CU(cudaStreamSynchronize(s->be.st));
    CU(cudaGetLastError());
    CU(cudaEventElapsedTime(&s->lastBuildMs, s->evStart, s->evStop));
    nb->launches += s->be.launches;
    s->be.launches = 0;
    /* gpu_init_pairlist without the host: device-to-device into the handle's list */
    if (nbnxm_b200_init_pairlist_device(nb, iloc, s->st.sci.p, s->st.nsci, s->st.cjp.p, s->st.ncjp, s->st.excl.p, s->st.nexcl, nbs::c_cl))
    {
        return 1;
    }
    /* the perturbed pairs that left the cluster list: gpu_init_feppairlist without the host */
    if (s->st.havePerturbed
        && nbnxm_b200_init_feppairlist_device(nb, iloc, s->st.numFepI, s->st.numFepPairs, s->st.fepIinr.p, s->st.fepShift.p,
                                              s->st.fepPairEntry.p, s->st.fepJjnr.p, s->st.fepInteracts.p))
    {
        return 1;
    }
    /* the copies run on the list's stream; the next build (local stream) overwrites their source */
    if (nb->stream[iloc] != s->be.st || s->st.havePerturbed) CU(cudaStreamSynchronize(nb->stream[iloc]));
    return 0;
}

int nbnxm_b200_gpu_search_set_perturbed(nbnxm_b200_gpu_search_t* s, int natoms, const unsigned char* perturbed)
{
    if (!s || natoms < 0) return fail("nbnxm_b200_gpu_search_set_perturbed: bad argument");
    CU(cudaSetDevice(s->nb->device));
    if (perturbed != nullptr && s->st.haveGrid && natoms != s->st.g.natoms)
    {
        return fail("nbnxm_b200_gpu_search_set_perturbed: %d flags for a grid of %d atoms", natoms, s->st.g.natoms);
    }
    if (nbs::setPerturbed(s->be, s->st, natoms, perturbed)) return 1;
    CU(cudaStreamSynchronize(s->be.st)); /* the caller's array may go out of scope */
    return 0;
}

int nbnxm_b200_gpu_search_fep_sizes(const nbnxm_b200_gpu_search_t* s, int* num_i, int* num_j)
{
    if (!s) return fail("null search handle");
    if (num_i) *num_i = s->st.numFepI;
    if (num_j) *num_j = s->st.numFepPairs;
    return 0;
}

int nbnxm_b200_gpu_search_fep_download(nbnxm_b200_gpu_search_t* s, int* iinr, int* jindex, int* jjnr, int* shift, unsigned char* interacts)
{
    if (!s || !jindex) return fail("nbnxm_b200_gpu_search_fep_download: null argument");
    CU(cudaSetDevice(s->nb->device));
    CU(cudaStreamSynchronize(s->be.st));
    const int ni = s->st.numFepI, np = s->st.numFepPairs;
    if (iinr && ni > 0) CU(cudaMemcpy(iinr, s->st.fepIinr.p, sizeof(int) * ni, cudaMemcpyDeviceToHost));
    if (shift && ni > 0) CU(cudaMemcpy(shift, s->st.fepShift.p, sizeof(int) * ni, cudaMemcpyDeviceToHost));
    if (jjnr && np > 0) CU(cudaMemcpy(jjnr, s->st.fepJjnr.p, sizeof(int) * np, cudaMemcpyDeviceToHost));
    if (interacts && np > 0) CU(cudaMemcpy(interacts, s->st.fepInteracts.p, np, cudaMemcpyDeviceToHost));
    /* jindex from the i-entry of every pair (entries are consecutive and none is empty) */
    std::vector<int> pairEntry(np);
    if (np > 0) CU(cudaMemcpy(pairEntry.data(), s->st.fepPairEntry.p, sizeof(int) * np, cudaMemcpyDeviceToHost));
    for (int n = 0; n <= ni; n++) jindex[n] = 0;
    for (int k = 0; k < np; k++)
    {
        if (pairEntry[k] < 0 || pairEntry[k] >= ni) return fail("nbnxm_b200_gpu_search_fep_download: pair %d names i-entry %d of %d", k, pairEntry[k], ni);
        jindex[pairEntry[k] + 1]++;
    }
    for (int n = 0; n < ni; n++) jindex[n + 1] += jindex[n];
    return 0;
}

/* ---- x-slab lists on the device (multi-GPU search step without the host) ---- */

namespace nbb
{
/* global bin / cluster indices -> one rank's order (home bins first, then halo bins): nbnxm_b200_pairlist_reindex on the device */
__global__ void __launch_bounds__(256) search_reindex_kernel(nbnxm_b200_sci_t* __restrict__ sci, int nsci, nbnxm_b200_cj_packed_t* __restrict__ cjp,
                                                             int ncjp, int firstHomeBin, long long cjOffset, int numClustersTotal)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < nsci)
    {
        sci[i].sci -= firstHomeBin;
    }
    if (i < ncjp)
    {
#pragma unroll
        for (int jm = 0; jm < 4; jm++)
        {
            /* unused slots of partially filled j-groups carry no mask bits and an unspecified index: keep them loadable */
            const long long cj = static_cast<long long>(cjp[i].cj[jm]) + cjOffset;
            cjp[i].cj[jm]      = static_cast<int>(cj < 0 ? 0 : (cj > numClustersTotal - 1 ? numClustersTotal - 1 : cj));
        }
    }
}
} // namespace nbb

int nbnxm_b200_gpu_search_gather_slab(nbnxm_b200_gpu_search_t* s, nbnxm_b200_t* target, int home_begin, int home_end, int halo_begin,
                                      int halo_end)
{
    if (!s || !target || !s->st.haveGrid) return fail("nbnxm_b200_gpu_search_gather_slab: bad argument");
    const nbs::Grid& g = s->st.g;
    if (home_begin < 0 || home_end > g.nbins || home_begin > home_end || halo_begin < 0 || halo_end > g.nbins || halo_begin > halo_end)
    {
        return fail("nbnxm_b200_gpu_search_gather_slab: bin range");
    }
    nbnxm_b200* nb = s->nb;
    if (nb->device != target->device) return fail("nbnxm_b200_gpu_search_gather_slab: both handles must live on the same device");
    CU(cudaSetDevice(nb->device));
    const int nhome = (home_end - home_begin) * nbs::c_binAtoms, nhalo = (halo_end - halo_begin) * nbs::c_binAtoms;
    if (nbnxm_b200_init_atomdata_device(target, nhome + nhalo, nhome)) return 1;
    CU(cudaStreamSynchronize(s->be.st)); /* the gridding wrote the source arrays on the search stream */
    cudaStream_t st = target->stream[0];
    struct Range
    {
        int src, dst, n;
    } ranges[2] = { { home_begin * nbs::c_binAtoms, 0, nhome }, { halo_begin * nbs::c_binAtoms, nhome, nhalo } };
    for (const Range& r : ranges)
    {
        if (r.n == 0) continue;
        CU(cudaMemcpyAsync(target->xq.p + r.dst, nb->xq.p + r.src, sizeof(float4) * r.n, cudaMemcpyDeviceToDevice, st));
        if (nb->atomType.p && target->atomType.p)
        {
            CU(cudaMemcpyAsync(target->atomType.p + r.dst, nb->atomType.p + r.src, sizeof(int) * r.n, cudaMemcpyDeviceToDevice, st));
        }
        if (nb->ljComb.p && target->ljComb.p)
        {
            CU(cudaMemcpyAsync(target->ljComb.p + r.dst, nb->ljComb.p + r.src, sizeof(float2) * r.n, cudaMemcpyDeviceToDevice, st));
        }
    }
    CU(cudaStreamSynchronize(st));
    return 0;
}

int nbnxm_b200_gpu_search_build_slab(nbnxm_b200_gpu_search_t* s, nbnxm_b200_t* target, int iloc, float rlist, int min_sci, int home_begin,
                                     int home_end, int halo_begin, int halo_end, int required_tx)
{
    if (!s || !target || iloc < 0 || iloc > 1) return fail("nbnxm_b200_gpu_search_build_slab: bad argument");
    if (!s->st.haveGrid) return fail("nbnxm_b200_gpu_search_build_slab: put the atoms on the grid first");
    if (s->st.havePerturbed) return fail("nbnxm_b200_gpu_search_build_slab: the perturbed-pair split is not available for slab lists");
    const nbs::Grid& g = s->st.g;
    if (home_begin < 0 || home_end > g.nbins || home_begin > home_end || halo_begin < 0 || halo_end > g.nbins || halo_begin > halo_end)
    {
        return fail("nbnxm_b200_gpu_search_build_slab: bin range");
    }
    for (int d = 0; d < 3; d++)
    {
        if (2 * rlist >= g.box[d]) return fail("nbnxm_b200_gpu_search_build_slab: rlist %g must be shorter than half the box (%g)", rlist, g.box[d]);
    }
    nbnxm_b200* nb = s->nb;
    if (nb->device != target->device) return fail("nbnxm_b200_gpu_search_build_slab: both handles must live on the same device");
    CU(cudaSetDevice(nb->device));
    CU(cudaEventRecord(s->evStart, s->be.st));
    /* local: home x home, half shell; non-local: home x halo, every y / z shift, the periodic x image in required_tx */
    const int jLo = iloc == 0 ? home_begin : halo_begin, jHi = iloc == 0 ? home_end : halo_end;
    if (nbs::buildPairlist(s->be, s->st, reinterpret_cast<const nbs::XQ*>(nb->xq.p), rlist, min_sci, home_begin, home_end, jLo, jHi,
                           iloc == 1 ? 1 : 0, iloc == 1 ? required_tx : 0))
    {
        return 1;
    }
    const int       numHome  = home_end - home_begin;
    const int       ncl      = (numHome + (halo_end - halo_begin)) * nbs::c_binCl;
    const long long cjOffset = iloc == 1 ? static_cast<long long>(numHome - halo_begin) * nbs::c_binCl
                                         : -static_cast<long long>(home_begin) * nbs::c_binCl;
    const int       n        = s->st.nsci > s->st.ncjp ? s->st.nsci : s->st.ncjp;
    if (n > 0)
    {
        nbb::search_reindex_kernel<<<(n + 255) / 256, 256, 0, s->be.st>>>(s->st.sci.p, s->st.nsci, s->st.cjp.p, s->st.ncjp, home_begin,
                                                                          cjOffset, ncl);
        s->be.launches++;
    }
    CU(cudaEventRecord(s->evStop, s->be.st));
    CU(cudaStreamSynchronize(s->be.st));
    CU(cudaGetLastError());
    CU(cudaEventElapsedTime(&s->lastBuildMs, s->evStart, s->evStop));
    target->launches += s->be.launches;
    s->be.launches = 0;
    if (nbnxm_b200_init_pairlist_device(target, iloc, s->st.sci.p, s->st.nsci, s->st.cjp.p, s->st.ncjp, s->st.excl.p, s->st.nexcl, nbs::c_cl))
    {
        return 1;
    }
    CU(cudaStreamSynchronize(target->stream[iloc]));
    return 0;
}

int nbnxm_b200_gpu_search_sizes(const nbnxm_b200_gpu_search_t* s, int* nsci, int* ncj_packed, int* nexcl, long long* ncluster_pairs,
                                float* build_ms)
{
    if (!s) return fail("null search handle");
    if (nsci) *nsci = s->st.nsci;
    if (ncj_packed) *ncj_packed = s->st.ncjp;
    if (nexcl) *nexcl = s->st.nexcl;
    if (ncluster_pairs) *ncluster_pairs = s->st.numClusterPairsHost;
    if (build_ms) *build_ms = s->lastBuildMs;
    return 0;
}

int nbnxm_b200_gpu_search_download(nbnxm_b200_gpu_search_t* s, nbnxm_b200_sci_t* sci, nbnxm_b200_cj_packed_t* cj_packed,
                                   nbnxm_b200_excl_t* excl)
{
    if (!s) return fail("null search handle");
    CU(cudaSetDevice(s->nb->device));
    CU(cudaStreamSynchronize(s->be.st));
    if (sci && s->st.nsci > 0) CU(cudaMemcpy(sci, s->st.sci.p, sizeof(*sci) * s->st.nsci, cudaMemcpyDeviceToHost));
    if (cj_packed && s->st.ncjp > 0) CU(cudaMemcpy(cj_packed, s->st.cjp.p, sizeof(*cj_packed) * s->st.ncjp, cudaMemcpyDeviceToHost));
    if (excl && s->st.nexcl > 0) CU(cudaMemcpy(excl, s->st.excl.p, sizeof(*excl) * s->st.nexcl, cudaMemcpyDeviceToHost));
    return 0;
}

} // extern "C"
