/* x-slab halo exchange of the nbnxm_b200 path: coordinates of the +x neighbour's first columns in,
 * forces on them back, over NCCL point-to-point on NVLink.
 *
 * Plays the role of gmx::GpuHaloExchange (src/gromacs/domdec/gpuhaloexchange.h:80-130;
 * communicateHaloCoordinates / communicateHaloForces, src/gromacs/domdec/gpuhaloexchange_impl_gpu.cpp:286-470)
 * for a 1-D decomposition with one pulse.  Differences by design:
 *   - atoms are kept in nbat (grid) order and slabs are whole grid columns, so the send and receive
 *     regions are contiguous ranges of the xq / f arrays: no index map, no pack kernel; coordinates are
 *     received straight into NBAtomDataGpu::xq;
 *   - the periodic image shift is carried by the pair list (nbnxm_sci_t::shift), not applied to the halo;
 *   - forces travel as the 16-byte internal accumulator and are added on receipt with one
 *     red.global.add.v4.f32 per atom (the local kernel may be adding to the same atoms concurrently);
 *   - transport is ncclSend / ncclRecv on the handle's non-local stream: stream-ordered with the
 *     kernels on either side, no host synchronisation.
 * libnccl is loaded at run time (dlopen) so that the single-GPU path has no NCCL dependency.
 */
#include <dlfcn.h>

#include <cstring>

#include "nbnxm_handle.cuh"

namespace nbb
{

/* the subset of nccl.h this file needs (NCCL's ABI for these entry points is stable across 2.x) */
typedef struct ncclComm* ncclComm_t;
typedef struct
{
    char internal[128];
} ncclUniqueId;
enum
{
    c_ncclSuccess = 0,
    c_ncclFloat32 = 7
};

struct NcclApi
{
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId*)                                          = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int)                   = nullptr;
    int (*CommDestroy)(ncclComm_t)                                             = nullptr;
    int (*GroupStart)()                                                        = nullptr;
    int (*GroupEnd)()                                                          = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t)       = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t)             = nullptr;
    const char* (*GetErrorString)(int)                                         = nullptr;
    const char* (*GetLastError)(ncclComm_t)                                    = nullptr;
};

static NcclApi g_nccl;

static int loadNccl()
{
    if (g_nccl.lib) return 0;
    /* a process that already carries an NCCL (e.g. the one bundled with torch) shares it */
    const char* names[] = { "libnccl.so.2", "libnccl.so" };
    void*       lib     = nullptr;
    for (const char* n : names)
    {
        lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) return fail("nbnxm_b200 halo: cannot load libnccl.so.2 (%s)", dlerror());
#define SYM(field, name)                                                              \
    *reinterpret_cast<void**>(&g_nccl.field) = dlsym(lib, name);                      \
    if (!g_nccl.field) return fail("nbnxm_b200 halo: libnccl lacks the symbol %s", name)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    *reinterpret_cast<void**>(&g_nccl.GetLastError) = dlsym(lib, "ncclGetLastError");
    g_nccl.lib = lib;
    return 0;
}

#define NC(call)                                                                                           \
    do                                                                                                     \
    {                                                                                                      \
        int r_ = (call);                                                                                   \
        if (r_ != c_ncclSuccess)                                                                           \
        {                                                                                                  \
            return nbb::fail("%s:%d %s failed: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); \
        }                                                                                                  \
    } while (0)

struct HaloState
{
    ncclComm_t     comm = nullptr;
    int            rank = 0, nranks = 1;
    int            sendFirst = 0, sendCount = 0; /* home atoms the -x neighbour needs */
    int            recvFirst = 0, recvCount = 0; /* halo atoms, owned by the +x neighbour */
    DevBuf<float4> fRecv;                        /* forces on our sendFirst.. atoms computed by the -x neighbour */
    bool           timing = false;
    cudaEvent_t    ev[4]  = { nullptr, nullptr, nullptr, nullptr };
    double         xMs = 0, fMs = 0;
    int            xCount = 0, fCount = 0;
};

__global__ void __launch_bounds__(256) halo_add_f_kernel(float4* __restrict__ f4, const float4* __restrict__ in, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
    {
        const float4 v = in[i];
        red_add_v4(f4 + i, v.x, v.y, v.z);
    }
}

} // namespace nbb

using namespace nbb;

extern "C" {

int nbnxm_b200_halo_get_unique_id(char* id, int nbytes)
{
    if (!id || nbytes < int(sizeof(ncclUniqueId))) return fail("nbnxm_b200_halo_get_unique_id: need a %d-byte buffer", int(sizeof(ncclUniqueId)));
    if (loadNccl()) return 1;
    ncclUniqueId u;
    NC(g_nccl.GetUniqueId(&u));
    memcpy(id, u.internal, sizeof(u.internal));
    return 0;
}

int nbnxm_b200_halo_init(nbnxm_b200_t* nb, const char* id, int rank, int nranks)
{
    if (!nb || !id || nranks < 1 || rank < 0 || rank >= nranks) return fail("nbnxm_b200_halo_init: bad argument");
    if (nb->halo) return fail("nbnxm_b200_halo_init: already initialised");
    if (!nb->localAndNonlocal && nranks > 1) return fail("nbnxm_b200_halo_init: the handle was created without a non-local stream");
    if (loadNccl()) return 1;
    CU(cudaSetDevice(nb->device));
    HaloState* h = new HaloState();
    h->rank      = rank;
    h->nranks    = nranks;
    ncclUniqueId u;
    memcpy(u.internal, id, sizeof(u.internal));
    int r = g_nccl.CommInitRank(&h->comm, nranks, u, rank);
    if (r != c_ncclSuccess)
    {
        delete h;
        return fail("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
    }
    for (cudaEvent_t& e : h->ev) CU(cudaEventCreate(&e));
    nb->halo = h;
    return 0;
}

int nbnxm_b200_halo_free(nbnxm_b200_t* nb)
{
    if (!nb || !nb->halo) return 0;
    cudaSetDevice(nb->device);
    cudaDeviceSynchronize();
    HaloState* h = nb->halo;
    if (h->comm) g_nccl.CommDestroy(h->comm);
    h->fRecv.release();
    for (cudaEvent_t e : h->ev)
        if (e) cudaEventDestroy(e);
    delete h;
    nb->halo = nullptr;
    return 0;
}

int nbnxm_b200_halo_set_ranges(nbnxm_b200_t* nb, int send_first, int send_count, int recv_first, int recv_count)
{
    if (!nb || !nb->halo) return fail("nbnxm_b200_halo_set_ranges: halo exchange not initialised");
    if (send_first < 0 || send_count < 0 || send_first + send_count > nb->natomsLocal)
    {
        return fail("nbnxm_b200_halo_set_ranges: send range [%d, %d) is not inside the %d local atoms", send_first,
                    send_first + send_count, nb->natomsLocal);
    }
    if (recv_first < nb->natomsLocal || recv_count < 0 || recv_first + recv_count > nb->natoms)
    {
        return fail("nbnxm_b200_halo_set_ranges: receive range [%d, %d) is not inside the non-local atoms [%d, %d)", recv_first,
                    recv_first + recv_count, nb->natomsLocal, nb->natoms);
    }
    CU(cudaSetDevice(nb->device));
    HaloState* h = nb->halo;
    h->sendFirst = send_first;
    h->sendCount = send_count;
    h->recvFirst = recv_first;
    h->recvCount = recv_count;
    if (size_t(send_count) > h->fRecv.alloc)
    {
        CU(cudaStreamSynchronize(nb->stream[1]));
    }
    CU(h->fRecv.reserve(send_count > 0 ? send_count : 1));
    return 0;
}

/* coordinates: our first columns to the -x neighbour, the +x neighbour's first columns into our halo */
int nbnxm_b200_halo_exchange_x(nbnxm_b200_t* nb)
{
    if (!nb || !nb->halo) return fail("nbnxm_b200_halo_exchange_x: halo exchange not initialised");
    HaloState* h = nb->halo;
    if (h->nranks == 1) return 0;
    CU(cudaSetDevice(nb->device));
    cudaStream_t st = nb->stream[1];
    /* the local coordinates must be on the device before they are sent */
    if (nbnxm_b200_insert_nonlocal_dependency(nb, 1)) return 1;
    const int down = (h->rank + h->nranks - 1) % h->nranks, up = (h->rank + 1) % h->nranks;
    if (h->timing) CU(cudaEventRecord(h->ev[0], st));
    NC(g_nccl.GroupStart());
    if (h->sendCount > 0) NC(g_nccl.Send(nb->xq.p + h->sendFirst, size_t(h->sendCount) * 4, c_ncclFloat32, down, h->comm, st));
    if (h->recvCount > 0) NC(g_nccl.Recv(nb->xq.p + h->recvFirst, size_t(h->recvCount) * 4, c_ncclFloat32, up, h->comm, st));
    NC(g_nccl.GroupEnd());
    if (h->timing) CU(cudaEventRecord(h->ev[1], st));
    return 0;
}

/* forces: what our non-local kernel put on the halo atoms goes back to their owner (+x), what the -x
 * neighbour computed on our first columns is added to our accumulator */
int nbnxm_b200_halo_exchange_f(nbnxm_b200_t* nb)
{
    if (!nb || !nb->halo) return fail("nbnxm_b200_halo_exchange_f: halo exchange not initialised");
    HaloState* h = nb->halo;
    if (h->nranks == 1) return 0;
    CU(cudaSetDevice(nb->device));
    cudaStream_t st   = nb->stream[1];
    const int    down = (h->rank + h->nranks - 1) % h->nranks, up = (h->rank + 1) % h->nranks;
    if (h->timing) CU(cudaEventRecord(h->ev[2], st));
    NC(g_nccl.GroupStart());
    if (h->recvCount > 0) NC(g_nccl.Send(nb->f4.p + h->recvFirst, size_t(h->recvCount) * 4, c_ncclFloat32, up, h->comm, st));
    if (h->sendCount > 0) NC(g_nccl.Recv(h->fRecv.p, size_t(h->sendCount) * 4, c_ncclFloat32, down, h->comm, st));
    NC(g_nccl.GroupEnd());
    if (h->sendCount > 0)
    {
        halo_add_f_kernel<<<(h->sendCount + 255) / 256, 256, 0, st>>>(nb->f4.p + h->sendFirst, h->fRecv.p, h->sendCount);
        nb->launches++;
    }
    if (h->timing) CU(cudaEventRecord(h->ev[3], st));
    CU(cudaGetLastError());
    return 0;
}

int nbnxm_b200_halo_set_timing(nbnxm_b200_t* nb, int enable)
{
    if (!nb || !nb->halo) return fail("halo exchange not initialised");
    nb->halo->timing = enable != 0;
    return 0;
}

/* accumulates the device time of the last x and f exchange (call after the step completed) */
int nbnxm_b200_halo_get_timings(nbnxm_b200_t* nb, double* x_ms, double* f_ms, int reset)
{
    if (!nb || !nb->halo) return fail("halo exchange not initialised");
    HaloState* h = nb->halo;
    CU(cudaSetDevice(nb->device));
    if (h->timing && h->nranks > 1)
    {
        float a = 0, b = 0;
        if (cudaEventSynchronize(h->ev[1]) == cudaSuccess && cudaEventElapsedTime(&a, h->ev[0], h->ev[1]) == cudaSuccess)
        {
            h->xMs += a;
            h->xCount++;
        }
        if (cudaEventSynchronize(h->ev[3]) == cudaSuccess && cudaEventElapsedTime(&b, h->ev[2], h->ev[3]) == cudaSuccess)
        {
            h->fMs += b;
            h->fCount++;
        }
        cudaGetLastError();
    }
    if (x_ms) *x_ms = h->xCount ? h->xMs / h->xCount : 0.0;
    if (f_ms) *f_ms = h->fCount ? h->fMs / h->fCount : 0.0;
    if (reset)
    {
        h->xMs = h->fMs = 0;
        h->xCount = h->fCount = 0;
    }
    return 0;
}

} // extern "C"

int nbnxm_b200_do_force_step(nbnxm_b200_t* nb, int step, const nbnxm_b200_step_flags_t* fl, const float* xq_host, float* f_host)
{
    if (!nb || !fl) return nbb::fail("nbnxm_b200_do_force_step: null argument");
    const int e = fl->compute_energy, v = fl->compute_virial;
    if (xq_host && nbnxm_b200_copy_xq_to_gpu(nb, 0, xq_host)) return 1;
    if (nbnxm_b200_clear_outputs(nb, v)) return 1;
    if (fl->have_halo)
    {
        /* clear + H2D done -> the non-local stream may start */
        if (nbnxm_b200_insert_nonlocal_dependency(nb, 0)) return 1;
        if (nbnxm_b200_halo_exchange_x(nb)) return 1;
    }
    if (nbnxm_b200_launch_kernel(nb, 0, e, v)) return 1;
    if (fl->have_halo)
    {
        if (nbnxm_b200_launch_kernel(nb, 1, e, v)) return 1;
        if (nbnxm_b200_halo_exchange_f(nb)) return 1;
    }
    if (fl->dynamic_pruning)
    {
        if (!fl->have_halo)
        {
            if (step % 2 == 1 && nbnxm_b200_launch_kernel_pruneonly(nb, 0, fl->rolling_prune_parts)) return 1;
        }
        else if (nbnxm_b200_launch_kernel_pruneonly(nb, step % 2 == 0 ? 0 : 1, fl->rolling_prune_parts))
        {
            return 1;
        }
    }
    const int onDevice = (f_host == nullptr);
    if (fl->have_halo && nbnxm_b200_launch_cpyback(nb, 1, f_host, 0, 0, 1)) return 1;
    return nbnxm_b200_launch_cpyback(nb, 0, f_host, e, v, onDevice);
}

