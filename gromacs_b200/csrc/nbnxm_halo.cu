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

    /* Peer-memory path: no transport calls at all.  The non-local kernel reads the +x neighbour's home coordinates
     * and reduces the forces on them straight into the neighbour's accumulator over NVLink (the arrays are mapped
     * with CUDA IPC); ranks synchronise through step counters in device memory:
     *   flags[0] of a rank = last step for which its coordinates are in place and its accumulator is cleared,
     *   flags[1] of a rank = last step for which its -x neighbour has finished adding forces to it. */
    bool        peerEnabled = false;
    int         peerStep    = 0;
    DevBuf<int> flags;               /* ours: [0] ready, [1] forces from -x neighbour done, [2] error */
    float4*     upXq    = nullptr;   /* +x neighbour's xq / f4 / flags, mapped */
    float4*     upF4    = nullptr;
    int*        upFlags = nullptr;
};

/* what a rank publishes to its -x neighbour */
struct PeerBlob
{
    cudaIpcMemHandle_t xq, f4, flags;
    int                homeFirst; /* first atom of the range the neighbour imports as its halo */
    int                natoms;
};

__global__ void flag_set_kernel(int* flag, int value)
{
    /* everything this stream did before (kernels writing our or the neighbour's memory) is ordered before the flag */
    __threadfence_system();
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}

__global__ void flag_wait_kernel(const int* flag, int value, int* errorFlag)
{
    wait_flag_geq(flag, value, errorFlag);
}

__global__ void __launch_bounds__(256) halo_add_f_kernel(float4* __restrict__ f4, const float4* __restrict__ in, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
    {
        const float4 v = in[i];
        red_add_v4(f4 + i, v.x, v.y, v.z);
    }
}

bool peer_halo_enabled(const nbnxm_b200* nb) { return nb && nb->halo && nb->halo->peerEnabled; }
int  peer_next_step(nbnxm_b200* nb) { return ++nb->halo->peerStep; }
int  peer_publish_ready(nbnxm_b200* nb, int n, cudaStream_t s)
{
    flag_set_kernel<<<1, 1, 0, s>>>(nb->halo->flags.p + 0, n);
    nb->launches++;
    CU(cudaGetLastError());
    return 0;
}
int peer_wait_neighbour_ready(nbnxm_b200* nb, int n, cudaStream_t s)
{
    flag_wait_kernel<<<1, 1, 0, s>>>(nb->halo->upFlags + 0, n, nb->halo->flags.p + 2);
    nb->launches++;
    CU(cudaGetLastError());
    return 0;
}
int peer_publish_forces_done(nbnxm_b200* nb, int n, cudaStream_t s)
{
    flag_set_kernel<<<1, 1, 0, s>>>(nb->halo->upFlags + 1, n);
    nb->launches++;
    CU(cudaGetLastError());
    return 0;
}
int peer_wait_forces_from_neighbour(nbnxm_b200* nb, int n, cudaStream_t s)
{
    flag_wait_kernel<<<1, 1, 0, s>>>(nb->halo->flags.p + 1, n, nb->halo->flags.p + 2);
    nb->launches++;
    CU(cudaGetLastError());
    return 0;
}
void peer_send_range(const nbnxm_b200* nb, int* first, int* count)
{
    *first = nb->halo->sendFirst;
    *count = nb->halo->sendCount;
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
    nbnxm_b200_peer_close(nb);
    HaloState* h = nb->halo;
    if (h->comm) g_nccl.CommDestroy(h->comm);
    h->fRecv.release();
    h->flags.release();
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

/* ---- peer-memory halo path ---- */
int nbnxm_b200_peer_blob_size(void) { return int(sizeof(PeerBlob)); }

int nbnxm_b200_peer_export(nbnxm_b200_t* nb, unsigned char* blob, int nbytes)
{
    if (!nb || !nb->halo || !blob || nbytes < int(sizeof(PeerBlob))) return fail("nbnxm_b200_peer_export: bad argument");
    if (!nb->xq.p || !nb->f4.p) return fail("nbnxm_b200_peer_export: call after gpu_init_atomdata");
    HaloState* h = nb->halo;
    CU(cudaSetDevice(nb->device));
    if (!h->flags.p)
    {
        CU(h->flags.reserve(8));
        CU(cudaMemset(h->flags.p, 0, sizeof(int) * 8));
    }
    PeerBlob b;
    memset(&b, 0, sizeof(b));
    CU(cudaIpcGetMemHandle(&b.xq, nb->xq.p));
    CU(cudaIpcGetMemHandle(&b.f4, nb->f4.p));
    CU(cudaIpcGetMemHandle(&b.flags, h->flags.p));
    b.homeFirst = h->sendFirst;
    b.natoms    = nb->natoms;
    memcpy(blob, &b, sizeof(b));
    return 0;
}

/* maps the +x neighbour's arrays; collective in the sense that every rank must have exported first */
int nbnxm_b200_peer_import(nbnxm_b200_t* nb, const unsigned char* up_blob, int nbytes)
{
    if (!nb || !nb->halo || !up_blob || nbytes < int(sizeof(PeerBlob))) return fail("nbnxm_b200_peer_import: bad argument");
    HaloState* h = nb->halo;
    if (h->nranks < 2) return fail("nbnxm_b200_peer_import: needs at least two ranks");
    if (!h->flags.p) return fail("nbnxm_b200_peer_import: export first");
    CU(cudaSetDevice(nb->device));
    PeerBlob b;
    memcpy(&b, up_blob, sizeof(b));
    void *xq = nullptr, *f4 = nullptr, *fl = nullptr;
    CU(cudaIpcOpenMemHandle(&xq, b.xq, cudaIpcMemLazyEnablePeerAccess));
    CU(cudaIpcOpenMemHandle(&f4, b.f4, cudaIpcMemLazyEnablePeerAccess));
    CU(cudaIpcOpenMemHandle(&fl, b.flags, cudaIpcMemLazyEnablePeerAccess));
    h->upXq    = static_cast<float4*>(xq);
    h->upF4    = static_cast<float4*>(f4);
    h->upFlags = static_cast<int*>(fl);
    /* our halo atoms [recvFirst, recvFirst + recvCount) are the neighbour's [homeFirst, ...) */
    nb->peerXqJ    = h->upXq + b.homeFirst - h->recvFirst;
    nb->peerF4J    = h->upF4 + b.homeFirst - h->recvFirst;
    h->peerEnabled = true;
    h->peerStep    = 0;
    return 0;
}

int nbnxm_b200_peer_close(nbnxm_b200_t* nb)
{
    if (!nb || !nb->halo) return 0;
    HaloState* h = nb->halo;
    cudaSetDevice(nb->device);
    cudaDeviceSynchronize();
    if (h->upXq) cudaIpcCloseMemHandle(h->upXq);
    if (h->upF4) cudaIpcCloseMemHandle(h->upF4);
    if (h->upFlags) cudaIpcCloseMemHandle(h->upFlags);
    h->upXq = h->upF4 = nullptr;
    h->upFlags     = nullptr;
    nb->peerXqJ    = nullptr;
    nb->peerF4J    = nullptr;
    h->peerEnabled = false;
    return 0;
}

/* 0 = fine, 1 = a wait on a neighbour's step counter timed out */
int nbnxm_b200_peer_error(nbnxm_b200_t* nb, int* error)
{
    if (!nb || !nb->halo || !error) return fail("nbnxm_b200_peer_error: bad argument");
    *error = 0;
    if (nb->halo->flags.p)
    {
        CU(cudaSetDevice(nb->device));
        CU(cudaMemcpy(error, nb->halo->flags.p + 2, sizeof(int), cudaMemcpyDeviceToHost));
    }
    return 0;
}

/* The do_force sequence with the peer-memory halo: no coordinate or force transport.
 *   local stream   : [H2D xq] clear -> flags[0] = step -> local kernel -> [local rolling prune] -> wait own non-local
 *                    stream, wait flags[1] >= step -> copy-back
 *   non-local strm : wait (clear done), wait +x neighbour's flags[0] >= step -> non-local kernel (j-atoms in the
 *                    neighbour's memory) -> [non-local rolling prune] -> neighbour's flags[1] = step */
static int peerForceStep(nbnxm_b200_t* nb, int step, const nbnxm_b200_step_flags_t* fl, const float* xq_host, float* f_host)
{
    HaloState*   h  = nb->halo;
    const int    e = fl->compute_energy, v = fl->compute_virial;
    cudaStream_t sl = nb->stream[0], sn = nb->stream[1];
    const int    n  = ++h->peerStep;
    CU(cudaSetDevice(nb->device));
    if (xq_host && nbnxm_b200_copy_xq_to_gpu(nb, 0, xq_host)) return 1;
    if (nbnxm_b200_clear_outputs(nb, v)) return 1;
    flag_set_kernel<<<1, 1, 0, sl>>>(h->flags.p + 0, n);
    if (nbnxm_b200_insert_nonlocal_dependency(nb, 0)) return 1;
    if (nbnxm_b200_insert_nonlocal_dependency(nb, 1)) return 1;
    flag_wait_kernel<<<1, 1, 0, sn>>>(h->upFlags + 0, n, h->flags.p + 2);
    nb->launches += 2;
    if (nbnxm_b200_launch_kernel(nb, 0, e, v)) return 1;
    if (nbnxm_b200_launch_kernel(nb, 1, e, v)) return 1;
    if (fl->dynamic_pruning)
    {
        if (nbnxm_b200_launch_kernel_pruneonly(nb, step % 2 == 0 ? 0 : 1, fl->rolling_prune_parts)) return 1;
    }
    /* the neighbour may use its forces (and overwrite its coordinates) once our non-local work on them is done */
    flag_set_kernel<<<1, 1, 0, sn>>>(h->upFlags + 1, n);
    nb->launches++;
    if (nbnxm_b200_launch_cpyback(nb, 1, f_host, 0, 0, 1)) return 1;
    /* the -x neighbour's non-local kernel adds to our accumulator: wait for it before reading the forces */
    flag_wait_kernel<<<1, 1, 0, sl>>>(h->flags.p + 1, n, h->flags.p + 2);
    nb->launches++;
    return nbnxm_b200_launch_cpyback(nb, 0, f_host, e, v, f_host == nullptr);
}

int nbnxm_b200_do_force_step(nbnxm_b200_t* nb, int step, const nbnxm_b200_step_flags_t* fl, const float* xq_host, float* f_host)
{
    if (!nb || !fl) return nbb::fail("nbnxm_b200_do_force_step: null argument");
    const int e = fl->compute_energy, v = fl->compute_virial;
    if (fl->have_halo == 3)
    {
        if (!nb->halo || !nb->halo->peerEnabled) return nbb::fail("nbnxm_b200_do_force_step: peer-memory halo not set up");
        return peerForceStep(nb, step, fl, xq_host, f_host);
    }
    if (xq_host && nbnxm_b200_copy_xq_to_gpu(nb, 0, xq_host)) return 1;
    if (nbnxm_b200_clear_outputs(nb, v)) return 1;
    if (fl->have_halo)
    {
        /* clear + H2D done -> the non-local stream may start */
        if (nbnxm_b200_insert_nonlocal_dependency(nb, 0)) return 1;
        if (fl->have_halo == 1 && nbnxm_b200_halo_exchange_x(nb)) return 1;
    }
    /* single rank: the rolling prune of odd steps runs beside the force kernel as background work (lowest-priority stream,
     * launched first so that both are pending together) instead of after it; in line after a fresh list's first-pass prune */
    const bool backgroundPrune = !fl->have_halo && fl->dynamic_pruning && step % 2 == 1 && nbb::background_prune_possible(nb);
    if (backgroundPrune && nbb::launch_background_prune(nb, fl->rolling_prune_parts)) return 1;
    if (nbnxm_b200_launch_kernel(nb, 0, e, v)) return 1;
    if (fl->have_halo)
    {
        if (nbnxm_b200_launch_kernel(nb, 1, e, v)) return 1;
        if (fl->have_halo == 1 && nbnxm_b200_halo_exchange_f(nb)) return 1;
    }
    if (backgroundPrune)
    {
        if (nbb::join_background_prune(nb)) return 1;
    }
    else if (fl->dynamic_pruning)
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

