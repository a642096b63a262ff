/* Private definition of the nbnxm_b200 handle (the reference's NbnxmGpu,
 * src/gromacs/nbnxm/cuda/nbnxm_cuda_types.h:69) shared by the translation units that implement the
 * C ABI (nbnxm_api.cu, nbnxm_halo.cu). Not part of the public interface. */
#ifndef NBNXM_B200_HANDLE_CUH
#define NBNXM_B200_HANDLE_CUH

#include <cstdarg>
#include <cstdio>

#include <set>
#include <vector>

#include "nbnxm_device.cuh"

namespace nbb
{

extern thread_local char g_lastError[512];
/* sets the thread's last-error message, returns 1 */
int fail(const char* fmt, ...);

#define CU(call)                                                                                      \
    do                                                                                                \
    {                                                                                                 \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
        {                                                                                             \
            return nbb::fail("%s:%d %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
        }                                                                                             \
    } while (0)

template<typename T>
struct DevBuf
{
    T*     p     = nullptr;
    size_t n     = 0;
    size_t alloc = 0;
    /* grow-only reallocation with 20 % slack, contents are not preserved
     * (reallocateDeviceBuffer, src/gromacs/gpu_utils/devicebuffer.h) */
    cudaError_t reserve(size_t count)
    {
        n = count;
        if (count <= alloc)
        {
            return cudaSuccess;
        }
        if (p)
        {
            cudaFree(p);
            p = nullptr;
        }
        alloc         = count + count / 5 + 64;
        cudaError_t e = cudaMalloc(&p, alloc * sizeof(T));
        if (e != cudaSuccess)
        {
            alloc = 0;
            n     = 0;
        }
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p     = nullptr;
        n     = 0;
        alloc = 0;
    }
};

struct PairList
{
    DevBuf<nbnxm_b200_sci_t>       sci, sciSorted;
    DevBuf<int>                    sciCount, sciHistogram, sciOffset, rollingPart;
    DevBuf<nbnxm_b200_cj_packed_t> cjPacked;
    DevBuf<unsigned int>           imaskOuter;
    DevBuf<nbnxm_b200_excl_t>      excl;
    DevBuf<unsigned long long>     pairCount;
    int                            numSci               = 0;
    int                            naCi                 = -1;
    bool                           haveFreshList        = false;
    int                            rollingNumParts      = 0;
    bool                           didPrune             = false;
    bool                           didRollingPrune      = false;

    PairlistDev dev(bool counting) const
    {
        PairlistDev d;
        d.sci          = sci.p;
        d.sciSorted    = sciSorted.p;
        d.sciCount     = sciCount.p;
        d.sciHistogram = sciHistogram.p;
        d.sciOffset    = sciOffset.p;
        d.cjPacked     = cjPacked.p;
        d.imaskOuter   = imaskOuter.p;
        d.excl         = excl.p;
        d.rollingPart  = rollingPart.p;
        d.pairCount    = counting ? pairCount.p : nullptr;
        d.numSci       = numSci;
        return d;
    }
};

struct TimedRegion
{
    cudaEvent_t start, stop;
    int         kind; /* 0..3 force[prune][energy], 4 prune, 5 rolling prune, 6 xq h2d, 7 f d2h, 8 pairlist h2d */
};

struct HaloState; /* nbnxm_halo.cu */

} // namespace nbb

struct nbnxm_b200
{
    int                 device = 0;
    cudaStream_t        stream[2]    = { nullptr, nullptr };
    bool                ownStream[2] = { false, false };
    bool                localAndNonlocal = false;
    nbnxm_b200_params_t params{};
    int                 numTypes = 0;
    int                 numSMs   = 0;

    nbb::DevBuf<float4> xq, f4;
    nbb::DevBuf<float>  f3;
    nbb::DevBuf<int>    atomType;
    nbb::DevBuf<float2> ljComb;
    nbb::DevBuf<float>  shiftVec;
    nbb::DevBuf<double> fshift, energy;
    /* gpuGetNBAtomData (nbnxm_gpu_data_mgmt.cpp:1823): f3 and a float3[45] shift-force buffer other device code (GPU listed
     * forces) adds into; once handed out, clear_outputs zeroes them and the copy-back stage adds the kernels' forces on top */
    bool               sharedOutputs = false;
    nbb::DevBuf<float> fshiftShared;
    float*             h_fshiftShared = nullptr;
    nbb::DevBuf<float2> nbfp, nbfpComb;
    nbb::DevBuf<float>  coulombTab;
    nbb::DevBuf<float>  packedConsts; /* ParamsDev::packedConsts */
    bool           shiftVecUploaded = false;
    int            natoms = 0, natomsLocal = 0;

    /* x buffer ops (one entry per grid) */
    struct XGrid
    {
        int first = 0, n = 0;
    };
    std::vector<XGrid> xgrids;
    nbb::DevBuf<int>        atomIndex;
    nbb::DevBuf<int>        cell; /* atom -> nbat slot, for the force reduction (GpuForceReduction::Impl::cellInfo_) */
    int                     numCells = 0;

    nbb::PairList plist[2];
    bool     haveWork[2] = { false, false };

    double* h_fshift = nullptr; /* pinned staging, NBStagingData (gpu_types_common.h:142) */
    double* h_energy = nullptr;

    cudaEvent_t nonlocalDone = nullptr, localH2DDone = nullptr;

    bool                     doTiming = false;
    std::vector<nbb::TimedRegion> regions;
    nbnxm_b200_timings_t     timings{};
    bool                     pairCounting = false;
    long long                launches     = 0;

    /* copy streams and events of the chunk-pipelined step (nbnxm_b200_do_force_step_pipelined), created on first use */
    cudaStream_t             h2dStream = nullptr, d2hStream = nullptr;
    cudaStream_t             pipeKernelStream[3] = { nullptr, nullptr, nullptr }; /* the chunk kernels rotate over the local stream and these */
    int                      pipeKernelStreams   = 1;                       /* how many of them are in use (NBNXM_B200_PIPE_STREAMS - 1) */
    std::vector<cudaEvent_t> chunkH2D, chunkKernel;
    cudaEvent_t              pipeStart = nullptr, pipeD2HDone = nullptr, pipePruneDone = nullptr, pipeAllH2D = nullptr;
    /* the rolling prune as background work (NBNXM_B200_BACKGROUND_PRUNE=1; off by default): a stream of the lowest priority next
     * to kernel streams one level above it.  Measured on a B200: CTAs of the lower-priority launch are not dispatched into the
     * registers the force CTAs leave on an SM while force CTAs are pending - the prune runs in the force kernel's tail only,
     * 0.1 ... 0.3 % per step (profiles/r02aa_background_prune_ab.txt); created on first use (nbb::backgroundPruneStream) */
    cudaStream_t             pruneStream = nullptr;
    cudaEvent_t              pruneFork   = nullptr;
    int                      kernelPriority = 0; /* priority of the streams we create for force kernels */
    bool                     flatPriorities = true;
    bool                     backgroundPrune = false;
    /* optional timeline of one pipelined step (nbnxm_b200_set_pipeline_timeline): per chunk the ends of its H2D copy, the start
     * and end of its kernel and the end of its D2H copy, as timing events against tlStart */
    bool                     pipeTimeline = false;
    int                      tlChunks     = 0;
    cudaEvent_t              tlStart      = nullptr;
    std::vector<cudaEvent_t> tlEvents; /* 4 per chunk */

    /* perturbed (FEP) pair kernels, nbnxm_fep.cu: end-state atom data, pair lists, coupling parameters, dV/dlambda */
    struct FepList
    {
        nbb::DevBuf<int>           pairEntry, iinr, shift, jjnr;
        nbb::DevBuf<unsigned char> exclFep;
        int                        numI = 0, numPairs = 0;
    };
    FepList             feplist[2];
    nbb::DevBuf<float>  fepQ;      /* 2 per atom */
    nbb::DevBuf<int>    fepType;   /* 2 per atom */
    nbb::DevBuf<float>  fepLjComb; /* 4 per atom */
    nbb::DevBuf<double> fepDvdl;   /* VdW, Coulomb */
    nbb::DevBuf<double> fepForeign; /* per foreign lambda: E_lj, E_el, dV/dlambda VdW, Coulomb */
    int                 fepNumForeign = 0;
    bool                haveFep = false, haveFepAtomdata = false;
    float               fepAlphaCoul = 0, fepAlphaVdw = 0, fepSigma6WithInvalidSigma = 0, fepSigma6Minimum = 0;
    float               fepLambdaCoul = 0, fepLambdaVdw = 0;
    int                 fepLambdaPower = 1;

    nbb::HaloState* halo = nullptr;
    std::set<const void*> carveoutSet; /* kernels whose shared-memory carve-out preference was set */

    /* peer-memory halo path: j-atom arrays of the non-local launch (null: our own arrays) */
    const float4* peerXqJ   = nullptr;
    float4*       peerF4J   = nullptr;

    nbb::ParamsDev   pd{};
    nbb::AtomDataDev ad(int iloc = 0) const
    {
        nbb::AtomDataDev a;
        a.xq       = xq.p;
        a.f4       = f4.p;
        const bool peer = (iloc == 1 && peerXqJ != nullptr);
        a.xqJ       = peer ? peerXqJ : xq.p;
        a.f4J       = peer ? peerF4J : f4.p;

        a.atomType = atomType.p;
        a.ljComb   = ljComb.p;
        a.shiftVec = shiftVec.p;
        a.fshift   = fshift.p;
        a.energy   = energy.p;
        a.numTypes = numTypes;
        return a;
    }
};

/* Hooks of the peer-memory halo (nbnxm_halo.cu) for the chunk-pipelined step of a slab (nbnxm_api.cu):
 * the step counter handshake of peerForceStep, one call per edge. */
namespace nbb
{
/* the rolling prune of the local list as background work beside the force kernel (nbnxm_api.cu) */
int  backgroundPruneStream(nbnxm_b200* nb);
bool background_prune_possible(const nbnxm_b200* nb);
int  launch_background_prune(nbnxm_b200* nb, int num_parts);
int  join_background_prune(nbnxm_b200* nb);
bool peer_halo_enabled(const nbnxm_b200* nb);
/* next step number of this rank's handshake (every rank counts its steps alike) */
int peer_next_step(nbnxm_b200* nb);
/* flags[0] = n on `s`: our coordinates are in place and our accumulator is cleared for step n */
int peer_publish_ready(nbnxm_b200* nb, int n, cudaStream_t s);
/* on `s`: wait until the +x neighbour has published step n */
int peer_wait_neighbour_ready(nbnxm_b200* nb, int n, cudaStream_t s);
/* on `s`: tell the +x neighbour that our additions to its forces for step n are complete */
int peer_publish_forces_done(nbnxm_b200* nb, int n, cudaStream_t s);
/* on `s`: wait until the -x neighbour has finished adding forces to us for step n */
int peer_wait_forces_from_neighbour(nbnxm_b200* nb, int n, cudaStream_t s);
/* the home atoms the -x neighbour adds forces to */
void peer_send_range(const nbnxm_b200* nb, int* first, int* count);
} // namespace nbb

#endif
