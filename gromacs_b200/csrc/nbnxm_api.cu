/* C-ABI implementation of nbnxm_b200 (include/nbnxm_b200.h): device-memory management, stream /
 * event protocol and kernel launch logic.  It plays the role of the reference's
 * src/gromacs/nbnxm/nbnxm_gpu_data_mgmt.cpp (buffers, H2D/D2H, events),
 * src/gromacs/nbnxm/cuda/nbnxm_cuda.cu (launch logic :516-786) and
 * src/gromacs/nbnxm/gpu_common.h (task completion :141-431), behind plain C entry points.
 * There is no CPU path: every entry point needs a CUDA device.
 */
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include <vector>

#include "nbnxm_device.cuh"
#include "nbnxm_force_kernel.cuh"
#include "nbnxm_handle.cuh"

namespace nbb
{
void launch_prune(bool fresh, const AtomDataDev& ad, const ParamsDev& p, const PairlistDev& pl, int numParts, cudaStream_t stream);
int  launch_sci_sort(const PairlistDev& pl, cudaStream_t stream);
void launch_count_pairs(const PairlistDev& pl, cudaStream_t stream);
void launch_x_to_nbat_x(float4* xq, const float* x, const int* atomIndex, int first, int n, cudaStream_t s);
void launch_f4_to_f3(const float4* f4, float* f3, int first, int n, cudaStream_t s, bool accumulate = false);
void launch_reduce_f(const float4* f4, const float* rvecToAdd, float* fTotal, const int* cell, int atomStart, int n, bool accumulate,
                     cudaStream_t s);
void launch_pack_xq(const float4* xq, const int* index, int n, const float* shift, float4* out, cudaStream_t s);
void launch_copy4(const float4* in, float4* out, int n, cudaStream_t s);
void launch_unpack_add_f(float4* f4, const int* index, int n, const float4* in, cudaStream_t s);

ForceKernelPtr select_force_kernel(int elec, int vdw, bool energy, bool prune, int numTypes)
{
    switch (elec)
    {
        case NBNXM_B200_ELEC_CUT: return select_force_kernel_elec<NBNXM_B200_ELEC_CUT>(vdw, energy, prune, numTypes);
        case NBNXM_B200_ELEC_RF: return select_force_kernel_elec<NBNXM_B200_ELEC_RF>(vdw, energy, prune, numTypes);
        case NBNXM_B200_ELEC_EWALD_TAB: return select_force_kernel_elec<NBNXM_B200_ELEC_EWALD_TAB>(vdw, energy, prune, numTypes);
        case NBNXM_B200_ELEC_EWALD_TAB_TWIN: return select_force_kernel_elec<NBNXM_B200_ELEC_EWALD_TAB_TWIN>(vdw, energy, prune, numTypes);
        case NBNXM_B200_ELEC_EWALD_ANA: return select_force_kernel_elec<NBNXM_B200_ELEC_EWALD_ANA>(vdw, energy, prune, numTypes);
        case NBNXM_B200_ELEC_EWALD_ANA_TWIN: return select_force_kernel_elec<NBNXM_B200_ELEC_EWALD_ANA_TWIN>(vdw, energy, prune, numTypes);
        default: return nullptr;
    }
}
} // namespace nbb

using namespace nbb;

namespace nbb
{
thread_local char g_lastError[512] = "";

int fail(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_lastError, sizeof(g_lastError), fmt, ap);
    va_end(ap);
    return 1;
}
} // namespace nbb

namespace
{

int fillParamsDev(nbnxm_b200* nb)
{
    const nbnxm_b200_params_t& s = nb->params;
    ParamsDev&                 d = nb->pd;
    /* ElecType::None: the plain cut-off kernels with the charges switched off (include/nbnxm_b200.h) */
    d.epsfac = (s.elec_type == NBNXM_B200_ELEC_NONE) ? 0.0f : s.epsfac;
    d.c_rf = s.c_rf; d.two_k_rf = s.two_k_rf; d.ewald_beta = s.ewald_beta;
    d.sh_ewald = s.sh_ewald; d.sh_lj_ewald = s.sh_lj_ewald; d.ewaldcoeff_lj = s.ewaldcoeff_lj;
    d.rcoulomb_sq = s.rcoulomb_sq; d.rvdw_sq = s.rvdw_sq; d.rvdw_switch = s.rvdw_switch;
    d.rlist_outer_sq = s.rlist_outer_sq; d.rlist_inner_sq = s.rlist_inner_sq;
    d.disp_c2 = s.disp_c2; d.disp_c3 = s.disp_c3; d.disp_cpot = s.disp_cpot;
    d.rep_c2 = s.rep_c2; d.rep_c3 = s.rep_c3; d.rep_cpot = s.rep_cpot;
    d.sw_c3 = s.sw_c3; d.sw_c4 = s.sw_c4; d.sw_c5 = s.sw_c5;
    d.coulomb_tab_scale = s.coulomb_tab_scale;
    double pmeNumD[7] = {}, pmeDenD[5] = {};
    {
        /* pmeCorrF coefficients (src/gromacs/nbnxm/nbnxm_kernel_utils.h:216-250), powers of beta folded in */
        const double cn[7] = { -0.75225204789749321333, 0.069670166153766424023, -0.019278317264888380590,
                               0.0010054721316683106153, -0.000053401640219807709149, 1.4703624142580877519e-6,
                               -1.7357322914161492954e-8 };
        const double cd[5] = { 1.0, 0.50736591960530292870, 0.11583842382862377919, 0.014866955030185295499,
                               0.0011193462567257629232 };
        const double b = s.ewald_beta, b2 = b * b;
        double       pw = 1.0;
        for (int k = 0; k < 7; k++)
        {
            pmeNumD[k]  = cn[k] * pw * b2 * b;
            d.pmeNum[k] = static_cast<float>(pmeNumD[k]);
            if (k < 5)
            {
                pmeDenD[k]  = cd[k] * pw;
                d.pmeDen[k] = static_cast<float>(pmeDenD[k]);
            }
            pw *= b2;
        }
        /* no Ewald electrostatics: beta = 0, the packed constants are not read */
        if (pmeNumD[6] == 0.0) pmeNumD[6] = 1.0;
    }
    d.nbfp       = nb->nbfp.p;
    d.nbfpComb   = nb->nbfpComb.p;
    d.coulombTab = nb->coulombTab.p;
    {
        /* the packed kernel reads its pair-body constants from global memory (nbnxm_force_kernel_packed.cuh);
         * kernels of earlier steps may still be reading the previous values */
        float h[pcCount] = {};
        h[pcRc2] = d.rcoulomb_sq;
#ifdef NBNXM_PACKED_PLAIN_PMECORR
        for (int k = 0; k < 7; k++) h[pcNum0 + k] = d.pmeNum[k];
        for (int k = 0; k < 5; k++) h[pcDen0 + k] = d.pmeDen[k];
#else
        /* the packed kernel evaluates num / den with both polynomials divided by num[6] (nbnxm_force_kernel_packed.cuh, pair_w) */
        for (int k = 0; k < 7; k++) h[pcNum0 + k] = static_cast<float>(pmeNumD[k] / pmeNumD[6]);
        for (int k = 0; k < 5; k++) h[pcDen0 + k] = static_cast<float>(pmeDenD[k] / pmeNumD[6]);
#endif
        h[pcRvdw2] = d.rvdw_sq; h[pcBeta] = d.ewald_beta; h[pcEpsfac] = d.epsfac; h[pcRvdwSwitch] = d.rvdw_switch;
        h[pcDispC2] = d.disp_c2; h[pcDispC3] = d.disp_c3; h[pcRepC2] = d.rep_c2; h[pcRepC3] = d.rep_c3;
        h[pcDispC2Third] = d.disp_c2 * (1.0f / 3.0f); h[pcDispC3Quarter] = d.disp_c3 * 0.25f;
        h[pcRepC2Third] = d.rep_c2 * (1.0f / 3.0f); h[pcRepC3Quarter] = d.rep_c3 * 0.25f;
        h[pcDispCpot] = d.disp_cpot; h[pcRepCpot] = d.rep_cpot;
        h[pcSwC3] = d.sw_c3; h[pcSwC4] = d.sw_c4; h[pcSwC5] = d.sw_c5;
        h[pcSwC3x3] = 3.0f * d.sw_c3; h[pcSwC4x4] = 4.0f * d.sw_c4; h[pcSwC5x5] = 5.0f * d.sw_c5;
        h[pcCrf] = d.c_rf; h[pcTwoKrf] = d.two_k_rf; h[pcHalfTwoKrf] = 0.5f * d.two_k_rf; h[pcShEwald] = d.sh_ewald;
        {
            const float c2 = d.ewaldcoeff_lj * d.ewaldcoeff_lj;
            h[pcLjeCoeff2] = c2; h[pcLjeCoeff6Sixth] = c2 * c2 * c2 * c_oneSixth; h[pcShLjEwald] = d.sh_lj_ewald;
            h[pcTwoBetaOverSqrtPi] = static_cast<float>(2.0 * double(d.ewald_beta) * 0.56418958354775628695);
        }
        CU(nb->packedConsts.reserve(pcCount));
        if (nb->stream[0]) CU(cudaStreamSynchronize(nb->stream[0]));
        if (nb->stream[1]) CU(cudaStreamSynchronize(nb->stream[1]));
        CU(cudaMemcpy(nb->packedConsts.p, h, sizeof(h), cudaMemcpyHostToDevice));
        d.packedConsts = nb->packedConsts.p;
    }
    return 0;
}

/* kernel flavor that evaluates the handle's electrostatics type */
int kernelElecType(const nbnxm_b200* nb)
{
    return nb->params.elec_type == NBNXM_B200_ELEC_NONE ? int(NBNXM_B200_ELEC_CUT) : nb->params.elec_type;
}

bool usesLjComb(int vdw) { return vdw == NBNXM_B200_VDW_CUT_COMB_GEOM || vdw == NBNXM_B200_VDW_CUT_COMB_LB; }

/* getGpuAtomRange, src/gromacs/nbnxm/gpu_common_utils.h:72-91 */
int atomRange(const nbnxm_b200* nb, int aloc, int* begin, int* count)
{
    switch (aloc)
    {
        case 0: *begin = 0; *count = nb->natomsLocal; return 0;
        case 1: *begin = nb->natomsLocal; *count = nb->natoms - nb->natomsLocal; return 0;
        case 2: *begin = 0; *count = nb->natoms; return 0;
        default: return fail("invalid atom locality %d", aloc);
    }
}

void beginRegion(nbnxm_b200* nb, int kind, cudaStream_t s)
{
    if (!nb->doTiming) return;
    TimedRegion r;
    cudaEventCreate(&r.start);
    cudaEventCreate(&r.stop);
    r.kind = kind;
    cudaEventRecord(r.start, s);
    nb->regions.push_back(r);
}
void endRegion(nbnxm_b200* nb, cudaStream_t s)
{
    if (!nb->doTiming) return;
    cudaEventRecord(nb->regions.back().stop, s);
}

int collectTimings(nbnxm_b200* nb)
{
    for (TimedRegion& r : nb->regions)
    {
        CU(cudaEventSynchronize(r.stop));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, r.start, r.stop));
        nbnxm_b200_timings_t& t = nb->timings;
        if (r.kind & 16)
        {
            t.force_nonlocal_ms += ms;
            t.force_nonlocal_count++;
            r.kind &= 15;
        }
        switch (r.kind)
        {
            case 0: case 1: case 2: case 3:
                t.force_ms[(r.kind >> 1) & 1][r.kind & 1] += ms;
                t.force_count[(r.kind >> 1) & 1][r.kind & 1]++;
                break;
            case 4: t.prune_ms += ms; t.prune_count++; break;
            case 5: t.rolling_prune_ms += ms; t.rolling_prune_count++; break;
            case 6: t.xq_h2d_ms += ms; break;
            case 7: t.f_d2h_ms += ms; break;
            case 8: t.pairlist_h2d_ms += ms; break;
        }
        cudaEventDestroy(r.start);
        cudaEventDestroy(r.stop);
    }
    nb->regions.clear();
    return 0;
}

/* canSkipNonbondedWork, src/gromacs/nbnxm/gpu_common_utils.h:60 */
bool canSkipNonbondedWork(const nbnxm_b200* nb, int iloc) { return iloc == 1 && nb->plist[iloc].numSci == 0; }

} // namespace

extern "C" {

const char* nbnxm_b200_last_error(void) { return nbb::g_lastError; }

int nbnxm_b200_init(nbnxm_b200_t** out, int device, const nbnxm_b200_params_t* params, int ntypes, const float* nbfp,
                    const float* nbfp_comb, const float* coulomb_tab, int coulomb_tab_size, int local_and_nonlocal,
                    void* local_stream, void* nonlocal_stream)
{
    if (!out || !params || !nbfp || ntypes <= 0) return fail("nbnxm_b200_init: null argument");
    if (params->elec_type < 0 || params->elec_type > NBNXM_B200_ELEC_NONE || params->vdw_type < 0 || params->vdw_type > NBNXM_B200_VDW_EWALD_LB)
    {
        return fail("The requested electrostatics / VdW type (%d / %d) is not implemented in the GPU accelerated kernels",
                    params->elec_type, params->vdw_type);
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    {
        return fail("nbnxm_b200_init: no CUDA device available (this library has no CPU fallback)");
    }
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
    {
        return fail("nbnxm_b200_init: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    }
    nbnxm_b200* nb       = new nbnxm_b200();
    nb->device           = device;
    nb->numSMs           = prop.multiProcessorCount;
    nb->localAndNonlocal = local_and_nonlocal != 0;
    nb->params           = *params;
    nb->numTypes         = ntypes;
    if (local_stream)
    {
        nb->stream[0] = static_cast<cudaStream_t>(local_stream);
    }
    else
    {
        /* one level above the lowest priority, which is left to the background rolling prune */
        int lo = 0, hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        nb->kernelPriority = (lo - 1 >= hi) ? lo - 1 : lo;
        /* Priority tiers (kernel streams above the lowest priority, the force copy-down stream at the highest) only with
         * NBNXM_B200_STREAM_PRIORITIES=1, which the background rolling prune needs to mean anything.  Measured at 2 ranks on the
         * 12.3 M-atom box: with tiers the end-to-end slab step takes 9.28 ms against 7.87 ms without - the one-thread kernels
         * that publish "coordinates in place" sit on the copy stream and queue behind every pending force CTA of a higher tier
         * (profiles/r02ag_slab_e2e_chunks_priorities_ab.txt) */
        const char* sp = getenv("NBNXM_B200_STREAM_PRIORITIES");
        nb->flatPriorities = !(sp && atoi(sp) != 0);
        if (nb->flatPriorities) nb->kernelPriority = lo;
        CU(cudaStreamCreateWithPriority(&nb->stream[0], cudaStreamNonBlocking, nb->kernelPriority));
        nb->ownStream[0] = true;
    }
    {
        /* measured about neutral (profiles/r02aa_background_prune_ab.txt): off unless asked for */
        const char* bp       = getenv("NBNXM_B200_BACKGROUND_PRUNE");
        nb->backgroundPrune = (bp && atoi(bp) != 0);
    }
    if (nb->localAndNonlocal)
    {
        if (nonlocal_stream)
        {
            nb->stream[1] = static_cast<cudaStream_t>(nonlocal_stream);
        }
        else
        {
            /* non-local work is on the critical path of the halo exchange: highest priority,
             * as in the reference's DeviceStreamManager */
            int lo = 0, hi = 0;
            CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            CU(cudaStreamCreateWithPriority(&nb->stream[1], cudaStreamNonBlocking, hi));
            nb->ownStream[1] = true;
        }
    }
    else
    {
        nb->stream[1] = nb->stream[0];
    }
    CU(cudaEventCreateWithFlags(&nb->nonlocalDone, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&nb->localH2DDone, cudaEventDisableTiming));
    CU(cudaMallocHost(&nb->h_fshift, sizeof(double) * 3 * c_numShiftVectors));
    CU(cudaMallocHost(&nb->h_energy, sizeof(double) * 2));

    CU(nb->nbfp.reserve(size_t(ntypes) * ntypes));
    CU(cudaMemcpyAsync(nb->nbfp.p, nbfp, sizeof(float2) * ntypes * ntypes, cudaMemcpyHostToDevice, nb->stream[0]));
    CU(nb->nbfpComb.reserve(ntypes));
    if (nbfp_comb)
    {
        CU(cudaMemcpyAsync(nb->nbfpComb.p, nbfp_comb, sizeof(float2) * ntypes, cudaMemcpyHostToDevice, nb->stream[0]));
    }
    else
    {
        CU(cudaMemsetAsync(nb->nbfpComb.p, 0, sizeof(float2) * ntypes, nb->stream[0]));
    }
    if (coulomb_tab && coulomb_tab_size > 0)
    {
        CU(nb->coulombTab.reserve(coulomb_tab_size));
        CU(cudaMemcpyAsync(nb->coulombTab.p, coulomb_tab, sizeof(float) * coulomb_tab_size, cudaMemcpyHostToDevice, nb->stream[0]));
    }
    /* initAtomdataFirst, nbnxm_gpu_data_mgmt.cpp:322 */
    CU(nb->shiftVec.reserve(3 * c_numShiftVectors));
    CU(nb->fshift.reserve(3 * c_numShiftVectors));
    CU(nb->energy.reserve(2));
    CU(cudaMemsetAsync(nb->fshift.p, 0, sizeof(double) * 3 * c_numShiftVectors, nb->stream[0]));
    CU(cudaMemsetAsync(nb->energy.p, 0, sizeof(double) * 2, nb->stream[0]));
    for (int l = 0; l < 2; l++)
    {
        CU(nb->plist[l].pairCount.reserve(1));
        CU(cudaMemsetAsync(nb->plist[l].pairCount.p, 0, sizeof(unsigned long long), nb->stream[0]));
    }
    CU(cudaStreamSynchronize(nb->stream[0]));
    if (fillParamsDev(nb)) return 1;
    *out = nb;
    return 0;
}

int nbnxm_b200_free(nbnxm_b200_t* nb)
{
    if (!nb) return 0;
    cudaSetDevice(nb->device);
    cudaDeviceSynchronize();
    nbnxm_b200_halo_free(nb);
    collectTimings(nb);
    nb->xq.release(); nb->f4.release(); nb->f3.release(); nb->atomType.release(); nb->ljComb.release();
    nb->shiftVec.release(); nb->fshift.release(); nb->energy.release(); nb->nbfp.release();
    nb->nbfpComb.release(); nb->coulombTab.release(); nb->atomIndex.release(); nb->packedConsts.release(); nb->cell.release();
    nb->fepQ.release(); nb->fepType.release(); nb->fepLjComb.release(); nb->fepDvdl.release(); nb->fepForeign.release();
    for (nbnxm_b200::FepList& fl : nb->feplist)
    {
        fl.pairEntry.release(); fl.iinr.release(); fl.shift.release(); fl.jjnr.release(); fl.exclFep.release();
    }
    for (PairList& pl : nb->plist)
    {
        pl.sci.release(); pl.sciSorted.release(); pl.sciCount.release(); pl.sciHistogram.release();
        pl.sciOffset.release(); pl.rollingPart.release(); pl.cjPacked.release(); pl.imaskOuter.release();
        pl.excl.release(); pl.pairCount.release();
    }
    for (cudaEvent_t ev : nb->chunkH2D) cudaEventDestroy(ev);
    for (cudaEvent_t ev : nb->chunkKernel) cudaEventDestroy(ev);
    if (nb->pipeStart) cudaEventDestroy(nb->pipeStart);
    if (nb->tlStart) cudaEventDestroy(nb->tlStart);
    if (nb->pipePruneDone) cudaEventDestroy(nb->pipePruneDone);
    if (nb->pipeAllH2D) cudaEventDestroy(nb->pipeAllH2D);
    for (cudaEvent_t ev : nb->tlEvents) cudaEventDestroy(ev);
    if (nb->pipeD2HDone) cudaEventDestroy(nb->pipeD2HDone);
    if (nb->pruneFork) cudaEventDestroy(nb->pruneFork);
    if (nb->pruneStream) cudaStreamDestroy(nb->pruneStream);
    if (nb->h2dStream) cudaStreamDestroy(nb->h2dStream);
    if (nb->d2hStream) cudaStreamDestroy(nb->d2hStream);
    for (cudaStream_t ps : nb->pipeKernelStream) if (ps) cudaStreamDestroy(ps);
    if (nb->h_fshift) cudaFreeHost(nb->h_fshift);
    if (nb->h_fshiftShared) cudaFreeHost(nb->h_fshiftShared);
    nb->fshiftShared.release();
    if (nb->h_energy) cudaFreeHost(nb->h_energy);
    if (nb->nonlocalDone) cudaEventDestroy(nb->nonlocalDone);
    if (nb->localH2DDone) cudaEventDestroy(nb->localH2DDone);
    if (nb->ownStream[0]) cudaStreamDestroy(nb->stream[0]);
    if (nb->ownStream[1]) cudaStreamDestroy(nb->stream[1]);
    delete nb;
    return 0;
}

int nbnxm_b200_update_params(nbnxm_b200_t* nb, const nbnxm_b200_params_t* params, const float* coulomb_tab, int coulomb_tab_size)
{
    if (!nb || !params) return fail("nbnxm_b200_update_params: null argument");
    CU(cudaSetDevice(nb->device));
    if (params->elec_type != nb->params.elec_type && !(params->elec_type >= 2 && nb->params.elec_type >= 2))
    {
        return fail("nbnxm_b200_update_params: the electrostatics family cannot change");
    }
    nb->params = *params;
    if (coulomb_tab && coulomb_tab_size > 0)
    {
        CU(cudaStreamSynchronize(nb->stream[0]));
        CU(cudaStreamSynchronize(nb->stream[1]));
        CU(nb->coulombTab.reserve(coulomb_tab_size));
        CU(cudaMemcpyAsync(nb->coulombTab.p, coulomb_tab, sizeof(float) * coulomb_tab_size, cudaMemcpyHostToDevice, nb->stream[0]));
    }
    return fillParamsDev(nb);
}

static int initPairlist(nbnxm_b200_t* nb, int iloc, const nbnxm_b200_sci_t* sci, int nsci, const nbnxm_b200_cj_packed_t* cj_packed,
                        int ncj_packed, const nbnxm_b200_excl_t* excl, int nexcl, int na_ci, cudaMemcpyKind kind)
{
    if (!nb || iloc < 0 || iloc > 1) return fail("nbnxm_b200_init_pairlist: bad argument");
    if (nsci < 0 || ncj_packed < 0 || nexcl < 0) return fail("nbnxm_b200_init_pairlist: negative size");
    if (na_ci != c_clusterSize) return fail("nbnxm_b200_init_pairlist: cluster size %d unsupported (need 8)", na_ci);
    CU(cudaSetDevice(nb->device));
    PairList&    pl = nb->plist[iloc];
    cudaStream_t st = nb->stream[iloc];
    if (pl.naCi >= 0 && pl.naCi != na_ci)
    {
        return fail("In init_plist: the #atoms per cluster has changed (from %d to %d)", pl.naCi, na_ci);
    }
    pl.naCi = na_ci;
    /* buffers may still be in use by the previous list's kernels */
    CU(cudaStreamSynchronize(st));
    if (kind == cudaMemcpyHostToDevice) beginRegion(nb, 8, st);
    pl.numSci = nsci;
    CU(pl.sci.reserve(nsci));
    CU(pl.sciSorted.reserve(nsci));
    CU(pl.sciCount.reserve(nsci));
    CU(pl.rollingPart.reserve(nsci));
    CU(pl.sciHistogram.reserve(c_sciHistogramSize + 1));
    CU(pl.sciOffset.reserve(c_sciHistogramSize));
    CU(pl.cjPacked.reserve(ncj_packed));
    CU(pl.imaskOuter.reserve(size_t(ncj_packed) * 2));
    CU(pl.excl.reserve(nexcl > 0 ? nexcl : 1));
    if (nsci > 0)
    {
        CU(cudaMemcpyAsync(pl.sci.p, sci, sizeof(*sci) * nsci, kind, st));
        CU(cudaMemcpyAsync(pl.sciSorted.p, sci, sizeof(*sci) * nsci, kind, st));
        CU(cudaMemsetAsync(pl.rollingPart.p, 0, sizeof(int) * nsci, st));
        CU(cudaMemsetAsync(pl.sciCount.p, 0, sizeof(int) * nsci, st));
    }
    CU(cudaMemsetAsync(pl.sciHistogram.p, 0, sizeof(int) * (c_sciHistogramSize + 1), st));
    if (ncj_packed > 0)
    {
        CU(cudaMemcpyAsync(pl.cjPacked.p, cj_packed, sizeof(*cj_packed) * ncj_packed, kind, st));
        CU(cudaMemsetAsync(pl.imaskOuter.p, 0, sizeof(unsigned int) * 2 * ncj_packed, st));
    }
    if (nexcl > 0)
    {
        CU(cudaMemcpyAsync(pl.excl.p, excl, sizeof(*excl) * nexcl, kind, st));
    }
    if (kind == cudaMemcpyHostToDevice) endRegion(nb, st);
    pl.haveFreshList   = true;
    pl.didPrune        = false;
    pl.didRollingPrune = false;
    pl.rollingNumParts = 0;
    return 0;
}

int nbnxm_b200_init_pairlist(nbnxm_b200_t* nb, int iloc, const nbnxm_b200_sci_t* sci, int nsci,
                             const nbnxm_b200_cj_packed_t* cj_packed, int ncj_packed, const nbnxm_b200_excl_t* excl,
                             int nexcl, int na_ci)
{
    return initPairlist(nb, iloc, sci, nsci, cj_packed, ncj_packed, excl, nexcl, na_ci, cudaMemcpyHostToDevice);
}

int nbnxm_b200_init_pairlist_device(nbnxm_b200_t* nb, int iloc, const nbnxm_b200_sci_t* d_sci, int nsci,
                                    const nbnxm_b200_cj_packed_t* d_cj_packed, int ncj_packed, const nbnxm_b200_excl_t* d_excl,
                                    int nexcl, int na_ci)
{
    return initPairlist(nb, iloc, d_sci, nsci, d_cj_packed, ncj_packed, d_excl, nexcl, na_ci, cudaMemcpyDeviceToDevice);
}

/* buffers of gpu_init_atomdata (nbnxm_gpu_data_mgmt.cpp:1006) for natoms nbat slots, contents left to the caller */
static int reserveAtomdata(nbnxm_b200_t* nb, int natoms, int natoms_local)
{
    if (natoms % (c_clusterSize * c_superClusterSize) != 0 && natoms % c_clusterSize != 0)
    {
        return fail("nbnxm_b200_init_atomdata: natoms (%d) must be a multiple of the cluster size", natoms);
    }
    CU(cudaSetDevice(nb->device));
    cudaStream_t st = nb->stream[0];
    const bool   realloc = size_t(natoms) > nb->xq.alloc;
    if (realloc)
    {
        CU(cudaStreamSynchronize(nb->stream[0]));
        CU(cudaStreamSynchronize(nb->stream[1]));
    }
    CU(nb->xq.reserve(natoms));
    CU(nb->f4.reserve(natoms));
    CU(nb->f3.reserve(size_t(natoms) * 3));
    CU(nb->atomType.reserve(natoms));
    CU(nb->ljComb.reserve(natoms));
    if (realloc && natoms > 0)
    {
        CU(cudaMemsetAsync(nb->f4.p, 0, sizeof(float4) * nb->f4.alloc, st));
        CU(cudaMemsetAsync(nb->xq.p, 0, sizeof(float4) * nb->xq.alloc, st));
    }
    nb->natoms      = natoms;
    nb->natomsLocal = natoms_local;
    return 0;
}

int nbnxm_b200_init_atomdata(nbnxm_b200_t* nb, int natoms, int natoms_local, const int* atom_type, const float* lj_comb)
{
    if (!nb || natoms < 0 || natoms_local < 0 || natoms_local > natoms) return fail("nbnxm_b200_init_atomdata: bad argument");
    const bool comb = usesLjComb(nb->params.vdw_type);
    if (comb && !lj_comb) return fail("nbnxm_b200_init_atomdata: this VdW flavor needs lj_comb");
    if (!comb && !atom_type) return fail("nbnxm_b200_init_atomdata: this VdW flavor needs atom types");
    if (reserveAtomdata(nb, natoms, natoms_local)) return 1;
    cudaStream_t st = nb->stream[0];
    if (natoms > 0)
    {
        if (atom_type) CU(cudaMemcpyAsync(nb->atomType.p, atom_type, sizeof(int) * natoms, cudaMemcpyHostToDevice, st));
        if (lj_comb) CU(cudaMemcpyAsync(nb->ljComb.p, lj_comb, sizeof(float2) * natoms, cudaMemcpyHostToDevice, st));
    }
    return 0;
}

int nbnxm_b200_init_atomdata_device(nbnxm_b200_t* nb, int natoms, int natoms_local)
{
    if (!nb || natoms < 0 || natoms_local < 0 || natoms_local > natoms) return fail("nbnxm_b200_init_atomdata_device: bad argument");
    return reserveAtomdata(nb, natoms, natoms_local);
}

int nbnxm_b200_upload_shiftvec(nbnxm_b200_t* nb, const float* shift_vec, int dynamic_box)
{
    if (!nb || !shift_vec) return fail("nbnxm_b200_upload_shiftvec: null argument");
    CU(cudaSetDevice(nb->device));
    /* only if we have a dynamic box or it was never uploaded (nbnxm_gpu_data_mgmt.cpp:719-737) */
    if (dynamic_box || !nb->shiftVecUploaded)
    {
        CU(cudaMemcpyAsync(nb->shiftVec.p, shift_vec, sizeof(float) * 3 * c_numShiftVectors, cudaMemcpyHostToDevice, nb->stream[0]));
        nb->shiftVecUploaded = true;
    }
    return 0;
}

int nbnxm_b200_copy_xq_to_gpu(nbnxm_b200_t* nb, int aloc, const float* xq)
{
    if (!nb || !xq) return fail("nbnxm_b200_copy_xq_to_gpu: null argument");
    if (aloc != 0 && aloc != 1) return fail("nbnxm_b200_copy_xq_to_gpu: locality must be Local or NonLocal");
    CU(cudaSetDevice(nb->device));
    int begin, count;
    if (atomRange(nb, aloc, &begin, &count)) return 1;
    cudaStream_t st = nb->stream[aloc];
    /* skip the non-local copy if there is no non-local work (nbnxm_gpu_data_mgmt.cpp:1498-1516) */
    if (aloc == 1 && !nb->haveWork[1] && nb->plist[1].numSci == 0 && nb->feplist[1].numPairs == 0)
    {
        nb->plist[1].haveFreshList = false;
        return 0;
    }
    beginRegion(nb, 6, st);
    if (count > 0)
    {
        CU(cudaMemcpyAsync(nb->xq.p + begin, xq + 4 * size_t(begin), sizeof(float4) * count, cudaMemcpyHostToDevice, st));
    }
    endRegion(nb, st);
    if (aloc == 0) nbnxm_b200_insert_nonlocal_dependency(nb, 0);
    return 0;
}

int nbnxm_b200_insert_nonlocal_dependency(nbnxm_b200_t* nb, int iloc)
{
    if (!nb) return fail("null handle");
    /* nbnxmInsertNonlocalGpuDependency, nbnxm_gpu_data_mgmt.cpp:1464-1487 */
    if (nb->localAndNonlocal)
    {
        if (iloc == 0)
        {
            CU(cudaEventRecord(nb->localH2DDone, nb->stream[0]));
        }
        else
        {
            CU(cudaStreamWaitEvent(nb->stream[1], nb->localH2DDone, 0));
        }
    }
    return 0;
}

int nbnxm_b200_init_x_to_nbat_x(nbnxm_b200_t* nb, int grid, int ngrids, const int* atom_index, int natoms_nbat,
                                const int* /*cxy_na*/, const int* /*cxy_ind*/, int /*ncolumns*/, int /*num_atoms_per_cell*/,
                                int atom_offset)
{
    if (!nb || !atom_index || grid < 0 || grid >= ngrids) return fail("nbnxm_b200_init_x_to_nbat_x: bad argument");
    CU(cudaSetDevice(nb->device));
    if (grid == 0)
    {
        nb->xgrids.assign(ngrids, nbnxm_b200::XGrid());
        if (size_t(nb->natoms) > nb->atomIndex.alloc)
        {
            CU(cudaStreamSynchronize(nb->stream[0]));
            CU(cudaStreamSynchronize(nb->stream[1]));
        }
        CU(nb->atomIndex.reserve(nb->natoms));
    }
    if (atom_offset < 0 || atom_offset + natoms_nbat > nb->natoms)
    {
        return fail("nbnxm_b200_init_x_to_nbat_x: grid range [%d, %d) outside the %d nbat atoms", atom_offset, atom_offset + natoms_nbat, nb->natoms);
    }
    nb->xgrids[grid].first = atom_offset;
    nb->xgrids[grid].n     = natoms_nbat;
    /* the reference launches one thread per cell slot from the column tables; with the atom-index
     * array already holding -1 for filler slots one thread per nbat slot is equivalent */
    CU(cudaMemcpyAsync(nb->atomIndex.p + atom_offset, atom_index, sizeof(int) * natoms_nbat, cudaMemcpyHostToDevice, nb->stream[0]));
    return 0;
}

int nbnxm_b200_x_to_nbat_x(nbnxm_b200_t* nb, const float* d_x, void* x_ready_event, int aloc)
{
    if (!nb || !d_x) return fail("nbnxm_b200_x_to_nbat_x: null argument");
    if (nb->xgrids.empty()) return fail("nbnxm_b200_x_to_nbat_x: call nbnxm_b200_init_x_to_nbat_x first");
    CU(cudaSetDevice(nb->device));
    cudaStream_t st = nb->stream[aloc == 1 ? 1 : 0];
    if (x_ready_event) CU(cudaStreamWaitEvent(st, static_cast<cudaEvent_t>(x_ready_event), 0));
    /* grid 0 holds the local atoms, grids 1.. the non-local zones (nbnxm_gpu_buffer_ops.cpp:59-92) */
    const int g0 = (aloc == 1) ? 1 : 0;
    const int g1 = (aloc == 0) ? 1 : int(nb->xgrids.size());
    for (int g = g0; g < g1; g++)
    {
        launch_x_to_nbat_x(nb->xq.p, d_x, nb->atomIndex.p, nb->xgrids[g].first, nb->xgrids[g].n, st);
        nb->launches++;
    }
    CU(cudaGetLastError());
    if (aloc == 0) nbnxm_b200_insert_nonlocal_dependency(nb, 0);
    return 0;
}

int nbnxm_b200_init_reduce_f(nbnxm_b200_t* nb, const int* cell, int natoms)
{
    if (!nb || !cell || natoms < 0) return fail("nbnxm_b200_init_reduce_f: bad argument");
    CU(cudaSetDevice(nb->device));
    if (size_t(natoms) > nb->cell.alloc)
    {
        CU(cudaStreamSynchronize(nb->stream[0]));
        CU(cudaStreamSynchronize(nb->stream[1]));
    }
    CU(nb->cell.reserve(natoms > 0 ? natoms : 1));
    nb->numCells = natoms;
    if (natoms > 0) CU(cudaMemcpyAsync(nb->cell.p, cell, sizeof(int) * natoms, cudaMemcpyHostToDevice, nb->stream[0]));
    /* the caller's array may be pageable and short-lived */
    CU(cudaStreamSynchronize(nb->stream[0]));
    return 0;
}

int nbnxm_b200_reduce_f(nbnxm_b200_t* nb, float* d_f_total, const float* d_rvec_force_to_add, int atom_start, int num_atoms,
                        int accumulate, void* stream)
{
    if (!nb || !d_f_total) return fail("nbnxm_b200_reduce_f: null argument");
    if (atom_start < 0 || num_atoms < 0 || atom_start + num_atoms > nb->numCells)
    {
        return fail("nbnxm_b200_reduce_f: atom range [%d, %d) outside the %d atoms of nbnxm_b200_init_reduce_f", atom_start,
                    atom_start + num_atoms, nb->numCells);
    }
    CU(cudaSetDevice(nb->device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : nb->stream[0];
    launch_reduce_f(nb->f4.p, d_rvec_force_to_add, d_f_total, nb->cell.p, atom_start, num_atoms, accumulate != 0, st);
    if (num_atoms > 0) nb->launches++;
    CU(cudaGetLastError());
    return 0;
}

static int launchPruneOnly(nbnxm_b200_t* nb, int iloc, int num_parts, cudaStream_t other);

int nbnxm_b200_launch_kernel_pruneonly(nbnxm_b200_t* nb, int iloc, int num_parts)
{
    return launchPruneOnly(nb, iloc, num_parts, nullptr);
}

} // extern "C"

namespace nbb
{
/* creates the stream of the background rolling prune (lowest priority) and its fork event on first use */
int backgroundPruneStream(nbnxm_b200* nb)
{
    if (nb->pruneStream) return 0;
    int lo = 0, hi = 0;
    CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CU(cudaStreamCreateWithPriority(&nb->pruneStream, cudaStreamNonBlocking, lo));
    CU(cudaEventCreateWithFlags(&nb->pruneFork, cudaEventDisableTiming));
    if (!nb->pipePruneDone) CU(cudaEventCreateWithFlags(&nb->pipePruneDone, cudaEventDisableTiming));
    return 0;
}

/* The rolling prune of the local list as background work of a force step (single rank, nbnxm_b200_do_force_step): forked from
 * the local stream - coordinates in place, previous step done - onto the lowest-priority stream, launched BEFORE the force
 * kernel so that both are pending together; the local stream joins it with nbb::join_background_prune before the copy-back.
 * Masks the force kernel reads while they are updated: see nbnxm_b200_do_force_step_pipelined. */
int launch_background_prune(nbnxm_b200* nb, int num_parts)
{
    if (backgroundPruneStream(nb)) return 1;
    CU(cudaEventRecord(nb->pruneFork, nb->stream[0]));
    CU(cudaStreamWaitEvent(nb->pruneStream, nb->pruneFork, 0));
    if (launchPruneOnly(nb, 0, num_parts, nb->pruneStream)) return 1;
    CU(cudaEventRecord(nb->pipePruneDone, nb->pruneStream));
    return 0;
}
int join_background_prune(nbnxm_b200* nb)
{
    CU(cudaStreamWaitEvent(nb->stream[0], nb->pipePruneDone, 0));
    return 0;
}
bool background_prune_possible(const nbnxm_b200* nb)
{
    return nb->backgroundPrune && !nb->plist[0].haveFreshList && nb->plist[0].numSci > 0;
}
} // namespace nbb

extern "C" {

/* on `other` instead of the locality's own stream when given (the pipelined step overlaps the rolling prune with its force kernels) */
static int launchPruneOnly(nbnxm_b200_t* nb, int iloc, int num_parts, cudaStream_t other)
{
    if (!nb || iloc < 0 || iloc > 1 || num_parts < 1) return fail("nbnxm_b200_launch_kernel_pruneonly: bad argument");
    CU(cudaSetDevice(nb->device));
    PairList&    pl = nb->plist[iloc];
    cudaStream_t st = other ? other : nb->stream[iloc];
    if (pl.haveFreshList)
    {
        if (num_parts != 1) return fail("With first pruning we expect 1 part");
        pl.rollingNumParts = 0;
    }
    else
    {
        if (pl.rollingNumParts == 0)
        {
            pl.rollingNumParts = num_parts;
        }
        else if (num_parts != pl.rollingNumParts)
        {
            return fail("It is not allowed to change numParts in between list generation steps");
        }
    }
    const int numSciInPartMax = (pl.numSci + num_parts - 1) / num_parts;
    if (numSciInPartMax <= 0)
    {
        pl.haveFreshList = false;
        return 0;
    }
    beginRegion(nb, pl.haveFreshList ? 4 : 5, st);
    launch_prune(pl.haveFreshList, nb->ad(iloc), nb->pd, pl.dev(false), num_parts, st);
    nb->launches++;
    if (pl.haveFreshList)
    {
        nb->launches += launch_sci_sort(pl.dev(false), st);
        pl.haveFreshList = false;
        pl.didPrune      = true;
    }
    else
    {
        pl.didRollingPrune = true;
    }
    endRegion(nb, st);
    CU(cudaGetLastError());
    return 0;
}

int nbnxm_b200_launch_kernel(nbnxm_b200_t* nb, int iloc, int compute_energy, int compute_virial)
{
    if (!nb || iloc < 0 || iloc > 1) return fail("nbnxm_b200_launch_kernel: bad argument");
    CU(cudaSetDevice(nb->device));
    PairList&    pl = nb->plist[iloc];
    cudaStream_t st = nb->stream[iloc];
    if (canSkipNonbondedWork(nb, iloc))
    {
        pl.haveFreshList = false;
        return 0;
    }
    if (nb->params.use_dynamic_pruning && pl.haveFreshList)
    {
        if (nbnxm_b200_launch_kernel_pruneonly(nb, iloc, 1)) return 1;
    }
    if (pl.numSci == 0)
    {
        return 0;
    }
    if (!nb->shiftVecUploaded) return fail("nbnxm_b200_launch_kernel: shift vectors were never uploaded");
    const bool     doPrune = pl.haveFreshList && !pl.didPrune;
    ForceKernelPtr kernel  = select_force_kernel(kernelElecType(nb), nb->params.vdw_type, compute_energy != 0, doPrune, nb->numTypes);
    if (!kernel)
    {
        return fail("nbnxm_b200_launch_kernel: no kernel for elec_type %d vdw_type %d", nb->params.elec_type, nb->params.vdw_type);
    }
    beginRegion(nb, (doPrune ? 2 : 0) + (compute_energy ? 1 : 0) + (iloc == 1 ? 16 : 0), st);
    if (nb->pairCounting)
    {
        /* diagnostics: the pairs this launch is about to evaluate (masks as the force kernel will read them) */
        launch_count_pairs(pl.dev(true), st);
        nb->launches++;
    }
    if (nb->carveoutSet.insert(reinterpret_cast<const void*>(kernel)).second)
    {
        /* 2.5 - 6.5 KB of shared memory per 32-thread CTA and up to 21 CTAs per SM: ask for the large carve-out */
        CU(cudaFuncSetAttribute(reinterpret_cast<const void*>(kernel), cudaFuncAttributePreferredSharedMemoryCarveout,
                                cudaSharedmemCarveoutMaxShared));
    }
    /* one 32-thread CTA per sci entry */
    kernel<<<pl.numSci, 32, 0, st>>>(nb->ad(iloc), nb->pd, pl.dev(false), compute_virial != 0);
    nb->launches++;
    if (doPrune)
    {
        nb->launches += launch_sci_sort(pl.dev(false), st);
        pl.didPrune      = true;
        pl.haveFreshList = false;
    }
    endRegion(nb, st);
    CU(cudaGetLastError());
    return 0;
}

/* Force kernel over a contiguous range of the sci array in the caller's order (not the count-sorted copy): the chunk
 * launches of the pipelined step. */
static int launchKernelRange(nbnxm_b200_t* nb, int firstSci, int numSci, int compute_energy, int compute_virial, cudaStream_t st)
{
    PairList& pl = nb->plist[0];
    if (numSci <= 0) return 0;
    ForceKernelPtr kernel = select_force_kernel(kernelElecType(nb), nb->params.vdw_type, compute_energy != 0, false, nb->numTypes);
    if (!kernel) return fail("no kernel for elec_type %d vdw_type %d", nb->params.elec_type, nb->params.vdw_type);
    if (nb->carveoutSet.insert(reinterpret_cast<const void*>(kernel)).second)
    {
        CU(cudaFuncSetAttribute(reinterpret_cast<const void*>(kernel), cudaFuncAttributePreferredSharedMemoryCarveout,
                                cudaSharedmemCarveoutMaxShared));
    }
    PairlistDev d = pl.dev(false);
    d.sciSorted   = const_cast<nbnxm_b200_sci_t*>(pl.sci.p) + firstSci;
    d.numSci      = numSci;
    kernel<<<numSci, 32, 0, st>>>(nb->ad(0), nb->pd, d, compute_virial != 0);
    nb->launches++;
    CU(cudaGetLastError());
    return 0;
}

int nbnxm_b200_do_force_step_pipelined(nbnxm_b200_t* nb, int step, const nbnxm_b200_step_flags_t* fl, const float* xq_host,
                                       float* f_host, int nchunks, const int* chunk_first_atom, const int* chunk_first_sci,
                                       const unsigned int* chunk_needs)
{
    if (!nb || !fl || !xq_host || !f_host || !chunk_first_atom || !chunk_first_sci || !chunk_needs)
    {
        return fail("nbnxm_b200_do_force_step_pipelined: null argument");
    }
    if (nchunks < 1 || nchunks > 32) return fail("nbnxm_b200_do_force_step_pipelined: 1..32 chunks");
    /* a slab of a multi-GPU run: the chunks cover the home atoms and the local list; the non-local list runs on its own stream
     * against the +x neighbour's memory (peer-memory halo only) */
    const bool slab = fl->have_halo != 0;
    if (slab && (fl->have_halo != 3 || !nbb::peer_halo_enabled(nb)))
    {
        return fail("nbnxm_b200_do_force_step_pipelined: with a halo only the peer-memory transport is pipelined");
    }
    PairList& pl = nb->plist[0];
    if (chunk_first_atom[0] != 0 || chunk_first_atom[nchunks] != (slab ? nb->natomsLocal : nb->natoms) || chunk_first_sci[0] != 0
        || chunk_first_sci[nchunks] != pl.numSci)
    {
        return fail("nbnxm_b200_do_force_step_pipelined: the chunks must cover all (home) atoms and all sci entries of the local list");
    }
    if (pl.haveFreshList || pl.numSci == 0 || (slab && nb->plist[1].haveFreshList))
    {
        /* the first-pass prune of a fresh list needs every coordinate: plain sequence on search steps */
        return nbnxm_b200_do_force_step(nb, step, fl, xq_host, f_host);
    }
    CU(cudaSetDevice(nb->device));
    const int    e = fl->compute_energy, v = fl->compute_virial;
    cudaStream_t st = nb->stream[0];
    const int    peerStep = slab ? nbb::peer_next_step(nb) : 0;
    if (!nb->h2dStream)
    {
        /* the pack kernels of the force copies must not queue behind pending force CTAs: highest priority; the chunk kernels
         * like the local stream */
        int lo = 0, hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CU(cudaStreamCreateWithFlags(&nb->h2dStream, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithPriority(&nb->d2hStream, cudaStreamNonBlocking, nb->flatPriorities ? lo : hi));
        for (cudaStream_t& ps : nb->pipeKernelStream) CU(cudaStreamCreateWithPriority(&ps, cudaStreamNonBlocking, nb->kernelPriority));
        /* streams the chunk kernels rotate over: 2 (default) ... 4, NBNXM_B200_PIPE_STREAMS for A/B runs */
        const char* ns       = getenv("NBNXM_B200_PIPE_STREAMS");
        nb->pipeKernelStreams = ns ? std::max(1, std::min(3, atoi(ns) - 1)) : 1;
        CU(cudaEventCreateWithFlags(&nb->pipeStart, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&nb->pipeD2HDone, cudaEventDisableTiming));
    }
    while (int(nb->chunkH2D.size()) < nchunks)
    {
        cudaEvent_t a, b;
        CU(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        nb->chunkH2D.push_back(a);
        nb->chunkKernel.push_back(b);
    }
    const bool tl = nb->pipeTimeline;
    if (tl)
    {
        if (!nb->tlStart) CU(cudaEventCreate(&nb->tlStart));
        while (int(nb->tlEvents.size()) < 4 * nchunks)
        {
            cudaEvent_t ev;
            CU(cudaEventCreate(&ev));
            nb->tlEvents.push_back(ev);
        }
        nb->tlChunks = nchunks;
        CU(cudaEventRecord(nb->tlStart, st));
    }
    /* everything of the previous step on the local stream (kernels reading xq, copies of f) comes first */
    if (nbnxm_b200_clear_outputs(nb, v)) return 1;
    CU(cudaEventRecord(nb->pipeStart, st));
    CU(cudaStreamWaitEvent(nb->h2dStream, nb->pipeStart, 0));
    CU(cudaStreamWaitEvent(nb->d2hStream, nb->pipeStart, 0));
    /* Coordinate chunks go up in the order in which the sci chunks 0, 1, 2, ... first need them: the chunks at the far end of x
     * that the first sci chunks read through the periodic boundary follow the first few directly, so that no sci chunk has to
     * wait for the end of the upload (with the plain order 0 ... n-1 the sci chunks next to the boundary ran last and kept a
     * quarter of the force copies behind the last kernel, profiles/r02j_timeline_12m_32.jsonl) */
    int h2dOrder[32], h2dPos[32], numOrdered = 0;
    for (int c = 0; c < nchunks; c++) h2dPos[c] = -1;
    for (int k = 0; k < nchunks; k++)
        for (int c = 0; c < nchunks; c++)
            if ((chunk_needs[k] & (1u << c)) && h2dPos[c] < 0)
            {
                h2dPos[c]             = numOrdered;
                h2dOrder[numOrdered++] = c;
            }
    for (int c = 0; c < nchunks; c++)
        if (h2dPos[c] < 0)
        {
            h2dPos[c]             = numOrdered;
            h2dOrder[numOrdered++] = c;
        }
    for (int n = 0; n < nchunks; n++)
    {
        const int c     = h2dOrder[n];
        const int first = chunk_first_atom[c], count = chunk_first_atom[c + 1] - first;
        if (count > 0)
        {
            CU(cudaMemcpyAsync(nb->xq.p + first, xq_host + 4 * size_t(first), sizeof(float4) * count, cudaMemcpyHostToDevice,
                               nb->h2dStream));
        }
        CU(cudaEventRecord(nb->chunkH2D[c], nb->h2dStream));
        if (tl) CU(cudaEventRecord(nb->tlEvents[4 * c], nb->h2dStream));
    }
    /* The rolling prune of this step goes behind the last coordinate copy on the copy stream, i.e. next to the force kernels
     * instead of after them, where it would sit between the last kernel and the last force copies.  Running it while force
     * kernels read the same masks is safe: it only drops cluster pairs whose atom pairs are all beyond rlistInner >= the
     * cut-off at these very coordinates, or re-admits pairs that came inside rlistInner - a kernel that sees a mask word
     * before or after the update evaluates the same interactions within the cut-off; mask words are written whole. */
    if (slab)
    {
        /* coordinates in place and accumulator cleared (the copy stream waited for the clear): the -x neighbour's non-local
         * kernel may read them and add to our forces */
        if (nbb::peer_publish_ready(nb, peerStep, nb->h2dStream)) return 1;
        if (!nb->pipeAllH2D) CU(cudaEventCreateWithFlags(&nb->pipeAllH2D, cudaEventDisableTiming));
        CU(cudaEventRecord(nb->pipeAllH2D, nb->h2dStream));
        /* the non-local stream: our i-atoms and the neighbour's j-atoms must be there; kernel, rolling prune of the non-local
         * list on odd steps (prunekerneldispatch.cpp:123-128), then the neighbour may use its forces */
        cudaStream_t sn = nb->stream[1];
        CU(cudaStreamWaitEvent(sn, nb->pipeStart, 0));
        CU(cudaStreamWaitEvent(sn, nb->pipeAllH2D, 0));
        if (nbb::peer_wait_neighbour_ready(nb, peerStep, sn)) return 1;
        if (nbnxm_b200_launch_kernel(nb, 1, e, v)) return 1;
        if (fl->dynamic_pruning && step % 2 == 1 && nbnxm_b200_launch_kernel_pruneonly(nb, 1, fl->rolling_prune_parts)) return 1;
        if (nbb::peer_publish_forces_done(nb, peerStep, sn)) return 1;
        CU(cudaEventRecord(nb->nonlocalDone, sn));
    }
    /* single rank: rolling prune on odd steps; slab: the local list on even steps */
    const bool pruneThisStep = fl->dynamic_pruning && (slab ? step % 2 == 0 : step % 2 == 1);
    if (pruneThisStep)
    {
        if (!nb->pipePruneDone) CU(cudaEventCreateWithFlags(&nb->pipePruneDone, cudaEventDisableTiming));
        /* behind the last coordinate copy; on the background stream (lowest priority) it runs beside the force kernels in what
         * they leave of every SM, on the copy stream (NBNXM_B200_BACKGROUND_PRUNE=0) it takes the GPU for its duration */
        cudaStream_t ps = nb->h2dStream;
        if (nb->backgroundPrune)
        {
            if (nbb::backgroundPruneStream(nb)) return 1;
            ps = nb->pruneStream;
            CU(cudaEventRecord(nb->pruneFork, nb->h2dStream));
            CU(cudaStreamWaitEvent(ps, nb->pruneFork, 0));
        }
        if (launchPruneOnly(nb, 0, fl->rolling_prune_parts, ps)) return 1;
        CU(cudaEventRecord(nb->pipePruneDone, ps));
    }
    /* sci chunks in the order in which their coordinates are complete: by the last of the atom chunks they need to arrive */
    int order[32], lastNeeded[32];
    for (int k = 0; k < nchunks; k++)
    {
        order[k]      = k;
        lastNeeded[k] = 0;
        for (int c = 0; c < nchunks; c++)
            if (chunk_needs[k] & (1u << c)) lastNeeded[k] = std::max(lastNeeded[k], h2dPos[c]);
    }
    for (int a = 1; a < nchunks; a++)
        for (int b = a; b > 0 && lastNeeded[order[b]] < lastNeeded[order[b - 1]]; b--) std::swap(order[b], order[b - 1]);
    /* the chunk kernels alternate between two streams so that the tail of one overlaps the head of the next (their
     * force reductions are atomic) */
    for (int i = 0; i < nb->pipeKernelStreams; i++) CU(cudaStreamWaitEvent(nb->pipeKernelStream[i], nb->pipeStart, 0));
    for (int n = 0; n < nchunks; n++)
    {
        const int    k  = order[n];
        const int    slot = n % (nb->pipeKernelStreams + 1);
        cudaStream_t ks   = slot == 0 ? st : nb->pipeKernelStream[slot - 1];
        for (int c = 0; c < nchunks; c++)
            if (chunk_needs[k] & (1u << c)) CU(cudaStreamWaitEvent(ks, nb->chunkH2D[c], 0));
        if (tl) CU(cudaEventRecord(nb->tlEvents[4 * k + 1], ks));
        if (launchKernelRange(nb, chunk_first_sci[k], chunk_first_sci[k + 1] - chunk_first_sci[k], e, v, ks)) return 1;
        CU(cudaEventRecord(nb->chunkKernel[k], ks));
        if (tl) CU(cudaEventRecord(nb->tlEvents[4 * k + 2], ks));
    }
    /* what follows on the local stream (rolling prune, energy copies) comes after every chunk kernel */
    for (int n = 0; n < nchunks; n++)
        if (n % (nb->pipeKernelStreams + 1) != 0) CU(cudaStreamWaitEvent(st, nb->chunkKernel[order[n]], 0));
    /* forces of an atom chunk are final once every sci chunk that touches it has run: atom chunks in that order */
    int position[32], readyAt[32], chunkOrder[32];
    for (int n = 0; n < nchunks; n++) position[order[n]] = n;
    for (int c = 0; c < nchunks; c++)
    {
        chunkOrder[c] = c;
        readyAt[c]    = 0;
        for (int k = 0; k < nchunks; k++)
            if (chunk_needs[k] & (1u << c)) readyAt[c] = std::max(readyAt[c], position[k]);
    }
    int sendFirst = 0, sendCount = 0;
    if (slab)
    {
        /* chunks the -x neighbour adds forces to come last, behind its "done" flag; our own non-local kernel adds to home i-atoms
         * near the +x face: every copy waits for it (it is short and starts early) */
        nbb::peer_send_range(nb, &sendFirst, &sendCount);
        for (int c = 0; c < nchunks; c++)
            if (chunk_first_atom[c] < sendFirst + sendCount && chunk_first_atom[c + 1] > sendFirst) readyAt[c] += nchunks;
        CU(cudaStreamWaitEvent(nb->d2hStream, nb->nonlocalDone, 0));
    }
    for (int a = 1; a < nchunks; a++)
        for (int b = a; b > 0 && readyAt[chunkOrder[b]] < readyAt[chunkOrder[b - 1]]; b--) std::swap(chunkOrder[b], chunkOrder[b - 1]);
    bool waitedForNeighbour = false;
    for (int n = 0; n < nchunks; n++)
    {
        const int c = chunkOrder[n];
        if (slab && !waitedForNeighbour && readyAt[c] >= nchunks)
        {
            if (nbb::peer_wait_forces_from_neighbour(nb, peerStep, nb->d2hStream)) return 1;
            waitedForNeighbour = true;
        }
        for (int k = 0; k < nchunks; k++)
            if (chunk_needs[k] & (1u << c)) CU(cudaStreamWaitEvent(nb->d2hStream, nb->chunkKernel[k], 0));
        const int first = chunk_first_atom[c], count = chunk_first_atom[c + 1] - first;
        if (count > 0)
        {
            launch_f4_to_f3(nb->f4.p, nb->f3.p, first, count, nb->d2hStream, nb->sharedOutputs);
            nb->launches++;
            CU(cudaMemcpyAsync(f_host + 3 * size_t(first), nb->f3.p + 3 * size_t(first), sizeof(float) * 3 * count,
                               cudaMemcpyDeviceToHost, nb->d2hStream));
        }
        if (tl) CU(cudaEventRecord(nb->tlEvents[4 * c + 3], nb->d2hStream));
    }
    /* the handshake completes every step, whatever the chunks look like */
    if (slab && !waitedForNeighbour && nbb::peer_wait_forces_from_neighbour(nb, peerStep, nb->d2hStream)) return 1;
    CU(cudaEventRecord(nb->pipeD2HDone, nb->d2hStream));
    /* energies and shift forces of the non-local kernel are in the same accumulators as the local ones */
    if (slab) CU(cudaStreamWaitEvent(st, nb->nonlocalDone, 0));
    if (pruneThisStep) CU(cudaStreamWaitEvent(st, nb->pipePruneDone, 0));
    if (v) CU(cudaMemcpyAsync(nb->h_fshift, nb->fshift.p, sizeof(double) * 3 * c_numShiftVectors, cudaMemcpyDeviceToHost, st));
    if (e) CU(cudaMemcpyAsync(nb->h_energy, nb->energy.p, sizeof(double) * 2, cudaMemcpyDeviceToHost, st));
    /* gpu_wait_finish_task synchronises the local stream: make it cover the force copies */
    CU(cudaStreamWaitEvent(st, nb->pipeD2HDone, 0));
    CU(cudaGetLastError());
    return 0;
}

int nbnxm_b200_launch_cpyback(nbnxm_b200_t* nb, int aloc, float* f, int compute_energy, int compute_virial, int use_gpu_f_buffer_ops)
{
    if (!nb) return fail("null handle");
    if (aloc != 0 && aloc != 1) return fail("nbnxm_b200_launch_cpyback: locality must be Local or NonLocal");
    if (!use_gpu_f_buffer_ops && !f) return fail("nbnxm_b200_launch_cpyback: null force buffer");
    CU(cudaSetDevice(nb->device));
    const int    iloc = aloc;
    cudaStream_t st   = nb->stream[iloc];
    /* don't launch non-local copy-back if there was no non-local work to do */
    if (aloc == 1 && !nb->haveWork[1] && nb->plist[1].numSci == 0 && nb->feplist[1].numPairs == 0)
    {
        return 0;
    }
    int begin, count;
    if (atomRange(nb, aloc, &begin, &count)) return 1;
    /* the non-local kernel also writes forces of local atoms: the local copy-back has to wait for it
     * (nbnxm_gpu_data_mgmt.cpp:1313-1346) */
    if (aloc == 0 && nb->localAndNonlocal && (nb->haveWork[1] || nb->plist[1].numSci > 0 || nb->feplist[1].numPairs > 0))
    {
        CU(cudaStreamWaitEvent(st, nb->nonlocalDone, 0));
    }
    beginRegion(nb, 7, st);
    launch_f4_to_f3(nb->f4.p, nb->f3.p, begin, count, st, nb->sharedOutputs);
    nb->launches++;
    if (!use_gpu_f_buffer_ops && count > 0)
    {
        CU(cudaMemcpyAsync(f + 3 * size_t(begin), nb->f3.p + 3 * size_t(begin), sizeof(float) * 3 * count, cudaMemcpyDeviceToHost, st));
    }
    if (aloc == 1 && nb->localAndNonlocal)
    {
        CU(cudaEventRecord(nb->nonlocalDone, st));
    }
    if (aloc == 0)
    {
        if (compute_virial)
        {
            CU(cudaMemcpyAsync(nb->h_fshift, nb->fshift.p, sizeof(double) * 3 * c_numShiftVectors, cudaMemcpyDeviceToHost, st));
            if (nb->sharedOutputs)
            {
                CU(cudaMemcpyAsync(nb->h_fshiftShared, nb->fshiftShared.p, sizeof(float) * 3 * c_numShiftVectors, cudaMemcpyDeviceToHost, st));
            }
        }
        if (compute_energy)
        {
            CU(cudaMemcpyAsync(nb->h_energy, nb->energy.p, sizeof(double) * 2, cudaMemcpyDeviceToHost, st));
        }
    }
    endRegion(nb, st);
    CU(cudaGetLastError());
    return 0;
}

static int finishTask(nbnxm_b200_t* nb, int aloc, int compute_energy, int compute_virial, float* e_lj, float* e_el, float* fshift)
{
    /* gpu_reduce_staged_outputs, gpu_common.h:141-171: only the local stream carries energies */
    if (aloc == 0)
    {
        if (compute_energy)
        {
            if (e_lj) *e_lj += static_cast<float>(nb->h_energy[0]);
            if (e_el) *e_el += static_cast<float>(nb->h_energy[1]);
        }
        if (compute_virial && fshift)
        {
            for (int i = 0; i < 3 * c_numShiftVectors; i++) fshift[i] += static_cast<float>(nb->h_fshift[i]);
            if (nb->sharedOutputs)
            {
                for (int i = 0; i < 3 * c_numShiftVectors; i++) fshift[i] += nb->h_fshiftShared[i];
            }
        }
    }
    PairList& pl       = nb->plist[aloc];
    pl.haveFreshList   = false;
    pl.didPrune        = false;
    pl.didRollingPrune = false;
    return 0;
}

int nbnxm_b200_try_finish_task(nbnxm_b200_t* nb, int aloc, int compute_energy, int compute_virial, float* e_lj, float* e_el,
                               float* fshift, int* done)
{
    if (!nb || !done) return fail("nbnxm_b200_try_finish_task: null argument");
    if (aloc != 0 && aloc != 1) return fail("nbnxm_b200_try_finish_task: locality must be Local or NonLocal");
    CU(cudaSetDevice(nb->device));
    cudaError_t q = cudaStreamQuery(nb->stream[aloc]);
    if (q == cudaErrorNotReady)
    {
        *done = 0;
        return 0;
    }
    CU(q);
    *done = 1;
    return finishTask(nb, aloc, compute_energy, compute_virial, e_lj, e_el, fshift);
}

int nbnxm_b200_wait_finish_task(nbnxm_b200_t* nb, int aloc, int compute_energy, int compute_virial, float* e_lj, float* e_el, float* fshift)
{
    if (!nb) return fail("null handle");
    if (aloc != 0 && aloc != 1) return fail("nbnxm_b200_wait_finish_task: locality must be Local or NonLocal");
    CU(cudaSetDevice(nb->device));
    CU(cudaStreamSynchronize(nb->stream[aloc]));
    return finishTask(nb, aloc, compute_energy, compute_virial, e_lj, e_el, fshift);
}

int nbnxm_b200_clear_outputs(nbnxm_b200_t* nb, int compute_virial)
{
    if (!nb) return fail("null handle");
    CU(cudaSetDevice(nb->device));
    cudaStream_t st = nb->stream[0];
    if (nb->natoms > 0) CU(cudaMemsetAsync(nb->f4.p, 0, sizeof(float4) * nb->natoms, st));
    if (nb->sharedOutputs)
    {
        if (nb->natoms > 0) CU(cudaMemsetAsync(nb->f3.p, 0, sizeof(float) * 3 * size_t(nb->natoms), st));
        if (compute_virial) CU(cudaMemsetAsync(nb->fshiftShared.p, 0, sizeof(float) * 3 * c_numShiftVectors, st));
    }
    if (compute_virial)
    {
        CU(cudaMemsetAsync(nb->fshift.p, 0, sizeof(double) * 3 * c_numShiftVectors, st));
        CU(cudaMemsetAsync(nb->energy.p, 0, sizeof(double) * 2, st));
        /* dV/dlambda of the perturbed kernels is cleared with the energies (clearing of dvdlLJ / dvdlElec next to eLJ / eElec) */
        if (nb->fepDvdl.p) CU(cudaMemsetAsync(nb->fepDvdl.p, 0, sizeof(double) * 2, st));
    }
    return 0;
}

int nbnxm_b200_setup_short_range_work(nbnxm_b200_t* nb, int iloc, int have_bonded_work)
{
    if (!nb || iloc < 0 || iloc > 1) return fail("bad argument");
    nb->haveWork[iloc] = (nb->plist[iloc].numSci > 0) || (nb->feplist[iloc].numPairs > 0) || have_bonded_work;
    return 0;
}
int nbnxm_b200_have_short_range_work(const nbnxm_b200_t* nb, int iloc) { return (nb && iloc >= 0 && iloc <= 1) ? nb->haveWork[iloc] : 0; }

int nbnxm_b200_min_ci_balanced(const nbnxm_b200_t* nb)
{
    /* one warp per sci entry, up to ~24 resident warps per SM: aim at several waves of entries so that the
     * sorted (largest first) schedule can even out the tail; the list builder splits long entries to get
     * there (the reference uses 61 x #SM 64-thread blocks, cuda/nbnxm_cuda_data_mgmt.cu:95-130) */
    return nb ? 128 * nb->numSMs : 0;
}
int nbnxm_b200_is_kernel_ewald_analytical(const nbnxm_b200_t* nb)
{
    return nb && (nb->params.elec_type == NBNXM_B200_ELEC_EWALD_ANA || nb->params.elec_type == NBNXM_B200_ELEC_EWALD_ANA_TWIN);
}

int nbnxm_b200_set_timing(nbnxm_b200_t* nb, int enable)
{
    if (!nb) return fail("null handle");
    nb->doTiming = enable != 0;
    return 0;
}
int nbnxm_b200_get_timings(nbnxm_b200_t* nb, nbnxm_b200_timings_t* out)
{
    if (!nb || !out) return fail("null argument");
    CU(cudaSetDevice(nb->device));
    if (collectTimings(nb)) return 1;
    *out = nb->timings;
    return 0;
}
int nbnxm_b200_reset_timings(nbnxm_b200_t* nb)
{
    if (!nb) return fail("null handle");
    CU(cudaSetDevice(nb->device));
    if (collectTimings(nb)) return 1;
    memset(&nb->timings, 0, sizeof(nb->timings));
    return 0;
}

int nbnxm_b200_get_device_buffers(nbnxm_b200_t* nb, float** d_xq, float** d_f, int* natoms)
{
    if (!nb) return fail("null handle");
    if (d_xq) *d_xq = reinterpret_cast<float*>(nb->xq.p);
    if (d_f) *d_f = nb->f3.p;
    if (natoms) *natoms = nb->natoms;
    return 0;
}
int nbnxm_b200_set_pipeline_timeline(nbnxm_b200_t* nb, int enable)
{
    if (!nb) return fail("null handle");
    nb->pipeTimeline = enable != 0;
    return 0;
}

int nbnxm_b200_get_pipeline_timeline(nbnxm_b200_t* nb, int max_chunks, int* nchunks, float* ms)
{
    if (!nb || !nchunks || !ms) return fail("nbnxm_b200_get_pipeline_timeline: null argument");
    if (!nb->tlStart || nb->tlChunks == 0 || nb->tlChunks > max_chunks) return fail("nbnxm_b200_get_pipeline_timeline: no timeline recorded");
    CU(cudaSetDevice(nb->device));
    CU(cudaDeviceSynchronize());
    *nchunks = nb->tlChunks;
    for (int n = 0; n < 4 * nb->tlChunks; n++)
    {
        CU(cudaEventElapsedTime(ms + n, nb->tlStart, nb->tlEvents[n]));
    }
    return 0;
}

int nbnxm_b200_get_shared_outputs(nbnxm_b200_t* nb, float** d_f, float** d_fshift)
{
    if (!nb) return fail("null handle");
    CU(cudaSetDevice(nb->device));
    if (!nb->sharedOutputs)
    {
        CU(nb->fshiftShared.reserve(3 * c_numShiftVectors));
        CU(cudaMallocHost(&nb->h_fshiftShared, sizeof(float) * 3 * c_numShiftVectors));
        CU(cudaMemsetAsync(nb->fshiftShared.p, 0, sizeof(float) * 3 * c_numShiftVectors, nb->stream[0]));
        if (nb->natoms > 0) CU(cudaMemsetAsync(nb->f3.p, 0, sizeof(float) * 3 * size_t(nb->natoms), nb->stream[0]));
        nb->sharedOutputs = true;
    }
    if (d_f) *d_f = nb->f3.p;
    if (d_fshift) *d_fshift = nb->fshiftShared.p;
    return 0;
}
int nbnxm_b200_get_streams(nbnxm_b200_t* nb, void** local_stream, void** nonlocal_stream)
{
    if (!nb) return fail("null handle");
    if (local_stream) *local_stream = nb->stream[0];
    if (nonlocal_stream) *nonlocal_stream = nb->stream[1];
    return 0;
}

int nbnxm_b200_download_pairlist(nbnxm_b200_t* nb, int iloc, nbnxm_b200_cj_packed_t* cj_packed, unsigned int* imask_outer,
                                 nbnxm_b200_sci_t* sci_sorted, int* sci_count, int* rolling_part)
{
    if (!nb || iloc < 0 || iloc > 1) return fail("bad argument");
    CU(cudaSetDevice(nb->device));
    PairList& pl = nb->plist[iloc];
    CU(cudaStreamSynchronize(nb->stream[iloc]));
    if (cj_packed && pl.cjPacked.n) CU(cudaMemcpy(cj_packed, pl.cjPacked.p, sizeof(*cj_packed) * pl.cjPacked.n, cudaMemcpyDeviceToHost));
    if (imask_outer && pl.imaskOuter.n) CU(cudaMemcpy(imask_outer, pl.imaskOuter.p, sizeof(unsigned) * pl.imaskOuter.n, cudaMemcpyDeviceToHost));
    if (sci_sorted && pl.numSci) CU(cudaMemcpy(sci_sorted, pl.sciSorted.p, sizeof(*sci_sorted) * pl.numSci, cudaMemcpyDeviceToHost));
    if (sci_count && pl.numSci) CU(cudaMemcpy(sci_count, pl.sciCount.p, sizeof(int) * pl.numSci, cudaMemcpyDeviceToHost));
    if (rolling_part && pl.numSci) CU(cudaMemcpy(rolling_part, pl.rollingPart.p, sizeof(int) * pl.numSci, cudaMemcpyDeviceToHost));
    return 0;
}

int nbnxm_b200_set_pair_counting(nbnxm_b200_t* nb, int enable)
{
    if (!nb) return fail("null handle");
    CU(cudaSetDevice(nb->device));
    nb->pairCounting = enable != 0;
    for (int l = 0; l < 2; l++) CU(cudaMemsetAsync(nb->plist[l].pairCount.p, 0, sizeof(unsigned long long), nb->stream[l]));
    return 0;
}
int nbnxm_b200_get_pair_count(nbnxm_b200_t* nb, int iloc, long long* npairs)
{
    if (!nb || !npairs || iloc < 0 || iloc > 1) return fail("bad argument");
    CU(cudaSetDevice(nb->device));
    CU(cudaStreamSynchronize(nb->stream[iloc]));
    unsigned long long v = 0;
    CU(cudaMemcpy(&v, nb->plist[iloc].pairCount.p, sizeof(v), cudaMemcpyDeviceToHost));
    CU(cudaMemset(nb->plist[iloc].pairCount.p, 0, sizeof(v)));
    *npairs = static_cast<long long>(v);
    return 0;
}
long long nbnxm_b200_launch_count(const nbnxm_b200_t* nb) { return nb ? nb->launches : 0; }

int nbnxm_b200_pack_xq(nbnxm_b200_t* nb, const int* d_index, int n, const float* shift3, float* d_send, void* stream)
{
    if (!nb || (n > 0 && (!d_index || !d_send || !shift3))) return fail("nbnxm_b200_pack_xq: null argument");
    CU(cudaSetDevice(nb->device));
    launch_pack_xq(nb->xq.p, d_index, n, shift3, reinterpret_cast<float4*>(d_send), stream ? static_cast<cudaStream_t>(stream) : nb->stream[1]);
    nb->launches++;
    CU(cudaGetLastError());
    return 0;
}
int nbnxm_b200_unpack_xq(nbnxm_b200_t* nb, int first, int n, const float* d_recv, void* stream)
{
    if (!nb || first < 0 || first + n > nb->natoms) return fail("nbnxm_b200_unpack_xq: range outside atom data");
    CU(cudaSetDevice(nb->device));
    launch_copy4(reinterpret_cast<const float4*>(d_recv), nb->xq.p + first, n, stream ? static_cast<cudaStream_t>(stream) : nb->stream[1]);
    nb->launches++;
    CU(cudaGetLastError());
    return 0;
}
int nbnxm_b200_pack_f(nbnxm_b200_t* nb, int first, int n, float* d_send, void* stream)
{
    if (!nb || first < 0 || first + n > nb->natoms) return fail("nbnxm_b200_pack_f: range outside atom data");
    CU(cudaSetDevice(nb->device));
    launch_copy4(nb->f4.p + first, reinterpret_cast<float4*>(d_send), n, stream ? static_cast<cudaStream_t>(stream) : nb->stream[1]);
    nb->launches++;
    CU(cudaGetLastError());
    return 0;
}
int nbnxm_b200_unpack_add_f(nbnxm_b200_t* nb, const int* d_index, int n, const float* d_recv, void* stream)
{
    if (!nb || (n > 0 && (!d_index || !d_recv))) return fail("nbnxm_b200_unpack_add_f: null argument");
    CU(cudaSetDevice(nb->device));
    launch_unpack_add_f(nb->f4.p, d_index, n, reinterpret_cast<const float4*>(d_recv), stream ? static_cast<cudaStream_t>(stream) : nb->stream[0]);
    nb->launches++;
    CU(cudaGetLastError());
    return 0;
}

} // extern "C"
