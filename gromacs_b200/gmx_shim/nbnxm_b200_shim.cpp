/* Reference-side binding of libnbnxm_b200: the free functions of the reference's NBNXM GPU-backend
 * boundary (src/gromacs/nbnxm/nbnxm_gpu.h:68-313, src/gromacs/nbnxm/gpu_data_mgmt.h:66-181) implemented on
 * top of the C ABI in include/nbnxm_b200.h.
 *
 * This file is what a maintainer adds to the reference tree INSTEAD of
 *   src/gromacs/nbnxm/nbnxm_gpu_data_mgmt.cpp, src/gromacs/nbnxm/cuda/nbnxm_cuda.cu,
 *   src/gromacs/nbnxm/cuda/nbnxm_cuda_data_mgmt.cu, src/gromacs/nbnxm/cuda/nbnxm_cuda_kernel_*.cu,
 *   src/gromacs/nbnxm/cuda/nbnxm_gpu_buffer_ops_internal.cu
 * in a GMX_GPU=CUDA build, linking libnbnxm_b200.so.  It is compiled against the reference's headers, so it
 * cannot be built without the reference tree (INTEGRATION.md has the recipe).  Everything above it - the
 * callers in nonbonded_verlet_t, init_nb_verlet and do_force - stays untouched.
 *
 * Conventions kept from the reference: these functions do not return errors; a non-zero status from the C
 * ABI becomes gmx_fatal, as CUDA errors do there (CU_RET_ERR).
 */
#include "gmxpre.h"

#include <optional>
#include <vector>

#include "gromacs/gpu_utils/device_stream_manager.h"
#include "gromacs/gpu_utils/devicebuffer_datatype.h"
#include "gromacs/gpu_utils/gpu_utils.h"
#include "gromacs/gpu_utils/gpueventsynchronizer.h"
#include "gromacs/hardware/device_information.h"
#include "gromacs/listed_forces/listed_forces_gpu.h"
#include "gromacs/mdtypes/enerdata.h"
#include "gromacs/mdtypes/interaction_const.h"
#include "gromacs/mdtypes/locality.h"
#include "gromacs/mdtypes/simulation_workload.h"
#include "gromacs/nbnxm/atomdata.h"
#include "gromacs/nbnxm/atompairlist.h"
#include "gromacs/nbnxm/gpu_data_mgmt.h"
#include "gromacs/nbnxm/gpu_types_common.h"
#include "gromacs/nbnxm/gridset.h"
#include "gromacs/nbnxm/nbnxm.h"
#include "gromacs/nbnxm/nbnxm_enums.h"
#include "gromacs/nbnxm/nbnxm_gpu.h"
#include "gromacs/nbnxm/pairlist.h"
#include "gromacs/nbnxm/pairlistparams.h"
#include "gromacs/nbnxm/pairlistsets.h"
#include "gromacs/nbnxm/grid.h"
#include "gromacs/tables/forcetable.h"
#include "gromacs/timing/gpu_timing.h"
#include "gromacs/timing/wallcycle.h"
#include "gromacs/utility/fatalerror.h"

#include "gromacs/utility/listoflists.h"

#include "nbnxm_b200.h"
#include "nbnxm_b200_search.h"

namespace gmx
{

/* The reference's callers only ever hold a pointer to this type (nonbonded_verlet_t::gpuNbv_,
 * src/gromacs/nbnxm/nbnxm.h:501); its definition belongs to the backend. */
struct NbnxmGpu
{
    nbnxm_b200_t*             handle = nullptr;
    bool                      useLjCombRule = false;
    /* the reference's streams the library runs on (owned by the DeviceStreamManager) */
    const DeviceStream* deviceStreams[2] = { nullptr, nullptr };
    gmx_wallclock_gpu_nbnxm_t timings;
    /* the search step on the device (nbnxm_b200_gpu_search_*), created on first use */
    nbnxm_b200_gpu_search_t* search = nullptr;
    /* perturbed (FEP) pair kernels in use (copy_gpu_fepparams) */
    bool haveFep = false;
    /* coupling parameters of the foreign-lambda energies: entry 0 is the current lambda, then all_lambda
     * (nbfe_foreign_cuda_kernel.cuh:218-229) */
    std::vector<float> foreignLambdaCoul, foreignLambdaVdw;
    bool               foreignLaunched = false;
    /* what gpuGetNBAtomData hands to the GPU listed forces: views of the library's buffers */
    NBAtomDataGpu atdat{};
};

namespace
{

void check(int status, const char* what)
{
    if (status != 0)
    {
        gmx_fatal(FARGS, "%s failed: %s", what, nbnxm_b200_last_error());
    }
}

int toInt(InteractionLocality l)
{
    return l == InteractionLocality::Local ? 0 : 1;
}
int toInt(AtomLocality l)
{
    return l == AtomLocality::Local ? 0 : (l == AtomLocality::NonLocal ? 1 : 2);
}

/* nbnxmGpuPickVdwKernelType + nbnxmGpuPickElectrostaticsKernelType + set_cutoff_parameters,
 * src/gromacs/nbnxm/nbnxm_gpu_data_mgmt.cpp:168-240, 368-460.  Analytical Ewald is the reference's choice on
 * every NVIDIA device except CC 7.0 / 8.0, so it is the choice on sm_100. */
nbnxm_b200_params_t makeParams(const interaction_const_t& ic, const PairlistParams& listParams, LJCombinationRule ljComb)
{
    nbnxm_b200_params_t p{};
    const bool          twin = (ic.coulomb.cutoff != ic.vdw.cutoff);
    if (ic.coulomb.type == CoulombInteractionType::Cut)
    {
        p.elec_type = NBNXM_B200_ELEC_CUT;
    }
    else if (ic.coulomb.type == CoulombInteractionType::Fmm && !ic.nbnxmIsDirectCoulombProvider)
    {
        p.elec_type = NBNXM_B200_ELEC_NONE;
    }
    else if (usingRF(ic.coulomb.type))
    {
        p.elec_type = NBNXM_B200_ELEC_RF;
    }
    else if (usingPme(ic.coulomb.type) || ic.coulomb.type == CoulombInteractionType::Ewald)
    {
        p.elec_type = twin ? NBNXM_B200_ELEC_EWALD_ANA_TWIN : NBNXM_B200_ELEC_EWALD_ANA;
    }
    else
    {
        gmx_fatal(FARGS, "The requested electrostatics type is not implemented in the GPU accelerated kernels");
    }
    if (ic.vdw.type == VanDerWaalsType::Cut)
    {
        switch (ic.vdw.modifier)
        {
            case InteractionModifiers::None:
            case InteractionModifiers::PotShift:
                p.vdw_type = (ljComb == LJCombinationRule::None)        ? NBNXM_B200_VDW_CUT
                             : (ljComb == LJCombinationRule::Geometric) ? NBNXM_B200_VDW_CUT_COMB_GEOM
                                                                        : NBNXM_B200_VDW_CUT_COMB_LB;
                break;
            case InteractionModifiers::ForceSwitch: p.vdw_type = NBNXM_B200_VDW_FSWITCH; break;
            case InteractionModifiers::PotSwitch: p.vdw_type = NBNXM_B200_VDW_PSWITCH; break;
            default: gmx_fatal(FARGS, "The requested VdW interaction modifier is not implemented in the GPU accelerated kernels");
        }
    }
    else if (ic.vdw.type == VanDerWaalsType::Pme)
    {
        p.vdw_type = (ic.vdw.pmeCombinationRule == LongRangeVdW::Geom) ? NBNXM_B200_VDW_EWALD_GEOM : NBNXM_B200_VDW_EWALD_LB;
    }
    else
    {
        gmx_fatal(FARGS, "The requested VdW type is not implemented in the GPU accelerated kernels");
    }
    p.epsfac              = ic.coulomb.epsfac;
    p.c_rf                = ic.coulomb.reactionFieldShift;
    p.two_k_rf            = 2.0 * ic.coulomb.reactionFieldCoefficient;
    p.ewald_beta          = ic.coulomb.ewaldCoeff;
    p.sh_ewald            = ic.coulomb.ewaldShift;
    p.sh_lj_ewald         = ic.vdw.ewaldShift;
    p.ewaldcoeff_lj       = ic.vdw.ewaldCoeff;
    p.rcoulomb_sq         = ic.coulomb.cutoff * ic.coulomb.cutoff;
    p.rvdw_sq             = ic.vdw.cutoff * ic.vdw.cutoff;
    p.rvdw_switch         = ic.vdw.switchDistance;
    p.rlist_outer_sq      = listParams.rlistOuter * listParams.rlistOuter;
    p.rlist_inner_sq      = listParams.rlistInner * listParams.rlistInner;
    p.disp_c2             = ic.vdw.dispersionShift.c2;
    p.disp_c3             = ic.vdw.dispersionShift.c3;
    p.disp_cpot           = ic.vdw.dispersionShift.cpot;
    p.rep_c2              = ic.vdw.repulsionShift.c2;
    p.rep_c3              = ic.vdw.repulsionShift.c3;
    p.rep_cpot            = ic.vdw.repulsionShift.cpot;
    p.sw_c3               = ic.vdw.switchConstants.c3;
    p.sw_c4               = ic.vdw.switchConstants.c4;
    p.sw_c5               = ic.vdw.switchConstants.c5;
    p.coulomb_tab_scale   = ic.coulombEwaldTables ? ic.coulombEwaldTables->scale : 0.0F;
    p.use_dynamic_pruning = listParams.useDynamicPruning ? 1 : 0;
    return p;
}

} // namespace

NbnxmGpu* gpu_init(const DeviceStreamManager& deviceStreamManager,
                   const interaction_const_t* ic,
                   const PairlistParams&      listParams,
                   const nbnxm_atomdata_t*    nbat,
                   bool                       bLocalAndNonlocal,
                   const std::optional<size_t> nLambda)
{
    auto*                      nb     = new NbnxmGpu();
    if (nLambda.has_value())
    {
        /* a perturbed run (nbnxm_setup.cpp:578): room for the current lambda + nLambda foreign ones; the values
         * arrive with copy_gpu_fepparams */
        nb->foreignLambdaCoul.assign(*nLambda + 1, 0.0F);
        nb->foreignLambdaVdw.assign(*nLambda + 1, 0.0F);
    }
    const auto&                params = nbat->params();
    const nbnxm_b200_params_t  p      = makeParams(*ic, listParams, params.ljCombinationRule);
    nb->useLjCombRule = (p.vdw_type == NBNXM_B200_VDW_CUT_COMB_GEOM || p.vdw_type == NBNXM_B200_VDW_CUT_COMB_LB);
    const float* tab     = ic->coulombEwaldTables ? ic->coulombEwaldTables->tableF.data() : nullptr;
    const int    tabSize = ic->coulombEwaldTables ? static_cast<int>(ic->coulombEwaldTables->tableF.size()) : 0;
    /* run on the reference's own streams so that its event protocol with PME / bonded / update work holds */
    nb->deviceStreams[0] = &deviceStreamManager.stream(DeviceStreamType::NonBondedLocal);
    nb->deviceStreams[1] = bLocalAndNonlocal ? &deviceStreamManager.stream(DeviceStreamType::NonBondedNonLocal)
                                             : nb->deviceStreams[0];
    void* localStream    = nb->deviceStreams[0]->stream();
    void* nonLocalStream = bLocalAndNonlocal ? nb->deviceStreams[1]->stream() : nullptr;
    check(nbnxm_b200_init(&nb->handle,
                          deviceStreamManager.deviceInfo().id,
                          &p,
                          params.numTypes,
                          params.nbfp.data(),
                          params.nbfp_comb.empty() ? nullptr : params.nbfp_comb.data(),
                          tab,
                          tabSize,
                          bLocalAndNonlocal ? 1 : 0,
                          localStream,
                          nonLocalStream),
          "nbnxm_b200_init");
    /* decideGpuTimingsUsage (gpu_utils/gpu_utils.cpp:63): event timing of the kernels is off unless asked for */
    check(nbnxm_b200_set_timing(nb->handle, decideGpuTimingsUsage() ? 1 : 0), "nbnxm_b200_set_timing");
    return nb;
}

void gpu_free(NbnxmGpu* nb)
{
    if (nb != nullptr && nb->search != nullptr)
    {
        nbnxm_b200_gpu_search_free(nb->search);
        nb->search = nullptr;
    }
    if (nb != nullptr)
    {
        nbnxm_b200_free(nb->handle);
        delete nb;
    }
}

void gpu_init_pairlist(NbnxmGpu* nb, const NbnxmPairlistGpu* h_nblist, InteractionLocality iloc)
{
    static_assert(sizeof(nbnxm_sci_t) == sizeof(nbnxm_b200_sci_t), "sci layout");
    static_assert(sizeof(nbnxm_cj_packed_t) == sizeof(nbnxm_b200_cj_packed_t), "cjPacked layout");
    static_assert(sizeof(nbnxm_excl_t) == sizeof(nbnxm_b200_excl_t), "excl layout");
    check(nbnxm_b200_init_pairlist(nb->handle,
                                   toInt(iloc),
                                   reinterpret_cast<const nbnxm_b200_sci_t*>(h_nblist->sci.data()),
                                   static_cast<int>(h_nblist->sci.size()),
                                   reinterpret_cast<const nbnxm_b200_cj_packed_t*>(h_nblist->cjPacked.list_.data()),
                                   static_cast<int>(h_nblist->cjPacked.size()),
                                   reinterpret_cast<const nbnxm_b200_excl_t*>(h_nblist->excl.data()),
                                   static_cast<int>(h_nblist->excl.size()),
                                   h_nblist->na_ci),
          "nbnxm_b200_init_pairlist");
}

void gpu_init_atomdata(NbnxmGpu* nb, const nbnxm_atomdata_t* nbat)
{
    const auto& params = nbat->params();
    check(nbnxm_b200_init_atomdata(nb->handle,
                                   nbat->numAtoms(),
                                   nbat->numLocalAtoms(),
                                   nb->useLjCombRule ? nullptr : params.type.data(),
                                   nb->useLjCombRule ? params.lj_comb.data() : nullptr),
          "nbnxm_b200_init_atomdata");
    /* end-state data of the perturbed kernels (NBAtomDataGpu::q4 / atomTypes4 / ljComb4), when the run has them */
    if (nb->haveFep && !params.qA.empty())
    {
        check(nbnxm_b200_init_fep_atomdata(nb->handle,
                                           params.qA.data(),
                                           params.qB.data(),
                                           nb->useLjCombRule ? nullptr : params.typeA.data(),
                                           nb->useLjCombRule ? nullptr : params.typeB.data(),
                                           nb->useLjCombRule ? params.ljCombA.data() : nullptr,
                                           nb->useLjCombRule ? params.ljCombB.data() : nullptr),
              "nbnxm_b200_init_fep_atomdata");
    }
}

void gpu_upload_shiftvec(NbnxmGpu* nb, const nbnxm_atomdata_t* nbatom)
{
    check(nbnxm_b200_upload_shiftvec(nb->handle, reinterpret_cast<const float*>(nbatom->shift_vec.data()), nbatom->bDynamicBox ? 1 : 0),
          "nbnxm_b200_upload_shiftvec");
}

void gpu_copy_xq_to_gpu(NbnxmGpu* nb, const nbnxm_atomdata_t* nbdata, AtomLocality aloc)
{
    check(nbnxm_b200_copy_xq_to_gpu(nb->handle, toInt(aloc), nbdata->x().data()), "nbnxm_b200_copy_xq_to_gpu");
}

void gpu_launch_kernel(NbnxmGpu* nb, const StepWorkload& stepWork, InteractionLocality iloc)
{
    check(nbnxm_b200_launch_kernel(nb->handle, toInt(iloc), stepWork.computeEnergy ? 1 : 0, stepWork.computeVirial ? 1 : 0),
          "nbnxm_b200_launch_kernel");
}

void gpu_launch_kernel_pruneonly(NbnxmGpu* nb, InteractionLocality iloc, int numParts)
{
    check(nbnxm_b200_launch_kernel_pruneonly(nb->handle, toInt(iloc), numParts), "nbnxm_b200_launch_kernel_pruneonly");
}

void gpu_launch_cpyback(NbnxmGpu* nb, nbnxm_atomdata_t* nbatom, const StepWorkload& stepWork, AtomLocality aloc)
{
    check(nbnxm_b200_launch_cpyback(nb->handle,
                                    toInt(aloc),
                                    nbatom->outputBuffer(0).f.data(),
                                    stepWork.computeEnergy ? 1 : 0,
                                    stepWork.computeVirial ? 1 : 0,
                                    stepWork.useGpuFBufferOps ? 1 : 0),
          "nbnxm_b200_launch_cpyback");
}

bool gpu_try_finish_task(NbnxmGpu*           nb,
                         const StepWorkload& stepWork,
                         AtomLocality        aloc,
                         real*               e_lj,
                         real*               e_el,
                         double*             dvdl_lj,
                         double*             dvdl_el,
                         ArrayRef<RVec>      shiftForces,
                         ForeignLambdaTerms* foreign_term,
                         GpuTaskCompletion   completionKind)
{
    float* fshift = shiftForces.empty() ? nullptr : reinterpret_cast<float*>(shiftForces.data());
    if (completionKind == GpuTaskCompletion::Check)
    {
        int done = 0;
        check(nbnxm_b200_try_finish_task(nb->handle, toInt(aloc), stepWork.computeEnergy, stepWork.computeVirial, e_lj, e_el, fshift, &done),
              "nbnxm_b200_try_finish_task");
        if (done == 0)
        {
            return false;
        }
    }
    else
    {
        check(nbnxm_b200_wait_finish_task(nb->handle, toInt(aloc), stepWork.computeEnergy, stepWork.computeVirial, e_lj, e_el, fshift),
              "nbnxm_b200_wait_finish_task");
    }
    /* the perturbed kernels' dV/dlambda and foreign-lambda terms are staged with the energies and added once, at
     * the local wait (gpu_reduce_staged_outputs / gpu_reduce_staged_foreign_term, gpu_common.h:141-199) */
    if (nb->haveFep && aloc == AtomLocality::Local)
    {
        if (stepWork.computeEnergy && dvdl_lj != nullptr && dvdl_el != nullptr)
        {
            float dLj = 0, dEl = 0;
            check(nbnxm_b200_get_fep_dvdl(nb->handle, &dLj, &dEl, 0), "nbnxm_b200_get_fep_dvdl");
            *dvdl_lj += dLj;
            *dvdl_el += dEl;
        }
        if (nb->foreignLaunched && foreign_term != nullptr)
        {
            const int           n = static_cast<int>(nb->foreignLambdaCoul.size());
            std::vector<double> terms(4 * n);
            check(nbnxm_b200_get_fep_foreign(nb->handle, n, terms.data()), "nbnxm_b200_get_fep_foreign");
            for (int idx = 0; idx < n; idx++)
            {
                foreign_term->accumulate(idx, FreeEnergyPerturbationCouplingType::Vdw, terms[4 * idx], terms[4 * idx + 2]);
                foreign_term->accumulate(idx, FreeEnergyPerturbationCouplingType::Coul, terms[4 * idx + 1], terms[4 * idx + 3]);
            }
        }
        nb->foreignLaunched = false;
    }
    return true;
}

float gpu_wait_finish_task(NbnxmGpu*           nb,
                           const StepWorkload& stepWork,
                           AtomLocality        aloc,
                           const bool          haveSoftCore,
                           gmx_enerdata_t*     enerd,
                           ArrayRef<RVec>  shiftForces,
                           gmx_wallcycle*  wcycle)
{
    auto cycleCounter = (aloc == AtomLocality::Local) ? WallCycleCounter::WaitGpuNbL : WallCycleCounter::WaitGpuNbNL;
    wallcycle_start(wcycle, cycleCounter);
    /* gpu_common.h:407-419: soft-core makes the lambda dependence non-linear */
    auto& dvdl = haveSoftCore ? enerd->dvdl_nonlin : enerd->dvdl_lin;
    gpu_try_finish_task(nb,
                        stepWork,
                        aloc,
                        enerd->grpp.energyGroupPairTerms[NonBondedEnergyTerms::LJSR].data(),
                        enerd->grpp.energyGroupPairTerms[NonBondedEnergyTerms::CoulombSR].data(),
                        &dvdl[FreeEnergyPerturbationCouplingType::Vdw],
                        &dvdl[FreeEnergyPerturbationCouplingType::Coul],
                        shiftForces,
                        &enerd->foreignLambdaTerms,
                        GpuTaskCompletion::Wait);
    return static_cast<float>(wallcycle_stop(wcycle, cycleCounter));
}

void gpu_clear_outputs(NbnxmGpu* nb, bool computeVirial)
{
    check(nbnxm_b200_clear_outputs(nb->handle, computeVirial ? 1 : 0), "nbnxm_b200_clear_outputs");
}

void nbnxmInsertNonlocalGpuDependency(NbnxmGpu* nb, InteractionLocality interactionLocality)
{
    check(nbnxm_b200_insert_nonlocal_dependency(nb->handle, toInt(interactionLocality)), "nbnxm_b200_insert_nonlocal_dependency");
}

void setupGpuShortRangeWorkLow(NbnxmGpu* nb, const ListedForcesGpu* listedForcesGpu, InteractionLocality iLocality)
{
    const bool haveBonded = (listedForcesGpu != nullptr && listedForcesGpu->haveInteractions());
    check(nbnxm_b200_setup_short_range_work(nb->handle, toInt(iLocality), haveBonded ? 1 : 0), "nbnxm_b200_setup_short_range_work");
}

bool haveGpuShortRangeWork(const NbnxmGpu* nb, InteractionLocality interactionLocality)
{
    return nbnxm_b200_have_short_range_work(nb->handle, toInt(interactionLocality)) != 0;
}

void nbnxm_gpu_init_x_to_nbat_x(const GridSet& gridSet, NbnxmGpu* gpu_nbv)
{
    const int numGrids = static_cast<int>(gridSet.grids().size());
    for (int g = 0; g < numGrids; g++)
    {
        const Grid& grid        = gridSet.grid(g);
        const int   atomOffset  = grid.firstAtomInCell(0); /* first nbat slot of this grid */
        const int   numNbatAtoms = grid.atomIndexEnd() - atomOffset;
        check(nbnxm_b200_init_x_to_nbat_x(gpu_nbv->handle,
                                          g,
                                          numGrids,
                                          gridSet.atomIndices().data() + atomOffset,
                                          numNbatAtoms,
                                          grid.numAtomsPerCell().data(),
                                          grid.cellToBin().data(),
                                          grid.numCells(),
                                          grid.numAtomsPerBin(),
                                          atomOffset),
              "nbnxm_b200_init_x_to_nbat_x");
    }
}

void nbnxm_gpu_x_to_nbat_x(NbnxmGpu* gpu_nbv, DeviceBuffer<RVec> d_x, GpuEventSynchronizer* xReadyOnDevice, AtomLocality locality)
{
    /* the library runs on the reference's own streams, so the reference's event object can enqueue the wait */
    if (xReadyOnDevice != nullptr)
    {
        xReadyOnDevice->enqueueWaitEvent(*gpu_nbv->deviceStreams[locality == AtomLocality::NonLocal ? 1 : 0]);
    }
    check(nbnxm_b200_x_to_nbat_x(gpu_nbv->handle, reinterpret_cast<const float*>(d_x), nullptr, toInt(locality)),
          "nbnxm_b200_x_to_nbat_x");
}

/* ---- the search step on the device: what nonbonded_verlet_t::putAtomsOnGrid (nbnxm.cpp:78),
 * nbnxm_atomdata_t::setAtomProperties (atomdata.cpp:1107), PairlistSets::construct (pairlist.cpp:4056),
 * gpu_init_atomdata and gpu_init_pairlist do between them, for a caller whose coordinates live in
 * StatePropagatorDataGpu::getCoordinates().  New entry points next to the reference's (nothing in the
 * reference calls them yet): nonbonded_verlet_t::putAtomsOnGrid / constructPairlist call them instead
 * of their CPU bodies when the run keeps x on the device (useGpuXBufferOps). ---- */

//! Static per-atom data of the topology: charges and types in atom order, LJ combination parameters per type, exclusions.
void nbnxm_b200_gpu_set_atoms(NbnxmGpu*                 nb,
                              ArrayRef<const real>      charges,
                              ArrayRef<const int>       atomTypes,
                              int                       numTypes,
                              ArrayRef<const real>      ljCombPerType,
                              const ListOfLists<int>&   exclusions)
{
    if (nb->search == nullptr)
    {
        check(nbnxm_b200_gpu_search_create(&nb->search, nb->handle), "nbnxm_b200_gpu_search_create");
    }
    static_assert(sizeof(real) == sizeof(float), "libnbnxm_b200 is a mixed-precision (float) backend");
    check(nbnxm_b200_gpu_search_set_atoms(nb->search,
                                          static_cast<int>(charges.size()),
                                          charges.data(),
                                          atomTypes.data(),
                                          numTypes,
                                          ljCombPerType.empty() ? nullptr : ljCombPerType.data(),
                                          exclusions.listRangesView().data(),
                                          exclusions.elementsView().data()),
          "nbnxm_b200_gpu_search_set_atoms");
}

//! putAtomsOnGrid + setAtomProperties + gpu_init_atomdata for a rectangular unit cell, from the device coordinate buffer.
void nbnxm_b200_gpu_put_atoms_on_grid(NbnxmGpu* nb, const matrix box, DeviceBuffer<RVec> d_x, GpuEventSynchronizer* xReadyOnDevice)
{
    GMX_RELEASE_ASSERT(nb->search != nullptr, "nbnxm_b200_gpu_set_atoms has to be called first");
    GMX_RELEASE_ASSERT(box[YY][XX] == 0 && box[ZZ][XX] == 0 && box[ZZ][YY] == 0,
                       "the device gridder handles rectangular unit cells");
    if (xReadyOnDevice != nullptr)
    {
        xReadyOnDevice->enqueueWaitEvent(*nb->deviceStreams[0]);
    }
    const float boxDiag[3] = { box[XX][XX], box[YY][YY], box[ZZ][ZZ] };
    check(nbnxm_b200_gpu_search_put_atoms_on_grid(
                  nb->search, boxDiag, 1, reinterpret_cast<const float*>(d_x), nullptr, nullptr, nullptr, nullptr, nullptr),
          "nbnxm_b200_gpu_search_put_atoms_on_grid");
}

//! PairlistSets::construct + gpu_init_pairlist for the local list of a single-domain run.
void nbnxm_b200_gpu_construct_pairlist(NbnxmGpu* nb, InteractionLocality iloc, real rlist)
{
    GMX_RELEASE_ASSERT(nb->search != nullptr, "nbnxm_b200_gpu_put_atoms_on_grid has to be called first");
    /* all bins against all bins (-1: up to the last bin), one zone */
    check(nbnxm_b200_gpu_search_build(nb->search, toInt(iloc), rlist, nbnxm_b200_min_ci_balanced(nb->handle), 0, -1, 0, -1, 0, 0),
          "nbnxm_b200_gpu_search_build");
}

//! GpuForceReduction::execute for the nonbonded forces: f_total[a] (+)= f_nbat[cell[a]] (+ f_other[a]); the atom -> slot
//! map is the one nbnxm_b200_gpu_put_atoms_on_grid left on the device (or nbnxm_b200_init_reduce_f from GridSet::cells()).
void nbnxm_b200_gpu_reduce_f(NbnxmGpu*          nb,
                             DeviceBuffer<RVec> d_fTotal,
                             DeviceBuffer<RVec> d_fToAdd,
                             int                atomStart,
                             int                numAtoms,
                             bool               accumulate)
{
    check(nbnxm_b200_reduce_f(nb->handle,
                              reinterpret_cast<float*>(d_fTotal),
                              reinterpret_cast<const float*>(d_fToAdd),
                              atomStart,
                              numAtoms,
                              accumulate ? 1 : 0,
                              nullptr),
          "nbnxm_b200_reduce_f");
}

int gpu_min_ci_balanced(NbnxmGpu* nb)
{
    return nb != nullptr ? nbnxm_b200_min_ci_balanced(nb->handle) : 0;
}

bool gpu_is_kernel_ewald_analytical(const NbnxmGpu* nb)
{
    return nbnxm_b200_is_kernel_ewald_analytical(nb->handle) != 0;
}

DeviceBuffer<RVec> gpu_get_f(NbnxmGpu* nb)
{
    float* d_f = nullptr;
    nbnxm_b200_get_device_buffers(nb->handle, nullptr, &d_f, nullptr);
    return reinterpret_cast<DeviceBuffer<RVec>>(d_f);
}

NBAtomDataGpu* gpuGetNBAtomData(NbnxmGpu* nb)
{
    /* the GPU listed forces read xq and add into f and fShift (listed_forces_gpu_impl: d_xq_, d_f_, d_fShift_) */
    float *d_xq = nullptr, *d_f = nullptr, *d_fshift = nullptr;
    int    natoms = 0;
    check(nbnxm_b200_get_device_buffers(nb->handle, &d_xq, nullptr, &natoms), "nbnxm_b200_get_device_buffers");
    check(nbnxm_b200_get_shared_outputs(nb->handle, &d_f, &d_fshift), "nbnxm_b200_get_shared_outputs");
    nb->atdat.numAtoms      = natoms;
    nb->atdat.numAtomsAlloc = natoms;
    nb->atdat.xq            = reinterpret_cast<DeviceBuffer<Float4>>(d_xq);
    nb->atdat.f             = reinterpret_cast<DeviceBuffer<Float3>>(d_f);
    nb->atdat.fShift        = reinterpret_cast<DeviceBuffer<Float3>>(d_fshift);
    return &nb->atdat;
}

gmx_wallclock_gpu_nbnxm_t* gpu_get_timings(NbnxmGpu* nb)
{
    if (nb == nullptr)
    {
        return nullptr;
    }
    nbnxm_b200_timings_t t;
    check(nbnxm_b200_get_timings(nb->handle, &t), "nbnxm_b200_get_timings");
    for (int prune = 0; prune < 2; prune++)
    {
        for (int energy = 0; energy < 2; energy++)
        {
            nb->timings.ktime[prune][energy].t = t.force_ms[prune][energy];
            nb->timings.ktime[prune][energy].c = t.force_count[prune][energy];
        }
    }
    nb->timings.pruneTime.t        = t.prune_ms;
    nb->timings.pruneTime.c        = t.prune_count;
    nb->timings.dynamicPruneTime.t = t.rolling_prune_ms;
    nb->timings.dynamicPruneTime.c = t.rolling_prune_count;
    nb->timings.nb_h2d_t           = t.xq_h2d_ms;
    nb->timings.nb_d2h_t           = t.f_d2h_ms;
    nb->timings.pl_h2d_t           = t.pairlist_h2d_ms;
    return &nb->timings;
}

void gpu_reset_timings(nonbonded_verlet_t* nbv)
{
    if (nbv != nullptr && nbv->gpuNbv() != nullptr)
    {
        check(nbnxm_b200_reset_timings(nbv->gpuNbv()->handle), "nbnxm_b200_reset_timings");
    }
}

void gpu_pme_loadbal_update_param(nonbonded_verlet_t* nbv, const interaction_const_t& ic)
{
    if (nbv == nullptr || !nbv->useGpu())
    {
        return;
    }
    NbnxmGpu*                 nb = nbv->gpuNbv();
    const nbnxm_b200_params_t p  = makeParams(ic, nbv->pairlistSets().params(), nbv->nbat().params().ljCombinationRule);
    const float* tab     = ic.coulombEwaldTables ? ic.coulombEwaldTables->tableF.data() : nullptr;
    const int    tabSize = ic.coulombEwaldTables ? static_cast<int>(ic.coulombEwaldTables->tableF.size()) : 0;
    check(nbnxm_b200_update_params(nb->handle, &p, tab, tabSize), "nbnxm_b200_update_params");
}

/* ---- perturbed (free-energy) pair kernels: gpu_data_mgmt.h:75, :106, nbnxm_gpu.h:115 ---- */

void copy_gpu_fepparams(NbnxmGpu*   nb,
                        bool        bFepGpuNonBonded,
                        float       alphaCoul,
                        float       alphaVdw,
                        int         lambdaPower,
                        float       sigma6WithInvalidSigma,
                        float       sigma6Minimum,
                        float       lambdaCoul,
                        float       lambdaVdw,
                        int         nLambda,
                        const EnumerationArray<FreeEnergyPerturbationCouplingType, std::vector<double>>& all_lambda)
{
    nb->haveFep = bFepGpuNonBonded;
    if (nLambda > 0)
    {
        nb->foreignLambdaCoul.assign(nLambda + 1, lambdaCoul);
        nb->foreignLambdaVdw.assign(nLambda + 1, lambdaVdw);
        for (int i = 0; i < nLambda; i++)
        {
            nb->foreignLambdaCoul[i + 1] = static_cast<float>(all_lambda[FreeEnergyPerturbationCouplingType::Coul][i]);
            nb->foreignLambdaVdw[i + 1]  = static_cast<float>(all_lambda[FreeEnergyPerturbationCouplingType::Vdw][i]);
        }
    }
    else
    {
        nb->foreignLambdaCoul.clear();
        nb->foreignLambdaVdw.clear();
    }
    check(nbnxm_b200_copy_fepparams(nb->handle,
                                    bFepGpuNonBonded ? 1 : 0,
                                    alphaCoul,
                                    alphaVdw,
                                    lambdaPower,
                                    sigma6WithInvalidSigma,
                                    sigma6Minimum,
                                    lambdaCoul,
                                    lambdaVdw),
          "nbnxm_b200_copy_fepparams");
}

void gpu_init_feppairlist(NbnxmGpu* nb, const AtomPairlist& h_feplist, InteractionLocality iloc, ArrayRef<const int> atomIndices)
{
    /* the list holds atom indices; the kernels work in nbat order (nbnxm_gpu_data_mgmt.cpp:887-915) */
    std::vector<int> slotOfAtom(atomIndices.size(), 0);
    for (int i = 0; i < atomIndices.ssize(); i++)
    {
        if (atomIndices[i] >= 0 && atomIndices[i] < atomIndices.ssize())
        {
            slotOfAtom[atomIndices[i]] = i;
        }
    }
    ArrayRef<const AtomPairlist::IEntry> iList = h_feplist.iList();
    ArrayRef<const AtomPairlist::JEntry> jList = h_feplist.flatJList();
    std::vector<int>                     iinr(iList.size()), shift(iList.size()), jindex(iList.size() + 1, 0), jjnr(jList.size());
    std::vector<unsigned char>           interacts(jList.size());
    for (int i = 0; i < iList.ssize(); i++)
    {
        iinr[i]       = slotOfAtom[iList[i].atom];
        shift[i]      = iList[i].shiftIndex;
        jindex[i + 1] = jindex[i] + static_cast<int>(h_feplist.jList(i).size());
    }
    for (int j = 0; j < jList.ssize(); j++)
    {
        jjnr[j]      = slotOfAtom[jList[j].atom];
        interacts[j] = jList[j].interacts ? 1 : 0;
    }
    check(nbnxm_b200_init_feppairlist(
                  nb->handle, toInt(iloc), static_cast<int>(iinr.size()), iinr.data(), jindex.data(), jjnr.data(), shift.data(), interacts.data()),
          "nbnxm_b200_init_feppairlist");
}

void gpu_launch_free_energy_kernel(NbnxmGpu* nb, const SimulationWorkload& simulationWork, const StepWorkload& stepWork, InteractionLocality iloc)
{
    check(nbnxm_b200_launch_free_energy_kernel(nb->handle, toInt(iloc), stepWork.computeEnergy ? 1 : 0, stepWork.computeVirial ? 1 : 0),
          "nbnxm_b200_launch_free_energy_kernel");
    /* the foreign-lambda launch (nbfe_cuda.cu:358-400): energies and dV/dlambda at the current and all foreign lambdas */
    if (simulationWork.useGpuForeignNonbondedFE && stepWork.computeDhdl && !nb->foreignLambdaCoul.empty())
    {
        check(nbnxm_b200_launch_foreign_energy_kernel(nb->handle,
                                                      toInt(iloc),
                                                      static_cast<int>(nb->foreignLambdaCoul.size()),
                                                      nb->foreignLambdaCoul.data(),
                                                      nb->foreignLambdaVdw.data()),
              "nbnxm_b200_launch_foreign_energy_kernel");
        nb->foreignLaunched = true;
    }
}

} // namespace gmx
