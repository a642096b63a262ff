/* Instantiates the force kernels of one electrostatics type (compile with -DNBNXM_ELEC=<0..5>);
 * the reference builds the same flavor matrix with macros in
 * src/gromacs/nbnxm/cuda/nbnxm_cuda_kernels.cuh and looks kernels up through four function-pointer
 * tables (src/gromacs/nbnxm/cuda/nbnxm_cuda.cu:169-440). */
#include <cstdlib>

#include "nbnxm_force_kernel.cuh"
#include "nbnxm_force_kernel_packed.cuh"

#ifndef NBNXM_ELEC
#    error "compile with -DNBNXM_ELEC=<electrostatics type>"
#endif

namespace nbb
{

/* the packed-FP32 kernel where the flavor has one (no fused prune); NBNXM_B200_SCALAR_KERNEL=1
 * forces the scalar kernel, for A/B measurements */
template<int VDW, bool AVAILABLE = PackedFlavor<NBNXM_ELEC, VDW>::available>
struct PackedPick
{
    static ForceKernelPtr get(bool) { return nullptr; }
};
template<int VDW>
struct PackedPick<VDW, true>
{
    static ForceKernelPtr get(bool energy)
    {
        return energy ? nbnxm_force_kernel_packed<NBNXM_ELEC, VDW, true> : nbnxm_force_kernel_packed<NBNXM_ELEC, VDW, false>;
    }
};

template<int VDW>
static ForceKernelPtr pick(bool energy, bool prune, int numTypes)
{
    const bool scalarOnly = (std::getenv("NBNXM_B200_SCALAR_KERNEL") != nullptr);
    /* the packed kernel stages the type table of the flavors that use one in shared memory */
    const bool typesFit = !PackedFlavor<NBNXM_ELEC, VDW>::typeTable || numTypes <= c_packedMaxTypes;
    if (!prune && !scalarOnly && typesFit && PackedPick<VDW>::get(energy) != nullptr)
    {
        return PackedPick<VDW>::get(energy);
    }
    if (energy)
    {
        return prune ? nbnxm_force_kernel<NBNXM_ELEC, VDW, true, true> : nbnxm_force_kernel<NBNXM_ELEC, VDW, true, false>;
    }
    return prune ? nbnxm_force_kernel<NBNXM_ELEC, VDW, false, true> : nbnxm_force_kernel<NBNXM_ELEC, VDW, false, false>;
}

template<>
ForceKernelPtr select_force_kernel_elec<NBNXM_ELEC>(int vdw, bool energy, bool prune, int numTypes)
{
    switch (vdw)
    {
        case NBNXM_B200_VDW_CUT: return pick<NBNXM_B200_VDW_CUT>(energy, prune, numTypes);
        case NBNXM_B200_VDW_CUT_COMB_GEOM: return pick<NBNXM_B200_VDW_CUT_COMB_GEOM>(energy, prune, numTypes);
        case NBNXM_B200_VDW_CUT_COMB_LB: return pick<NBNXM_B200_VDW_CUT_COMB_LB>(energy, prune, numTypes);
        case NBNXM_B200_VDW_FSWITCH: return pick<NBNXM_B200_VDW_FSWITCH>(energy, prune, numTypes);
        case NBNXM_B200_VDW_PSWITCH: return pick<NBNXM_B200_VDW_PSWITCH>(energy, prune, numTypes);
        case NBNXM_B200_VDW_EWALD_GEOM: return pick<NBNXM_B200_VDW_EWALD_GEOM>(energy, prune, numTypes);
        case NBNXM_B200_VDW_EWALD_LB: return pick<NBNXM_B200_VDW_EWALD_LB>(energy, prune, numTypes);
        default: return nullptr;
    }
}

} // namespace nbb
