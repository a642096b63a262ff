/* Device-side data layout and small PTX helpers shared by the nbnxm_b200 kernels.
 *
 * Layout in HBM (one set per handle):
 *   xq        float4[natoms]   x,y,z,q in nbat (grid) order       (NBAtomDataGpu::xq, gpu_types_common.h:176)
 *   f4        float4[natoms]   force accumulator, w unused, 16-byte aligned so i/j force
 *                              reductions are one red.global.add.v4.f32 per atom
 *   f3        float [natoms*3] packed forces handed to the caller  (NBAtomDataGpu::f)
 *   atomType  int   [natoms]   or  ljComb float2[natoms]           (NBAtomDataGpu::atomTypes / ljComb)
 *   shiftVec  float [45*3]
 *   fshift    double[45*3], energy double[2] (eLJ, eElec): accumulated in double so that the totals
 *                              stay within 1e-6 relative at 10^7 atoms
 *   per list: sci, sciSorted, sciCount, cjPacked, imaskOuter[2*ncjPacked], excl, rollingPart
 *                              (GpuPairlist, gpu_types_common.h:413)
 */
#ifndef NBNXM_B200_DEVICE_CUH
#define NBNXM_B200_DEVICE_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nbnxm_b200.h"

namespace nbb
{

constexpr int   c_clusterSize        = 8;  // atoms per cluster            (nbnxm_enums.h:176)
constexpr int   c_superClusterSize   = 8;  // clusters per super-cluster   (nbnxm_enums.h:246)
constexpr int   c_jGroupSize         = 4;  // j-clusters per cjPacked      (nbnxm_enums.h:289)
constexpr int   c_centralShiftIndex  = 22; // pbcutil/ishift.h:43-56
constexpr int   c_numShiftVectors    = 45;
constexpr int   c_sciHistogramSize   = 8192;       // gpu_types_common.h:83
constexpr float c_minDistanceSquared = 3.82e-07f;  // pairlist.h:154
constexpr float c_oneSixth           = 0.16666667f;
constexpr float c_oneTwelfth         = 0.08333333f;

struct AtomDataDev
{
    const float4* xq;
    float4*       f4;
    /* j-atoms: the same arrays, except in the non-local launch of the peer-memory halo path, where the j-atoms of the
     * list are the +x neighbour's home atoms, read from and reduced into its memory over NVLink (nbnxm_halo.cu);
     * offset so that the list's j-atom indices apply */
    const float4* xqJ;
    float4*       f4J;

    const int*    atomType;
    const float2* ljComb;
    const float*  shiftVec;
    double*       fshift;
    double*       energy; // [0] = LJ, [1] = electrostatics
    int           numTypes;
};

struct ParamsDev
{
    float epsfac, c_rf, two_k_rf, ewald_beta, sh_ewald, sh_lj_ewald, ewaldcoeff_lj;
    float rcoulomb_sq, rvdw_sq, rvdw_switch, rlist_outer_sq, rlist_inner_sq;
    float disp_c2, disp_c3, disp_cpot, rep_c2, rep_c3, rep_cpot, sw_c3, sw_c4, sw_c5;
    float coulomb_tab_scale;
    /* pmeCorrF coefficients with beta folded in: num[k] = cn_k beta^(2k+3), den[k] = cd_k beta^(2k) */
    float pmeNum[7], pmeDen[5];
    const float2* nbfp;
    const float2* nbfpComb;
    const float*  coulombTab;
    /* device copy of {rcoulomb_sq, pmeNum[7], pmeDen[5]} for the packed kernel (nbnxm_force_kernel_packed.cuh) */
    const float*  packedConsts;
};

/* layout of ParamsDev::packedConsts */
enum PackedConstIndex
{
    pcRc2 = 0, pcNum0 = 1, pcDen0 = 8, pcRvdw2 = 13, pcBeta, pcEpsfac, pcRvdwSwitch, pcDispC2, pcDispC3, pcRepC2, pcRepC3,
    pcDispC2Third, pcDispC3Quarter, pcRepC2Third, pcRepC3Quarter, pcDispCpot, pcRepCpot, pcSwC3, pcSwC4, pcSwC5, pcSwC3x3,
    pcSwC4x4, pcSwC5x5, pcCrf, pcTwoKrf, pcHalfTwoKrf, pcShEwald, pcLjeCoeff2, pcLjeCoeff6Sixth, pcShLjEwald, pcTwoBetaOverSqrtPi, pcCount
};

struct PairlistDev
{
    const nbnxm_b200_sci_t*  sci;
    nbnxm_b200_sci_t*        sciSorted;
    int*                     sciCount;
    int*                     sciHistogram;
    int*                     sciOffset;
    nbnxm_b200_cj_packed_t*  cjPacked;
    unsigned int*            imaskOuter;
    const nbnxm_b200_excl_t* excl;
    int*                     rollingPart;
    unsigned long long*      pairCount; // optional diagnostics counter (may be null)
    int                      numSci;
};

__device__ __forceinline__ void red_add_v4(float4* addr, float x, float y, float z)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(x), "f"(y), "f"(z), "f"(0.0f)
                 : "memory");
}

/* the same, issued only where `doIt` is set: a predicated instruction instead of a divergent branch */
__device__ __forceinline__ void red_add_v4_if(const bool doIt, float4* addr, float x, float y, float z)
{
    asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.s32 p, %5, 0;\n\t"
            "@p red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n\t"
            "}" ::"l"(addr),
            "f"(x), "f"(y), "f"(z), "f"(0.0f), "r"(static_cast<int>(doIt))
            : "memory");
}

/* Spin until the 32-bit flag (written by another GPU with a system-scope release) reaches `value`; gives up after about
 * two seconds and raises *errorFlag instead of hanging the device. */
__device__ __forceinline__ void wait_flag_geq(const int* flag, const int value, int* errorFlag)
{
    const long long t0 = clock64();
    int             v;
    do
    {
        asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (v >= value) return;
        __nanosleep(200);
    } while (clock64() - t0 < 4000000000ll);
    if (errorFlag) atomicExch(errorFlag, 1);
}

__device__ __forceinline__ float norm2_fma(float dx, float dy, float dz)
{
    // fixed evaluation order; the prune masks are defined on exactly this expression
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

} // namespace nbb

#endif
