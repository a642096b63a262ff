/* Force kernel of the nbnxm_b200 path (all electrostatics x VdW flavors, F / F+E, optional
 * outer-list pruning).  Replaces the reference's macro-generated kernels
 * src/gromacs/nbnxm/cuda/nbnxm_cuda_kernel.cuh:138-746 (physics from
 * src/gromacs/nbnxm/nbnxm_kernel_utils.h:56-289); written from scratch for sm_100a.
 *
 * Work decomposition (differs from the reference on purpose):
 *   - one WARP per sci entry (i-super-cluster work unit), several independent warps per CTA, no
 *     block-level synchronisation;
 *   - lane = il + 8*jl: the lane owns i-atom `il` of every i-cluster and the TWO j-atoms jl and jl+4 of
 *     the current j-cluster, i.e. both halves of the reference's "cluster-pair split"
 *     (imei[0], imei[1]) are evaluated by the same warp as two independent pair streams;
 *   - i-atom coordinates/parameters are staged once per sci entry in shared memory (one 128-byte
 *     conflict-free LDS.128 per i-cluster step, shared by both pair streams);
 *   - j forces are reduced over the 8 il-lanes by shuffles and added with one
 *     red.global.add.v4.f32 per j-atom; i forces are kept in registers for the whole sci entry.
 */
#ifndef NBNXM_B200_FORCE_KERNEL_CUH
#define NBNXM_B200_FORCE_KERNEL_CUH

#include "nbnxm_device.cuh"

namespace nbb
{

constexpr int c_forceWarpsPerBlock = 4;
constexpr int c_forceThreads       = c_forceWarpsPerBlock * 32;

template<int ELEC, int VDW, bool ENERGY>
struct Flavor
{
    static constexpr bool elecCut   = (ELEC == NBNXM_B200_ELEC_CUT);
    static constexpr bool elecRF    = (ELEC == NBNXM_B200_ELEC_RF);
    static constexpr bool ewaldTab  = (ELEC == NBNXM_B200_ELEC_EWALD_TAB || ELEC == NBNXM_B200_ELEC_EWALD_TAB_TWIN);
    static constexpr bool ewaldAna  = (ELEC == NBNXM_B200_ELEC_EWALD_ANA || ELEC == NBNXM_B200_ELEC_EWALD_ANA_TWIN);
    static constexpr bool ewaldAny  = ewaldTab || ewaldAna;
    static constexpr bool vdwCutoffCheck = (ELEC == NBNXM_B200_ELEC_EWALD_TAB_TWIN || ELEC == NBNXM_B200_ELEC_EWALD_ANA_TWIN);
    static constexpr bool ljCombGeom = (VDW == NBNXM_B200_VDW_CUT_COMB_GEOM);
    static constexpr bool ljCombLB   = (VDW == NBNXM_B200_VDW_CUT_COMB_LB);
    static constexpr bool ljComb     = ljCombGeom || ljCombLB;
    static constexpr bool ljFSwitch  = (VDW == NBNXM_B200_VDW_FSWITCH);
    static constexpr bool ljPSwitch  = (VDW == NBNXM_B200_VDW_PSWITCH);
    static constexpr bool ljEwaldGeom = (VDW == NBNXM_B200_VDW_EWALD_GEOM);
    static constexpr bool ljEwaldLB   = (VDW == NBNXM_B200_VDW_EWALD_LB);
    static constexpr bool ljEwald     = ljEwaldGeom || ljEwaldLB;
    // excluded pairs inside the cut-off still get the Ewald / RF correction (nbnxm_cuda_kernel.cuh:70-79)
    static constexpr bool exclusionForces = ewaldAny || elecRF || ljEwald || (elecCut && ENERGY);
};

/* (2/sqrt(pi) z exp(-z^2) - erf(z)) / z^3 as a rational minimax approximation in z^2; the
 * coefficients are the ones the reference uses in pmeCorrF (nbnxm_kernel_utils.h:216-250). */
__device__ __forceinline__ float pme_corr_f(const float z2)
{
    float den = 0.0011193462567257629232f;
    den       = fmaf(den, z2, 0.014866955030185295499f);
    den       = fmaf(den, z2, 0.11583842382862377919f);
    den       = fmaf(den, z2, 0.50736591960530292870f);
    den       = fmaf(den, z2, 1.0f);
    float num = -1.7357322914161492954e-8f;
    num       = fmaf(num, z2, 1.4703624142580877519e-6f);
    num       = fmaf(num, z2, -0.000053401640219807709149f);
    num       = fmaf(num, z2, 0.0010054721316683106153f);
    num       = fmaf(num, z2, -0.019278317264888380590f);
    num       = fmaf(num, z2, 0.069670166153766424023f);
    num       = fmaf(num, z2, -0.75225204789749321333f);
    return __fdividef(num, den);
}

struct PairConsts
{
    float rc2, rcoulomb, rvdw2, beta, beta2, beta3, ljeCoeff2, ljeCoeff6_6;
};

/* One atom pair. Returns F/r (already zero-masked by the caller), accumulates energies.
 * c6/c12 are 6*C6, 12*C12; qq = epsfac*qi*qj; intBit = 0 for topology-excluded pairs. */
template<int ELEC, int VDW, bool ENERGY>
__device__ __forceinline__ float pair_force(const ParamsDev& p,
                                            const PairConsts& k,
                                            float             r2,
                                            const float       qq,
                                            const float       c6,
                                            const float       c12,
                                            const float       c6grid,
                                            const float       intBit,
                                            float&            eLJ,
                                            float&            eEl)
{
    using Fl = Flavor<ELEC, VDW, ENERGY>;
    r2                = fmaxf(r2, c_minDistanceSquared);
    const float invR  = rsqrtf(r2);
    const float invR2 = invR * invR;
    float       invR6 = invR2 * invR2 * invR2;
    if (Fl::exclusionForces)
    {
        invR6 *= intBit;
    }
    float fInvR = invR6 * (c12 * invR6 - c6) * invR2;
    float eLJp  = 0.0f;
    if (ENERGY || Fl::ljPSwitch)
    {
        eLJp = c12 * (invR6 * invR6 + p.rep_cpot) * c_oneTwelfth - c6 * (invR6 + p.disp_cpot) * c_oneSixth;
        if (Fl::exclusionForces)
        {
            eLJp *= intBit;
        }
    }
    if (Fl::ljFSwitch)
    {
        const float r   = r2 * invR;
        const float rsw = fmaxf(r - p.rvdw_switch, 0.0f);
        fInvR += (-c6 * (p.disp_c2 + p.disp_c3 * rsw) + c12 * (p.rep_c2 + p.rep_c3 * rsw)) * rsw * rsw * invR;
        if (ENERGY)
        {
            eLJp += (c6 * (p.disp_c2 * (1.0f / 3.0f) + p.disp_c3 * 0.25f * rsw)
                     - c12 * (p.rep_c2 * (1.0f / 3.0f) + p.rep_c3 * 0.25f * rsw))
                    * rsw * rsw * rsw;
        }
    }
    if (Fl::ljEwald)
    {
        const float invR6nm = invR2 * invR2 * invR2;
        const float cr2     = k.ljeCoeff2 * r2;
        const float expmcr2 = __expf(-cr2);
        const float poly    = 1.0f + cr2 + 0.5f * cr2 * cr2;
        fInvR += c6grid * (invR6nm - expmcr2 * (invR6nm * poly + k.ljeCoeff6_6)) * invR2;
        if (ENERGY)
        {
            eLJp += c_oneSixth * c6grid * (invR6nm * (1.0f - expmcr2 * poly) + p.sh_lj_ewald * intBit);
        }
    }
    if (Fl::ljPSwitch)
    {
        const float r   = r2 * invR;
        const float rsw = r - p.rvdw_switch;
        if (rsw > 0.0f)
        {
            const float sw  = 1.0f + (p.sw_c3 + (p.sw_c4 + p.sw_c5 * rsw) * rsw) * rsw * rsw * rsw;
            const float dsw = (3.0f * p.sw_c3 + (4.0f * p.sw_c4 + 5.0f * p.sw_c5 * rsw) * rsw) * rsw * rsw;
            fInvR           = fInvR * sw - invR * eLJp * dsw;
            eLJp *= sw;
        }
    }
    if (Fl::vdwCutoffCheck)
    {
        const float inRange = (r2 < k.rvdw2) ? 1.0f : 0.0f;
        fInvR *= inRange;
        eLJp *= inRange;
    }
    if (ENERGY)
    {
        eLJ += eLJp;
    }

    if (Fl::elecCut)
    {
        fInvR += qq * (Fl::exclusionForces ? intBit : 1.0f) * invR2 * invR;
        if (ENERGY) eEl += qq * (intBit * invR - p.c_rf);
    }
    if (Fl::elecRF)
    {
        fInvR += qq * (intBit * invR2 * invR - p.two_k_rf);
        if (ENERGY) eEl += qq * (intBit * invR + 0.5f * p.two_k_rf * r2 - p.c_rf);
    }
    if (Fl::ewaldAna)
    {
        fInvR += qq * (intBit * invR2 * invR + pme_corr_f(k.beta2 * r2) * k.beta3);
    }
    if (Fl::ewaldTab)
    {
        /* pairs beyond the cut-off are evaluated too (and masked afterwards): keep them inside the table */
        const float normalized = p.coulomb_tab_scale * fminf(r2 * invR, k.rcoulomb);
        const int   index      = static_cast<int>(normalized);
        const float fraction   = normalized - index;
        const float left       = __ldg(p.coulombTab + index);
        const float right      = __ldg(p.coulombTab + index + 1);
        fInvR += qq * (intBit * invR2 - fmaf(fraction, right - left, left)) * invR;
    }
    if (Fl::ewaldAny && ENERGY)
    {
        /* erfc keeps the relative accuracy of the (small) real-space term; excluded pairs (intBit = 0)
         * get -erf(beta r)/r = (erfc - 1)/r */
        eEl += qq * (invR * (erfcf(r2 * invR * k.beta) - (1.0f - intBit)) - intBit * p.sh_ewald);
    }
    return fInvR;
}

/* LJ parameters of one pair from the flavor's parameter source. pi/pj: per-atom float2
 * (comb-rule flavors) ; ti/tj: type indices (ti already multiplied by numTypes). */
template<int ELEC, int VDW, bool ENERGY>
__device__ __forceinline__ void lj_pair_params(const ParamsDev& p,
                                               const float2     pi,
                                               const float2     pj,
                                               const int        tiTimesN,
                                               const int        ti,
                                               const int        tj,
                                               float&           c6,
                                               float&           c12,
                                               float&           c6grid)
{
    using Fl = Flavor<ELEC, VDW, ENERGY>;
    c6grid   = 0.0f;
    if (Fl::ljCombGeom)
    {
        c6  = pi.x * pj.x;
        c12 = pi.y * pj.y;
    }
    else if (Fl::ljCombLB)
    {
        const float sigma   = pi.x + pj.x;
        const float epsilon = pi.y * pj.y;
        const float sigma2  = sigma * sigma;
        const float sigma6  = sigma2 * sigma2 * sigma2;
        c6                  = epsilon * sigma6;
        c12                 = c6 * sigma6;
    }
    else
    {
        const float2 c = __ldg(p.nbfp + tiTimesN + tj);
        c6             = c.x;
        c12            = c.y;
        if (Fl::ljEwaldGeom)
        {
            c6grid = __ldg(p.nbfpComb + ti).x * __ldg(p.nbfpComb + tj).x;
        }
        if (Fl::ljEwaldLB)
        {
            const float2 a = __ldg(p.nbfpComb + ti), b = __ldg(p.nbfpComb + tj);
            const float  sigma = a.x + b.x, epsilon = a.y * b.y, sigma2 = sigma * sigma;
            c6grid = epsilon * sigma2 * sigma2 * sigma2;
        }
    }
}

template<int ELEC, int VDW, bool ENERGY, bool PRUNE>
__global__ void __launch_bounds__(c_forceThreads)
        nbnxm_force_kernel(const AtomDataDev ad, const ParamsDev p, const PairlistDev pl, const int calcFshift)
{
    using Fl                 = Flavor<ELEC, VDW, ENERGY>;
    constexpr unsigned c_full = 0xffffffffu;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int il   = lane & 7;
    const int jl   = lane >> 3;

    const int sciIdx = blockIdx.x * c_forceWarpsPerBlock + warp;
    if (sciIdx >= pl.numSci)
    {
        return;
    }
    /* The fused force+prune variant produces the counts the sort consumes, so it walks the
     * unsorted list (same rule as nbnxm_cuda_kernel.cuh:158-166). */
    const nbnxm_b200_sci_t s = PRUNE ? pl.sci[sciIdx] : pl.sciSorted[sciIdx];

    __shared__ float4 sm_xqi[c_forceWarpsPerBlock][64];
    __shared__ float2 sm_lji[c_forceWarpsPerBlock][64]; // comb params, or (type*numTypes, type) as int bits

    const float shx = ad.shiftVec[3 * s.shift], shy = ad.shiftVec[3 * s.shift + 1], shz = ad.shiftVec[3 * s.shift + 2];

    PairConsts k;
    k.rc2         = p.rcoulomb_sq;
    k.rcoulomb    = sqrtf(p.rcoulomb_sq);
    k.rvdw2       = p.rvdw_sq;
    k.beta        = p.ewald_beta;
    k.beta2       = p.ewald_beta * p.ewald_beta;
    k.beta3       = k.beta2 * p.ewald_beta;
    k.ljeCoeff2   = p.ewaldcoeff_lj * p.ewaldcoeff_lj;
    k.ljeCoeff6_6 = k.ljeCoeff2 * k.ljeCoeff2 * k.ljeCoeff2 * c_oneSixth;

    /* energies: float partial sums per j-cluster, double across the sci entry, so that the totals keep
     * 1e-6 relative accuracy (the reference accumulates in float, gpu_common.h:151-161) */
    double eLJ = 0.0, eEl = 0.0;
    const bool diagonalEntry = (s.shift == c_centralShiftIndex && s.cj_packed_begin < s.cj_packed_end
                                && pl.cjPacked[s.cj_packed_begin].cj[0] == s.sci * c_superClusterSize);

#pragma unroll
    for (int h = 0; h < 2; h++)
    {
        const int ai = s.sci * 64 + lane + 32 * h;
        float4    v  = ad.xq[ai];
        if (ENERGY && Fl::exclusionForces && diagonalEntry)
        {
            /* self terms, once per diagonal sci entry (nbnxm_cuda_kernel.cuh:383-417) */
            const float q2 = p.epsfac * v.w * v.w;
            if (Fl::ewaldAny) eEl -= q2 * p.ewald_beta * 0.56418958354775628695f;
            if (Fl::elecRF || Fl::elecCut) eEl -= q2 * 0.5f * p.c_rf;
            if (Fl::ljEwald)
            {
                eLJ += __ldg(p.nbfp + ad.atomType[ai] * (ad.numTypes + 1)).x * 0.5f * c_oneSixth * k.ljeCoeff6_6;
            }
        }
        v.x += shx;
        v.y += shy;
        v.z += shz;
        v.w *= p.epsfac;
        sm_xqi[warp][lane + 32 * h] = v;
        if (Fl::ljComb)
        {
            sm_lji[warp][lane + 32 * h] = ad.ljComb[ai];
        }
        else
        {
            const int t                 = ad.atomType[ai];
            sm_lji[warp][lane + 32 * h] = make_float2(__int_as_float(t * ad.numTypes), __int_as_float(t));
        }
    }
    __syncwarp();

    float3 fi[c_superClusterSize];
#pragma unroll
    for (int ci = 0; ci < c_superClusterSize; ci++)
    {
        fi[ci] = make_float3(0.0f, 0.0f, 0.0f);
    }

    /* j <= i within the same cluster on the central shift is the "Newton" half of the diagonal
     * cluster pair and the self pair (nbnxm_cuda_kernel.cuh:421-423) */
    const bool nonSelf0 = !(s.shift == c_centralShiftIndex && jl <= il);
    const bool nonSelf1 = !(s.shift == c_centralShiftIndex && jl + 4 <= il);

    const float        rlistOuter2 = p.rlist_outer_sq;
    int                prunedCount = 0;
    unsigned long long pairCount   = 0;

    for (int jp = s.cj_packed_begin; jp < s.cj_packed_end; jp++)
    {
        const int4 cjv = *reinterpret_cast<const int4*>(pl.cjPacked[jp].cj);
        const int4 mev = *reinterpret_cast<const int4*>(pl.cjPacked[jp].imei);
        unsigned   imask0 = static_cast<unsigned>(mev.x), imask1 = static_cast<unsigned>(mev.z);
        const unsigned imaskAny = imask0 | imask1;
        if (imaskAny == 0u)
        {
            continue;
        }
        const unsigned wexcl0 = pl.excl[mev.y].pair[lane];
        const unsigned wexcl1 = pl.excl[mev.w].pair[lane];
        const unsigned on0all = imask0, on1all = imask1;
        if (pl.pairCount != nullptr)
        {
            pairCount += 32ull * (__popc(imask0) + __popc(imask1));
        }
        const int cjs[4] = { cjv.x, cjv.y, cjv.z, cjv.w };

#pragma unroll
        for (int jm = 0; jm < c_jGroupSize; jm++)
        {
            if (imaskAny & (0xffu << (jm * 8)))
            {
                const int    cj  = cjs[jm];
                const int    aj0 = cj * c_clusterSize + jl;
                const int    aj1 = aj0 + 4;
                const float4 xj0 = ad.xq[aj0];
                const float4 xj1 = ad.xq[aj1];
                float2       pj0 = make_float2(0.0f, 0.0f), pj1 = make_float2(0.0f, 0.0f);
                int          tj0 = 0, tj1 = 0;
                if (Fl::ljComb)
                {
                    pj0 = ad.ljComb[aj0];
                    pj1 = ad.ljComb[aj1];
                }
                else
                {
                    tj0 = ad.atomType[aj0];
                    tj1 = ad.atomType[aj1];
                }
                float3 fj0 = make_float3(0.0f, 0.0f, 0.0f), fj1 = make_float3(0.0f, 0.0f, 0.0f);
                float  eLJj = 0.0f, eElj = 0.0f;

#pragma unroll
                for (int ci = 0; ci < c_superClusterSize; ci++)
                {
                    const unsigned bit = 1u << (jm * 8 + ci);
                    if (imaskAny & bit)
                    {
                        const float4 xi  = sm_xqi[warp][ci * 8 + il];
                        const float  dx0 = xi.x - xj0.x, dy0 = xi.y - xj0.y, dz0 = xi.z - xj0.z;
                        const float  dx1 = xi.x - xj1.x, dy1 = xi.y - xj1.y, dz1 = xi.z - xj1.z;
                        const float  r20 = norm2_fma(dx0, dy0, dz0);
                        const float  r21 = norm2_fma(dx1, dy1, dz1);
                        const bool   on0 = (on0all & bit) != 0u, on1 = (on1all & bit) != 0u;
                        if (PRUNE)
                        {
                            /* clear the bit of a (cluster pair, half) with no atom pair inside
                             * rlistOuter (nbnxm_cuda_kernel.cuh:495-503) */
                            if (on0 && !__any_sync(c_full, r20 < rlistOuter2)) imask0 &= ~bit;
                            if (on1 && !__any_sync(c_full, r21 < rlistOuter2)) imask1 &= ~bit;
                        }
                        const float intBit0 = (wexcl0 & bit) ? 1.0f : 0.0f;
                        const float intBit1 = (wexcl1 & bit) ? 1.0f : 0.0f;
                        bool        within0, within1;
                        if (Fl::exclusionForces)
                        {
                            const bool offDiagonal = (cj != s.sci * c_superClusterSize + ci);
                            within0                = on0 && (r20 < k.rc2) && (nonSelf0 || offDiagonal);
                            within1                = on1 && (r21 < k.rc2) && (nonSelf1 || offDiagonal);
                        }
                        else
                        {
                            within0 = on0 && (r20 < k.rc2) && (intBit0 != 0.0f);
                            within1 = on1 && (r21 < k.rc2) && (intBit1 != 0.0f);
                        }
                        if (within0 || within1)
                        {
                            const float2 pi  = sm_lji[warp][ci * 8 + il];
                            const int    tiN = __float_as_int(pi.x), ti = __float_as_int(pi.y);
                            float        c60, c120, c6g0, c61, c121, c6g1;
                            lj_pair_params<ELEC, VDW, ENERGY>(p, pi, pj0, tiN, ti, tj0, c60, c120, c6g0);
                            lj_pair_params<ELEC, VDW, ENERGY>(p, pi, pj1, tiN, ti, tj1, c61, c121, c6g1);
                            float e0lj = 0.0f, e0el = 0.0f, e1lj = 0.0f, e1el = 0.0f;
                            float F0 = pair_force<ELEC, VDW, ENERGY>(p, k, r20, xi.w * xj0.w, c60, c120, c6g0, intBit0, e0lj, e0el);
                            float F1 = pair_force<ELEC, VDW, ENERGY>(p, k, r21, xi.w * xj1.w, c61, c121, c6g1, intBit1, e1lj, e1el);
                            F0 = within0 ? F0 : 0.0f;
                            F1 = within1 ? F1 : 0.0f;
                            if (ENERGY)
                            {
                                eLJj += (within0 ? e0lj : 0.0f) + (within1 ? e1lj : 0.0f);
                                eElj += (within0 ? e0el : 0.0f) + (within1 ? e1el : 0.0f);
                            }
                            fi[ci].x = fmaf(F0, dx0, fmaf(F1, dx1, fi[ci].x));
                            fi[ci].y = fmaf(F0, dy0, fmaf(F1, dy1, fi[ci].y));
                            fi[ci].z = fmaf(F0, dz0, fmaf(F1, dz1, fi[ci].z));
                            fj0.x    = fmaf(-F0, dx0, fj0.x);
                            fj0.y    = fmaf(-F0, dy0, fj0.y);
                            fj0.z    = fmaf(-F0, dz0, fj0.z);
                            fj1.x    = fmaf(-F1, dx1, fj1.x);
                            fj1.y    = fmaf(-F1, dy1, fj1.y);
                            fj1.z    = fmaf(-F1, dz1, fj1.z);
                        }
                    }
                }

                if (ENERGY)
                {
                    eLJ += eLJj;
                    eEl += eElj;
                }
                /* reduce the two j-atom forces over the 8 il-lanes, one v4 reduction per j-atom */
#pragma unroll
                for (int m = 1; m < 8; m <<= 1)
                {
                    fj0.x += __shfl_xor_sync(c_full, fj0.x, m);
                    fj0.y += __shfl_xor_sync(c_full, fj0.y, m);
                    fj0.z += __shfl_xor_sync(c_full, fj0.z, m);
                    fj1.x += __shfl_xor_sync(c_full, fj1.x, m);
                    fj1.y += __shfl_xor_sync(c_full, fj1.y, m);
                    fj1.z += __shfl_xor_sync(c_full, fj1.z, m);
                }
                if (il == 0)
                {
                    red_add_v4(ad.f4 + aj0, fj0.x, fj0.y, fj0.z);
                }
                if (il == 1)
                {
                    red_add_v4(ad.f4 + aj1, fj1.x, fj1.y, fj1.z);
                }
            }
        }
        if (PRUNE)
        {
            if (lane == 0)
            {
                pl.cjPacked[jp].imei[0].imask = imask0;
                pl.cjPacked[jp].imei[1].imask = imask1;
            }
            prunedCount += __popc(imask0) + __popc(imask1);
        }
    }

    /* i forces: reduce over the 4 jl-lanes, one v4 reduction per i-atom; shift force from the
     * per-lane partial sums (central shift skipped, nbnxm_cuda_kernel.cuh:697-717) */
    float fsx = 0.0f, fsy = 0.0f, fsz = 0.0f;
#pragma unroll
    for (int ci = 0; ci < c_superClusterSize; ci++)
    {
        float x = fi[ci].x, y = fi[ci].y, z = fi[ci].z;
        fsx += x;
        fsy += y;
        fsz += z;
        x += __shfl_xor_sync(c_full, x, 8);
        y += __shfl_xor_sync(c_full, y, 8);
        z += __shfl_xor_sync(c_full, z, 8);
        x += __shfl_xor_sync(c_full, x, 16);
        y += __shfl_xor_sync(c_full, y, 16);
        z += __shfl_xor_sync(c_full, z, 16);
        if (jl == (ci & 3))
        {
            red_add_v4(ad.f4 + (s.sci * c_superClusterSize + ci) * c_clusterSize + il, x, y, z);
        }
    }
    if (calcFshift && s.shift != c_centralShiftIndex)
    {
#pragma unroll
        for (int m = 16; m > 0; m >>= 1)
        {
            fsx += __shfl_xor_sync(c_full, fsx, m);
            fsy += __shfl_xor_sync(c_full, fsy, m);
            fsz += __shfl_xor_sync(c_full, fsz, m);
        }
        if (lane < 3)
        {
            atomicAdd(ad.fshift + 3 * s.shift + lane, static_cast<double>(lane == 0 ? fsx : (lane == 1 ? fsy : fsz)));
        }
    }
    if (ENERGY)
    {
#pragma unroll
        for (int m = 16; m > 0; m >>= 1)
        {
            eLJ += __shfl_xor_sync(c_full, eLJ, m);
            eEl += __shfl_xor_sync(c_full, eEl, m);
        }
        if (lane < 2)
        {
            atomicAdd(ad.energy + lane, lane == 0 ? eLJ : eEl);
        }
    }
    if (PRUNE && lane == 0)
    {
        /* histogram index for the sci sort (nbnxm_cuda_kernel.cuh:725-744) */
        const int index = max(c_sciHistogramSize - prunedCount - 1, 0);
        atomicAdd(pl.sciHistogram + index, 1);
        pl.sciCount[sciIdx] = index;
    }
    if (pl.pairCount != nullptr && lane == 0)
    {
        atomicAdd(pl.pairCount, pairCount);
    }
}

typedef void (*ForceKernelPtr)(const AtomDataDev, const ParamsDev, const PairlistDev, const int);

/* one translation unit per electrostatics type instantiates its 7 x 2 x 2 kernels */
template<int ELEC>
ForceKernelPtr select_force_kernel_elec(int vdw, bool energy, bool prune);

ForceKernelPtr select_force_kernel(int elec, int vdw, bool energy, bool prune);

} // namespace nbb

#endif
