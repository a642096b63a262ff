/* Force kernel of the nbnxm_b200 path (all electrostatics x VdW flavors, F / F+E, optional
 * outer-list pruning).  Replaces the reference's macro-generated kernels
 * src/gromacs/nbnxm/cuda/nbnxm_cuda_kernel.cuh:138-746 (physics from
 * src/gromacs/nbnxm/nbnxm_kernel_utils.h:56-289); written from scratch for sm_100a.
 *
 * Work decomposition (differs from the reference on purpose):
 *   - one WARP = one 32-thread CTA per sci entry (i-super-cluster work unit): per-entry data is
 *     CTA-uniform (uniform registers / predicates, no convergence barriers), no block-level
 *     synchronisation;
 *   - lane = il + 8*jl owns i-atom `il` of every i-cluster; the unit of work is one (j-cluster, half)
 *     = 8 i-atoms x 4 j-atoms per i-cluster, i.e. exactly one imask bit per i-cluster, one pair per
 *     lane.  The two halves of the reference's "cluster-pair split" (imei[0], imei[1]) are walked one
 *     after the other by the same warp, so a half whose bit was pruned costs nothing;
 *   - the (j-cluster, half) loop is NOT unrolled and the i-cluster loop is (i forces live in registers):
 *     the hot loop stays inside the 32 KB instruction cache;
 *   - i-atom coordinates/parameters are staged once per sci entry in shared memory; the 32 j-atoms of a
 *     cjPacked group are fetched by one coalesced 16-byte load per lane, one group ahead, and parked in
 *     shared memory (conflict-free LDS.128 broadcasts in the pair loop);
 *   - j forces are reduced over the 8 il-lanes by shuffles and added with one
 *     red.global.add.v4.f32 per j-atom; i forces are kept in registers for the whole sci entry.
 */
#ifndef NBNXM_B200_FORCE_KERNEL_CUH
#define NBNXM_B200_FORCE_KERNEL_CUH

#include "nbnxm_device.cuh"

namespace nbb
{

/* one warp per CTA; resident CTAs per SM the kernels are compiled for (register budget 65536 / 32 / N) */
#ifndef NBNXM_FORCE_MIN_BLOCKS
#    define NBNXM_FORCE_MIN_BLOCKS 20
#endif
constexpr int c_forceMinBlocksPerSM = NBNXM_FORCE_MIN_BLOCKS;

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

/* beta^3 (2/sqrt(pi) z exp(-z^2) - erf(z)) / z^3 with z = beta r as a rational minimax approximation in r^2.
 * The coefficients are the reference's pmeCorrF ones (nbnxm_kernel_utils.h:216-250) with the powers of beta
 * folded in on the host (ParamsDev::pmeNum / pmeDen, fillParamsDev): they sit in the constant bank and feed
 * the FFMAs directly, which saves the z^2 = beta^2 r^2 and the * beta^3 multiplications of every pair. */
__device__ __forceinline__ float pme_corr_f_beta3(const ParamsDev& p, const float r2)
{
    float den = fmaf(p.pmeDen[4], r2, p.pmeDen[3]);
    den       = fmaf(den, r2, p.pmeDen[2]);
    den       = fmaf(den, r2, p.pmeDen[1]);
    den       = fmaf(den, r2, 1.0f);
    float num = fmaf(p.pmeNum[6], r2, p.pmeNum[5]);
    num       = fmaf(num, r2, p.pmeNum[4]);
    num       = fmaf(num, r2, p.pmeNum[3]);
    num       = fmaf(num, r2, p.pmeNum[2]);
    num       = fmaf(num, r2, p.pmeNum[1]);
    num       = fmaf(num, r2, p.pmeNum[0]);
    /* den >= 1: the plain approximate reciprocal needs no range fix-up */
    float rden;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rden) : "f"(den));
    return num * rden;
}

/* Dynamic-index read of one of four registers without local memory. */
__device__ __forceinline__ unsigned select4(const unsigned a, const unsigned b, const unsigned c, const unsigned d, const int i)
{
    const unsigned lo = (i & 1) ? b : a;
    const unsigned hi = (i & 1) ? d : c;
    return (i & 2) ? hi : lo;
}

__device__ __forceinline__ float rsqrt_approx(const float x)
{
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
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
    /* In the energy kernels qq = qi*qj comes without the electric conversion factor: the Coulomb energy is
     * a small residual of large terms of both signs, and a factor rounded into every i-charge would be a
     * systematic relative error on those terms.  The factor is applied to the double-precision total. */
    const float qqF   = ENERGY ? qq * p.epsfac : qq;
    r2                = fmaxf(r2, c_minDistanceSquared);
    float invR = rsqrt_approx(r2);
    if (ENERGY)
    {
        /* one Newton-Raphson step: the energy totals are sums with heavy cancellation and need pair
         * energies good to an ulp to stay within 1e-6 relative */
        invR = invR * fmaf(-0.5f * r2 * invR, invR, 1.5f);
    }
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
        fInvR += qqF * (Fl::exclusionForces ? intBit : 1.0f) * invR2 * invR;
        if (ENERGY) eEl += qq * (intBit * invR - p.c_rf);
    }
    if (Fl::elecRF)
    {
        fInvR += qqF * (intBit * invR2 * invR - p.two_k_rf);
        if (ENERGY) eEl += qq * (intBit * invR + 0.5f * p.two_k_rf * r2 - p.c_rf);
    }
    if (Fl::ewaldAna)
    {
        fInvR += qqF * (intBit * invR2 * invR + pme_corr_f_beta3(p, r2));
    }
    if (Fl::ewaldTab)
    {
        /* pairs beyond the cut-off are evaluated too (and masked afterwards): keep them inside the table */
        const float normalized = p.coulomb_tab_scale * fminf(r2 * invR, k.rcoulomb);
        const int   index      = static_cast<int>(normalized);
        const float fraction   = normalized - index;
        const float left       = __ldg(p.coulombTab + index);
        const float right      = __ldg(p.coulombTab + index + 1);
        fInvR += qqF * (intBit * invR2 - fmaf(fraction, right - left, left)) * invR;
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

/* One (j-cluster, half) against the i-clusters whose bits are set in m8: one atom pair per lane and
 * i-cluster.  EXCL = false is the fast path for groups without exclusion masks (excl_ind == 0, entry 0
 * of the exclusion array is all ones): no per-pair mask evaluation and no diagonal test - the diagonal
 * cluster pair always carries an exclusion mask (pairlist.cpp:651-688). */
template<int ELEC, int VDW, bool ENERGY, bool PRUNE, bool EXCL>
__device__ __forceinline__ void cluster_half(const ParamsDev&  p,
                                             const PairConsts& k,
                                             const float4*     xqi,
                                             const float2*     lji,
                                             const float4      xj,
                                             const float2      pj,
                                             const unsigned    m8,
                                             const unsigned    wex8,
                                             const bool        nonSelf,
                                             const int         ciDiag,
                                             const float       rlistOuter2,
                                             float3 (&fi)[c_superClusterSize],
                                             float3&   fj,
                                             float&    eLJj,
                                             float&    eElj,
                                             unsigned& keep8)
{
    using Fl                  = Flavor<ELEC, VDW, ENERGY>;
    constexpr unsigned c_full = 0xffffffffu;
    const int          tj     = __float_as_int(pj.x);
#pragma unroll
    for (int ci = 0; ci < c_superClusterSize; ci++)
    {
        if (m8 & (1u << ci))
        {
            const float4 xi = xqi[ci * c_clusterSize];
            const float  dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
            const float  r2 = norm2_fma(dx, dy, dz);
            if (PRUNE)
            {
                /* clear the bit of a (cluster pair, half) with no atom pair inside rlistOuter
                 * (nbnxm_cuda_kernel.cuh:495-503) */
                if (!__any_sync(c_full, r2 < rlistOuter2)) keep8 &= ~(1u << ci);
            }
            const float intBit = (!EXCL || (wex8 & (1u << ci))) ? 1.0f : 0.0f;
            bool        within;
            if (!EXCL)
            {
                within = (r2 < k.rc2);
            }
            else if (Fl::exclusionForces)
            {
                within = (r2 < k.rc2) && (nonSelf || ciDiag != ci);
            }
            else
            {
                within = (r2 < k.rc2) && (intBit != 0.0f);
            }
            const float2 pi  = lji[ci * c_clusterSize];
            const int    tiN = __float_as_int(pi.x), ti = __float_as_int(pi.y);
            float        c6, c12, c6g;
            lj_pair_params<ELEC, VDW, ENERGY>(p, pi, pj, tiN, ti, tj, c6, c12, c6g);
            float ePairLJ = 0.0f, ePairEl = 0.0f;
            float F = pair_force<ELEC, VDW, ENERGY>(p, k, r2, xi.w * xj.w, c6, c12, c6g, intBit, ePairLJ, ePairEl);
            F       = within ? F : 0.0f;
            if (ENERGY)
            {
                eLJj += within ? ePairLJ : 0.0f;
                eElj += within ? ePairEl : 0.0f;
            }
            fi[ci].x = fmaf(F, dx, fi[ci].x);
            fi[ci].y = fmaf(F, dy, fi[ci].y);
            fi[ci].z = fmaf(F, dz, fi[ci].z);
            fj.x     = fmaf(-F, dx, fj.x);
            fj.y     = fmaf(-F, dy, fj.y);
            fj.z     = fmaf(-F, dz, fj.z);
        }
    }
}

/* One warp = one CTA = one sci entry.  With 32-thread CTAs everything that is per-entry (list masks,
 * cluster indices, loop trip counts) is CTA-uniform, which lets the compiler keep it in uniform
 * registers and branch on uniform predicates without convergence barriers. */
template<int ELEC, int VDW, bool ENERGY, bool PRUNE>
__global__ void __launch_bounds__(32, (ENERGY || PRUNE) ? c_forceMinBlocksPerSM - 4 : c_forceMinBlocksPerSM)
        nbnxm_force_kernel(const AtomDataDev ad, const ParamsDev p, const PairlistDev pl, const int calcFshift)
{
    using Fl                  = Flavor<ELEC, VDW, ENERGY>;
    constexpr unsigned c_full = 0xffffffffu;
    /* the energy kernels carry one general pair path only: two copies would not fit the instruction cache */
    constexpr bool c_fastPath = !ENERGY && !PRUNE;

    const int lane   = threadIdx.x;
    const int il     = lane & 7;
    const int jl     = lane >> 3;
    const int sciIdx = blockIdx.x;
    /* The fused force+prune variant produces the counts the sort consumes, so it walks the
     * unsorted list (same rule as nbnxm_cuda_kernel.cuh:158-166). */
    const nbnxm_b200_sci_t s = PRUNE ? pl.sci[sciIdx] : pl.sciSorted[sciIdx];

    /* staging areas: the 64 i-atoms of the entry (for its whole lifetime) and the 32 j-atoms of the
     * current cjPacked group */
    __shared__ float4 sm_xqi[64];
    __shared__ float2 sm_lji[64]; // comb params, or (type*numTypes, type) as int bits
    __shared__ float4 sm_xqj[32];
    __shared__ float4 sm_ljj[32]; // (c6, c12) or (type as int bits, -), then the global atom index
    /* per-lane partial j forces of the current group: [j-atom slot][il ^ (slot & 7)] so that both the stores
     * (8 different 16-byte columns of one row per quarter warp) and the column sums (8 different columns per
     * quarter warp) are bank-conflict free */
    __shared__ __align__(128) float4 sm_fj[32 * c_clusterSize];

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

    /* energies: float partial sums per j-cluster half, double across the sci entry, so that the totals
     * keep 1e-6 relative accuracy (the reference accumulates in float, gpu_common.h:151-161) */
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
            const double q2 = static_cast<double>(v.w) * v.w;
            if (Fl::ewaldAny) eEl -= q2 * p.ewald_beta * 0.56418958354775628695;
            if (Fl::elecRF || Fl::elecCut) eEl -= q2 * 0.5 * p.c_rf;
            if (Fl::ljEwald)
            {
                eLJ += __ldg(p.nbfp + ad.atomType[ai] * (ad.numTypes + 1)).x * 0.5f * c_oneSixth * k.ljeCoeff6_6;
            }
        }
        v.x += shx;
        v.y += shy;
        v.z += shz;
        if (!ENERGY)
        {
            v.w *= p.epsfac;
        }
        sm_xqi[lane + 32 * h] = v;
        if (Fl::ljComb)
        {
            sm_lji[lane + 32 * h] = ad.ljComb[ai];
        }
        else
        {
            const int t           = ad.atomType[ai];
            sm_lji[lane + 32 * h] = make_float2(__int_as_float(t * ad.numTypes), __int_as_float(t));
        }
    }

    float3 fi[c_superClusterSize];
#pragma unroll
    for (int ci = 0; ci < c_superClusterSize; ci++)
    {
        fi[ci] = make_float3(0.0f, 0.0f, 0.0f);
    }

    const bool         centralShift = (s.shift == c_centralShiftIndex);
    const float        rlistOuter2  = p.rlist_outer_sq;
    int                prunedCount  = 0;

    /* Software pipeline over the cjPacked groups of the entry: group descriptors (32 bytes, CTA-uniform
     * loads) are fetched two groups ahead, the 32 j-atoms of a group one group ahead - lane L fetches
     * atom (L & 7) of j-cluster (L >> 3), one coalesced 16-byte load per lane - and parked in shared
     * memory, together with their LJ parameters and global index, for the (half, j-cluster) loop. */
    const uint4* cjGroups = reinterpret_cast<const uint4*>(pl.cjPacked);
    int          jp       = s.cj_packed_begin;
    const uint4  zero4    = make_uint4(0u, 0u, 0u, 0u);
    uint4        cjNext = zero4, meNext = zero4, cjNext2 = zero4, meNext2 = zero4;
    if (jp < s.cj_packed_end)
    {
        cjNext = cjGroups[2 * jp];
        meNext = cjGroups[2 * jp + 1];
    }
    if (jp + 1 < s.cj_packed_end)
    {
        cjNext2 = cjGroups[2 * jp + 2];
        meNext2 = cjGroups[2 * jp + 3];
    }
    float4 xjNext = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    float4 pjNext = make_float4(0.0f, 0.0f, 0.0f, 0.0f); /* (c6, c12) or (type, -), then the atom index */

    auto fetchAtoms = [&](const uint4 cjv, const uint4 mev, float4& xj, float4& pj) {
        /* unused slots of a partially filled group have no mask bits and an unspecified index */
        if ((mev.x | mev.z) & (0xffu << (8 * jl)))
        {
            const int aj = static_cast<int>(select4(cjv.x, cjv.y, cjv.z, cjv.w, jl)) * c_clusterSize + il;
            xj           = ad.xqJ[aj];
            if (Fl::ljComb)
            {
                const float2 c = ad.ljComb[aj];
                pj.x           = c.x;
                pj.y           = c.y;
            }
            else
            {
                pj.x = __int_as_float(ad.atomType[aj]);
            }
            pj.z = __int_as_float(aj);
        }
    };
    fetchAtoms(cjNext, meNext, xjNext, pjNext);

    const float4* xqiLane = sm_xqi + il;
    const float2* ljiLane = sm_lji + il;

    for (; jp < s.cj_packed_end; jp++)
    {
        const uint4 mev   = meNext;
        const int   ajOwn = __float_as_int(pjNext.z); /* the atom this lane fetched for the group */
        __syncwarp();
        sm_xqj[lane] = xjNext;
        sm_ljj[lane] = pjNext;
        __syncwarp();
        cjNext = cjNext2;
        meNext = meNext2;
        if (jp + 2 < s.cj_packed_end)
        {
            cjNext2 = cjGroups[2 * jp + 4];
            meNext2 = cjGroups[2 * jp + 5];
        }
        else
        {
            meNext2 = zero4;
        }
        fetchAtoms(cjNext, meNext, xjNext, pjNext);

        if ((mev.x | mev.z) == 0u)
        {
            continue;
        }
        unsigned keep0 = mev.x, keep1 = mev.z;

        /* the two halves of the cluster-pair split one after the other; within a half one iteration per
         * j-cluster: 8 i-atoms x 4 j-atoms per i-cluster, one pair per lane */
#pragma unroll 1
        for (int half = 0; half < 2; half++)
        {
            unsigned  cur     = half ? mev.z : mev.x;
            const int exclInd = static_cast<int>(half ? mev.w : mev.y);
            /* entry 0 of the exclusion array is all ones (pairlist.h:274-287) */
            unsigned      wex    = (exclInd != 0 && cur != 0u) ? pl.excl[exclInd].pair[lane] : c_full;
            const float4* xqjPtr = sm_xqj + half * 4 + jl;
            const float4* ljjPtr = sm_ljj + half * 4 + jl;
            /* j <= i within the same cluster on the central shift is the "Newton" half of the
             * diagonal cluster pair and the self pair (nbnxm_cuda_kernel.cuh:421-423) */
            const bool nonSelf = !(centralShift && (half * 4 + jl) <= il);
            unsigned   cleared = 0u;
#pragma unroll 1
            for (int shift = 0; cur != 0u; cur >>= 8, wex >>= 8, xqjPtr += c_clusterSize, ljjPtr += c_clusterSize, shift += 8)
            {
                const unsigned m8 = cur & 0xffu;
                if (m8 == 0u)
                {
                    continue;
                }
                const float4 xj  = *xqjPtr;
                const float4 pjv = *ljjPtr;
                const float2 pj  = make_float2(pjv.x, pjv.y);
                const int    aj  = __float_as_int(pjv.z);
                float3       fj  = make_float3(0.0f, 0.0f, 0.0f);
                float        eLJj = 0.0f, eElj = 0.0f;
                unsigned     keep8 = m8;

                if (c_fastPath && exclInd == 0)
                {
                    cluster_half<ELEC, VDW, ENERGY, PRUNE, false>(p, k, xqiLane, ljiLane, xj, pj, m8, 0xffu, true, -1,
                                                                   rlistOuter2, fi, fj, eLJj, eElj, keep8);
                }
                else
                {
                    /* the i-cluster this j-cluster is, if any */
                    const int ciDiag = (aj >> 3) - s.sci * c_superClusterSize;
                    cluster_half<ELEC, VDW, ENERGY, PRUNE, true>(p, k, xqiLane, ljiLane, xj, pj, m8, wex, nonSelf, ciDiag,
                                                                  rlistOuter2, fi, fj, eLJj, eElj, keep8);
                }

                if (ENERGY)
                {
                    eLJ += eLJj;
                    eEl += eElj;
                }
                /* park the lane's partial j force; the group's 32 j-atoms are reduced together below */
                {
                    const int slot = static_cast<int>(xqjPtr - sm_xqj);
                    sm_fj[slot * c_clusterSize + (il ^ (slot & 7))] = make_float4(fj.x, fj.y, fj.z, 0.0f);
                }
                if (PRUNE)
                {
                    cleared |= (m8 & ~keep8) << shift;
                }
            }
            if (PRUNE)
            {
                if (half) keep1 &= ~cleared; else keep0 &= ~cleared;
            }
        }
        /* j forces of the group: lane L sums the 8 partial forces of j-atom slot L (j-cluster L >> 3, atom
         * L & 7, i.e. half (L >> 2) & 1) and adds them with ONE v4 reduction; slots of halves that were not
         * visited hold stale data and are skipped */
        __syncwarp();
        {
            const unsigned visited = (((lane >> 2) & 1) ? mev.z : mev.x) & (0xffu << (8 * (lane >> 3)));
            if (visited != 0u)
            {
                const float4* col = sm_fj + lane * c_clusterSize;
                const float4  v0  = col[il];
                float         sx = v0.x, sy = v0.y, sz = v0.z;
#pragma unroll
                for (int kx = 1; kx < c_clusterSize; kx++)
                {
                    const float4 v = col[kx ^ il];
                    sx += v.x;
                    sy += v.y;
                    sz += v.z;
                }
                red_add_v4(ad.f4J + ajOwn, sx, sy, sz);
            }
        }
        if (PRUNE)
        {
            if (lane == 0)
            {
                pl.cjPacked[jp].imei[0].imask = keep0;
                pl.cjPacked[jp].imei[1].imask = keep1;
            }
            prunedCount += __popc(keep0) + __popc(keep1);
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
        red_add_v4_if(jl == (ci & 3), ad.f4 + (s.sci * c_superClusterSize + ci) * c_clusterSize + il, x, y, z);
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
            atomicAdd(ad.energy + lane, lane == 0 ? eLJ : eEl * static_cast<double>(p.epsfac));
        }
    }
    if (PRUNE && lane == 0)
    {
        /* histogram index for the sci sort (nbnxm_cuda_kernel.cuh:725-744) */
        const int index = max(c_sciHistogramSize - prunedCount - 1, 0);
        atomicAdd(pl.sciHistogram + index, 1);
        pl.sciCount[sciIdx] = index;
    }
}

typedef void (*ForceKernelPtr)(const AtomDataDev, const ParamsDev, const PairlistDev, const int);

/* one translation unit per electrostatics type instantiates its 7 x 2 x 2 kernels */
template<int ELEC>
ForceKernelPtr select_force_kernel_elec(int vdw, bool energy, bool prune, int numTypes);

ForceKernelPtr select_force_kernel(int elec, int vdw, bool energy, bool prune, int numTypes);

} // namespace nbb

#endif
