/* Packed-FP32 (FFMA2 / FMUL2 / FADD2) force-only kernel of the nbnxm_b200 path.
 *
 * sm_100a can issue one instruction that does two FP32 FMAs on a 64-bit register pair (PTX
 * fma.rn.f32x2, SASS FFMA2; scalar registers, immediates and negations broadcast as operands for free).
 * It has the FLOP throughput of FFMA but half its issue-slot cost (profiles/microbench/fp32_peak.cu), and the
 * scalar kernel is issue bound, not FMA-pipe bound.  So this kernel evaluates the TWO halves of a cluster
 * pair - j-atoms jl and jl+4 of a j-cluster against the same i-atom - as one packed pair stream: every
 * quantity of the pair physics is a (half 0, half 1) register pair.  A cluster pair of which only one half
 * survived pruning (about one in six) runs with the other half masked off.
 *
 * Same list walk, staging and reductions as nbnxm_force_kernel (nbnxm_force_kernel.cuh), which remains the
 * kernel for the energy and fused-prune variants and for the flavors that need per-pair table look-ups or
 * expf (tabulated Ewald, LJ-PME).  Physics from src/gromacs/nbnxm/nbnxm_kernel_utils.h:56-289 and
 * src/gromacs/nbnxm/cuda/nbnxm_cuda_kernel.cuh:505-660.
 */
#ifndef NBNXM_B200_FORCE_KERNEL_PACKED_CUH
#define NBNXM_B200_FORCE_KERNEL_PACKED_CUH

#include "nbnxm_force_kernel.cuh"

namespace nbb
{

/* resident CTAs per SM the packed kernel is compiled for: two pair streams per lane want ~120 registers (ptxas rematerialises addresses in every pair body below that) */
#ifndef NBNXM_PACKED_MIN_BLOCKS
#    define NBNXM_PACKED_MIN_BLOCKS 14
#endif
constexpr int c_packedMinBlocksPerSM = NBNXM_PACKED_MIN_BLOCKS;
/* i-force accumulators: packed (half 0, half 1) pairs (48 registers, 3 FFMA2 per pair body) or scalars (24
 * registers, 6 FFMA) */
#ifndef NBNXM_PACKED_FI
#    define NBNXM_PACKED_FI 1
#endif

typedef unsigned long long f32x2; /* (lo, hi) = (half 0, half 1) */

__device__ __forceinline__ f32x2 pk(const float lo, const float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f32x2 bc(const float a)
{
    return pk(a, a);
}
__device__ __forceinline__ float lo(const f32x2 v)
{
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return a;
}
__device__ __forceinline__ float hi(const f32x2 v)
{
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return b;
}
__device__ __forceinline__ f32x2 fma2(const f32x2 a, const f32x2 b, const f32x2 c)
{
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 mul2(const f32x2 a, const f32x2 b)
{
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 add2(const f32x2 a, const f32x2 b)
{
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 sub2(const f32x2 a, const f32x2 b)
{
    f32x2 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

/* flavors the packed kernel covers */
template<int ELEC, int VDW>
struct PackedFlavor
{
    using Fl                        = Flavor<ELEC, VDW, false>;
    static constexpr bool available = !Fl::ewaldTab && !Fl::ljEwald;
};

/* beta^3 pmeCorrF(beta^2 r^2) for both halves (coefficients: nbnxm_kernel_utils.h:216-250) */
__device__ __forceinline__ f32x2 pme_corr_f_packed(const f32x2 z2)
{
    f32x2 den = fma2(z2, bc(0.0011193462567257629232f), bc(0.014866955030185295499f));
    den       = fma2(den, z2, bc(0.11583842382862377919f));
    den       = fma2(den, z2, bc(0.50736591960530292870f));
    den       = fma2(den, z2, bc(1.0f));
    f32x2 num = fma2(z2, bc(-1.7357322914161492954e-8f), bc(1.4703624142580877519e-6f));
    num       = fma2(num, z2, bc(-0.000053401640219807709149f));
    num       = fma2(num, z2, bc(0.0010054721316683106153f));
    num       = fma2(num, z2, bc(-0.019278317264888380590f));
    num       = fma2(num, z2, bc(0.069670166153766424023f));
    num       = fma2(num, z2, bc(-0.75225204789749321333f));
    float r0, r1;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(lo(den)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(hi(den)));
    return mul2(num, pk(r0, r1));
}

/* erfc(x), x >= 0, for both halves: t P(t) form with t = 1 / (1 + x/2), P a degree-9 fit of erfc(x) exp(x^2)
 * (relative fit error 2.6e-9 on [0, 3.6]; in float32 about 3e-7 absolute, like erfcf), times exp(-x^2) with the
 * rounding error of x^2 compensated.  15 packed FP32 operations and 4 MUFU for two values, against ~55
 * instructions per value for erfcf. */
__device__ __forceinline__ f32x2 erfc_packed(const f32x2 x)
{
    const f32x2 d = fma2(x, bc(0.5f), bc(1.0f));
    float       t0, t1;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(lo(d)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(hi(d)));
    const f32x2 t = pk(t0, t1);
    f32x2       q = fma2(t, bc(-1.744380291e-02f), bc(4.996527726e-02f));
    q             = fma2(q, t, bc(8.959100685e-02f));
    q             = fma2(q, t, bc(-5.334397185e-01f));
    q             = fma2(q, t, bc(6.932961627e-01f));
    q             = fma2(q, t, bc(-2.149867186e-01f));
    q             = fma2(q, t, bc(4.016416182e-01f));
    q             = fma2(q, t, bc(2.443820402e-01f));
    q             = fma2(q, t, bc(2.873080888e-01f));
    q             = fma2(q, t, bc(-3.139566102e-04f));
    const f32x2 x2hi = mul2(x, x);
    const f32x2 x2lo = fma2(x, x, sub2(bc(0.0f), x2hi));
    const f32x2 arg  = mul2(x2hi, bc(-1.4426950408889634f));
    float       e0, e1;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(lo(arg)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(hi(arg)));
    f32x2 e = pk(e0, e1);
    e       = fma2(sub2(bc(0.0f), e), x2lo, e);
    return mul2(q, e);
}

/* F/r of the two pairs (not yet masked). c6n = -6*C6, c12 = 12*C12, qq = epsfac*qi*qj; intBit = 1/0 per
 * half, only read when EXCL. r2 already clamped to c_minDistanceSquared. */
template<int ELEC, int VDW, bool ENERGY, bool EXCL>
__device__ __forceinline__ f32x2 pair_force_packed(const ParamsDev&  p,
                                                   const PairConsts& k,
                                                   const f32x2       r2,
                                                   const f32x2       qq,
                                                   const f32x2       c6n,
                                                   const f32x2       c12,
                                                   const f32x2       intBit,
                                                   f32x2&            eLJout,
                                                   f32x2&            eElout)
{
    using Fl   = Flavor<ELEC, VDW, ENERGY>;
    f32x2 invR = pk(rsqrt_approx(lo(r2)), rsqrt_approx(hi(r2)));
    if (ENERGY)
    {
        /* one Newton-Raphson step, see pair_force(); qq comes without epsfac in the energy kernels */
        invR = mul2(invR, fma2(mul2(mul2(r2, bc(-0.5f)), invR), invR, bc(1.5f)));
    }
    const f32x2 qqF   = ENERGY ? mul2(qq, bc(p.epsfac)) : qq;
    const f32x2 invR2 = mul2(invR, invR);
    f32x2       invR6 = mul2(mul2(invR2, invR2), invR2);
    if (EXCL && Fl::exclusionForces)
    {
        invR6 = mul2(invR6, intBit);
    }
    /* invR6 (c12 invR6 - c6) invR2 */
    f32x2 fInvR = mul2(mul2(invR6, fma2(c12, invR6, c6n)), invR2);
    f32x2 eLJ   = 0ull;
    if (ENERGY || Fl::ljPSwitch)
    {
        /* c12 (invR6^2 + rep.cpot) / 12 - c6 (invR6 + disp.cpot) / 6 */
        eLJ = fma2(mul2(c12, bc(c_oneTwelfth)), fma2(invR6, invR6, bc(p.rep_cpot)),
                   mul2(mul2(c6n, bc(c_oneSixth)), add2(invR6, bc(p.disp_cpot))));
        if (EXCL && Fl::exclusionForces)
        {
            eLJ = mul2(eLJ, intBit);
        }
    }
    if (Fl::ljFSwitch || Fl::ljPSwitch)
    {
        const f32x2 r     = mul2(r2, invR);
        const f32x2 rswRaw = sub2(r, bc(p.rvdw_switch));
        const f32x2 rsw   = pk(fmaxf(lo(rswRaw), 0.0f), fmaxf(hi(rswRaw), 0.0f));
        if (Fl::ljFSwitch)
        {
            /* (-c6 (d2 + d3 rsw) + c12 (r2 + r3 rsw)) rsw^2 / r */
            const f32x2 disp = fma2(rsw, bc(p.disp_c3), bc(p.disp_c2));
            const f32x2 rep  = fma2(rsw, bc(p.rep_c3), bc(p.rep_c2));
            const f32x2 t    = fma2(c12, rep, mul2(c6n, disp));
            const f32x2 rsw2 = mul2(rsw, rsw);
            fInvR            = fma2(mul2(t, rsw2), invR, fInvR);
            if (ENERGY)
            {
                /* (c6 (d2/3 + d3/4 rsw) - c12 (r2/3 + r3/4 rsw)) rsw^3, with c6n = -c6 */
                const f32x2 dispE = fma2(rsw, bc(p.disp_c3 * 0.25f), bc(p.disp_c2 * (1.0f / 3.0f)));
                const f32x2 repE  = fma2(rsw, bc(p.rep_c3 * 0.25f), bc(p.rep_c2 * (1.0f / 3.0f)));
                const f32x2 u     = fma2(c12, repE, mul2(c6n, dispE));
                eLJ               = fma2(sub2(bc(0.0f), u), mul2(rsw2, rsw), eLJ);
            }
        }
        else
        {
            /* potential switch needs the pair energy even in the force-only kernel */
            const f32x2 rsw2 = mul2(rsw, rsw);
            const f32x2 sw   = fma2(mul2(rsw2, rsw), fma2(fma2(rsw, bc(p.sw_c5), bc(p.sw_c4)), rsw, bc(p.sw_c3)), bc(1.0f));
            const f32x2 dsw  = mul2(rsw2, fma2(fma2(rsw, bc(5.0f * p.sw_c5), bc(4.0f * p.sw_c4)), rsw, bc(3.0f * p.sw_c3)));
            /* fInvR sw - invR eLJ dsw (rsw = 0 gives sw = 1, dsw = 0) */
            fInvR = fma2(fInvR, sw, mul2(mul2(invR, eLJ), sub2(bc(0.0f), dsw)));
            eLJ   = mul2(eLJ, sw);
        }
    }
    if (Fl::vdwCutoffCheck)
    {
        const f32x2 inRange = pk(lo(r2) < k.rvdw2 ? 1.0f : 0.0f, hi(r2) < k.rvdw2 ? 1.0f : 0.0f);
        fInvR               = mul2(fInvR, inRange);
        eLJ                 = mul2(eLJ, inRange);
    }
    eLJout = eLJ;
    f32x2 invR3 = mul2(invR2, invR);
    if (EXCL && Fl::exclusionForces)
    {
        invR3 = mul2(invR3, intBit);
    }
    /* intBit * invR for the energies */
    const f32x2 invRi = (EXCL) ? mul2(invR, intBit) : invR;
    f32x2       eEl   = 0ull;
    if (Fl::elecCut)
    {
        fInvR = fma2(qqF, invR3, fInvR);
        if (ENERGY) eEl = mul2(qq, sub2(invRi, bc(p.c_rf)));
    }
    if (Fl::elecRF)
    {
        fInvR = fma2(qqF, sub2(invR3, bc(p.two_k_rf)), fInvR);
        if (ENERGY) eEl = mul2(qq, add2(invRi, fma2(r2, bc(0.5f * p.two_k_rf), bc(-p.c_rf))));
    }
    if (Fl::ewaldAna)
    {
        const f32x2 corr = pme_corr_f_packed(mul2(r2, bc(k.beta2)));
        fInvR            = fma2(qqF, fma2(corr, bc(k.beta3), invR3), fInvR);
        if (ENERGY)
        {
            /* qq (invR (erfc(beta r) - (1 - intBit)) - intBit sh_ewald); excluded pairs get -erf(beta r)/r */
            const f32x2 ec = erfc_packed(mul2(mul2(r2, invR), bc(k.beta)));
            if (EXCL)
            {
                eEl = mul2(qq, fma2(invR, add2(ec, sub2(intBit, bc(1.0f))), mul2(intBit, bc(-p.sh_ewald))));
            }
            else
            {
                eEl = mul2(qq, fma2(invR, ec, bc(-p.sh_ewald)));
            }
        }
    }
    eElout = eEl;
    return fInvR;
}

#if NBNXM_PACKED_FI
typedef f32x2 FiAcc;
__device__ __forceinline__ void fi_add(FiAcc& acc, const f32x2 F, const f32x2 d) { acc = fma2(F, d, acc); }
__device__ __forceinline__ float fi_total(const FiAcc acc) { return lo(acc) + hi(acc); }
__device__ __forceinline__ FiAcc fi_zero() { return 0ull; }
#else
typedef float FiAcc;
__device__ __forceinline__ void fi_add(FiAcc& acc, const f32x2 F, const f32x2 d) { acc = fmaf(lo(F), lo(d), fmaf(hi(F), hi(d), acc)); }
__device__ __forceinline__ float fi_total(const FiAcc acc) { return acc; }
__device__ __forceinline__ FiAcc fi_zero() { return 0.0f; }
#endif

/* One j-cluster (both halves) against the i-clusters whose bits are set in mAny = m0 | m1. */
template<int ELEC, int VDW, bool ENERGY, bool EXCL>
__device__ __forceinline__ void cluster_pair_packed(const ParamsDev&  p,
                                                    const PairConsts& k,
                                                    const float4*     xqi,
                                                    const float2*     lji,
                                                    const f32x2       xj,
                                                    const f32x2       yj,
                                                    const f32x2       zj,
                                                    const f32x2       qj,
                                                    const f32x2       ljj0, /* (c6, c12 comb) or types as int bits */
                                                    const f32x2       ljj1,
                                                    const unsigned    m0,
                                                    const unsigned    m1,
                                                    const unsigned    wex0,
                                                    const unsigned    wex1,
                                                    const bool        nonSelf0,
                                                    const bool        nonSelf1,
                                                    const int         ciDiag,
                                                    FiAcc (&fi)[c_superClusterSize][3],
                                                    f32x2 (&fj)[3],
                                                    f32x2& eLJacc,
                                                    f32x2& eElacc)
{
    using Fl            = Flavor<ELEC, VDW, ENERGY>;
    const unsigned mAny = m0 | m1;
#pragma unroll
    for (int ci = 0; ci < c_superClusterSize; ci++)
    {
        if (mAny & (1u << ci))
        {
            const float4 xi = xqi[ci * c_clusterSize];
            const f32x2  dx = sub2(bc(xi.x), xj), dy = sub2(bc(xi.y), yj), dz = sub2(bc(xi.z), zj);
            const f32x2  r2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
            bool         w0 = (m0 & (1u << ci)) && (lo(r2) < k.rc2);
            bool         w1 = (m1 & (1u << ci)) && (hi(r2) < k.rc2);
            f32x2        intBit = bc(1.0f);
            if (EXCL)
            {
                const bool i0 = (wex0 & (1u << ci)) != 0u, i1 = (wex1 & (1u << ci)) != 0u;
                intBit        = pk(i0 ? 1.0f : 0.0f, i1 ? 1.0f : 0.0f);
                if (Fl::exclusionForces)
                {
                    const bool offDiagonal = (ciDiag != ci);
                    w0                     = w0 && (nonSelf0 || offDiagonal);
                    w1                     = w1 && (nonSelf1 || offDiagonal);
                }
                else
                {
                    w0 = w0 && i0;
                    w1 = w1 && i1;
                }
            }
            const float2 pi = lji[ci * c_clusterSize];
            f32x2        c6n, c12;
            if (Fl::ljCombGeom)
            {
                c6n = mul2(bc(-pi.x), ljj0);
                c12 = mul2(bc(pi.y), ljj1);
            }
            else if (Fl::ljCombLB)
            {
                const f32x2 sigma  = add2(bc(pi.x), ljj0);
                const f32x2 eps    = mul2(bc(pi.y), ljj1);
                const f32x2 sigma2 = mul2(sigma, sigma);
                const f32x2 sigma6 = mul2(mul2(sigma2, sigma2), sigma2);
                const f32x2 c6     = mul2(eps, sigma6);
                c12                = mul2(c6, sigma6);
                c6n                = sub2(bc(0.0f), c6);
            }
            else
            {
                const int    tiN = __float_as_int(pi.x);
                const float2 a   = __ldg(p.nbfp + tiN + __float_as_int(lo(ljj0)));
                const float2 b   = __ldg(p.nbfp + tiN + __float_as_int(hi(ljj0)));
                c6n              = pk(-a.x, -b.x);
                c12              = pk(a.y, b.y);
            }
            const f32x2 r2c = pk(fmaxf(lo(r2), c_minDistanceSquared), fmaxf(hi(r2), c_minDistanceSquared));
            f32x2       ePairLJ, ePairEl;
            f32x2       F = pair_force_packed<ELEC, VDW, ENERGY, EXCL>(p, k, r2c, mul2(bc(xi.w), qj), c6n, c12, intBit, ePairLJ, ePairEl);
            F             = pk(w0 ? lo(F) : 0.0f, w1 ? hi(F) : 0.0f);
            if (ENERGY)
            {
                eLJacc = add2(eLJacc, pk(w0 ? lo(ePairLJ) : 0.0f, w1 ? hi(ePairLJ) : 0.0f));
                eElacc = add2(eElacc, pk(w0 ? lo(ePairEl) : 0.0f, w1 ? hi(ePairEl) : 0.0f));
            }
            fi_add(fi[ci][0], F, dx);
            fi_add(fi[ci][1], F, dy);
            fi_add(fi[ci][2], F, dz);
            fj[0]           = fma2(F, dx, fj[0]);
            fj[1]           = fma2(F, dy, fj[1]);
            fj[2]           = fma2(F, dz, fj[2]);
        }
    }
}

template<int ELEC, int VDW, bool ENERGY>
__global__ void __launch_bounds__(32, ENERGY ? c_packedMinBlocksPerSM - 2 : c_packedMinBlocksPerSM)
        nbnxm_force_kernel_packed(const AtomDataDev ad, const ParamsDev p, const PairlistDev pl, const int calcFshift)
{
    using Fl                  = Flavor<ELEC, VDW, ENERGY>;
    constexpr unsigned c_full = 0xffffffffu;

    const int lane   = threadIdx.x;
    const int il     = lane & 7;
    const int jl     = lane >> 3;
    const int sciIdx = blockIdx.x;
    const nbnxm_b200_sci_t s = pl.sciSorted[sciIdx];

    /* i-atoms of the entry; j-atoms of the current group as (half 0, half 1) pairs: entry u = 4*jm + jl holds
     * atoms jl and jl+4 of j-cluster jm */
    __shared__ float4 sm_xqi[64];
    __shared__ float2 sm_lji[64];
    __shared__ float4 sm_jxy[16]; /* xA xB yA yB */
    __shared__ float4 sm_jzq[16]; /* zA zB qA qB */
    __shared__ float4 sm_jlj[16]; /* c6A c6B c12A c12B, or typeA typeB - - */
    __shared__ __align__(128) float4 sm_fj[32 * c_clusterSize];

    const float shx = ad.shiftVec[3 * s.shift], shy = ad.shiftVec[3 * s.shift + 1], shz = ad.shiftVec[3 * s.shift + 2];

    PairConsts k;
    k.rc2         = p.rcoulomb_sq;
    k.rcoulomb    = sqrtf(p.rcoulomb_sq);
    k.rvdw2       = p.rvdw_sq;
    k.beta        = p.ewald_beta;
    k.beta2       = p.ewald_beta * p.ewald_beta;
    k.beta3       = k.beta2 * p.ewald_beta;
    k.ljeCoeff2   = 0.0f;
    k.ljeCoeff6_6 = 0.0f;

    /* energies: float partial sums per j-cluster, double across the sci entry (see nbnxm_force_kernel) */
    double     eLJ = 0.0, eEl = 0.0;
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

    FiAcc fi[c_superClusterSize][3];
#pragma unroll
    for (int ci = 0; ci < c_superClusterSize; ci++)
    {
        fi[ci][0] = fi[ci][1] = fi[ci][2] = fi_zero();
    }

    const bool centralShift = (s.shift == c_centralShiftIndex);

    /* same software pipeline as nbnxm_force_kernel: descriptors two groups ahead, atoms one group ahead */
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
    float2 pjNext = make_float2(0.0f, 0.0f);
    int    ajNext = 0;

    auto fetchAtoms = [&](const uint4 cjv, const uint4 mev, float4& xj, float2& pj, int& ajOut) {
        if ((mev.x | mev.z) & (0xffu << (8 * jl)))
        {
            const int aj = static_cast<int>(select4(cjv.x, cjv.y, cjv.z, cjv.w, jl)) * c_clusterSize + il;
            xj           = ad.xq[aj];
            if (Fl::ljComb)
            {
                pj = ad.ljComb[aj];
            }
            else
            {
                pj.x = __int_as_float(ad.atomType[aj]);
            }
            ajOut = aj;
        }
    };
    fetchAtoms(cjNext, meNext, xjNext, pjNext, ajNext);

    const float4* xqiLane = sm_xqi + il;
    const float2* ljiLane = sm_lji + il;
    /* where this lane parks the atom it fetched: lane = 8*jm + atom, atom = 4*half + jl' */
    float* const stageXY = reinterpret_cast<float*>(sm_jxy + ((lane >> 3) * 4 + (lane & 3))) + ((lane >> 2) & 1);
    float* const stageZQ = reinterpret_cast<float*>(sm_jzq + ((lane >> 3) * 4 + (lane & 3))) + ((lane >> 2) & 1);
    float* const stageLJ = reinterpret_cast<float*>(sm_jlj + ((lane >> 3) * 4 + (lane & 3))) + ((lane >> 2) & 1);
    const bool    nonSelf0 = !(centralShift && jl <= il);
    const bool    nonSelf1 = !(centralShift && (jl + 4) <= il);

    for (; jp < s.cj_packed_end; jp++)
    {
        const uint4 cjv = cjNext, mev = meNext;
        const int   ajOwn = ajNext;
        __syncwarp();
        stageXY[0] = xjNext.x;
        stageXY[2] = xjNext.y;
        stageZQ[0] = xjNext.z;
        stageZQ[2] = xjNext.w;
        stageLJ[0] = pjNext.x;
        stageLJ[2] = pjNext.y;
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
        fetchAtoms(cjNext, meNext, xjNext, pjNext, ajNext);

        unsigned cur0 = mev.x, cur1 = mev.z;
        if ((cur0 | cur1) == 0u)
        {
            continue;
        }
        const int exclInd0 = static_cast<int>(mev.y), exclInd1 = static_cast<int>(mev.w);
        /* entry 0 of the exclusion array is all ones (pairlist.h:274-287) */
        unsigned  wex0 = (exclInd0 != 0) ? pl.excl[exclInd0].pair[lane] : c_full;
        unsigned  wex1 = (exclInd1 != 0) ? pl.excl[exclInd1].pair[lane] : c_full;
        /* the energy kernel carries the general pair path only: two copies would not fit the instruction cache */
        const bool haveExcl = ENERGY || (exclInd0 | exclInd1) != 0;

        const float4* jxy  = sm_jxy + jl;
        const float4* jzq  = sm_jzq + jl;
        const float4* jlj  = sm_jlj + jl;
        int           slot = jl; /* j-atom slot of half 0; half 1 is slot + 4 */
#pragma unroll 1
        for (int jm = 0; (cur0 | cur1) != 0u;
             jm++, cur0 >>= 8, cur1 >>= 8, wex0 >>= 8, wex1 >>= 8, jxy += 4, jzq += 4, jlj += 4, slot += c_clusterSize)
        {
            const unsigned m0 = cur0 & 0xffu, m1 = cur1 & 0xffu;
            if ((m0 | m1) == 0u)
            {
                continue;
            }
            const float4 xy = *jxy, zq = *jzq, lj = *jlj;
            const f32x2  xj = pk(xy.x, xy.y), yj = pk(xy.z, xy.w), zj = pk(zq.x, zq.y), qj = pk(zq.z, zq.w);
            const f32x2  ljj0 = pk(lj.x, lj.y), ljj1 = pk(lj.z, lj.w);
            f32x2        fj[3] = { 0ull, 0ull, 0ull };
            f32x2        eLJj = 0ull, eElj = 0ull;
            if (!haveExcl)
            {
                cluster_pair_packed<ELEC, VDW, ENERGY, false>(p, k, xqiLane, ljiLane, xj, yj, zj, qj, ljj0, ljj1, m0, m1, 0xffu,
                                                              0xffu, true, true, -1, fi, fj, eLJj, eElj);
            }
            else
            {
                const int cj     = static_cast<int>(select4(cjv.x, cjv.y, cjv.z, cjv.w, jm));
                const int ciDiag = cj - s.sci * c_superClusterSize;
                cluster_pair_packed<ELEC, VDW, ENERGY, true>(p, k, xqiLane, ljiLane, xj, yj, zj, qj, ljj0, ljj1, m0, m1, wex0,
                                                             wex1, nonSelf0, nonSelf1, ciDiag, fi, fj, eLJj, eElj);
            }
            if (ENERGY)
            {
                eLJ += lo(eLJj) + hi(eLJj);
                eEl += lo(eElj) + hi(eElj);
            }
            /* park the partial j forces (sign: the j-atom gets -F d) */
            if (m0 != 0u)
            {
                sm_fj[slot * c_clusterSize + (il ^ (slot & 7))] = make_float4(-lo(fj[0]), -lo(fj[1]), -lo(fj[2]), 0.0f);
            }
            if (m1 != 0u)
            {
                const int slot1                                     = slot + 4;
                sm_fj[slot1 * c_clusterSize + (il ^ (slot1 & 7))] = make_float4(-hi(fj[0]), -hi(fj[1]), -hi(fj[2]), 0.0f);
            }
        }

        /* j forces of the group: lane L sums the 8 partial forces of j-atom slot L and adds them with one
         * v4 reduction; slots of halves that were not visited hold stale data and are skipped */
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
                red_add_v4(ad.f4 + ajOwn, sx, sy, sz);
            }
        }
    }

    /* i forces: add the two halves, reduce over the 4 jl-lanes, one v4 reduction per i-atom; shift force from
     * the per-lane partial sums (central shift skipped, nbnxm_cuda_kernel.cuh:697-717) */
    float fsx = 0.0f, fsy = 0.0f, fsz = 0.0f;
#pragma unroll
    for (int ci = 0; ci < c_superClusterSize; ci++)
    {
        float x = fi_total(fi[ci][0]), y = fi_total(fi[ci][1]), z = fi_total(fi[ci][2]);
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
}

} // namespace nbb

#endif
