/* Packed-FP32 (FFMA2 / FMUL2 / FADD2) force kernel of the nbnxm_b200 path, force-only and force+energy.
 *
 * On sm_100a the FP32 pipe retires one warp-wide FFMA per scheduler per clock, i.e. the FP32 peak IS the issue
 * peak: every instruction that is not an FMA takes a slot away from it (profiles/r01d_force_packed.txt,
 * profiles/microbench).  This kernel is therefore built around the instruction count per atom pair:
 *   - the two halves of a cluster pair (j-atoms jl and jl+4 against the same i-atom) are one packed pair stream
 *     (fma.rn.f32x2): ALU, MUFU-feeding, shared-memory and control instructions are shared by two pairs.  A
 *     cluster pair of which only one half survived pruning (about one in four) runs a scalar body on that half, so a
 *     pruned half costs nothing (unrolled per i-cluster in the force-only kernels, one loop body where the pair body
 *     is long: energies, potential switch);
 *   - the pair force is evaluated as W = (F/r) r^2 (one multiplication by r^-2 at the end, no r^-3), the Ewald
 *     correction polynomial runs on r^2 with the powers of beta folded into its coefficients, and the r^2 = 0
 *     guard of filler atoms is an additive 1e-12 inside the first FMA instead of a max;
 *   - the path with exclusion masks is chosen per (j-cluster, i-cluster) from a warp-wide OR of the mask words,
 *     not per cjPacked group, and is one non-unrolled loop: about 3 % of the cluster pairs of a water box carry
 *     exclusions, but 28 % of the groups do;
 *   - per-group bookkeeping is kept small: the cjPacked groups of an entry are staged in shared memory 64 at a time, the
 *     pair-body constants live in uniform registers (loaded from global memory once), the lane-dependent
 *     shared-memory addresses are computed once (and kept from being re-derived from the thread index in every
 *     body), partial j forces are parked without sign change in rows padded to 9 float4 (conflict-free without a
 *     swizzle) and reduced once per group.
 *
 * Same list walk and reductions as nbnxm_force_kernel (nbnxm_force_kernel.cuh), which remains the kernel for the
 * fused-prune variant, for tabulated Ewald (per-pair table look-ups), LJ-PME with the Lorentz-Berthelot grid and for
 * type-table flavors with more than c_packedMaxTypes atom types.  Physics from
 * src/gromacs/nbnxm/nbnxm_kernel_utils.h:56-289 and src/gromacs/nbnxm/cuda/nbnxm_cuda_kernel.cuh:505-660.
 */
#ifndef NBNXM_B200_FORCE_KERNEL_PACKED_CUH
#define NBNXM_B200_FORCE_KERNEL_PACKED_CUH

#include "nbnxm_force_kernel.cuh"

namespace nbb
{

/* resident CTAs (= warps) per SM the kernels are compiled for */
#ifndef NBNXM_PACKED_MIN_BLOCKS
#    define NBNXM_PACKED_MIN_BLOCKS 20
#endif
#ifndef NBNXM_PACKED_MIN_BLOCKS_ENERGY
#    define NBNXM_PACKED_MIN_BLOCKS_ENERGY 16
#endif
/* type-table flavors keep nbfp in shared memory: at most this many atom types */
constexpr int c_packedMaxTypes = 8;
/* added to r^2 in the force-only path without exclusions: keeps r^-6 finite for coinciding filler atoms
 * (whose parameters are zero) and is below half an ulp of any r^2 > 2e-5 nm^2 */
constexpr float c_r2Guard = 1.0e-12f;
/* upper clamp of r^2 in the bodies that mask by multiplication (nm^2) */
constexpr float c_maxDistanceSquared = 1.0e4f;

typedef unsigned long long f32x2; /* (lo, hi) = (half 0, half 1) */

#pragma nv_diag_suppress 550 /* the unused half of an unpacked register pair */

__device__ __forceinline__ f32x2 pk(const float lo, const float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float lo(const f32x2 v)
{
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    (void)b;
    return a;
}
__device__ __forceinline__ float hi(const f32x2 v)
{
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    (void)a;
    return b;
}

/* The pair physics is written once for V = f32x2 (two pairs per lane) and V = float (one pair). */
__device__ __forceinline__ f32x2 vfma(const f32x2 a, const f32x2 b, const f32x2 c)
{
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
/* acc += a b with the accumulator as a read-write operand (one virtual register in the PTX) */
__device__ __forceinline__ void vfma_acc(f32x2& acc, const f32x2 a, const f32x2 b)
{
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ f32x2 vmul(const f32x2 a, const f32x2 b)
{
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 vadd(const f32x2 a, const f32x2 b)
{
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 vsub(const f32x2 a, const f32x2 b)
{
    f32x2 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float vfma(const float a, const float b, const float c) { return fmaf(a, b, c); }
__device__ __forceinline__ float vmul(const float a, const float b) { return a * b; }
__device__ __forceinline__ float vadd(const float a, const float b) { return a + b; }
__device__ __forceinline__ float vsub(const float a, const float b) { return a - b; }

template<typename V>
__device__ __forceinline__ V vbc(const float a);
template<>
__device__ __forceinline__ f32x2 vbc<f32x2>(const float a) { return pk(a, a); }
template<>
__device__ __forceinline__ float vbc<float>(const float a) { return a; }

__device__ __forceinline__ float rcp_approx(const float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_approx(const float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float vrsqrt(const float x) { return rsqrt_approx(x); }
__device__ __forceinline__ f32x2 vrsqrt(const f32x2 x) { return pk(rsqrt_approx(lo(x)), rsqrt_approx(hi(x))); }
__device__ __forceinline__ float vrcp(const float x) { return rcp_approx(x); }
__device__ __forceinline__ f32x2 vrcp(const f32x2 x) { return pk(rcp_approx(lo(x)), rcp_approx(hi(x))); }
__device__ __forceinline__ float vex2(const float x) { return ex2_approx(x); }
__device__ __forceinline__ f32x2 vex2(const f32x2 x) { return pk(ex2_approx(lo(x)), ex2_approx(hi(x))); }
__device__ __forceinline__ float vmax0(const float x) { return fmaxf(x, 0.0f); }
__device__ __forceinline__ f32x2 vmax0(const f32x2 x) { return pk(fmaxf(lo(x), 0.0f), fmaxf(hi(x), 0.0f)); }
/* 1 where x < c, else 0 */
__device__ __forceinline__ float vless(const float x, const float c) { return x < c ? 1.0f : 0.0f; }
__device__ __forceinline__ f32x2 vless(const f32x2 x, const float c) { return pk(lo(x) < c ? 1.0f : 0.0f, hi(x) < c ? 1.0f : 0.0f); }

/* v where a < b, else 0: spelled as setp + selp so that ptxas emits FSETP + FSEL (the C conditional compiles to a zeroed
 * register, a predicate and a predicated move per value) */
__device__ __forceinline__ float sel_lt(const float a, const float b, const float v)
{
#ifdef NBNXM_PACKED_NO_SELP
    return a < b ? v : 0.0f;
#else
    float r;
    asm("{\n.reg .pred p;\nsetp.lt.f32 p, %1, %2;\nselp.f32 %0, %3, 0f00000000, p;\n}" : "=f"(r) : "f"(a), "f"(b), "f"(v));
    return r;
#endif
}

/* v where a < b and the pair's half is in the list, else 0 */
__device__ __forceinline__ float sel_lt_listed(const float a, const float b, const float v, const unsigned listed)
{
    float r;
    asm("{\n.reg .pred p, q;\nsetp.ne.u32 q, %4, 0;\nsetp.lt.and.f32 p, %1, %2, q;\nselp.f32 %0, %3, 0f00000000, p;\n}"
        : "=f"(r)
        : "f"(a), "f"(b), "f"(v), "r"(listed));
    return r;
}

/* flavors the packed kernel covers */
template<int ELEC, int VDW>
struct PackedFlavor
{
    using Fl                        = Flavor<ELEC, VDW, false>;
    static constexpr bool available = !Fl::ewaldTab && !Fl::ljEwaldLB;
    /* LJ parameters from the type table (staged in shared memory) instead of per-atom combination parameters */
    static constexpr bool typeTable = !Fl::ljComb;
};

/* erfc(x), x >= 0: t P(t) form with t = 1 / (1 + x/2), P a degree-9 fit of erfc(x) exp(x^2) (relative fit error
 * 2.6e-9 on [0, 3.6]; in float32 about 3e-7 absolute, like erfcf), times exp(-x^2) with the rounding error of x^2
 * compensated.  15 FP32 operations and 2 MUFU per value, against ~55 instructions for erfcf. */
template<typename V>
__device__ __forceinline__ V erfc_poly(const V x, V& expMinusX2)
{
    const V t = vrcp(vfma(x, vbc<V>(0.5f), vbc<V>(1.0f)));
    V       q = vfma(t, vbc<V>(-1.744380291e-02f), vbc<V>(4.996527726e-02f));
    q         = vfma(q, t, vbc<V>(8.959100685e-02f));
    q         = vfma(q, t, vbc<V>(-5.334397185e-01f));
    q         = vfma(q, t, vbc<V>(6.932961627e-01f));
    q         = vfma(q, t, vbc<V>(-2.149867186e-01f));
    q         = vfma(q, t, vbc<V>(4.016416182e-01f));
    q         = vfma(q, t, vbc<V>(2.443820402e-01f));
    q         = vfma(q, t, vbc<V>(2.873080888e-01f));
    q         = vfma(q, t, vbc<V>(-3.139566102e-04f));
    /* exp(-x^2) with the rounding error of x^2 compensated: h = fl(x^2), nl = h - x^2 exactly, e (1 + nl) */
    const V nx = vmul(x, vbc<V>(-1.0f));
    const V h  = vmul(x, x);
    const V nl = vfma(nx, x, h);
    V       e  = vex2(vmul(h, vbc<V>(-1.4426950408889634f)));
    e          = vfma(e, nl, e);
    expMinusX2 = e;
    return vmul(q, e);
}

/* Constants of the pair bodies that are the same for every pair of a launch.  The kernel loads them from a small
 * global-memory array (ParamsDev::packedConsts, filled by fillParamsDev), not from the kernel parameters: values
 * read from the constant bank are re-loaded by ptxas in front of every use (FFMA2 takes no constant-bank operand),
 * a dozen instructions per pair body, while loaded values stay in uniform registers. */
struct PackedConsts
{
    float rc2, rvdw2, beta, epsfac;
    float num[7], den[5]; /* pmeCorrF with beta folded in: beta^3 pmeCorrF(beta^2 r^2) = num(r^2) / den(r^2) */
    float rvdw_switch, disp_c2, disp_c3, rep_c2, rep_c3, disp_c2_3, disp_c3_4, rep_c2_3, rep_c3_4, disp_cpot, rep_cpot;
    float sw_c3, sw_c4, sw_c5, sw_c3x3, sw_c4x4, sw_c5x5, c_rf, two_k_rf, half_two_k_rf, sh_ewald;
    float lje_coeff2, lje_coeff6_6, sh_lj_ewald; /* LJ-PME: ewaldcoeff_lj^2, ewaldcoeff_lj^6 / 6, potential shift */
    float two_beta_over_sqrt_pi;                  /* energy kernels: coefficient of exp(-beta^2 r^2) in the real-space force */
};

/* energy kernels take the Ewald real-space force of pairs without exclusions from the erfc / exp of the energy
 * (-DNBNXM_PACKED_ENERGY_PMECORR: the rational correction everywhere, for A/B runs) */
#ifdef NBNXM_PACKED_ENERGY_PMECORR
constexpr bool c_ewaldForceFromErfc = false;
#else
constexpr bool c_ewaldForceFromErfc = true;
#endif

template<int ELEC, int VDW, bool ENERGY>
__device__ __forceinline__ void load_packed_consts(PackedConsts& k, const float* __restrict__ g)
{
    using Fl = Flavor<ELEC, VDW, ENERGY>;
    k.rc2    = __ldg(g + pcRc2);
    k.epsfac = __ldg(g + pcEpsfac);
    if (Fl::vdwCutoffCheck) k.rvdw2 = __ldg(g + pcRvdw2);
    if (Fl::ewaldAna)
    {
#pragma unroll
        for (int n = 0; n < 7; n++) k.num[n] = __ldg(g + pcNum0 + n);
#pragma unroll
        for (int n = 0; n < 5; n++) k.den[n] = __ldg(g + pcDen0 + n);
        if (ENERGY)
        {
            k.beta     = __ldg(g + pcBeta);
            k.sh_ewald = __ldg(g + pcShEwald);
            k.two_beta_over_sqrt_pi = __ldg(g + pcTwoBetaOverSqrtPi);
        }
    }
    if (ENERGY || Fl::ljPSwitch)
    {
        k.disp_cpot = __ldg(g + pcDispCpot);
        k.rep_cpot  = __ldg(g + pcRepCpot);
    }
    if (Fl::ljFSwitch || Fl::ljPSwitch) k.rvdw_switch = __ldg(g + pcRvdwSwitch);
    if (Fl::ljFSwitch)
    {
        k.disp_c2 = __ldg(g + pcDispC2);
        k.disp_c3 = __ldg(g + pcDispC3);
        k.rep_c2  = __ldg(g + pcRepC2);
        k.rep_c3  = __ldg(g + pcRepC3);
        if (ENERGY)
        {
            k.disp_c2_3 = __ldg(g + pcDispC2Third);
            k.disp_c3_4 = __ldg(g + pcDispC3Quarter);
            k.rep_c2_3  = __ldg(g + pcRepC2Third);
            k.rep_c3_4  = __ldg(g + pcRepC3Quarter);
        }
    }
    if (Fl::ljPSwitch)
    {
        k.sw_c3   = __ldg(g + pcSwC3);
        k.sw_c4   = __ldg(g + pcSwC4);
        k.sw_c5   = __ldg(g + pcSwC5);
        k.sw_c3x3 = __ldg(g + pcSwC3x3);
        k.sw_c4x4 = __ldg(g + pcSwC4x4);
        k.sw_c5x5 = __ldg(g + pcSwC5x5);
    }
    if (Fl::ljEwald)
    {
        k.lje_coeff2   = __ldg(g + pcLjeCoeff2);
        k.lje_coeff6_6 = __ldg(g + pcLjeCoeff6Sixth);
        if (ENERGY) k.sh_lj_ewald = __ldg(g + pcShLjEwald);
    }
    if (Fl::elecCut || Fl::elecRF) k.c_rf = __ldg(g + pcCrf);
    if (Fl::elecRF)
    {
        k.two_k_rf      = __ldg(g + pcTwoKrf);
        k.half_two_k_rf = __ldg(g + pcHalfTwoKrf);
    }
}

/* W = (F/r) r^2 and r^-2 of one pair (V = float) or two (V = f32x2); F/r = W r^-2 is left to the caller, which
 * masks it.  c6n = -6*C6, c12 = 12*C12, qq = qi*qj (with epsfac unless ENERGY); intBit = 1/0, only read when EXCL.
 * r2 already guarded against 0. */
template<typename V, int ELEC, int VDW, bool ENERGY, bool EXCL>
__device__ __forceinline__ V pair_w(const ParamsDev&    p,
                                    const PackedConsts& k,
                                    const V             r2,
                                    const V             qq,
                                    const V             c6n,
                                    const V             c12,
                                    const V             c6grid, /* LJ-PME only */
                                    const V             intBit,
                                    V&                  invR2out,
                                    V&                  eLJout,
                                    V&                  eElout)
{
    using Fl = Flavor<ELEC, VDW, ENERGY>;
    V invR   = vrsqrt(r2);
    if (ENERGY)
    {
        /* one Newton-Raphson step, see pair_force(); qq comes without epsfac in the energy kernels */
        invR = vmul(invR, vfma(vmul(vmul(r2, vbc<V>(-0.5f)), invR), invR, vbc<V>(1.5f)));
    }
    const V qqF   = ENERGY ? vmul(qq, vbc<V>(k.epsfac)) : qq;
    const V invR2 = vmul(invR, invR);
    V       invR6 = vmul(vmul(invR2, invR2), invR2);
    if (EXCL && Fl::exclusionForces)
    {
        invR6 = vmul(invR6, intBit);
    }
    /* LJ: invR6 (c12 invR6 - c6) */
    V W   = vmul(invR6, vfma(c12, invR6, c6n));
    V eLJ = vbc<V>(0.0f);
    if (ENERGY || Fl::ljPSwitch)
    {
        /* c12 (invR6^2 + rep.cpot) / 12 - c6 (invR6 + disp.cpot) / 6 */
        eLJ = vfma(vmul(c12, vbc<V>(c_oneTwelfth)), vfma(invR6, invR6, vbc<V>(k.rep_cpot)),
                   vmul(vmul(c6n, vbc<V>(c_oneSixth)), vadd(invR6, vbc<V>(k.disp_cpot))));
        if (EXCL && Fl::exclusionForces)
        {
            eLJ = vmul(eLJ, intBit);
        }
    }
    V r = vbc<V>(0.0f);
    if (Fl::ljFSwitch || Fl::ljPSwitch || (Fl::ewaldAna && ENERGY))
    {
        r = vmul(r2, invR);
    }
    if (Fl::ljEwald)
    {
        /* LJ-PME grid correction with the geometric grid coefficient (nbnxm_cuda_kernel_utils.cuh:194-245):
         * F/r += c6grid (r^-6 - exp(-c^2 r^2) (r^-6 (1 + c^2 r^2 + c^4 r^4 / 2) + c^6 / 6)) r^-2; excluded pairs get it too */
        const V invR6nm = vmul(vmul(invR2, invR2), invR2);
        const V cr2     = vmul(r2, vbc<V>(k.lje_coeff2));
        const V expmcr2 = vex2(vmul(cr2, vbc<V>(-1.4426950408889634f)));
        const V npoly   = vfma(vfma(cr2, vbc<V>(-0.5f), vbc<V>(-1.0f)), cr2, vbc<V>(-1.0f)); /* -(1 + cr2 + cr2^2 / 2) */
        W               = vfma(c6grid, vfma(expmcr2, vfma(invR6nm, npoly, vbc<V>(-k.lje_coeff6_6)), invR6nm), W);
        if (ENERGY)
        {
            /* c6grid / 6 (r^-6 (1 - exp(-c^2 r^2) poly) + sh_lj_ewald intBit) */
            const V sh = EXCL ? vmul(intBit, vbc<V>(k.sh_lj_ewald)) : vbc<V>(k.sh_lj_ewald);
            eLJ        = vfma(vmul(c6grid, vbc<V>(c_oneSixth)), vfma(invR6nm, vfma(expmcr2, npoly, vbc<V>(1.0f)), sh), eLJ);
        }
    }
    if (Fl::ljFSwitch || Fl::ljPSwitch)
    {
        const V rsw = vmax0(vsub(r, vbc<V>(k.rvdw_switch)));
        if (Fl::ljFSwitch)
        {
            /* F/r += (-c6 (d2 + d3 rsw) + c12 (r2 + r3 rsw)) rsw^2 / r, i.e. W += (...) rsw^2 r */
            const V disp = vfma(rsw, vbc<V>(k.disp_c3), vbc<V>(k.disp_c2));
            const V rep  = vfma(rsw, vbc<V>(k.rep_c3), vbc<V>(k.rep_c2));
            const V t    = vfma(c12, rep, vmul(c6n, disp));
            const V rsw2 = vmul(rsw, rsw);
            W            = vfma(vmul(t, rsw2), r, W);
            if (ENERGY)
            {
                /* (c6 (d2/3 + d3/4 rsw) - c12 (r2/3 + r3/4 rsw)) rsw^3, with c6n = -c6 */
                const V dispE = vfma(rsw, vbc<V>(-k.disp_c3_4), vbc<V>(-k.disp_c2_3));
                const V repE  = vfma(rsw, vbc<V>(-k.rep_c3_4), vbc<V>(-k.rep_c2_3));
                const V u     = vfma(c12, repE, vmul(c6n, dispE)); /* minus the switch energy per rsw^3 */
                eLJ           = vfma(u, vmul(rsw2, rsw), eLJ);
            }
        }
        else
        {
            /* potential switch needs the pair energy even in the force-only kernel */
            const V rsw2 = vmul(rsw, rsw);
            const V sw   = vfma(vmul(rsw2, rsw), vfma(vfma(rsw, vbc<V>(k.sw_c5), vbc<V>(k.sw_c4)), rsw, vbc<V>(k.sw_c3)), vbc<V>(1.0f));
            /* minus the derivative of the switch function */
            const V ndsw = vmul(rsw2, vfma(vfma(rsw, vbc<V>(-k.sw_c5x5), vbc<V>(-k.sw_c4x4)), rsw, vbc<V>(-k.sw_c3x3)));
            /* F/r = F/r sw - eLJ dsw / r, i.e. W = W sw - r eLJ dsw (rsw = 0 gives sw = 1, dsw = 0) */
            W   = vfma(W, sw, vmul(vmul(r, eLJ), ndsw));
            eLJ = vmul(eLJ, sw);
        }
    }
    if (Fl::vdwCutoffCheck)
    {
        const V inRange = vless(r2, k.rvdw2);
        W               = vmul(W, inRange);
        eLJ             = vmul(eLJ, inRange);
    }
    eLJout = eLJ;
    /* intBit * invR: the plain Coulomb part of excluded pairs is absent */
    const V invRi = (EXCL && (Fl::exclusionForces || ENERGY)) ? vmul(invR, intBit) : invR;
    V       eEl   = vbc<V>(0.0f);
    if (Fl::elecCut)
    {
        W = vfma(qqF, invRi, W);
        if (ENERGY) eEl = vmul(qq, vsub(invRi, vbc<V>(k.c_rf)));
    }
    if (Fl::elecRF)
    {
        /* F/r += qq (invR3 - 2k), i.e. W += qq (invR - 2k r2) */
        W = vfma(qqF, vfma(r2, vbc<V>(-k.two_k_rf), invRi), W);
        if (ENERGY) eEl = vmul(qq, vadd(invRi, vfma(r2, vbc<V>(k.half_two_k_rf), vbc<V>(-k.c_rf))));
    }
    if (Fl::ewaldAna && ENERGY && !EXCL && c_ewaldForceFromErfc)
    {
        /* Energy kernels, pairs without exclusions: erfc(beta r) and exp(-beta^2 r^2) are evaluated for the energy anyway, and
         * the real-space force is made of the same two: F/r = qq (erfc(beta r) / r + 2 beta / sqrt(pi) exp(-beta^2 r^2)) / r^2,
         * i.e. W += qq (erfc invR + 2 beta / sqrt(pi) exp).  Saves the rational correction (10 packed FMAs and a reciprocal per
         * two pairs); erfc_poly is good to 3e-7 absolute, which this term inherits relative to qq / r^3 (the force-only kernels
         * carry the 1e-7 of their approximate rsqrt in the same place).  Excluded pairs keep the correction form, which has no
         * cancellation at small r. */
        V           expm;
        const V     ec = erfc_poly(vmul(r, vbc<V>(k.beta)), expm);
        W              = vfma(qqF, vfma(ec, invR, vmul(expm, vbc<V>(k.two_beta_over_sqrt_pi))), W);
        eEl            = vmul(qq, vfma(invR, ec, vbc<V>(-k.sh_ewald)));
    }
    else if (Fl::ewaldAna)
    {
        /* F/r += qq (invR3 + beta^3 pmeCorrF(beta^2 r2)), i.e. W += qq (invR + r2 num(r2) / den(r2)) */
#ifdef NBNXM_PACKED_PLAIN_PMECORR
        V den = vfma(r2, vbc<V>(k.den[4]), vbc<V>(k.den[3]));
        den   = vfma(den, r2, vbc<V>(k.den[2]));
        den   = vfma(den, r2, vbc<V>(k.den[1]));
        den   = vfma(den, r2, vbc<V>(1.0f));
        V num = vfma(r2, vbc<V>(k.num[6]), vbc<V>(k.num[5]));
#else
        /* both polynomials divided by the leading coefficient of the numerator (host, fillParamsDev): its first Horner step
         * is an addition with ONE constant operand - an FFMA2 takes a single uniform-register operand, the second constant
         * of r2 num[6] + num[5] cost a move into a vector register per pair body */
        V den = vfma(r2, vbc<V>(k.den[4]), vbc<V>(k.den[3]));
        den   = vfma(den, r2, vbc<V>(k.den[2]));
        den   = vfma(den, r2, vbc<V>(k.den[1]));
        den   = vfma(den, r2, vbc<V>(k.den[0]));
        V num = vadd(r2, vbc<V>(k.num[5]));
#endif
        num   = vfma(num, r2, vbc<V>(k.num[4]));
        num   = vfma(num, r2, vbc<V>(k.num[3]));
        num   = vfma(num, r2, vbc<V>(k.num[2]));
        num   = vfma(num, r2, vbc<V>(k.num[1]));
        num   = vfma(num, r2, vbc<V>(k.num[0]));
        /* |den| >= 1 / |num[6]| > 1 (den >= 1 in the plain form): the plain approximate reciprocal needs no range fix-up */
        const V corr = vmul(num, vrcp(den));
        W            = vfma(qqF, vfma(corr, r2, invRi), W);
        if (ENERGY)
        {
            /* qq (invR (erfc(beta r) - (1 - intBit)) - intBit sh_ewald); excluded pairs get -erf(beta r)/r */
            V       expm;
            const V ec = erfc_poly(vmul(r, vbc<V>(k.beta)), expm);
            if (EXCL)
            {
                eEl = vmul(qq, vfma(invR, vadd(ec, vsub(intBit, vbc<V>(1.0f))), vmul(intBit, vbc<V>(-k.sh_ewald))));
            }
            else
            {
                eEl = vmul(qq, vfma(invR, ec, vbc<V>(-k.sh_ewald)));
            }
        }
    }
    eElout   = eEl;
    invR2out = invR2;
    return W;
}

/* Staging of the cjPacked groups of an entry: through the bulk-copy engine (cp.async.bulk + mbarrier, two-stage ring; the
 * default - north star: "staged in shared memory via TMA or cp.async.bulk") or, with -DNBNXM_PACKED_LDG_DESC, by coalesced
 * 16-byte loads + st.shared of all lanes.  A/B on a B200 (profiles/r02k_*): 12.3 M atoms 12 847.5 -> 12 835.9 us, 1.5 M atoms
 * 1096.4 -> 1096.2 us, 96 k atoms F+E 176.0 -> 173.2 us; long_scoreboard 0.12 -> 0.08 per issue, +0.7 % instructions. */
#if !defined(NBNXM_PACKED_LDG_DESC) && !defined(NBNXM_PACKED_TMA_DESC)
#    define NBNXM_PACKED_TMA_DESC 1
#endif
constexpr int c_descChunk = 64; /* cjPacked groups staged in shared memory at a time */
constexpr int c_fjRow = 9; /* float4 per j-atom slot in PackedShared::fj: 8 partial forces + 1 of padding */

/* shared memory of one CTA (= one warp = one sci entry) */
struct PackedShared
{
    float4 xqi[64];  /* i-atoms: shifted coordinates, charge (times epsfac in the force-only kernel) */
    float4 lji[64];  /* i-atoms: (c6, c12, -, -) combination parameters or (type * numTypes as int bits, -, -, -); same
                        16-byte stride as xqi so that one lane offset addresses both */
    /* j-atoms of the current cjPacked group as (half 0, half 1) pairs: entry u = 4*jm + jl holds atoms jl and jl+4
     * of j-cluster jm */
    float4 jxy[16];  /* xA xB yA yB */
    float4 jzq[16];  /* zA zB qA qB */
    float4 jlj[16];  /* c6A c6B c12A c12B, or typeA typeB - - */
    /* per-lane partial j forces of the current group: [j-atom slot][il], rows padded to 9 float4 so that both the
     * stores (8 consecutive float4 of one row per quarter warp) and the row sums (8 rows, same column, per
     * quarter warp) are bank-conflict free with plain offsets */
    float4 fj[32 * c_fjRow];
    /* type table, -6*C6 and 12*C12 */
    float nbC6n[c_packedMaxTypes * c_packedMaxTypes];
    float nbC12[c_packedMaxTypes * c_packedMaxTypes];
    /* the cjPacked groups of the current chunk of the sci entry, as in the list: cj[4], (imask, excl_ind) x 2 */
    uint4 desc[2 * c_descChunk];
#ifdef NBNXM_PACKED_TMA_DESC
    /* one mbarrier per half of desc: the halves form a two-stage ring filled by cp.async.bulk */
    unsigned long long descBar[2];
#endif
};

#ifdef NBNXM_PACKED_TMA_DESC
/* The cjPacked groups of an entry reach shared memory through the bulk-copy engine: cp.async.bulk global -> shared with
 * mbarrier completion, two stages of c_descStage groups, issued by one lane, the first two stages as soon as the entry is
 * known (before the i-atoms are staged). */
constexpr int c_descStage = c_descChunk / 2;
__device__ __forceinline__ void mbar_init(const unsigned bar, const unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load(const unsigned dst, const void* src, const unsigned bytes, const unsigned bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(const unsigned bar, const unsigned parity)
{
    asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "NBNXM_MBAR_WAIT:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra NBNXM_MBAR_DONE;\n"
            "bra NBNXM_MBAR_WAIT;\n"
            "NBNXM_MBAR_DONE:\n"
            "}\n" ::"r"(bar),
            "r"(parity)
            : "memory");
}
#endif

/* 32-bit shared-memory addresses and explicit ld/st.shared: with generic pointers into the shared struct ptxas
 * re-derives the shared window base (S2UR SR_CgaCtaId, ULEA) and the lane offsets in front of the accesses of every
 * loop iteration. */
__device__ __forceinline__ unsigned smem_u32(const void* p)
{
    return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ float4 lds128(const unsigned a)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint4 lds128u(const unsigned a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float2 lds64(const unsigned a)
{
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ int lds32i(const unsigned a)
{
    int v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float lds32f(const unsigned a)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(const unsigned a, const float x, const float y, const float z, const float w)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void sts128u(const unsigned a, const uint4 v)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts32(const unsigned a, const float x)
{
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(x) : "memory");
}

/* LJ parameters of an i-atom: two floats, four with LJ-PME (the grid factor) */
template<int ELEC, int VDW>
__device__ __forceinline__ float4 load_lji(const unsigned a)
{
    if (Flavor<ELEC, VDW, false>::ljEwald)
    {
        return lds128(a);
    }
    const float2 v = lds64(a);
    return make_float4(v.x, v.y, 0.0f, 0.0f);
}

/* byte distance of the two type tables in PackedShared */
constexpr unsigned c_nbC12FromC6n = sizeof(float) * c_packedMaxTypes * c_packedMaxTypes;

/* j-atoms of one j-cluster as this lane sees them: atoms jl (lo) and jl+4 (hi) */
struct PackedJ
{
    f32x2 x, y, z, q, lj0, lj1;
};

/* scalar partial j forces of the current j-cluster: atoms jl (A) and jl+4 (B) */
struct FjAcc
{
    float xA, yA, zA, xB, yB, zB;
};

/* LJ parameters of two pairs: c6n = -6*C6, c12 = 12*C12 */
template<int ELEC, int VDW>
__device__ __forceinline__ void lj_params_packed(const PackedShared& sm, const float4 pi, const PackedJ& j, f32x2& c6n, f32x2& c12, f32x2& c6grid)
{
    using Fl = Flavor<ELEC, VDW, false>;
    if (Fl::ljCombGeom)
    {
        c6n = vmul(pk(-pi.x, -pi.x), j.lj0);
        c12 = vmul(pk(pi.y, pi.y), j.lj1);
    }
    else if (Fl::ljCombLB)
    {
        const f32x2 sigma  = vadd(pk(pi.x, pi.x), j.lj0);
        const f32x2 eps    = vmul(pk(pi.y, pi.y), j.lj1);
        const f32x2 sigma2 = vmul(sigma, sigma);
        const f32x2 sigma6 = vmul(vmul(sigma2, sigma2), sigma2);
        const f32x2 c6     = vmul(eps, sigma6);
        c12                = vmul(c6, sigma6);
        c6n                = vsub(pk(0.0f, 0.0f), c6);
    }
    else
    {
        /* pi.x: shared-memory address of the i-type's row of nbC6n, j.lj0: 4 * j-type */
        const unsigned ia = __float_as_uint(pi.x) + __float_as_uint(lo(j.lj0)), ib = __float_as_uint(pi.x) + __float_as_uint(hi(j.lj0));
        c6n               = pk(lds32f(ia), lds32f(ib));
        c12               = pk(lds32f(ia + c_nbC12FromC6n), lds32f(ib + c_nbC12FromC6n));
    }
    /* LJ-PME, geometric grid coefficient: product of the per-atom factors (pi.z, j.lj1) */
    c6grid = Fl::ljEwaldGeom ? vmul(pk(pi.z, pi.z), j.lj1) : 0ull;
}

/* One (i-cluster, j-cluster) pair with BOTH halves in the list and no exclusion masks, force only: the hot body. */
template<int ELEC, int VDW, bool LISTBITS = false>
__device__ __forceinline__ void body_both(const ParamsDev&    p,
                                          const PackedConsts& k,
                                          const PackedShared& sm,
                                          const float4        xi,
                                          const float4        pi,
                                          const PackedJ&      j,
                                          float (&fi)[3],
                                          FjAcc& fj,
                                          const unsigned listed0 = 1u, /* LISTBITS: non-zero where the half is in the list */
                                          const unsigned listed1 = 1u)
{
    const f32x2 dx = vsub(pk(xi.x, xi.x), j.x), dy = vsub(pk(xi.y, xi.y), j.y), dz = vsub(pk(xi.z, xi.z), j.z);
    f32x2       r2;
    if (Flavor<ELEC, VDW, false>::ljPSwitch)
    {
        /* the potential switch evaluates r^-12 also in the force-only kernel: the reference's clamp keeps it finite */
        r2 = vfma(dz, dz, vfma(dy, dy, vmul(dx, dx)));
        r2 = pk(fmaxf(lo(r2), c_minDistanceSquared), fmaxf(hi(r2), c_minDistanceSquared));
    }
    else
    {
        r2 = vfma(dz, dz, vfma(dy, dy, vfma(dx, dx, pk(c_r2Guard, c_r2Guard))));
    }
    f32x2 c6n, c12, c6grid;
    lj_params_packed<ELEC, VDW>(sm, pi, j, c6n, c12, c6grid);
    f32x2       invR2, e0, e1;
    const f32x2 W  = pair_w<f32x2, ELEC, VDW, false, false>(p, k, r2, vmul(pk(xi.w, xi.w), j.q), c6n, c12, c6grid, 0ull, invR2, e0, e1);
    const f32x2 F  = vmul(W, invR2);
    const float F0 = LISTBITS ? sel_lt_listed(lo(r2), k.rc2, lo(F), listed0) : sel_lt(lo(r2), k.rc2, lo(F));
    const float F1 = LISTBITS ? sel_lt_listed(hi(r2), k.rc2, hi(F), listed1) : sel_lt(hi(r2), k.rc2, hi(F));
    fi[0]          = fmaf(F0, lo(dx), fmaf(F1, hi(dx), fi[0]));
    fi[1]          = fmaf(F0, lo(dy), fmaf(F1, hi(dy), fi[1]));
    fi[2]          = fmaf(F0, lo(dz), fmaf(F1, hi(dz), fi[2]));
    fj.xA          = fmaf(F0, lo(dx), fj.xA);
    fj.yA          = fmaf(F0, lo(dy), fj.yA);
    fj.zA          = fmaf(F0, lo(dz), fj.zA);
    fj.xB          = fmaf(F1, hi(dx), fj.xB);
    fj.yB          = fmaf(F1, hi(dy), fj.yB);
    fj.zB          = fmaf(F1, hi(dz), fj.zB);
}

/* One (i-cluster, j-cluster) pair with only half HALF in the list and no exclusion masks, force only: a scalar pair
 * body on that half's j-atom, so that a pruned half costs nothing. */
template<int ELEC, int VDW, int HALF>
__device__ __forceinline__ void body_single(const ParamsDev&    p,
                                            const PackedConsts& k,
                                            const PackedShared& sm,
                                            const float4        xi,
                                            const float4        pi,
                                            const PackedJ&      j,
                                            float (&fi)[3],
                                            FjAcc& fj)
{
    using Fl       = Flavor<ELEC, VDW, false>;
    const float xj = HALF ? hi(j.x) : lo(j.x), yj = HALF ? hi(j.y) : lo(j.y), zj = HALF ? hi(j.z) : lo(j.z);
    const float qj = HALF ? hi(j.q) : lo(j.q), l0 = HALF ? hi(j.lj0) : lo(j.lj0), l1 = HALF ? hi(j.lj1) : lo(j.lj1);
    const float dx = xi.x - xj, dy = xi.y - yj, dz = xi.z - zj;
    const float r2 = Fl::ljPSwitch ? fmaxf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)), c_minDistanceSquared)
                                   : fmaf(dz, dz, fmaf(dy, dy, fmaf(dx, dx, c_r2Guard)));
    float       c6n, c12;
    if (Fl::ljCombGeom)
    {
        c6n = -pi.x * l0;
        c12 = pi.y * l1;
    }
    else if (Fl::ljCombLB)
    {
        const float sigma = pi.x + l0, eps = pi.y * l1, sigma2 = sigma * sigma, sigma6 = sigma2 * sigma2 * sigma2;
        const float c6 = eps * sigma6;
        c12 = c6 * sigma6;
        c6n = -c6;
    }
    else
    {
        const unsigned ia = __float_as_uint(pi.x) + __float_as_uint(l0);
        c6n               = lds32f(ia);
        c12               = lds32f(ia + c_nbC12FromC6n);
    }
    float       invR2, e0, e1;
    const float c6grid = Fl::ljEwaldGeom ? pi.z * l1 : 0.0f;
    const float W = pair_w<float, ELEC, VDW, false, false>(p, k, r2, xi.w * qj, c6n, c12, c6grid, 0.0f, invR2, e0, e1);
    const float F = sel_lt(r2, k.rc2, W * invR2);
    fi[0]         = fmaf(F, dx, fi[0]);
    fi[1]         = fmaf(F, dy, fi[1]);
    fi[2]         = fmaf(F, dz, fi[2]);
    if (HALF)
    {
        fj.xB = fmaf(F, dx, fj.xB);
        fj.yB = fmaf(F, dy, fj.yB);
        fj.zB = fmaf(F, dz, fj.zB);
    }
    else
    {
        fj.xA = fmaf(F, dx, fj.xA);
        fj.yA = fmaf(F, dy, fj.yA);
        fj.zA = fmaf(F, dz, fj.zA);
    }
}

/* One (i-cluster, j-cluster) pair with one half in the list and no exclusion masks, with energies: a scalar pair body
 * on the j-atom of the half chosen at run time.  Returns the masked F/r and the distance vector. */
template<int ELEC, int VDW, bool ENERGY>
__device__ __forceinline__ float body_single_general(const ParamsDev&    p,
                                                     const PackedConsts& k,
                                                     const float4        xi,
                                                     const float4        pi,
                                                     const PackedJ&      j,
                                                     const bool          half1,
                                                     float&              dx,
                                                     float&              dy,
                                                     float&              dz,
                                                     float&              eLJacc,
                                                     float&              eElacc)
{
    using Fl       = Flavor<ELEC, VDW, ENERGY>;
    const float xj = half1 ? hi(j.x) : lo(j.x), yj = half1 ? hi(j.y) : lo(j.y), zj = half1 ? hi(j.z) : lo(j.z);
    const float qj = half1 ? hi(j.q) : lo(j.q), l0 = half1 ? hi(j.lj0) : lo(j.lj0), l1 = half1 ? hi(j.lj1) : lo(j.lj1);
    dx             = xi.x - xj;
    dy             = xi.y - yj;
    dz             = xi.z - zj;
    const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    float       c6n, c12;
    if (Fl::ljCombGeom)
    {
        c6n = -pi.x * l0;
        c12 = pi.y * l1;
    }
    else if (Fl::ljCombLB)
    {
        const float sigma = pi.x + l0, eps = pi.y * l1, sigma2 = sigma * sigma, sigma6 = sigma2 * sigma2 * sigma2;
        const float c6 = eps * sigma6;
        c12 = c6 * sigma6;
        c6n = -c6;
    }
    else
    {
        const unsigned ia = __float_as_uint(pi.x) + __float_as_uint(l0);
        c6n               = lds32f(ia);
        c12               = lds32f(ia + c_nbC12FromC6n);
    }
    float       invR2, ePairLJ, ePairEl;
    const float c6grid = Fl::ljEwaldGeom ? pi.z * l1 : 0.0f;
    const float W = pair_w<float, ELEC, VDW, ENERGY, false>(p, k, fmaxf(r2, c_minDistanceSquared), xi.w * qj, c6n, c12, c6grid, 1.0f, invR2,
                                                            ePairLJ, ePairEl);
    const bool  w = r2 < k.rc2;
    if (ENERGY)
    {
        eLJacc += w ? ePairLJ : 0.0f;
        eElacc += w ? ePairEl : 0.0f;
    }
    return w ? W * invR2 : 0.0f;
}

/* One (i-cluster, j-cluster) pair, general: list mask bits per half, optional exclusion masks, optional energies.
 * Returns the masked F/r of the two pairs and the distance vectors. */
template<int ELEC, int VDW, bool ENERGY, bool EXCL, bool RAW_ENERGIES = false>
__device__ __forceinline__ f32x2 body_general(const ParamsDev&    p,
                                              const PackedConsts& k,
                                              const PackedShared& sm,
                                              const float4        xi,
                                              const float4        pi,
                                              const PackedJ&      j,
                                              const bool          m0, /* list mask bits of the two halves */
                                              const bool          m1,
                                              const bool          i0, /* exclusion mask bits (EXCL only) */
                                              const bool          i1,
                                              const bool          self0, /* pair gets no exclusion correction (EXCL only) */
                                              const bool          self1,
                                              f32x2&              dx,
                                              f32x2&              dy,
                                              f32x2&              dz,
                                              f32x2&              eLJacc, /* RAW_ENERGIES: the pair energies, unmasked */
                                              f32x2&              eElacc,
                                              f32x2*              wmOut = nullptr) /* RAW_ENERGIES: the 1 / 0 masks of the pairs */
{
    using Fl = Flavor<ELEC, VDW, ENERGY>;
    dx       = vsub(pk(xi.x, xi.x), j.x);
    dy       = vsub(pk(xi.y, xi.y), j.y);
    dz       = vsub(pk(xi.z, xi.z), j.z);
    const f32x2 r2 = vfma(dz, dz, vfma(dy, dy, vmul(dx, dx)));
    bool        w0 = m0 && (lo(r2) < k.rc2);
    bool        w1 = m1 && (hi(r2) < k.rc2);
    f32x2       intBit = pk(1.0f, 1.0f);
    if (EXCL)
    {
        intBit = pk(i0 ? 1.0f : 0.0f, i1 ? 1.0f : 0.0f);
        if (Fl::exclusionForces)
        {
            w0 = w0 && !self0;
            w1 = w1 && !self1;
        }
        else
        {
            w0 = w0 && i0;
            w1 = w1 && i1;
        }
    }
    f32x2 c6n, c12, c6grid;
    lj_params_packed<ELEC, VDW>(sm, pi, j, c6n, c12, c6grid);
    /* clamped from both sides: the masking below is a multiplication, so pairs with filler atoms (parked at -1e6 nm)
     * must stay finite in the Ewald polynomials too; real pairs of a listed cluster pair are a few nm apart */
    const f32x2 r2c = pk(fminf(fmaxf(lo(r2), c_minDistanceSquared), c_maxDistanceSquared),
                         fminf(fmaxf(hi(r2), c_minDistanceSquared), c_maxDistanceSquared));
    f32x2       invR2, ePairLJ, ePairEl;
    const f32x2 W = pair_w<f32x2, ELEC, VDW, ENERGY, EXCL>(p, k, r2c, vmul(pk(xi.w, xi.w), j.q), c6n, c12, c6grid, intBit, invR2, ePairLJ, ePairEl);
    /* masking by multiplication: every quantity is finite (r2 is clamped), and the energy sums become FMAs */
    const f32x2 wm = pk(w0 ? 1.0f : 0.0f, w1 ? 1.0f : 0.0f);
    if (ENERGY && RAW_ENERGIES)
    {
        eLJacc = ePairLJ;
        eElacc = ePairEl;
        *wmOut = wm;
    }
    else if (ENERGY)
    {
        vfma_acc(eLJacc, ePairLJ, wm);
        vfma_acc(eElacc, ePairEl, wm);
    }
    return vmul(W, vmul(invR2, wm));
}

#if defined(NBNXM_PACKED_MAXNREG_ENERGY) || defined(NBNXM_PACKED_MAXNREG_FORCE)
#    ifndef NBNXM_PACKED_MAXNREG_ENERGY
#        define NBNXM_PACKED_MAXNREG_ENERGY 128
#    endif
#    ifndef NBNXM_PACKED_MAXNREG_FORCE
#        define NBNXM_PACKED_MAXNREG_FORCE 96
#    endif
template<int ELEC, int VDW, bool ENERGY>
__global__ void __maxnreg__(ENERGY ? NBNXM_PACKED_MAXNREG_ENERGY : NBNXM_PACKED_MAXNREG_FORCE)
        nbnxm_force_kernel_packed(const AtomDataDev ad, const ParamsDev p, const PairlistDev pl, const int calcFshift)
#else
template<int ELEC, int VDW, bool ENERGY>
__global__ void __launch_bounds__(32, ENERGY ? NBNXM_PACKED_MIN_BLOCKS_ENERGY : NBNXM_PACKED_MIN_BLOCKS)
        nbnxm_force_kernel_packed(const AtomDataDev ad, const ParamsDev p, const PairlistDev pl, const int calcFshift)
#endif
{
    using Fl                   = Flavor<ELEC, VDW, ENERGY>;
    constexpr unsigned c_full  = 0xffffffffu;
    constexpr bool     c_types = PackedFlavor<ELEC, VDW>::typeTable;
    /* Single-half cluster pairs: two unrolled scalar bodies per i-cluster in the force-only kernels, one loop body for all
     * i-clusters where the pair body is long (energies, potential switch): the unrolled form exceeds the 32 KB
     * instruction cache by too much (measured: potential switch 515 -> 433 us with the loop, LJ-PME 365 -> 386 us) */
#ifdef NBNXM_PACKED_NO_UNROLLED_SINGLE
    constexpr bool c_unrollSingle = false;
#else
    constexpr bool c_unrollSingle = !ENERGY && !Fl::ljPSwitch;
#endif

    const int lane = threadIdx.x;
    const int il   = lane & 7;
    const int jl   = lane >> 3;
    const nbnxm_b200_sci_t s = pl.sciSorted[blockIdx.x];

    __shared__ __align__(128) PackedShared sm;

    const float shx = ad.shiftVec[3 * s.shift], shy = ad.shiftVec[3 * s.shift + 1], shz = ad.shiftVec[3 * s.shift + 2];

    PackedConsts k;
    load_packed_consts<ELEC, VDW, ENERGY>(k, p.packedConsts);

    /* Lane-dependent shared-memory addresses, passed through a shuffle: ptxas would otherwise re-derive each of
     * them from the thread index (S2R, shifts, masks) in front of every use instead of keeping a register. */
    /* i-atom il of i-cluster 0 in xqi (lji: + sizeof(xqi)) */
    const unsigned xqiAddr = __shfl_sync(c_full, smem_u32(sm.xqi + il), lane);
    constexpr unsigned c_ljiFromXqi = sizeof(float4) * 64;
    /* this lane's entry of j-cluster 0 in jxy (jzq: + 256, jlj: + 512) */
    const unsigned jAddr0 = __shfl_sync(c_full, smem_u32(sm.jxy + jl), lane);
    /* where it parks its partial j forces of j-cluster 0: slot jl (half 1: slot jl + 4), column il */
    const unsigned parkAddr0 = __shfl_sync(c_full, smem_u32(sm.fj + jl * c_fjRow + il), lane);
    /* the j-atom slot it reduces (the atom it fetched) */
    const unsigned sumAddr = __shfl_sync(c_full, smem_u32(sm.fj + lane * c_fjRow), lane);
    /* where it stages the atom it fetched (lane = 8*jm + atom, atom = 4*half + jl'): a float of jxy (+8: y; jzq, jlj alike) */
    const unsigned stageAddr =
            __shfl_sync(c_full, smem_u32(reinterpret_cast<float*>(sm.jxy) + ((lane >> 3) * 4 + (lane & 3)) * 4 + ((lane >> 2) & 1)), lane);
    /* the cjPacked groups of the current chunk: cj[4], then (imask, excl_ind) x 2 */
    const unsigned descAddr = __shfl_sync(c_full, smem_u32(sm.desc), lane);
#ifdef NBNXM_PACKED_TMA_DESC
    const unsigned barAddr   = __shfl_sync(c_full, smem_u32(sm.descBar), lane);
    const int      numStages = (s.cj_packed_end - s.cj_packed_begin + c_descStage - 1) / c_descStage;
    auto           issueStage = [&](const int c) {
        const int begin = s.cj_packed_begin + c * c_descStage;
        const int n     = min(c_descStage, s.cj_packed_end - begin);
        bulk_load(descAddr + (c & 1) * (32 * c_descStage), pl.cjPacked + begin, 32u * n, barAddr + 8 * (c & 1));
    };
    if (lane == 0)
    {
        mbar_init(barAddr, 1);
        mbar_init(barAddr + 8, 1);
        mbar_fence_init();
        /* the first two stages are on their way while the i-atoms are staged */
        if (numStages > 0) issueStage(0);
        if (numStages > 1) issueStage(1);
    }
    __syncwarp();
#endif
    /* il and jl can be read back from the low bits of these addresses (the struct is 128-byte aligned): cheaper than
     * holding them in registers for the few places that need them */
#define NBNXM_IL_FROM_ADDR ((xqiAddr >> 4) & 7u)
#define NBNXM_JL_FROM_ADDR ((jAddr0 >> 4) & 3u)

    /* energies: float partial sums per j-cluster, double across the sci entry (see nbnxm_force_kernel) */
    double     eLJ = 0.0, eEl = 0.0;
    const bool diagonalEntry = (s.shift == c_centralShiftIndex && s.cj_packed_begin < s.cj_packed_end
                                && pl.cjPacked[s.cj_packed_begin].cj[0] == s.sci * c_superClusterSize);

    if (c_types)
    {
        const int nt2 = ad.numTypes * ad.numTypes;
        for (int n = lane; n < nt2; n += 32)
        {
            const float2 c = __ldg(p.nbfp + n);
            sm.nbC6n[n]    = -c.x;
            sm.nbC12[n]    = c.y;
        }
    }
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
        sm.xqi[lane + 32 * h] = v;
        if (Fl::ljComb)
        {
            const float2 c        = ad.ljComb[ai];
            sm.lji[lane + 32 * h] = make_float4(c.x, c.y, 0.0f, 0.0f);
        }
        else
        {
            const int   t      = ad.atomType[ai];
            const float c6grid = Fl::ljEwald ? __ldg(p.nbfpComb + t).x : 0.0f;
            sm.lji[lane + 32 * h] = make_float4(__uint_as_float(smem_u32(sm.nbC6n + t * ad.numTypes)), 0.0f, c6grid, 0.0f);
            if (ENERGY && Fl::ljEwald && diagonalEntry)
            {
                /* LJ-PME self term, once per diagonal sci entry (nbnxm_cuda_kernel.cuh:395-408) */
                eLJ += __ldg(p.nbfp + t * (ad.numTypes + 1)).x * 0.5f * c_oneSixth * k.lje_coeff6_6;
            }
        }
    }

    float fi[c_superClusterSize][3];
#pragma unroll
    for (int ci = 0; ci < c_superClusterSize; ci++)
    {
        fi[ci][0] = fi[ci][1] = fi[ci][2] = 0.0f;
    }

    const bool centralShift = (s.shift == c_centralShiftIndex);

    const uint4* cjGroups = reinterpret_cast<const uint4*>(pl.cjPacked);

    /* The cjPacked groups of the entry are staged in shared memory c_descChunk at a time (coalesced 16-byte loads),
     * so that masks and j-cluster indices of the coming groups are a shared-memory read away.  Within a chunk the
     * 32 j-atoms of a group - lane L fetches atom (L & 7) of j-cluster (L >> 3), one coalesced 16-byte load per lane -
     * and the exclusion mask words are fetched one group ahead. */
#ifdef NBNXM_PACKED_TMA_DESC
    for (int stageIdx = 0; stageIdx < numStages; stageIdx++)
    {
        const int      chunkBegin = s.cj_packed_begin + stageIdx * c_descStage;
        const int      numInChunk = min(c_descStage, s.cj_packed_end - chunkBegin);
        const unsigned descCur    = descAddr + (stageIdx & 1) * (32 * c_descStage);
        (void)cjGroups;
        mbar_wait(barAddr + 8 * (stageIdx & 1), (stageIdx >> 1) & 1);
#else
    for (int chunkBegin = s.cj_packed_begin; chunkBegin < s.cj_packed_end; chunkBegin += c_descChunk)
    {
        const int      numInChunk = min(c_descChunk, s.cj_packed_end - chunkBegin);
        const unsigned descCur    = descAddr;
        __syncwarp();
        for (int t = lane; t < 2 * numInChunk; t += 32)
        {
            sts128u(descAddr + 16 * t, cjGroups[2 * chunkBegin + t]);
        }
        __syncwarp();
#endif

        float4   xjNext = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        float2   pjNext = make_float2(0.0f, 0.0f);
        unsigned wex0Next = c_full, wex1Next = c_full;
        bool     fetchedNext = false;
        auto     fetchGroup = [&](const int g) {
            const uint4 me = lds128u(descCur + 32 * g + 16);
            /* unused slots of a partially filled group have no mask bits and an unspecified index */
            fetchedNext = (((me.x | me.z) >> (8u * NBNXM_JL_FROM_ADDR)) & 0xffu) != 0u;
            if (fetchedNext)
            {
                const int aj = lds32i(descCur + 32 * g + 4u * NBNXM_JL_FROM_ADDR) * c_clusterSize + NBNXM_IL_FROM_ADDR;
                xjNext       = ad.xqJ[aj];
                if (Fl::ljComb)
                {
                    pjNext = ad.ljComb[aj];
                }
                else
                {
                    const int t = ad.atomType[aj];
                    pjNext.x    = __int_as_float(4 * t);
                    if (Fl::ljEwald) pjNext.y = __ldg(p.nbfpComb + t).x;
                }
            }
            /* entry 0 of the exclusion array is all ones (pairlist.h:274-287) */
            wex0Next = c_full;
            wex1Next = c_full;
            if ((me.y | me.w) != 0u)
            {
                const unsigned exclLane = 8u * NBNXM_JL_FROM_ADDR + NBNXM_IL_FROM_ADDR;
                if (me.y != 0u) wex0Next = pl.excl[me.y].pair[exclLane];
                if (me.w != 0u) wex1Next = pl.excl[me.w].pair[exclLane];
            }
        };
        fetchGroup(0);

        for (int g = 0; g < numInChunk; g++)
        {
            const uint4    mev   = lds128u(descCur + 32 * g + 16);
            const unsigned wex0 = wex0Next, wex1 = wex1Next;
            const bool     fetched = fetchedNext;
            __syncwarp();
            sts32(stageAddr, xjNext.x);
            sts32(stageAddr + 8, xjNext.y);
            sts32(stageAddr + 256, xjNext.z);
            sts32(stageAddr + 256 + 8, xjNext.w);
            sts32(stageAddr + 512, pjNext.x);
            if (Fl::ljComb || Fl::ljEwald)
            {
                sts32(stageAddr + 512 + 8, pjNext.y);
            }
            __syncwarp();
            if (g + 1 < numInChunk)
            {
                fetchGroup(g + 1);
            }

            unsigned cur0 = mev.x, cur1 = mev.z;
            if ((cur0 | cur1) == 0u)
            {
                continue;
            }
            /* curEx gets the bits of the (j-cluster, i-cluster) pairs in which any atom pair of either half is excluded */
            unsigned curEx = 0u;
            if ((mev.y | mev.w) != 0u)
            {
                curEx = __reduce_or_sync(c_full, ~(wex0 & wex1));
            }

            unsigned jAddr = jAddr0, parkAddr = parkAddr0;
#pragma unroll 1
            for (int jm = 0; (cur0 | cur1) != 0u;
                 jm++, cur0 >>= 8, cur1 >>= 8, curEx >>= 8, jAddr += 64, parkAddr += 16 * c_clusterSize * c_fjRow)
            {
                const unsigned m0 = cur0 & 0xffu, m1 = cur1 & 0xffu;
                if ((m0 | m1) == 0u)
                {
                    continue;
                }
                const float4 xy = lds128(jAddr), zq = lds128(jAddr + 256), lj = lds128(jAddr + 512);
                PackedJ      j;
                j.x   = pk(xy.x, xy.y);
                j.y   = pk(xy.z, xy.w);
                j.z   = pk(zq.x, zq.y);
                j.q   = pk(zq.z, zq.w);
                j.lj0 = pk(lj.x, lj.y);
                j.lj1 = pk(lj.z, lj.w);
                FjAcc fj = { 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f };
                f32x2 eLJj = 0ull, eElj = 0ull;
                const unsigned mFast = (m0 | m1) & ~curEx;
                if (c_unrollSingle)
                {
                    const unsigned mBoth = m0 & m1;
#pragma unroll
                    for (int ci = 0; ci < c_superClusterSize; ci++)
                    {
                        if (mFast & (1u << ci))
                        {
                            const float4 xi = lds128(xqiAddr + ci * (16 * c_clusterSize));
                            const float4 pi = load_lji<ELEC, VDW>(xqiAddr + c_ljiFromXqi + ci * (16 * c_clusterSize));
                            if (__builtin_expect((mBoth & (1u << ci)) != 0u, 1))
                            {
                                body_both<ELEC, VDW>(p, k, sm, xi, pi, j, fi[ci], fj);
                            }
                            else if (m0 & (1u << ci))
                            {
                                body_single<ELEC, VDW, 0>(p, k, sm, xi, pi, j, fi[ci], fj);
                            }
                            else
                            {
                                body_single<ELEC, VDW, 1>(p, k, sm, xi, pi, j, fi[ci], fj);
                            }
                        }
                    }
                }
                else
                {
                    /* Kernels with a long pair body (energies, potential switch): every cluster pair without exclusions runs the
                     * packed body, a half that is not in the list masked like a pair beyond the cut-off (3 ... 5 instructions on top
                     * of 81 ... 103) - the loop body for cluster pairs with one half that stood here took 157 ... 162 instructions
                     * per trip for 32 pairs (-DNBNXM_PACKED_SINGLE_LOOP brings it back for A/B runs) */
#ifdef NBNXM_PACKED_SINGLE_LOOP
                    constexpr bool c_singleAsBoth = false;
#else
                    constexpr bool c_singleAsBoth = true;
#endif
                    const unsigned mBoth = c_singleAsBoth ? mFast : (m0 & m1 & mFast);
                    f32x2          pjx = 0ull, pjy = 0ull, pjz = 0ull;
                    float          eLJb = 0.0f, eElb = 0.0f;
#pragma unroll
                    for (int ci = 0; ci < c_superClusterSize; ci++)
                    {
                        if (mBoth & (1u << ci))
                        {
                            const float4 xi = lds128(xqiAddr + ci * (16 * c_clusterSize));
                            const float4 pi = load_lji<ELEC, VDW>(xqiAddr + c_ljiFromXqi + ci * (16 * c_clusterSize));
                            if (ENERGY)
                            {
                                f32x2       dx, dy, dz;
#ifdef NBNXM_PACKED_ENERGY_PACKED_ACC
                                const f32x2 F = body_general<ELEC, VDW, ENERGY, false>(p, k, sm, xi, pi, j, true, true, true, true,
                                                                                      false, false, dx, dy, dz, eLJj, eElj);
                                fi[ci][0] = fmaf(lo(F), lo(dx), fmaf(hi(F), hi(dx), fi[ci][0]));
                                fi[ci][1] = fmaf(lo(F), lo(dy), fmaf(hi(F), hi(dy), fi[ci][1]));
                                fi[ci][2] = fmaf(lo(F), lo(dz), fmaf(hi(F), hi(dz), fi[ci][2]));
                                vfma_acc(pjx, F, dx);
                                vfma_acc(pjy, F, dy);
                                vfma_acc(pjz, F, dz);
#else
                                /* scalar accumulators for the j forces and the energies of the j-cluster: ptxas gives packed
                                 * accumulators of this conditional chain a temporary and two moves per update (10 MOV per body) */
                                f32x2       eLJp, eElp, wm;
                                const bool  l0 = !c_singleAsBoth || ((m0 >> ci) & 1u) != 0u, l1 = !c_singleAsBoth || ((m1 >> ci) & 1u) != 0u;
                                const f32x2 F = body_general<ELEC, VDW, ENERGY, false, true>(p, k, sm, xi, pi, j, l0, l1, true, true,
                                                                                            false, false, dx, dy, dz, eLJp, eElp, &wm);
                                const float F0 = lo(F), F1 = hi(F);
                                fi[ci][0] = fmaf(F0, lo(dx), fmaf(F1, hi(dx), fi[ci][0]));
                                fi[ci][1] = fmaf(F0, lo(dy), fmaf(F1, hi(dy), fi[ci][1]));
                                fi[ci][2] = fmaf(F0, lo(dz), fmaf(F1, hi(dz), fi[ci][2]));
                                fj.xA     = fmaf(F0, lo(dx), fj.xA);
                                fj.yA     = fmaf(F0, lo(dy), fj.yA);
                                fj.zA     = fmaf(F0, lo(dz), fj.zA);
                                fj.xB     = fmaf(F1, hi(dx), fj.xB);
                                fj.yB     = fmaf(F1, hi(dy), fj.yB);
                                fj.zB     = fmaf(F1, hi(dz), fj.zB);
                                eLJb      = fmaf(lo(eLJp), lo(wm), fmaf(hi(eLJp), hi(wm), eLJb));
                                eElb      = fmaf(lo(eElp), lo(wm), fmaf(hi(eElp), hi(wm), eElb));
#endif
                            }
                            else
                            {
                                body_both<ELEC, VDW, c_singleAsBoth>(p, k, sm, xi, pi, j, fi[ci], fj, m0 & (1u << ci), m1 & (1u << ci));
                            }
                        }
                    }
#ifdef NBNXM_PACKED_ENERGY_PACKED_ACC
                    if (ENERGY)
                    {
                        fj.xA = lo(pjx);
                        fj.yA = lo(pjy);
                        fj.zA = lo(pjz);
                        fj.xB = hi(pjx);
                        fj.yB = hi(pjy);
                        fj.zB = hi(pjz);
                    }
#else
                    (void)pjx;
                    (void)pjy;
                    (void)pjz;
                    if (ENERGY)
                    {
                        eLJ += eLJb;
                        eEl += eElb;
                    }
#endif
                    /* cluster pairs with one half only (about one in four after pruning): a scalar body on that half,
                     * one loop body for all i-clusters (the unrolled form would not fit the instruction cache) */
                    unsigned mSingle = mFast & ~mBoth;
                    if (mSingle != 0u)
                    {
                        float eLJs = 0.0f, eEls = 0.0f;
#pragma unroll 1
                        for (; mSingle != 0u; mSingle &= mSingle - 1u)
                        {
                            const int    ci    = __ffs(mSingle) - 1;
                            const bool   half1 = ((m0 >> ci) & 1u) == 0u;
                            const float4 xi    = lds128(xqiAddr + ci * (16 * c_clusterSize));
                            const float4 pi    = load_lji<ELEC, VDW>(xqiAddr + c_ljiFromXqi + ci * (16 * c_clusterSize));
                            float        dx, dy, dz;
                            const float  F = body_single_general<ELEC, VDW, ENERGY>(p, k, xi, pi, j, half1, dx, dy, dz, eLJs, eEls);
                            const float  fx = F * dx, fy = F * dy, fz = F * dz;
                            switch (ci)
                            {
#define NBNXM_FI_CASE(c) \
    case c:              \
        fi[c][0] += fx;  \
        fi[c][1] += fy;  \
        fi[c][2] += fz;  \
        break;
                                NBNXM_FI_CASE(0)
                                NBNXM_FI_CASE(1)
                                NBNXM_FI_CASE(2)
                                NBNXM_FI_CASE(3)
                                NBNXM_FI_CASE(4)
                                NBNXM_FI_CASE(5)
                                NBNXM_FI_CASE(6)
                                NBNXM_FI_CASE(7)
#undef NBNXM_FI_CASE
                            }
                            if (half1)
                            {
                                fj.xB += fx;
                                fj.yB += fy;
                                fj.zB += fz;
                            }
                            else
                            {
                                fj.xA += fx;
                                fj.yA += fy;
                                fj.zA += fz;
                            }
                        }
                        eLJ += eLJs;
                        eEl += eEls;
                    }
                }
                /* the few cluster pairs with exclusion masks: one loop body for all i-clusters, own accumulators */
                unsigned mEx = (m0 | m1) & curEx;
                if (mEx != 0u)
                {
                    /* the i-cluster this j-cluster is, if any */
                    const int ciDiag = lds32i(descCur + 32 * g + 4 * jm) - s.sci * c_superClusterSize;
                    /* j <= i within the same cluster on the central shift: the "Newton" half of the diagonal cluster
                     * pair and the self pair (nbnxm_cuda_kernel.cuh:421-423) */
                    const bool selfLo = centralShift && NBNXM_JL_FROM_ADDR <= NBNXM_IL_FROM_ADDR;
                    const bool selfHi = centralShift && (NBNXM_JL_FROM_ADDR + 4u) <= NBNXM_IL_FROM_ADDR;
                    f32x2      gx = 0ull, gy = 0ull, gz = 0ull;
#pragma unroll 1
                    for (; mEx != 0u; mEx &= mEx - 1u)
                    {
                        const int    ci = __ffs(mEx) - 1;
                        const float4 xi = lds128(xqiAddr + ci * (16 * c_clusterSize));
                        const float4 pi = load_lji<ELEC, VDW>(xqiAddr + c_ljiFromXqi + ci * (16 * c_clusterSize));
                        const bool   onDiagonal = (ciDiag == ci);
                        const int    bit        = 8 * jm + ci;
                        f32x2        dx, dy, dz;
                        const f32x2  F = body_general<ELEC, VDW, ENERGY, true>(
                                p, k, sm, xi, pi, j, ((m0 >> ci) & 1u) != 0u, ((m1 >> ci) & 1u) != 0u, ((wex0 >> bit) & 1u) != 0u,
                                ((wex1 >> bit) & 1u) != 0u, onDiagonal && selfLo, onDiagonal && selfHi, dx, dy, dz, eLJj, eElj);
                        const float fx = fmaf(lo(F), lo(dx), hi(F) * hi(dx));
                        const float fy = fmaf(lo(F), lo(dy), hi(F) * hi(dy));
                        const float fz = fmaf(lo(F), lo(dz), hi(F) * hi(dz));
                        switch (ci)
                        {
#define NBNXM_FI_CASE(c) \
    case c:              \
        fi[c][0] += fx;  \
        fi[c][1] += fy;  \
        fi[c][2] += fz;  \
        break;
                            NBNXM_FI_CASE(0)
                            NBNXM_FI_CASE(1)
                            NBNXM_FI_CASE(2)
                            NBNXM_FI_CASE(3)
                            NBNXM_FI_CASE(4)
                            NBNXM_FI_CASE(5)
                            NBNXM_FI_CASE(6)
                            NBNXM_FI_CASE(7)
#undef NBNXM_FI_CASE
                        }
                        gx = vfma(F, dx, gx);
                        gy = vfma(F, dy, gy);
                        gz = vfma(F, dz, gz);
                    }
                    fj.xA += lo(gx);
                    fj.yA += lo(gy);
                    fj.zA += lo(gz);
                    fj.xB += hi(gx);
                    fj.yB += hi(gy);
                    fj.zB += hi(gz);
                }
                if (ENERGY)
                {
                    eLJ += lo(eLJj) + hi(eLJj);
                    eEl += lo(eElj) + hi(eElj);
                }
                /* park the partial j forces of both halves (sum of F d: the j-atom gets minus that, applied after the
                 * reduction); a half without list bits parks zeros */
                sts128(parkAddr, fj.xA, fj.yA, fj.zA, 0.0f);
                sts128(parkAddr + 16 * 4 * c_fjRow, fj.xB, fj.yB, fj.zB, 0.0f);
            }

            /* j forces of the group: lane L sums the 8 partial forces of j-atom slot L (the atom it fetched) and adds
             * them with one v4 reduction; slots of j-clusters that were not visited hold stale data and are skipped */
            __syncwarp();
            {
                if (fetched)
                {
                    /* x and y as one packed sum: two instead of three additions per partial force, same roundings */
                    const float4 v0  = lds128(sumAddr);
                    f32x2        sxy = pk(v0.x, v0.y);
                    float        sz  = v0.z;
#pragma unroll
                    for (int kx = 1; kx < c_clusterSize; kx++)
                    {
                        const float4 v = lds128(sumAddr + 16 * kx);
                        sxy            = vadd(sxy, pk(v.x, v.y));
                        sz += v.z;
                    }
                    /* the atom this lane fetched for the group */
                    const int ajOwn = lds32i(descCur + 32 * g + 4u * NBNXM_JL_FROM_ADDR) * c_clusterSize + NBNXM_IL_FROM_ADDR;
                    red_add_v4(ad.f4J + ajOwn, -lo(sxy), -hi(sxy), -sz);
                }
            }
        }
#ifdef NBNXM_PACKED_TMA_DESC
        /* every lane is done with this stage: refill it with the stage after the next */
        __syncwarp();
        if (lane == 0 && stageIdx + 2 < numStages) issueStage(stageIdx + 2);
#endif
    }

    /* i forces: reduce over the 4 jl-lanes, one v4 reduction per i-atom; shift force from the per-lane partial
     * sums (central shift skipped, nbnxm_cuda_kernel.cuh:697-717) */
    float fsx = 0.0f, fsy = 0.0f, fsz = 0.0f;
#pragma unroll
    for (int ci = 0; ci < c_superClusterSize; ci++)
    {
        float x = fi[ci][0], y = fi[ci][1], z = fi[ci][2];
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

#undef NBNXM_IL_FROM_ADDR
#undef NBNXM_JL_FROM_ADDR

} // namespace nbb

#endif
