/* Per-pair body of the perturbed (free-energy) nonbonded kernel, host+device (float32).
 *
 * Replaces nbfe_kernel_*_cuda (src/gromacs/nbnxm/cuda/nbfe_cuda_kernel.cuh:93-664): every atom pair of the perturbed
 * atom-pair list (gpu_init_feppairlist: iinr / jIndex / jjnr / shift / exclFep) is evaluated in both end states A and B
 * and mixed with the coupling parameters lambda (Coulomb and VdW separately), with the Beutler soft-core (r-power 6)
 * where one end state has no repulsion; excluded pairs get the reaction-field / Ewald exclusion correction.
 * Outputs: forces, shift forces, E_lj, E_el, dV/dlambda for VdW and Coulomb.
 *
 * One work item per pair (the reference runs one warp per i-entry): the perturbed lists are short (the perturbed
 * atoms against their neighbours), so the i-force reduction is left to the atomics.  The CUDA kernel (nbnxm_fep.cu) is
 * a thin wrapper; tests/kernel_emu/fep_emu.cpp runs the same body in a host loop against the pinned oracle
 * oracle/nbfe_oracle.py (test infrastructure only).
 */
#ifndef NBNXM_B200_NBFE_BODIES_H
#define NBNXM_B200_NBFE_BODIES_H

#include <math.h>

#if defined(__CUDACC__)
#    define NBFE_HD __host__ __device__ __forceinline__
#else
#    define NBFE_HD inline
#endif

namespace nbfe
{

constexpr float c_minDistanceSquared = 3.82e-07f; /* c_nbnxmMinDistanceSquared, pairlist.h:154 */
constexpr float c_oneSixth           = 0.16666667f;
constexpr float c_oneTwelfth         = 0.08333333f;
constexpr int   c_centralShift       = 22;

enum Elec
{
    ElecCut   = 0,
    ElecRF    = 1,
    ElecEwald = 2
};
enum Vdw
{
    VdwCut      = 0,
    VdwCombGeom = 1,
    VdwCombLB   = 2,
    VdwFSwitch  = 3,
    VdwPSwitch  = 4
};

struct Params
{
    int   elec, vdw, twin;
    float epsfac, c_rf, two_k_rf, beta, sh_ewald, rcoulomb_sq, rvdw_sq, rvdw_switch;
    float disp_c2, disp_c3, disp_cpot, rep_c2, rep_c3, rep_cpot, sw_c3, sw_c4, sw_c5;
    /* copy_gpu_fepparams */
    float alphaCoul, alphaVdw, sigma6WithInvalidSigma, sigma6Minimum, lambdaCoul, lambdaVdw;
    int   lambdaPower;
    int   calcEnergy, calcFshift;
    int   numTypes;
    int   calcForces; /* 0: energies / dV/dlambda only (the foreign-lambda evaluation, nbfe_foreign_cuda_kernel.cuh) */
};

struct Atoms
{
    const float* xq;       /* 4 per atom */
    const float* qAB;      /* 2 per atom: charge in state A, B */
    const int*   typeAB;   /* 2 per atom (type-table flavors) */
    const float* ljCombAB; /* 4 per atom: combination parameters of state A (x, y), B (x, y) */
    const float* nbfp;     /* 2 per type pair: 6*C6, 12*C12 */
    const float* shiftVec; /* 3 per shift */
    float*       f4;       /* 4 per atom, force accumulator */
    double*      fshift;   /* 3 per shift */
    double*      energy;   /* E_lj, E_el */
    double*      dvdl;     /* dV/dlambda: VdW, Coulomb */
};

struct List
{
    int                  numPairs;
    const int*           pairEntry; /* per j-entry: its i-entry */
    const int*           iinr;      /* per i-entry: i-atom */
    const int*           shift;     /* per i-entry: shift index */
    const int*           jjnr;      /* per j-entry: j-atom */
    const unsigned char* exclFep;   /* per j-entry: 1 = pair interacts, 0 = excluded (correction only); may be null */
};

NBFE_HD void addFloat(float* p, float v)
{
#if defined(__CUDA_ARCH__)
    atomicAdd(p, v);
#else
    *p += v;
#endif
}

NBFE_HD void addDouble(double* p, double v)
{
#if defined(__CUDA_ARCH__)
    atomicAdd(p, v);
#else
    *p += v;
#endif
}

/* rational approximation of (2/sqrt(pi) z exp(-z^2) - erf(z)) / z^3 in z^2 (pmeCorrF, nbnxm_kernel_utils.h:216-250) */
NBFE_HD float pmeCorrF(float z2)
{
    const float FN6 = -1.7357322914161492954e-8f, FN5 = 1.4703624142580877519e-6f, FN4 = -0.000053401640219807709149f,
                FN3 = 0.0010054721316683106153f, FN2 = -0.019278317264888380590f, FN1 = 0.069670166153766424023f,
                FN0 = -0.75225204789749321333f;
    const float FD4 = 0.0011193462567257629232f, FD3 = 0.014866955030185295499f, FD2 = 0.11583842382862377919f,
                FD1 = 0.50736591960530292870f, FD0 = 1.0f;
    const float z4  = z2 * z2;
    float       pd0 = FD4 * z4 + FD2;
    const float pd1 = FD3 * z4 + FD1;
    pd0             = pd0 * z4 + FD0;
    pd0             = pd1 * z2 + pd0;
    pd0             = 1.0f / pd0;
    float pn0       = FN6 * z4 + FN4;
    float pn1       = FN5 * z4 + FN3;
    pn0             = pn0 * z4 + FN2;
    pn1             = pn1 * z4 + FN1;
    pn0             = pn0 * z4 + FN0;
    pn0             = pn1 * z2 + pn0;
    return pn0 * pd0;
}

NBFE_HD float sigma6FromC6C12(float c6, float c12, float sigma6Min, float sigma6Def)
{
    if (c6 > 0.0f && c12 > 0.0f)
    {
        const float s = 0.5f * c12 / c6;
        return s < sigma6Min ? sigma6Min : s;
    }
    return sigma6Def;
}

/* j-entry `j` of the perturbed pair list */
NBFE_HD void pair(const Params& p, const Atoms& a, const List& l, int j)
{
    const int   n  = l.pairEntry[j];
    const int   ai = l.iinr[n];
    const int   aj = l.jjnr[j];
    const int   sh = l.shift[n];
    const bool  included = (l.exclFep == nullptr || l.exclFep[j] != 0);
    const float xi = a.xq[4 * ai] + a.shiftVec[3 * sh], yi = a.xq[4 * ai + 1] + a.shiftVec[3 * sh + 1],
                zi = a.xq[4 * ai + 2] + a.shiftVec[3 * sh + 2];
    const float rx = xi - a.xq[4 * aj], ry = yi - a.xq[4 * aj + 1], rz = zi - a.xq[4 * aj + 2];
    float       r2 = rx * rx + ry * ry + rz * rz;

    const float rc2Max = p.twin ? fmaxf(p.rcoulomb_sq, p.rvdw_sq) : p.rcoulomb_sq;
    if (!(r2 < rc2Max) && included)
    {
        return;
    }
    float qq[2];
    qq[0] = a.qAB[2 * ai] * p.epsfac * a.qAB[2 * aj];
    qq[1] = a.qAB[2 * ai + 1] * p.epsfac * a.qAB[2 * aj + 1];

    const float lambdaFactorCoul[2] = { 1.0f - p.lambdaCoul, p.lambdaCoul };
    const float lambdaFactorVdw[2]  = { 1.0f - p.lambdaVdw, p.lambdaVdw };
    const float dLambdaFactor[2]    = { -1.0f, 1.0f };
    const bool  useSoftCore         = (p.alphaVdw != 0.0f);

    r2                 = fmaxf(r2, c_minDistanceSquared);
    const float invR   = 1.0f / sqrtf(r2);
    const float invR2  = invR * invR;
    float       fScalar = 0.0f;
    float       eLJ = 0.0f, eEl = 0.0f, dvdlLJ = 0.0f, dvdlEl = 0.0f;

    if (included)
    {
        float rpm2, rp;
        if (useSoftCore)
        {
            rpm2 = r2 * r2;
            rp   = rpm2 * r2;
        }
        else
        {
            rpm2 = invR2;
            rp   = 1.0f;
        }
        float c6[2], c12[2], sigma6[2] = { 0.0f, 0.0f };
        for (int k = 0; k < 2; k++)
        {
            if (p.vdw == VdwCombGeom)
            {
                c6[k]  = a.ljCombAB[4 * ai + 2 * k] * a.ljCombAB[4 * aj + 2 * k];
                c12[k] = a.ljCombAB[4 * ai + 2 * k + 1] * a.ljCombAB[4 * aj + 2 * k + 1];
                if (useSoftCore)
                {
                    sigma6[k] = sigma6FromC6C12(c6[k], c12[k], p.sigma6Minimum, p.sigma6WithInvalidSigma);
                }
            }
            else if (p.vdw == VdwCombLB)
            {
                const float si = a.ljCombAB[4 * ai + 2 * k], sj = a.ljCombAB[4 * aj + 2 * k];
                const float sigma   = (si == 0.0f || sj == 0.0f) ? 0.0f : si + sj;
                const float epsilon = a.ljCombAB[4 * ai + 2 * k + 1] * a.ljCombAB[4 * aj + 2 * k + 1];
                const float sigma2  = sigma * sigma;
                const float s6      = sigma2 * sigma2 * sigma2;
                c6[k]               = epsilon * s6;
                c12[k]              = c6[k] * s6;
                if (useSoftCore)
                {
                    sigma6[k] = (c6[k] > 0.0f && c12[k] > 0.0f) ? fmaxf(s6 * 0.5f, p.sigma6Minimum) : p.sigma6WithInvalidSigma;
                }
            }
            else
            {
                const int t = p.numTypes * a.typeAB[2 * ai + k] + a.typeAB[2 * aj + k];
                c6[k]       = a.nbfp[2 * t];
                c12[k]      = a.nbfp[2 * t + 1];
                if (useSoftCore)
                {
                    sigma6[k] = sigma6FromC6C12(c6[k], c12[k], p.sigma6Minimum, p.sigma6WithInvalidSigma);
                }
            }
        }
        float alphaVdwEff = p.alphaVdw, alphaCoulEff = p.alphaCoul;
        if (useSoftCore && c12[0] > 0.0f && c12[1] > 0.0f)
        {
            /* soft-core only where an end state has no repulsion: it is there to avoid infinities */
            alphaVdwEff  = 0.0f;
            alphaCoulEff = 0.0f;
        }
        float softcoreLambdaFactorCoul[2], softcoreLambdaFactorVdw[2], softcoreDlFactorCoul[2], softcoreDlFactorVdw[2];
        for (int k = 0; k < 2; k++)
        {
            const bool  sq               = (p.lambdaPower == 2);
            softcoreLambdaFactorCoul[k] = sq ? (1.0f - lambdaFactorCoul[k]) * (1.0f - lambdaFactorCoul[k]) : (1.0f - lambdaFactorCoul[k]);
            softcoreDlFactorCoul[k]     = dLambdaFactor[k] * p.lambdaPower / 6.0f * (sq ? (1.0f - lambdaFactorCoul[k]) : 1.0f);
            softcoreLambdaFactorVdw[k]  = sq ? (1.0f - lambdaFactorVdw[k]) * (1.0f - lambdaFactorVdw[k]) : (1.0f - lambdaFactorVdw[k]);
            softcoreDlFactorVdw[k]      = dLambdaFactor[k] * p.lambdaPower / 6.0f * (sq ? (1.0f - lambdaFactorVdw[k]) : 1.0f);
        }
        float fsCoul[2] = { 0.0f, 0.0f }, fsVdw[2] = { 0.0f, 0.0f }, vCoul[2] = { 0.0f, 0.0f }, vVdw[2] = { 0.0f, 0.0f };
        for (int k = 0; k < 2; k++)
        {
            if (qq[k] == 0.0f && c6[k] == 0.0f && c12[k] == 0.0f)
            {
                continue;
            }
            float rPInvC, r2C, rInvC, rPInvV, r2V, rInvV;
            if (useSoftCore)
            {
                rPInvC = 1.0f / (alphaCoulEff * softcoreLambdaFactorCoul[k] * sigma6[k] + rp);
                r2C    = 1.0f / cbrtf(rPInvC);
                rInvC  = 1.0f / sqrtf(r2C);
                if (alphaCoulEff != alphaVdwEff || softcoreLambdaFactorVdw[k] != softcoreLambdaFactorCoul[k])
                {
                    rPInvV = 1.0f / (alphaVdwEff * softcoreLambdaFactorVdw[k] * sigma6[k] + rp);
                    r2V    = 1.0f / cbrtf(rPInvV);
                    rInvV  = 1.0f / sqrtf(r2V);
                }
                else
                {
                    rPInvV = rPInvC;
                    r2V    = r2C;
                    rInvV  = rInvC;
                }
            }
            else
            {
                rPInvC = 1.0f;
                r2C    = r2;
                rInvC  = invR;
                rPInvV = 1.0f;
                r2V    = r2;
                rInvV  = invR;
            }
            if (c6[k] != 0.0f || c12[k] != 0.0f)
            {
                const float rInv6  = useSoftCore ? rPInvV : invR2 * invR2 * invR2;
                const float vVdw6  = c6[k] * rInv6;
                const float vVdw12 = c12[k] * rInv6 * rInv6;
                fsVdw[k]           = vVdw12 - vVdw6;
                vVdw[k] = (vVdw12 + c12[k] * p.rep_cpot) * c_oneTwelfth - (vVdw6 + c6[k] * p.disp_cpot) * c_oneSixth;
                if (p.vdw == VdwFSwitch)
                {
                    /* ljForceSwitch<energies, F*r>, nbnxm_kernel_utils.h:67-110 */
                    const float r   = r2V * rInvV;
                    const float rsw = fmaxf(r - p.rvdw_switch, 0.0f);
                    fsVdw[k] += -c6[k] * (p.disp_c2 + p.disp_c3 * rsw) * rsw * rsw * r
                                + c12[k] * (p.rep_c2 + p.rep_c3 * rsw) * rsw * rsw * r;
                    vVdw[k] += c6[k] * (p.disp_c2 / 3.0f + p.disp_c3 / 4.0f * rsw) * rsw * rsw * rsw
                               - c12[k] * (p.rep_c2 / 3.0f + p.rep_c3 / 4.0f * rsw) * rsw * rsw * rsw;
                }
                else if (p.vdw == VdwPSwitch)
                {
                    /* ljPotentialSwitch<energies, F*r>, nbnxm_kernel_utils.h:174-213 */
                    const float r   = r2V * rInvV;
                    const float rsw = r - p.rvdw_switch;
                    if (rsw > 0.0f)
                    {
                        const float sw  = 1.0f + (p.sw_c3 + (p.sw_c4 + p.sw_c5 * rsw) * rsw) * rsw * rsw * rsw;
                        const float dsw = (3.0f * p.sw_c3 + (4.0f * p.sw_c4 + 5.0f * p.sw_c5 * rsw) * rsw) * rsw * rsw;
                        fsVdw[k]        = fsVdw[k] * sw - r * vVdw[k] * dsw;
                        vVdw[k] *= sw;
                    }
                }
                if (p.twin && !(r2 < p.rvdw_sq))
                {
                    fsVdw[k] = 0.0f;
                    vVdw[k]  = 0.0f;
                }
            }
            if (qq[k] != 0.0f)
            {
                if (p.elec == ElecRF)
                {
                    fsCoul[k] = qq[k] * (rInvC - p.two_k_rf * r2C);
                    vCoul[k]  = qq[k] * (rInvC + 0.5f * p.two_k_rf * r2C - p.c_rf);
                }
                else if (p.elec == ElecCut)
                {
                    fsCoul[k] = qq[k] * rInvC;
                    vCoul[k]  = qq[k] * (rInvC - p.c_rf);
                }
                else
                {
                    fsCoul[k] = qq[k] * rInvC;
                    vCoul[k]  = qq[k] * (rInvC - p.sh_ewald);
                }
            }
            fsCoul[k] *= rPInvC;
            fsVdw[k] *= rPInvV;
        }
        for (int k = 0; k < 2; k++)
        {
            eEl += lambdaFactorCoul[k] * vCoul[k];
            eLJ += lambdaFactorVdw[k] * vVdw[k];
            dvdlEl += vCoul[k] * dLambdaFactor[k];
            dvdlLJ += vVdw[k] * dLambdaFactor[k];
            if (useSoftCore)
            {
                dvdlEl += lambdaFactorCoul[k] * alphaCoulEff * softcoreDlFactorCoul[k] * fsCoul[k] * sigma6[k];
                dvdlLJ += lambdaFactorVdw[k] * alphaVdwEff * softcoreDlFactorVdw[k] * fsVdw[k] * sigma6[k];
            }
            fScalar += lambdaFactorCoul[k] * fsCoul[k] * rpm2;
            fScalar += lambdaFactorVdw[k] * fsVdw[k] * rpm2;
        }
    }

    /* excluded pairs: reaction-field / plain cut-off exclusion correction */
    if ((p.elec == ElecCut || p.elec == ElecRF) && !included)
    {
        float ff, vv;
        if (p.elec == ElecCut)
        {
            ff = 0.0f;
            vv = -p.c_rf;
        }
        else
        {
            ff = -p.two_k_rf;
            vv = 0.5f * p.two_k_rf * r2 - p.c_rf;
        }
        if (ai == aj)
        {
            vv *= 0.5f;
        }
        for (int k = 0; k < 2; k++)
        {
            eEl += lambdaFactorCoul[k] * qq[k] * vv;
            dvdlEl += dLambdaFactor[k] * qq[k] * vv;
            fScalar += lambdaFactorCoul[k] * qq[k] * ff;
        }
    }
    /* Ewald: the long-range part is removed for excluded pairs and for pairs inside the Coulomb cut-off */
    if (p.elec == ElecEwald && (!included || r2 < p.rcoulomb_sq))
    {
        float vLr = invR * erff(r2 * invR * p.beta);
        if (ai == aj)
        {
            vLr *= 0.5f;
        }
        const float fLr = -pmeCorrF(p.beta * p.beta * r2) * p.beta * p.beta * p.beta;
        for (int k = 0; k < 2; k++)
        {
            eEl -= lambdaFactorCoul[k] * qq[k] * vLr;
            dvdlEl -= dLambdaFactor[k] * qq[k] * vLr;
            fScalar -= lambdaFactorCoul[k] * qq[k] * fLr;
        }
    }

    if (p.calcForces && fScalar != 0.0f)
    {
        const float fx = rx * fScalar, fy = ry * fScalar, fz = rz * fScalar;
        addFloat(&a.f4[4 * aj], -fx);
        addFloat(&a.f4[4 * aj + 1], -fy);
        addFloat(&a.f4[4 * aj + 2], -fz);
        addFloat(&a.f4[4 * ai], fx);
        addFloat(&a.f4[4 * ai + 1], fy);
        addFloat(&a.f4[4 * ai + 2], fz);
        if (p.calcFshift && sh != c_centralShift)
        {
            addDouble(&a.fshift[3 * sh], fx);
            addDouble(&a.fshift[3 * sh + 1], fy);
            addDouble(&a.fshift[3 * sh + 2], fz);
        }
    }
    if (p.calcEnergy)
    {
        if (eLJ != 0.0f) addDouble(&a.energy[0], eLJ);
        if (eEl != 0.0f) addDouble(&a.energy[1], eEl);
        if (dvdlLJ != 0.0f) addDouble(&a.dvdl[0], dvdlLJ);
        if (dvdlEl != 0.0f) addDouble(&a.dvdl[1], dvdlEl);
    }
}

} // namespace nbfe

#endif
