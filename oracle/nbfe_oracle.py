"""TEST INFRASTRUCTURE — CPU restatement (numpy / pure Python, double precision) of the reference's perturbed
(free-energy) nonbonded pair kernel for the GPU: src/gromacs/nbnxm/cuda/nbfe_cuda_kernel.cuh:93-664, with the helpers
it calls (nbnxm_cuda_kernel_utils.cuh:73-98 sigma/epsilon and sigma6 conversions, nbnxm_kernel_utils.h:67-110 force
switch, :174-213 potential switch, :216-250 pmeCorrF).  It walks the atom-pair list the reference uploads with
gpu_init_feppairlist (iinr / jIndex / jjnr / shift / exclFep, nbnxm_gpu_data_mgmt.cpp) — one i-atom against its
j-atoms, both end states A and B evaluated per pair, Beutler soft-core with r-power 6.

Pinned by tests/test_oracle_fep.py against the reference's 312 golden files for this kernel
(nbnxm/tests/refdata/NBInteraction_NonbondedFepGpuTest_*.xml).  Small systems only (Python loops).
This is the oracle for SURVEY section 8f #4 (perturbed pair kernels); the CUDA kernel it will check is not built yet.

May be imported only by tests/.
"""
import math
from dataclasses import dataclass, field

import numpy as np

C_MIN_DISTANCE_SQUARED = 3.82e-07   # c_nbnxmMinDistanceSquared, pairlist.h:154
ONE_SIXTH = 1.0 / 6.0
ONE_TWELFTH = 1.0 / 12.0
CENTRAL_SHIFT = 22


@dataclass
class FepParams:
    """NBParamGpu fields the kernel reads (gpu_types_common.h:222-262 + the FEP block set by copy_gpu_fepparams)."""
    elec: str = "cut"               # 'cut' | 'rf' | 'ewald'
    vdw: str = "cut"                # 'cut' | 'cutgeom' | 'cutlb' | 'fswitch' | 'pswitch'
    twin: bool = False              # VDW_CUTOFF_CHECK
    epsfac: float = 1.0
    c_rf: float = 0.0
    two_k_rf: float = 0.0
    ewald_beta: float = 0.0
    sh_ewald: float = 0.0
    rcoulomb_sq: float = 1.0
    rvdw_sq: float = 1.0
    rvdw_switch: float = 0.0
    disp: tuple = (0.0, 0.0, 0.0)   # c2, c3, cpot
    rep: tuple = (0.0, 0.0, 0.0)
    sw: tuple = (0.0, 0.0, 0.0)     # c3, c4, c5
    alpha_coul: float = 0.0
    alpha_vdw: float = 0.0
    lambda_power: int = 1
    sigma6_with_invalid_sigma: float = 0.0
    sigma6_minimum: float = 0.0
    lambda_coul: float = 0.0
    lambda_vdw: float = 0.0
    nbfp: np.ndarray = field(default=None)       # [ntypes*ntypes, 2]: 6*C6, 12*C12
    ntypes: int = 0


def pme_corr_f(z2):
    """(2/sqrt(pi) z exp(-z^2) - erf(z)) / z^3: the function the reference's rational pmeCorrF approximates."""
    if z2 < 1e-8:
        return -4.0 / (3.0 * math.sqrt(math.pi)) + z2 * 4.0 / (5.0 * math.sqrt(math.pi))
    z = math.sqrt(z2)
    return (2.0 / math.sqrt(math.pi) * z * math.exp(-z2) - math.erf(z)) / (z * z2)


def _sigma6_from_c6_c12(c6, c12, sigma6_min, sigma6_def):
    if c6 > 0.0 and c12 > 0.0:
        return max(0.5 * c12 / c6, sigma6_min)
    return sigma6_def


def nbfe_forces(p: FepParams, x, q_a, q_b, type_a, type_b, lj_comb_a, lj_comb_b, shift_vec, iinr, jindex, jjnr, shift,
                excl_fep, calc_fshift=True):
    """Returns (f[natoms, 3], fshift[nshift, 3], e_lj, e_el, dvdl_lj, dvdl_el).  lj_comb_a / lj_comb_b: per-atom
    combination-rule parameters of the two end states ([natoms, 2]) for the 'cutgeom' / 'cutlb' flavors."""
    x = np.asarray(x, np.float64)
    natoms = x.shape[0]
    f = np.zeros((natoms, 3))
    fshift = np.zeros((np.asarray(shift_vec).reshape(-1, 3).shape[0], 3))
    shift_vec = np.asarray(shift_vec, np.float64).reshape(-1, 3)
    e_lj = e_el = dvdl_lj = dvdl_el = 0.0

    use_soft_core = p.alpha_vdw != 0.0
    lam_c, lam_v = p.lambda_coul, p.lambda_vdw
    lfac_c = (1.0 - lam_c, lam_c)
    lfac_v = (1.0 - lam_v, lam_v)
    dlfac = (-1.0, 1.0)
    rpower = 6.0
    sc_lfac_c, sc_lfac_v, sc_dl_c, sc_dl_v = [0, 0], [0, 0], [0, 0], [0, 0]
    for k in range(2):
        sq = p.lambda_power == 2
        sc_lfac_c[k] = (1.0 - lfac_c[k]) ** 2 if sq else (1.0 - lfac_c[k])
        sc_dl_c[k] = dlfac[k] * p.lambda_power / rpower * ((1.0 - lfac_c[k]) if sq else 1.0)
        sc_lfac_v[k] = (1.0 - lfac_v[k]) ** 2 if sq else (1.0 - lfac_v[k])
        sc_dl_v[k] = dlfac[k] * p.lambda_power / rpower * ((1.0 - lfac_v[k]) if sq else 1.0)

    rc2_coul = p.rcoulomb_sq
    rc2_max = max(rc2_coul, p.rvdw_sq) if p.twin else rc2_coul
    beta = p.ewald_beta
    comb = p.vdw in ("cutgeom", "cutlb")
    switch = p.vdw in ("fswitch", "pswitch")

    for n, ai in enumerate(iinr):
        xi = x[ai] + shift_vec[shift[n]]
        qi = (q_a[ai] * p.epsfac, q_b[ai] * p.epsfac)
        fci = np.zeros(3)
        for j in range(jindex[n], jindex[n + 1]):
            aj = jjnr[j]
            included = bool(excl_fep[j]) if excl_fep is not None else True
            qq = (qi[0] * q_a[aj], qi[1] * q_b[aj])
            rv = xi - x[aj]
            r2 = float(rv @ rv)
            if not (r2 < rc2_max) and included:
                continue
            r2 = max(r2, C_MIN_DISTANCE_SQUARED)
            inv_r = 1.0 / math.sqrt(r2)
            inv_r2 = inv_r * inv_r
            f_scalar = 0.0
            if included:
                if use_soft_core:
                    rpm2 = r2 * r2
                    rp = rpm2 * r2
                else:
                    rpm2 = inv_r2
                    rp = 1.0
                c6, c12, sigma6 = [0.0, 0.0], [0.0, 0.0], [0.0, 0.0]
                for k, (types, combs) in enumerate(((type_a, lj_comb_a), (type_b, lj_comb_b))):
                    if not comb:
                        c6[k], c12[k] = p.nbfp[p.ntypes * types[ai] + types[aj]]
                        if use_soft_core:
                            sigma6[k] = _sigma6_from_c6_c12(c6[k], c12[k], p.sigma6_minimum, p.sigma6_with_invalid_sigma)
                    elif p.vdw == "cutgeom":
                        c6[k] = combs[ai][0] * combs[aj][0]
                        c12[k] = combs[ai][1] * combs[aj][1]
                        if use_soft_core:
                            sigma6[k] = _sigma6_from_c6_c12(c6[k], c12[k], p.sigma6_minimum, p.sigma6_with_invalid_sigma)
                    else:
                        sigma = combs[ai][0] + combs[aj][0]
                        if combs[ai][0] == 0.0 or combs[aj][0] == 0.0:
                            sigma = 0.0
                        eps = combs[ai][1] * combs[aj][1]
                        s6 = (sigma * sigma) ** 3
                        c6[k] = eps * s6
                        c12[k] = c6[k] * s6
                        if use_soft_core:
                            if c6[k] > 0.0 and c12[k] > 0.0:
                                sigma6[k] = max(s6 * 0.5, p.sigma6_minimum)
                            else:
                                sigma6[k] = p.sigma6_with_invalid_sigma
                alpha_v_eff, alpha_c_eff = p.alpha_vdw, p.alpha_coul
                if use_soft_core and c12[0] > 0.0 and c12[1] > 0.0:
                    # soft-core only where one end state has no repulsion: it is there to avoid infinities
                    alpha_v_eff = alpha_c_eff = 0.0
                fs_c, fs_v, v_c, v_v = [0.0, 0.0], [0.0, 0.0], [0.0, 0.0], [0.0, 0.0]
                for k in range(2):
                    if qq[k] == 0.0 and c6[k] == 0.0 and c12[k] == 0.0:
                        continue
                    if use_soft_core:
                        rpinv_c = 1.0 / (alpha_c_eff * sc_lfac_c[k] * sigma6[k] + rp)
                        r2_c = rpinv_c ** (-1.0 / 3.0)
                        rinv_c = 1.0 / math.sqrt(r2_c)
                        if alpha_c_eff != alpha_v_eff or sc_lfac_v[k] != sc_lfac_c[k]:
                            rpinv_v = 1.0 / (alpha_v_eff * sc_lfac_v[k] * sigma6[k] + rp)
                            r2_v = rpinv_v ** (-1.0 / 3.0)
                            rinv_v = 1.0 / math.sqrt(r2_v)
                        else:
                            rpinv_v, r2_v, rinv_v = rpinv_c, r2_c, rinv_c
                    else:
                        rpinv_c, r2_c, rinv_c = 1.0, r2, inv_r
                        rpinv_v, r2_v, rinv_v = 1.0, r2, inv_r
                    if c6[k] != 0.0 or c12[k] != 0.0:
                        rinv6 = rpinv_v if use_soft_core else inv_r2 ** 3
                        v6 = c6[k] * rinv6
                        v12 = c12[k] * rinv6 * rinv6
                        fs_v[k] = v12 - v6
                        v_v[k] = (v12 + c12[k] * p.rep[2]) * ONE_TWELFTH - (v6 + c6[k] * p.disp[2]) * ONE_SIXTH
                        if switch:
                            r = r2_v * rinv_v
                            rsw = max(r - p.rvdw_switch, 0.0)
                            if p.vdw == "fswitch":
                                # ljForceSwitch<true, calcFr = true>, nbnxm_kernel_utils.h:67-110
                                fs_v[k] += (-c6[k] * (p.disp[0] + p.disp[1] * rsw) * rsw * rsw * r
                                            + c12[k] * (p.rep[0] + p.rep[1] * rsw) * rsw * rsw * r)
                                v_v[k] += (c6[k] * (p.disp[0] / 3.0 + p.disp[1] / 4.0 * rsw) * rsw ** 3
                                           - c12[k] * (p.rep[0] / 3.0 + p.rep[1] / 4.0 * rsw) * rsw ** 3)
                            else:
                                # ljPotentialSwitch<true, calcFr = true>, nbnxm_kernel_utils.h:174-213
                                sw = 1.0 + (p.sw[0] + (p.sw[1] + p.sw[2] * rsw) * rsw) * rsw ** 3
                                dsw = (3.0 * p.sw[0] + (4.0 * p.sw[1] + 5.0 * p.sw[2] * rsw) * rsw) * rsw * rsw
                                fs_v[k] = fs_v[k] * sw - r * v_v[k] * dsw
                                v_v[k] *= sw
                        if p.twin and not (r2 < p.rvdw_sq):
                            fs_v[k] = 0.0
                            v_v[k] = 0.0
                    if qq[k] != 0.0:
                        if p.elec == "rf":
                            fs_c[k] = qq[k] * (rinv_c - p.two_k_rf * r2_c)
                            v_c[k] = qq[k] * (rinv_c + 0.5 * p.two_k_rf * r2_c - p.c_rf)
                        elif p.elec == "cut":
                            fs_c[k] = qq[k] * rinv_c
                            v_c[k] = qq[k] * (rinv_c - p.c_rf)
                        else:
                            fs_c[k] = qq[k] * rinv_c
                            v_c[k] = qq[k] * (rinv_c - p.sh_ewald)
                    fs_c[k] *= rpinv_c
                    fs_v[k] *= rpinv_v
                for k in range(2):
                    e_el += lfac_c[k] * v_c[k]
                    e_lj += lfac_v[k] * v_v[k]
                    dvdl_el += v_c[k] * dlfac[k]
                    dvdl_lj += v_v[k] * dlfac[k]
                    if use_soft_core:
                        dvdl_el += lfac_c[k] * alpha_c_eff * sc_dl_c[k] * fs_c[k] * sigma6[k]
                        dvdl_lj += lfac_v[k] * alpha_v_eff * sc_dl_v[k] * fs_v[k] * sigma6[k]
                    f_scalar += lfac_c[k] * fs_c[k] * rpm2
                    f_scalar += lfac_v[k] * fs_v[k] * rpm2
            # excluded pairs: reaction-field / plain cut-off exclusion correction
            if p.elec in ("cut", "rf") and not included:
                if p.elec == "cut":
                    ff, vv = 0.0, -p.c_rf
                else:
                    ff, vv = -p.two_k_rf, 0.5 * p.two_k_rf * r2 - p.c_rf
                if ai == aj:
                    vv *= 0.5
                for k in range(2):
                    e_el += lfac_c[k] * qq[k] * vv
                    dvdl_el += dlfac[k] * qq[k] * vv
                    f_scalar += lfac_c[k] * qq[k] * ff
            # Ewald: the long-range part is removed for excluded pairs and for pairs inside the Coulomb cut-off
            if p.elec == "ewald" and (not included or r2 < rc2_coul):
                v_lr = inv_r * math.erf(r2 * inv_r * beta)
                if ai == aj:
                    v_lr *= 0.5
                f_lr = -pme_corr_f(beta * beta * r2) * beta ** 3
                for k in range(2):
                    e_el -= lfac_c[k] * qq[k] * v_lr
                    dvdl_el -= dlfac[k] * qq[k] * v_lr
                    f_scalar -= lfac_c[k] * qq[k] * f_lr
            if f_scalar != 0.0:
                fij = rv * f_scalar
                f[aj] -= fij
                fci += fij
        f[ai] += fci
        if calc_fshift and shift[n] != CENTRAL_SHIFT:
            fshift[shift[n]] += fci
    return f, fshift, e_lj, e_el, dvdl_lj, dvdl_el
