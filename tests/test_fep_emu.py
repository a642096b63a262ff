"""CPU tests of the perturbed (FEP) pair kernel's body (gromacs_b200/csrc/nbfe_bodies.h, float32, the code the CUDA
kernel nbnxm_fep.cu wraps) run pair by pair on the host (tests/kernel_emu/fep_emu.cpp) against the pinned oracle
(oracle/nbfe_oracle.py, double): the reference's 312 golden configurations, and a larger random perturbed system with
several i-entries, periodic shifts and exclusions in every flavor."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from test_oracle_fep import cases, fep_test_system

HERE = os.path.dirname(os.path.abspath(__file__))


class EmuParams(C.Structure):
    """nbfe::Params"""
    _fields_ = ([(n, C.c_int) for n in ("elec", "vdw", "twin")]
                + [(n, C.c_float) for n in ("epsfac", "c_rf", "two_k_rf", "beta", "sh_ewald", "rcoulomb_sq", "rvdw_sq", "rvdw_switch",
                                            "disp_c2", "disp_c3", "disp_cpot", "rep_c2", "rep_c3", "rep_cpot", "sw_c3", "sw_c4", "sw_c5",
                                            "alphaCoul", "alphaVdw", "sigma6WithInvalidSigma", "sigma6Minimum", "lambdaCoul", "lambdaVdw")]
                + [(n, C.c_int) for n in ("lambdaPower", "calcEnergy", "calcFshift", "numTypes", "calcForces")])


@pytest.fixture(scope="module")
def emu():
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "kernel_emu"), "libfep_emu.so"], check=True)
    return C.CDLL(os.path.join(HERE, "kernel_emu", "libfep_emu.so"))


def emu_params(p):
    e = EmuParams()
    e.elec = {"cut": 0, "rf": 1, "ewald": 2}[p.elec]
    e.vdw = {"cut": 0, "cutgeom": 1, "cutlb": 2, "fswitch": 3, "pswitch": 4}[p.vdw]
    e.twin = int(p.twin)
    e.epsfac, e.c_rf, e.two_k_rf, e.beta, e.sh_ewald = p.epsfac, p.c_rf, p.two_k_rf, p.ewald_beta, p.sh_ewald
    e.rcoulomb_sq, e.rvdw_sq, e.rvdw_switch = p.rcoulomb_sq, p.rvdw_sq, p.rvdw_switch
    e.disp_c2, e.disp_c3, e.disp_cpot = p.disp
    e.rep_c2, e.rep_c3, e.rep_cpot = p.rep
    e.sw_c3, e.sw_c4, e.sw_c5 = p.sw
    e.alphaCoul, e.alphaVdw = p.alpha_coul, p.alpha_vdw
    e.sigma6WithInvalidSigma, e.sigma6Minimum = p.sigma6_with_invalid_sigma, p.sigma6_minimum
    e.lambdaCoul, e.lambdaVdw, e.lambdaPower = p.lambda_coul, p.lambda_vdw, p.lambda_power
    e.calcEnergy = e.calcFshift = e.calcForces = 1
    e.numTypes = p.ntypes
    return e


def run_emu(emu, p, x, q_a, q_b, type_a, type_b, lj_a, lj_b, shift_vec, iinr, jindex, jjnr, shift, excl_fep, energies_only=False):
    n = len(x)
    xq = np.zeros((n, 4), np.float32)
    xq[:, :3] = x
    q = np.ascontiguousarray(np.stack([q_a, q_b], 1), np.float32)
    t = np.ascontiguousarray(np.stack([type_a, type_b], 1), np.int32)
    lj = np.ascontiguousarray(np.concatenate([lj_a, lj_b], 1), np.float32)
    nbfp = np.ascontiguousarray(p.nbfp, np.float32)
    sv = np.ascontiguousarray(shift_vec, np.float32).reshape(-1, 3)
    f4 = np.zeros((n, 4), np.float32)
    fsh = np.zeros((sv.shape[0], 3), np.float64)
    en, dv = np.zeros(2), np.zeros(2)
    ia = lambda a: np.ascontiguousarray(a, np.int32)
    iinr, jindex, jjnr, shift = ia(iinr), ia(jindex), ia(jjnr), ia(shift)
    ex = np.ascontiguousarray(excl_fep, np.uint8)
    ptr = lambda a, ct: a.ctypes.data_as(C.POINTER(ct))
    e = emu_params(p)
    if energies_only:
        e.calcForces = e.calcFshift = 0
    assert emu.fep_emu_run(C.byref(e), ptr(xq, C.c_float), ptr(q, C.c_float), ptr(t, C.c_int), ptr(lj, C.c_float), ptr(nbfp, C.c_float),
                           ptr(sv, C.c_float), C.c_int(len(iinr)), ptr(iinr, C.c_int), ptr(jindex, C.c_int), ptr(jjnr, C.c_int),
                           ptr(shift, C.c_int), ptr(ex, C.c_ubyte), ptr(f4, C.c_float), ptr(fsh, C.c_double), ptr(en, C.c_double),
                           ptr(dv, C.c_double)) == 0
    return f4[:, :3].astype(np.float64), fsh, en[0], en[1], dv[0], dv[1]


@pytest.mark.parametrize("name,ref", cases()[::3], ids=[c[0] for c in cases()[::3]])
def test_body_matches_oracle_on_the_reference_configurations(emu, name, ref):
    from oracle.nbfe_oracle import nbfe_forces
    p, x, q_a, q_b, type_a, type_b, lj_a, lj_b, lst = fep_test_system(name)
    want = nbfe_forces(p, x, q_a, q_b, type_a, type_b, lj_a, lj_b, np.zeros((1, 3)), **lst)
    got = run_emu(emu, p, x, q_a, q_b, type_a, type_b, lj_a, lj_b, np.zeros((1, 3)), **lst)
    scale = np.abs(want[0]).max()
    assert np.abs(got[0] - want[0]).max() <= 3e-6 * scale
    assert np.abs(got[1] - want[1]).max() <= 3e-6 * scale
    for g, w, floor in zip(got[2:], want[2:], (1.0, 100.0, 1.0, 100.0)):
        assert abs(g - w) <= 3e-6 * max(abs(w), floor), (name, g, w)


def random_perturbed_system(seed, natoms=60, ni=12):
    rng = np.random.default_rng(seed)
    box = np.array([2.4, 2.6, 2.8])
    x = rng.random((natoms, 3)) * box
    q_a = rng.uniform(-0.8, 0.8, natoms)
    q_b = np.where(rng.random(natoms) < 0.4, 0.0, q_a * rng.uniform(0.5, 1.5, natoms))
    type_a = rng.integers(0, 3, natoms)
    type_b = np.where(rng.random(natoms) < 0.5, rng.integers(0, 3, natoms), type_a)
    sv = np.zeros((45, 3))
    for z in (-1, 0, 1):
        for y in (-1, 0, 1):
            for xx in (-2, -1, 0, 1, 2):
                sv[((z + 1) * 3 + (y + 1)) * 5 + (xx + 2)] = (xx * box[0], y * box[1], z * box[2])
    iinr, jindex, jjnr, shift, excl = [], [0], [], [], []
    for n in range(ni):
        ai = int(rng.integers(0, natoms))
        iinr.append(ai)
        shift.append(int(rng.choice([22, 22, 21, 23, 17, 7, 37])))
        js = [a for a in rng.choice(natoms, size=int(rng.integers(1, 40)), replace=False).tolist() if a != ai]
        if n % 3 == 0 and shift[-1] == 22 and ai not in js:
            js.append(ai)                                 # the excluded self pair (central image only)
        for aj in js:
            jjnr.append(int(aj))
            # exclusions are between bonded neighbours: inside the cut-off, where the kernel's rational approximation
            # of the Ewald correction is valid (nbnxm_kernel_utils.h:216-250)
            r = np.linalg.norm(x[ai] + sv[shift[-1]] - x[aj])
            excl.append(0 if aj == ai else int(not (r < 0.9 and rng.random() < 0.5)))
        jindex.append(len(jjnr))
    return x, q_a, q_b, type_a, type_b, sv, dict(iinr=iinr, jindex=jindex, jjnr=jjnr, shift=shift, excl_fep=excl)


@pytest.mark.parametrize("elec", ["cut", "rf", "ewald"])
@pytest.mark.parametrize("vdw", ["cut", "cutgeom", "cutlb", "fswitch", "pswitch"])
@pytest.mark.parametrize("lam,alpha,power,twin", [(0.3, 0.5, 1, False), (0.7, 0.0, 1, True), (0.45, 0.3, 2, True)])
def test_body_matches_oracle_on_a_random_perturbed_system(emu, elec, vdw, lam, alpha, power, twin):
    import math
    from gromacs_b200.system import ewald_beta, force_switch_constants, potential_switch_constants
    from oracle.nbfe_oracle import FepParams, nbfe_forces
    x, q_a, q_b, type_a, type_b, sv, lst = random_perturbed_system(11)
    nt = 3
    sig = np.array([0.30, 0.0, 0.34])
    eps = np.array([0.6, 0.0, 0.3])
    nbfp = np.zeros((nt * nt, 2))
    comb = np.zeros((nt, 2))
    for i in range(nt):
        for j in range(nt):
            if sig[i] > 0 and sig[j] > 0:
                if vdw == "cutlb":
                    s, e = 0.5 * (sig[i] + sig[j]), math.sqrt(eps[i] * eps[j])
                else:
                    s, e = math.sqrt(sig[i] * sig[j]), math.sqrt(eps[i] * eps[j])
                nbfp[i * nt + j] = (6.0 * 4 * e * s ** 6, 12.0 * 4 * e * s ** 12)
        a, b = nbfp[i * nt + i]
        if a > 0:
            comb[i] = (math.sqrt(a), math.sqrt(b)) if vdw != "cutlb" else (0.5 * (b / a) ** (1.0 / 6.0), math.sqrt(a * a / b))
    rc, rvdw = 1.1, (0.9 if twin else 1.1)
    rsw = rvdw - 0.2
    p = FepParams(nbfp=nbfp, ntypes=nt, elec=elec, vdw=vdw, twin=twin, epsfac=138.935458, c_rf=0.7, two_k_rf=0.4,
                  ewald_beta=ewald_beta(rc, 1e-5), sh_ewald=0.01, rcoulomb_sq=rc * rc, rvdw_sq=rvdw * rvdw, rvdw_switch=rsw,
                  alpha_coul=alpha * 0.8, alpha_vdw=alpha, lambda_power=power, sigma6_with_invalid_sigma=0.3 ** 6,
                  sigma6_minimum=0.25 ** 6, lambda_coul=lam, lambda_vdw=min(1.0, lam + 0.1))
    if vdw == "fswitch":
        d, r = force_switch_constants(6.0, rsw, rvdw), force_switch_constants(12.0, rsw, rvdw)
        p.disp, p.rep = tuple(d), tuple(r)
    elif vdw == "pswitch":
        p.sw = tuple(potential_switch_constants(rsw, rvdw))
    else:
        p.disp, p.rep = (0.0, 0.0, -1.0 / rvdw ** 6), (0.0, 0.0, -1.0 / rvdw ** 12)
    lj_a, lj_b = comb[type_a], comb[type_b]
    want = nbfe_forces(p, x, q_a, q_b, type_a, type_b, lj_a, lj_b, sv, **lst)
    # the oracle gets the float32 roundings of the inputs the body sees
    x32 = x.astype(np.float32).astype(np.float64)
    want = nbfe_forces(p, x32, q_a.astype(np.float32).astype(np.float64), q_b.astype(np.float32).astype(np.float64), type_a,
                       type_b, lj_a.astype(np.float32).astype(np.float64), lj_b.astype(np.float32).astype(np.float64),
                       sv.astype(np.float32).astype(np.float64), **lst)
    got = run_emu(emu, p, x, q_a, q_b, type_a, type_b, lj_a, lj_b, sv, **lst)
    scale = np.abs(want[0]).max()
    assert np.abs(got[0] - want[0]).max() <= 2e-5 * scale                 # float32 pair arithmetic, close contacts
    assert np.abs(got[1] - want[1]).max() <= 2e-5 * scale
    for g, w in zip(got[2:], want[2:]):
        assert abs(g - w) <= 2e-5 * max(abs(w), 100.0), (g, w)
    assert np.abs(want[1]).max() > 0 and abs(want[4]) > 0


def test_energy_only_mode_is_the_foreign_lambda_evaluation(emu):
    """calcForces = 0 (nbnxm_b200_launch_foreign_energy_kernel, the reference's nbfe_foreign kernel): the energies and
    dV/dlambda of the full evaluation at that lambda, no forces"""
    import copy
    name = [c[0] for c in cases() if "PME" in c[0] and "0_5_0_3_scCoulomb_Yes" in c[0] and "ljrule_None" in c[0]][0]
    p, x, q_a, q_b, type_a, type_b, lj_a, lj_b, lst = fep_test_system(name)
    for lam in (0.0, 0.25, 0.8):
        q = copy.copy(p)
        q.lambda_coul, q.lambda_vdw = lam, min(1.0, lam + 0.1)
        full = run_emu(emu, q, x, q_a, q_b, type_a, type_b, lj_a, lj_b, np.zeros((1, 3)), **lst)
        only = run_emu(emu, q, x, q_a, q_b, type_a, type_b, lj_a, lj_b, np.zeros((1, 3)), energies_only=True, **lst)
        assert not only[0].any() and not only[1].any() and full[0].any()
        assert only[2:] == full[2:]
