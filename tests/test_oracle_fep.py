"""CPU test: pins the perturbed-pair (FEP) oracle (oracle/nbfe_oracle.py) against the reference's own golden data
for its GPU FEP kernel — all 312 files of NonbondedFepGpuTest (tests/golden/refdata/fep_gpu.npz, made by
tests/golden/make_golden_fep.py): a 4-atom system, one i-atom against an excluded self pair and three perturbed
pairs (Coulomb- and/or VdW-perturbed), for 3 Coulomb types x 3 VdW modifiers x 3 LJ combination rules x
lambda in {0, 0.5, 1} x soft-core alpha in {0, 0.3} x soft-core Coulomb on/off
(src/gromacs/nbnxm/tests/freeenergygpukernel.cpp:330-372, 690-840)."""
import math
import os

import numpy as np
import pytest

from util import GOLDEN

ONE_4PI_EPS0 = 138.93545764438198          # c_one4PiEps0, src/gromacs/math/include/gromacs/math/units.h
LAMBDA_ALPHA = {"0_0": (0.0, 0.0), "0_0_3": (0.0, 0.3), "0_5_0": (0.5, 0.0), "0_5_0_3": (0.5, 0.3), "1_0": (1.0, 0.0),
                "1_0_3": (1.0, 0.3)}


def cases():
    d = np.load(os.path.join(GOLDEN, "refdata", "fep_gpu.npz"))
    return list(zip(d["names"].tolist(), d["values"]))


def fep_test_system(name):
    """AtomData / ForcerecHelper / InteractionConstHelper of the reference test, as plain arrays"""
    from gromacs_b200.system import ewald_beta
    from oracle.nbfe_oracle import FepParams
    coul, rest = name[len("coul_"):].split("_vdw_Cut_off_vdwmod_")
    mod, rest = rest.split("_ljrule_")
    rule, rest = rest.split("_coords_A_")
    la, sc = rest.split("_scCoulomb_")
    lam, alpha = LAMBDA_ALPHA[la]
    sc_coul = sc == "Yes"
    c6, c12 = 0.001458, 1.0062882e-6
    lj = {(0, 0): (c6, c12), (0, 2): (c6, c12), (2, 0): (c6, c12), (2, 2): (c6, c12)}
    nt = 3
    nbfp = np.zeros((nt * nt, 2))
    for (i, j), (a, b) in lj.items():
        nbfp[i * nt + j] = (6.0 * a, 12.0 * b)                        # makeNonBondedParameterLists + the 6 / 12 prefactors
    p = FepParams(nbfp=nbfp, ntypes=nt)
    p.elec = {"Cut_off": "cut", "Reaction_Field": "rf", "PME": "ewald"}[coul]
    # nbnxmGpuPickVdwKernelType (nbnxm_gpu_data_mgmt.cpp:368-396): force switch ignores the combination rule
    p.vdw = "fswitch" if mod == "Force_switch" else {"None": "cut", "Geometric": "cutgeom", "Lorentz_Berthelot": "cutlb"}[rule]
    p.epsfac = ONE_4PI_EPS0 * 0.25
    p.c_rf, p.two_k_rf = 1.0, 0.0
    p.ewald_beta = ewald_beta(1.0, 1.0e-5)
    p.sh_ewald = 1.0e-5
    p.rcoulomb_sq = p.rvdw_sq = 1.0
    p.rvdw_switch = 0.0
    p.disp = p.rep = (0.0, 0.0, -1.0)
    p.alpha_vdw = alpha
    p.alpha_coul = alpha if sc_coul else 0.0
    p.lambda_power = 1
    p.sigma6_with_invalid_sigma = 0.3 ** 6
    p.sigma6_minimum = 0.3 ** 6 if sc_coul else 0.0
    p.lambda_coul = p.lambda_vdw = lam
    x = np.array([[1.0, 1.0, 1.0], [1.1, 1.15, 1.2], [0.9, 0.85, 0.8], [1.1, 1.15, 0.8]])
    q_a, q_b = [1.0, -1.0, -1.0, 1.0], [1.0, 0.0, 0.0, 1.0]
    type_a, type_b = [0, 0, 0, 0], [0, 1, 2, 1]
    # per-type combination parameters (set_lj_parameter_data, atomdata.cpp:365-395), then per atom and end state
    comb = np.zeros((nt, 2))
    for t in range(nt):
        a, b = nbfp[t * nt + t]
        if rule == "Geometric":
            comb[t] = (math.sqrt(a), math.sqrt(b))
        elif rule == "Lorentz_Berthelot" and a > 0 and b > 0:
            comb[t] = (0.5 * (b / a) ** (1.0 / 6.0), math.sqrt(a * a / b))
    lj_a, lj_b = comb[type_a], comb[type_b]
    lst = dict(iinr=[0], jindex=[0, 4], jjnr=[0, 1, 2, 3], shift=[0], excl_fep=[False, True, True, True])
    return p, x, q_a, q_b, type_a, type_b, lj_a, lj_b, lst


@pytest.mark.parametrize("name,ref", cases(), ids=[c[0] for c in cases()])
def test_fep_oracle_matches_reference_gpu_refdata(name, ref):
    from oracle.nbfe_oracle import nbfe_forces
    p, x, q_a, q_b, type_a, type_b, lj_a, lj_b, lst = fep_test_system(name)
    f, fshift, e_lj, e_el, dvdl_lj, dvdl_el = nbfe_forces(p, x, q_a, q_b, type_a, type_b, lj_a, lj_b, np.zeros((1, 3)), **lst)
    # the refdata holds float32 GPU results printed with 8 digits: 2e-6 of the magnitude, plus the rounding of sums of
    # large cancelling terms (energies of several hundred kJ/mol per pair)
    scale_f = np.abs(ref[4:16]).max()
    assert np.abs(f.reshape(-1) - ref[4:16]).max() <= 3e-6 * scale_f, (f, ref[4:16])
    assert np.abs(fshift[0] - ref[16:19]).max() <= 3e-6 * scale_f
    for got, want, other in ((e_lj, ref[0], 1.0), (e_el, ref[1], 100.0), (dvdl_el, ref[2], 100.0), (dvdl_lj, ref[3], 1.0)):
        assert abs(got - want) <= 3e-6 * max(abs(want), other), (name, got, want)
