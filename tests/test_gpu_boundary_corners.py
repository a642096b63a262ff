"""Corners of the drop-in boundary no other test calls: gpu_try_finish_task (nbnxm/gpu_common.h:290), the PME load
balancing update gpu_pme_loadbal_update_param (nbnxm_gpu_data_mgmt.cpp:701), force + virial without energies
(stepWork.computeVirial && !computeEnergy, the reference still runs its energy kernel), gpuGetNBAtomData's shared
outputs (nbnxm_gpu_data_mgmt.cpp:1823) and the flavor pickers driving a launch."""
import copy
import time

import numpy as np
import pytest

from util import load_golden, oracle_forces, oracle_params, product_inputs, product_params, relrms

pytestmark = pytest.mark.gpu


def setup(nb, nbat, plist):
    from gromacs_b200 import LOCAL
    nb.gpu_init_atomdata(nbat)
    nb.gpu_init_pairlist(plist, LOCAL)
    nb.setupGpuShortRangeWork(LOCAL)
    nb.gpu_upload_shiftvec(nbat)
    nb.gpu_copy_xq_to_gpu(nbat, LOCAL)


def launch(nb, nbat, sw):
    from gromacs_b200 import LOCAL
    nb.gpu_clear_outputs(True)
    nb.gpu_launch_kernel(sw, LOCAL)
    nbat.f[:] = 0
    nb.gpu_launch_cpyback(nbat, sw, LOCAL)


def test_try_finish_task_polls_until_done(oracle):
    """GpuTaskCompletion::Check: returns false while the stream is busy without touching the outputs, then true once,
    adding energies and shift forces exactly like the waiting form"""
    from gromacs_b200 import LOCAL, NbnxmGpu, StepWorkload
    from gromacs_b200.workload import make_workload
    wl = make_workload("water96k_fswitch")
    plist = wl.pairlist(min_sci=4000)
    nbat = wl.nbat
    sw = StepWorkload(computeEnergy=True, computeVirial=True)
    nb = NbnxmGpu(wl.params, nbat)
    try:
        setup(nb, nbat, plist)
        launch(nb, nbat, sw)
        fsh_wait = np.zeros((45, 3), np.float32)
        e_wait = nb.gpu_wait_finish_task(sw, LOCAL, shiftForces=fsh_wait)
        f_wait = nbat.f.copy()
        # a queue of launches keeps the stream busy for ~10 ms: the first polls must come back "not done".  The forces stay
        # on the device (useGpuFBufferOps): a copy into pageable host memory would block the host until the stream is idle
        swq = StepWorkload(computeEnergy=True, computeVirial=True, useGpuFBufferOps=True)
        for _ in range(60):
            launch(nb, nbat, swq)
        fsh = np.zeros((45, 3), np.float32)
        polls, done = 0, False
        t0 = time.perf_counter()
        while not done:
            done, e_lj, e_el = nb.gpu_try_finish_task(sw, LOCAL, shiftForces=fsh)
            polls += 1
            if not done:
                assert e_lj == 0.0 and e_el == 0.0 and not fsh.any()
            assert time.perf_counter() - t0 < 30
        assert polls > 1, "the task was already complete at the first poll: the busy path was not exercised"
        assert abs(e_lj - e_wait[0]) <= 1e-6 * abs(e_wait[0]) and abs(e_el - e_wait[1]) <= 1e-6 * abs(e_wait[1])
        assert np.abs(fsh - fsh_wait).max() <= 1e-5 * np.abs(fsh_wait).max()
        # and the forces of a following ordinary step are the same as before
        launch(nb, nbat, sw)
        nb.gpu_wait_finish_task(sw, LOCAL)
        assert relrms(nbat.f.astype(np.float64), f_wait.astype(np.float64)) < 1e-6
    finally:
        nb.gpu_free()


def test_force_and_virial_without_energy(oracle):
    """computeVirial without computeEnergy: shift forces come back, energies are not added"""
    from gromacs_b200 import LOCAL, NbnxmGpu, StepWorkload
    d = load_golden("bench1_ewald_cutgeom")
    nbat, plist = product_inputs(d)
    f_ref, fsh_ref, _, _ = oracle_forces(oracle, d, oracle_params(oracle, d))
    nb = NbnxmGpu(product_params(d), nbat)
    try:
        setup(nb, nbat, plist)
        sw = StepWorkload(computeEnergy=False, computeVirial=True)
        launch(nb, nbat, sw)              # fresh list: fused prune
        nb.gpu_wait_finish_task(sw, LOCAL)
        launch(nb, nbat, sw)
        fsh = np.zeros((45, 3), np.float32)
        e_lj, e_el = nb.gpu_wait_finish_task(sw, LOCAL, shiftForces=fsh)
        assert e_lj == 0.0 and e_el == 0.0
        assert relrms(nbat.f.astype(np.float64), f_ref) <= 5e-6
        vir = -0.5 * np.einsum("si,sj->ij", d["shift_vec"].astype(np.float64), fsh.astype(np.float64))
        vir_ref = -0.5 * np.einsum("si,sj->ij", d["shift_vec"].astype(np.float64), fsh_ref)
        assert np.abs(vir - vir_ref).max() <= 5e-6 * np.abs(vir_ref).max()
    finally:
        nb.gpu_free()


def test_pme_loadbal_update_param(oracle):
    """PME load balancing changes rcoulomb, the Ewald coefficient and the list radii between steps (pme_load_balancing.cpp
    -> gpu_pme_loadbal_update_param); afterwards the kernels must use the new values - checked against the oracle with
    those values - and the next list must be treated as fresh"""
    from gromacs_b200 import LOCAL, NbnxmGpu, StepWorkload
    from gromacs_b200 import system as S
    import math
    d = load_golden("bench1_ewald_cutgeom")
    nbat, plist = product_inputs(d)
    p0 = product_params(d)
    nb = NbnxmGpu(p0, nbat)
    sw = StepWorkload(computeEnergy=True, computeVirial=True)
    try:
        setup(nb, nbat, plist)
        launch(nb, nbat, sw)
        nb.gpu_wait_finish_task(sw, LOCAL)
        # a shorter Coulomb cut-off with the matching Ewald splitting, twin-range against the unchanged VdW cut-off
        rc_new = 0.8
        beta = S.ewald_beta(rc_new, 1e-5)
        p1 = copy.copy(p0)
        p1.elec_type = 5                     # EwaldAnaTwin: rvdw != rcoulomb now
        p1.rcoulomb_sq = rc_new * rc_new
        p1.ewald_beta = beta
        p1.sh_ewald = math.erfc(beta * rc_new) / rc_new
        nb.gpu_pme_loadbal_update_param(p1)
        assert nb.gpu_is_kernel_ewald_analytical()
        launch(nb, nbat, sw)
        e_lj, e_el = nb.gpu_wait_finish_task(sw, LOCAL)
        po = oracle_params(oracle, d)
        po.elec_type, po.rcoulomb_sq, po.ewald_beta, po.sh_ewald = 5, p1.rcoulomb_sq, p1.ewald_beta, p1.sh_ewald
        f_ref, _, e_ref, _ = oracle_forces(oracle, d, po)
        assert relrms(nbat.f.astype(np.float64), f_ref) <= 5e-6
        assert abs(e_el - e_ref[1]) <= 1e-6 * abs(e_ref[1]) and abs(e_lj - e_ref[0]) <= 1e-6 * abs(e_ref[0]) + 2e-6
        # F-only (packed kernel) with the new constants
        swf = StepWorkload()
        launch(nb, nbat, swf)
        nb.gpu_wait_finish_task(swf, LOCAL)
        assert relrms(nbat.f.astype(np.float64), f_ref) <= 5e-6
    finally:
        nb.gpu_free()


def test_shared_outputs_keep_what_other_kernels_added(oracle):
    """gpuGetNBAtomData: the GPU listed forces add into f and fShift between gpu_clear_outputs and the copy-back
    (sim_util.cpp:1450, listed_forces_gpu_impl); those contributions must come back with the nonbonded ones"""
    import torch
    from gromacs_b200 import LOCAL, NbnxmGpu, StepWorkload
    d = load_golden("bench1_ewald_cutgeom")
    nbat, plist = product_inputs(d)
    f_ref, fsh_ref, _, _ = oracle_forces(oracle, d, oracle_params(oracle, d))
    nb = NbnxmGpu(product_params(d), nbat)
    sw = StepWorkload(computeEnergy=True, computeVirial=True)
    try:
        setup(nb, nbat, plist)
        d_f, d_fshift = nb.gpuGetNBAtomData()
        n = nbat.numAtoms()
        rng = np.random.default_rng(3)
        extra_f = rng.normal(0, 100.0, (n, 3)).astype(np.float32)
        extra_fs = rng.normal(0, 10.0, (45, 3)).astype(np.float32)
        stream = torch.cuda.ExternalStream(nb.streams()[0])
        for step in range(2):
            nb.gpu_clear_outputs(True)
            # "bonded kernel": adds on the nonbonded local stream, after the clear
            with torch.cuda.stream(stream):
                f_view = torch.as_tensor(_DevArray(d_f, (n, 3)), device="cuda")
                fs_view = torch.as_tensor(_DevArray(d_fshift, (45, 3)), device="cuda")
                f_view += torch.from_numpy(extra_f).cuda()
                fs_view += torch.from_numpy(extra_fs).cuda()
            stream.synchronize()
            nb.gpu_launch_kernel(sw, LOCAL)
            nbat.f[:] = 0
            nb.gpu_launch_cpyback(nbat, sw, LOCAL)
            fsh = np.zeros((45, 3), np.float32)
            nb.gpu_wait_finish_task(sw, LOCAL, shiftForces=fsh)
            assert relrms(nbat.f.astype(np.float64), f_ref + extra_f) <= 5e-6
            # the kernels leave the central shift (index 22) alone (nbnxm_cuda_kernel.cuh:700-717); the oracle, like the CPU
            # kernels, sums it up: compare the other 44, and expect only the foreign contribution at 22
            want = fsh_ref + extra_fs
            want[22] = extra_fs[22]
            assert np.abs(fsh - want).max() <= 1e-5 * np.abs(want).max()
    finally:
        nb.gpu_free()


class _DevArray:
    """__cuda_array_interface__ view of a raw float32 device buffer"""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "strides": None, "version": 3}


def test_pickers_select_the_kernels_the_reference_would():
    """interaction.py's pickers feed make_params; the flavor that runs is the one nbnxmGpuPick{Vdw,Electrostatics}KernelType
    gives for the same inputrec settings (nbnxm_gpu_data_mgmt.cpp:168-216, 368-460)"""
    from gromacs_b200 import NbnxmGpu
    from gromacs_b200.interaction import make_params, pick_elec_type, pick_vdw_type
    from gromacs_b200.nbnxm import ELEC_TYPES, VDW_TYPES
    d = load_golden("test243_ewald_cutnone")
    nbat, _ = product_inputs(d)
    table = [
        # (coulombtype, rcoulomb, rvdw, vdwtype, modifier, comb rule, ljpme rule) -> (elec, vdw) as the reference's enums
        (("Pme", 0.9, 0.9, "Cut", "PotShift", "Geometric", "Geom"), ("EwaldAna", "CutCombGeom")),
        (("Pme", 1.0, 0.9, "Cut", "PotShift", "LorentzBerthelot", "Geom"), ("EwaldAnaTwin", "CutCombLB")),
        (("Ewald", 0.9, 0.9, "Cut", "None", "None", "Geom"), ("EwaldAna", "Cut")),
        (("RF", 0.9, 0.9, "Cut", "ForceSwitch", "Geometric", "Geom"), ("RF", "FSwitch")),
        (("Cut", 0.9, 0.9, "Cut", "PotSwitch", "None", "Geom"), ("Cut", "PSwitch")),
        (("Pme", 0.9, 0.9, "Pme", "PotShift", "None", "Geom"), ("EwaldAna", "EwaldGeom")),
        (("Pme", 0.9, 0.9, "Pme", "PotShift", "None", "LB"), ("EwaldAna", "EwaldLB")),
        (("Fmm", 0.9, 0.9, "Cut", "PotShift", "None", "Geom"), ("None", "Cut")),
    ]
    for (ct, rcoul, rvdw, vt, mod, comb, ljpme), (elec_want, vdw_want) in table:
        elec, vdw = pick_elec_type(ct, rcoul, rvdw), pick_vdw_type(vt, mod, comb, ljpme)
        assert (elec, vdw) == (elec_want, vdw_want)
        p = make_params(elec, vdw, epsfac=138.9, rcoulomb=rcoul, rvdw=rvdw, rlist_outer=1.0, ewald_beta=3.1)
        nb = NbnxmGpu(p, nbat)
        try:
            assert nb.params.elec_type == ELEC_TYPES[elec_want] and nb.params.vdw_type == VDW_TYPES[vdw_want]
            assert nb.gpu_is_kernel_ewald_analytical() == elec_want.startswith("EwaldAna")
        finally:
            nb.gpu_free()
    assert pick_elec_type("Pme", 0.9, 0.9, analytical=False) == "EwaldTab"
    with pytest.raises(ValueError):
        pick_vdw_type("Cut", "ExactCutoff", "None")
    with pytest.raises(ValueError):
        pick_elec_type("GeneralizedRF", 0.9, 0.9)
