"""GPU parity tests (run on the B200 box): the CUDA path, called through the C ABI, against the CPU
oracle on the reference's own pair lists (tests/golden).  Tolerances are BASELINE.json's:
forces rel-RMS <= 5e-6 and max-component <= 1e-4 (relative to the largest force), energies and shift
forces within 1e-6 relative, pruned masks bit-exact."""
import numpy as np
import pytest

from test_tolerance_evidence import reference_simd_errors
from util import (golden_cases, load_golden, maxrel, oracle_forces, oracle_params, product_inputs, product_params,
                  relrms)

pytestmark = pytest.mark.gpu

F_RELRMS = 5e-6
F_MAXREL = 1e-4
E_REL = 1e-6
# With the Lorentz-Berthelot rule C6/C12 are rebuilt per pair in float32 from sigma and epsilon, exactly as
# the reference GPU kernel does; the fixed rounding of the O-O parameters is a systematic ~1e-6 term.
E_REL_LJ_LB = 2e-6
# The shift forces are sums of float32 pair forces with heavy cancellation: the reference's own float SIMD kernel is
# 0.6e-6 ... 4.4e-6 away from the double oracle on these fixtures (tests/test_tolerance_evidence.py).  The CUDA kernels are
# held to 1.5 x the reference float kernel's own error on the same fixture (measured on a B200, profiles/r02d_parity_errors.jsonl:
# below the reference's error on 48 of 52 fixture x kernel combinations, at most 1.9 x on the two split lists), with 1.5e-6 as
# the floor where the reference happens to be more exact than that.
VIR_REL = 1.5e-6
VIR_VS_REFERENCE_FLOAT = 1.5


def virial(shift_vec, fshift):
    return -0.5 * np.einsum("si,sj->ij", shift_vec.astype(np.float64), fshift)


def run_step(nb, nbat, plist, energy, virial, fresh_list=True):
    from gromacs_b200 import LOCAL, StepWorkload
    sw = StepWorkload(computeEnergy=energy, computeVirial=virial)
    if fresh_list:
        nb.gpu_init_atomdata(nbat)
        nb.gpu_init_pairlist(plist, LOCAL)
        nb.setupGpuShortRangeWork(LOCAL)
    nb.gpu_upload_shiftvec(nbat)
    nb.gpu_clear_outputs(computeVirial=True)
    nb.gpu_copy_xq_to_gpu(nbat, LOCAL)
    nb.gpu_launch_kernel(sw, LOCAL)
    nbat.f[:] = 0
    nb.gpu_launch_cpyback(nbat, sw, LOCAL)
    fshift = np.zeros((45, 3), np.float32)
    e_lj, e_el = nb.gpu_wait_finish_task(sw, LOCAL, shiftForces=fshift)
    return nbat.f.astype(np.float64).copy(), e_lj, e_el, fshift.astype(np.float64)


def record_tolerance(case, variant, vir, e_lj, e_el, simd):
    """keeps the measured errors next to the reference float kernel's own (gpurun_out/parity_errors.jsonl)"""
    import json
    import os
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_errors.jsonl"), "a") as fh:
        fh.write(json.dumps({"case": case, "kernel": variant, "virial_rel": vir, "e_lj_rel": e_lj, "e_el_rel": e_el,
                             "ref_simd_virial_rel": simd[0], "ref_simd_e_lj_rel": simd[1], "ref_simd_e_el_rel": simd[2]}) + "\n")


def check_forces(f, ref):
    assert relrms(f, ref) <= F_RELRMS, relrms(f, ref)
    assert maxrel(f, ref) <= F_MAXREL, maxrel(f, ref)


@pytest.fixture(params=["packed", "scalar"])
def kernel_variant(request, monkeypatch):
    """Force-only launches use the packed-FP32 (FFMA2) kernel where the flavor has one; NBNXM_B200_SCALAR_KERNEL
    selects the scalar kernel instead (read at every launch), so both are checked against the oracle."""
    if request.param == "scalar":
        monkeypatch.setenv("NBNXM_B200_SCALAR_KERNEL", "1")
    else:
        monkeypatch.delenv("NBNXM_B200_SCALAR_KERNEL", raising=False)
    return request.param


@pytest.mark.parametrize("case", golden_cases())
def test_force_energy_virial_parity(oracle, case, kernel_variant):
    from gromacs_b200 import NbnxmGpu
    d = load_golden(case)
    nbat, plist = product_inputs(d)
    po = oracle_params(oracle, d)
    f_ref, fsh_ref, e_ref, npairs_ref = oracle_forces(oracle, d, po)
    nb = NbnxmGpu(product_params(d), nbat)
    nb.set_pair_counting(True)
    try:
        # step 1: fresh list, static pruning -> fused force+prune kernel on the unsorted list, F only
        f, _, _, _ = run_step(nb, nbat, plist, energy=False, virial=False)
        check_forces(f, f_ref)
        assert nb.get_pair_count() == npairs_ref
        # step 2: sorted list, F+E+virial kernel
        f, e_lj, e_el, fsh = run_step(nb, nbat, plist, energy=True, virial=True, fresh_list=False)
        check_forces(f, f_ref)
        e_rel_lj = E_REL_LJ_LB if "cutlb" in case else E_REL
        assert abs(e_lj - e_ref[0]) <= e_rel_lj * abs(e_ref[0]) + 2e-6, (e_lj, e_ref[0])
        assert abs(e_el - e_ref[1]) <= E_REL * abs(e_ref[1]), (e_el, e_ref[1])
        # virial contribution of the shift forces: -1/2 sum_s shift_vec[s] (x) fshift[s]
        vir, vir_ref = virial(d["shift_vec"], fsh), virial(d["shift_vec"], fsh_ref)
        vir_err    = float(np.abs(vir - vir_ref).max() / np.abs(vir_ref).max())
        simd_errs  = reference_simd_errors(oracle, d)
        record_tolerance(case, kernel_variant, vir_err, abs(e_lj - e_ref[0]) / abs(e_ref[0]), abs(e_el - e_ref[1]) / abs(e_ref[1]), simd_errs)
        assert vir_err <= max(VIR_REL, VIR_VS_REFERENCE_FLOAT * simd_errs[0]), (vir_err, simd_errs[0])
        assert np.all(fsh[22] == 0)
        # step 3: F-only kernel on the sorted list
        f, _, _, _ = run_step(nb, nbat, plist, energy=False, virial=False, fresh_list=False)
        check_forces(f, f_ref)
    finally:
        nb.gpu_free()


@pytest.mark.parametrize("case", ["test243_ewald_cutnone", "bench1_ewald_cutnone", "bench1_ewald_fswitch"])
def test_tabulated_ewald_and_plain_cutoff_flavors(oracle, case):
    """flavors the golden set has no dedicated file for: tabulated Ewald (+twin), plain cut-off Coulomb, LJ-PME LB."""
    from gromacs_b200 import NbnxmGpu
    d = load_golden(case)
    nbat, plist = product_inputs(d)
    vdw_native = None
    for elec, vdw in (("EwaldTab", None), ("EwaldTabTwin", None), ("Cut", None), ("EwaldAna", "EwaldLB")):
        if vdw == "EwaldLB" and "cutnone" not in case:
            continue
        po = oracle_params(oracle, d, elec=elec, vdw=vdw)
        f_ref, fsh_ref, e_ref, _ = oracle_forces(oracle, d, po)
        nb = NbnxmGpu(product_params(d, elec=elec, vdw=vdw), nbat, coulomb_tab=d["ic_coulomb_tab_F"])
        try:
            f, e_lj, e_el, fsh = run_step(nb, nbat, plist, energy=True, virial=True)
            check_forces(f, f_ref)
            assert abs(e_lj - e_ref[0]) <= E_REL * abs(e_ref[0]) + 2e-6
            assert abs(e_el - e_ref[1]) <= E_REL * abs(e_ref[1])
            f, _, _, _ = run_step(nb, nbat, plist, energy=False, virial=False, fresh_list=False)
            check_forces(f, f_ref)
        finally:
            nb.gpu_free()


def masks_of(cj):
    return np.stack([cj[:, 4], cj[:, 6]], 1).reshape(-1)


@pytest.mark.parametrize("case", ["test243_ewald_cutnone_rl1.0_split", "bench1_ewald_cutgeom", "bench1_rf_cutnone_split",
                                  "bench1_ewald_ljpmegeom"])
def test_prune_masks_bit_exact(oracle, case):
    from gromacs_b200 import LOCAL, NbnxmGpu, StepWorkload
    d = load_golden(case)
    nbat, plist = product_inputs(d)
    rl, rc = float(d["rlist"][0]), float(d["ic_rcoulomb"][0])
    rin = 0.5 * (rl + rc)
    po = oracle_params(oracle, d, rlist_inner=rin)
    f_ref, _, _, _ = oracle_forces(oracle, d, po)

    # --- static pruning: the fused force+prune kernel clears bits beyond rlistOuter
    nb = NbnxmGpu(product_params(d), nbat)
    try:
        run_step(nb, nbat, plist, energy=False, virial=False)
        cj_gpu, _, sci_sorted, sci_count, _ = nb.download_pairlist()
        cj_o = d["pl_cjPacked"].copy()
        outer_o = np.zeros(2 * cj_o.shape[0], np.uint32)
        oracle.prune(oracle_params(oracle, d), d["pl_sci"], cj_o, outer_o, d["nbat_xq"], d["shift_vec"], fresh=True)
        assert np.array_equal(masks_of(cj_gpu), outer_o)
        assert np.array_equal(cj_gpu[:, :4], d["pl_cjPacked"][:, :4])
        check_sorted(d["pl_sci"], sci_sorted, sci_count, masks_of(cj_gpu))
    finally:
        nb.gpu_free()

    # --- dynamic pruning: first pass (outer + inner), then rolling passes on moved coordinates
    nb = NbnxmGpu(product_params(d, rlist_inner=rin, dynamic_pruning=True), nbat)
    try:
        f, _, _, _ = run_step(nb, nbat, plist, energy=False, virial=False)
        assert relrms(f, f_ref) <= F_RELRMS
        cj_gpu, outer_gpu, sci_sorted, sci_count, _ = nb.download_pairlist()
        cj_o = d["pl_cjPacked"].copy()
        outer_o = np.zeros(2 * cj_o.shape[0], np.uint32)
        cnt_o = oracle.prune(po, d["pl_sci"], cj_o, outer_o, d["nbat_xq"], d["shift_vec"], fresh=True)
        assert np.array_equal(outer_gpu, outer_o)
        assert np.array_equal(cj_gpu, cj_o)
        assert np.array_equal(sci_count, cnt_o)
        check_sorted(d["pl_sci"], sci_sorted, sci_count, masks_of(cj_gpu))

        rng = np.random.default_rng(7)
        num_parts = 3
        xq = nbat.xq.copy()
        for step in range(2 * num_parts):
            xq[:, :3] += rng.normal(0, 0.01, (xq.shape[0], 3)).astype(np.float32)
            nbat.xq[:] = xq
            nb.gpu_copy_xq_to_gpu(nbat, LOCAL)
            nb.gpu_launch_kernel_pruneonly(LOCAL, num_parts)
            oracle.prune(po, sci_sorted, cj_o, outer_o, xq, d["shift_vec"], fresh=False, part=step % num_parts,
                         nparts=num_parts)
            cj_gpu, outer_gpu, _, _, rolling_part = nb.download_pairlist()
            assert np.array_equal(cj_gpu, cj_o), "rolling prune step %d" % step
            assert np.array_equal(outer_gpu, outer_o)
            nunits = (d["pl_sci"].shape[0] + num_parts - 1) // num_parts
            assert np.all(rolling_part[:nunits] == (step + 1) % num_parts)
        # forces on the rolling-pruned list are still the physical forces for the moved coordinates
        d2 = dict(d)
        d2["nbat_xq"] = xq
        f_ref2, _, _, _ = oracle_forces(oracle, d2, po)
        sw = StepWorkload()
        nb.gpu_clear_outputs(True)
        nb.gpu_launch_kernel(sw, LOCAL)
        nb.gpu_launch_cpyback(nbat, sw, LOCAL)
        nb.gpu_wait_finish_task(sw, LOCAL)
        check_forces(nbat.f.astype(np.float64), f_ref2)
    finally:
        nb.gpu_free()


def check_sorted(sci, sci_sorted, sci_count, masks):
    """sciSorted is a permutation of sci ordered by decreasing pruned pair count (bucket order)."""
    a = sorted(map(tuple, sci.tolist()))
    b = sorted(map(tuple, sci_sorted.tolist()))
    assert a == b
    key = {tuple(s): c for s, c in zip(sci.tolist(), sci_count.tolist())}
    keys = [key[tuple(s)] for s in sci_sorted.tolist()]
    assert keys == sorted(keys)
    for s, c in zip(sci.tolist(), sci_count.tolist()):
        bits = int(sum(bin(int(v)).count("1") for v in masks[2 * s[2]:2 * s[3]]))
        assert c == max(8192 - bits - 1, 0)


def test_empty_and_degenerate_lists(oracle):
    from gromacs_b200 import LOCAL, NONLOCAL, NbnxmGpu, PairlistGpu, StepWorkload
    d = load_golden("test243_ewald_cutnone")
    nbat, plist = product_inputs(d)
    nb = NbnxmGpu(product_params(d, dynamic_pruning=True), nbat, bLocalAndNonlocal=True)
    try:
        empty = PairlistGpu(sci=np.zeros((0, 4), np.int32), cjPacked=np.zeros((0, 8), np.uint32),
                            excl=np.full((1, 32), 0xffffffff, np.uint32))
        nb.gpu_init_atomdata(nbat)
        nb.gpu_init_pairlist(empty, LOCAL)
        nb.gpu_init_pairlist(empty, NONLOCAL)
        nb.setupGpuShortRangeWork(LOCAL)
        nb.setupGpuShortRangeWork(NONLOCAL)
        assert not nb.haveGpuShortRangeWork(NONLOCAL)
        sw = StepWorkload(computeEnergy=True, computeVirial=True)
        nb.gpu_upload_shiftvec(nbat)
        nb.gpu_clear_outputs(True)
        nb.gpu_copy_xq_to_gpu(nbat, LOCAL)
        nb.gpu_copy_xq_to_gpu(nbat, NONLOCAL)
        nb.gpu_launch_kernel(sw, LOCAL)
        nb.gpu_launch_kernel(sw, NONLOCAL)
        nb.gpu_launch_kernel_pruneonly(LOCAL, 2)
        nb.gpu_launch_cpyback(nbat, sw, NONLOCAL)
        nb.gpu_launch_cpyback(nbat, sw, LOCAL)
        e = nb.gpu_wait_finish_task(sw, LOCAL)
        assert e == (0.0, 0.0) and not nbat.f.any()
        # a list whose masks are all zero and one with a single sci entry
        z = PairlistGpu(sci=d["pl_sci"], cjPacked=d["pl_cjPacked"].copy(), excl=d["pl_excl"])
        z.cjPacked[:, 4] = 0
        z.cjPacked[:, 6] = 0
        nb.gpu_init_pairlist(z, LOCAL)
        nb.gpu_launch_kernel(sw, LOCAL)
        nb.gpu_launch_cpyback(nbat, sw, LOCAL)
        nb.gpu_wait_finish_task(sw, LOCAL)
        assert not nbat.f.any()
    finally:
        nb.gpu_free()


def test_x_to_nbat_x_and_local_nonlocal_streams(oracle):
    """x buffer op + the local / non-local split of one list over two streams gives the same forces."""
    import torch
    from gromacs_b200 import LOCAL, NONLOCAL, NbnxmGpu, PairlistGpu, StepWorkload
    d = load_golden("bench1_ewald_cutgeom")
    nbat, plist = product_inputs(d)
    po = oracle_params(oracle, d)
    f_ref, _, e_ref, _ = oracle_forces(oracle, d, po)
    nsci = plist.sci.shape[0]
    half = nsci // 2
    loc = PairlistGpu(sci=plist.sci[:half], cjPacked=plist.cjPacked, excl=plist.excl)
    nonloc = PairlistGpu(sci=plist.sci[half:], cjPacked=plist.cjPacked, excl=plist.excl)
    nbat.numLocalAtoms = (nbat.numAtoms() // 128) * 64
    nb = NbnxmGpu(product_params(d), nbat, bLocalAndNonlocal=True)
    try:
        nb.gpu_init_atomdata(nbat)
        nb.gpu_init_pairlist(loc, LOCAL)
        nb.gpu_init_pairlist(nonloc, NONLOCAL)
        nb.setupGpuShortRangeWork(LOCAL)
        nb.setupGpuShortRangeWork(NONLOCAL)
        nb.gpu_upload_shiftvec(nbat)
        # coordinates arrive as a device rvec array in atom order; charges are set by a first xq upload
        xq0 = nbat.xq.copy()
        nbat.xq[:, :3] = 0
        nb.gpu_copy_xq_to_gpu(nbat, LOCAL)
        nb.gpu_copy_xq_to_gpu(nbat, NONLOCAL)
        nb.gpu_wait_finish_task(StepWorkload(), LOCAL)
        nb.gpu_wait_finish_task(StepWorkload(), NONLOCAL)
        nbat.xq[:] = xq0
        ai = d["nbat_atom_index"]
        nloc = nbat.numLocalAtoms
        nb.nbnxm_gpu_init_x_to_nbat_x(ai[:nloc], atom_offset=0, grid=0, ngrids=2)
        nb.nbnxm_gpu_init_x_to_nbat_x(ai[nloc:], atom_offset=nloc, grid=1, ngrids=2)
        x_dev = torch.from_numpy(d["sys_x"]).cuda().contiguous()
        torch.cuda.synchronize()
        sw = StepWorkload(computeEnergy=True, computeVirial=True)
        nb.gpu_clear_outputs(True)
        nb.nbnxm_gpu_x_to_nbat_x(x_dev.data_ptr(), None, LOCAL)
        nb.nbnxmInsertNonlocalGpuDependency(NONLOCAL)
        nb.nbnxm_gpu_x_to_nbat_x(x_dev.data_ptr(), None, NONLOCAL)
        # this test's "local" list also reads non-local atoms (it is half of one list, not a DD list):
        # wait for the non-local coordinates before launching it
        nb.gpu_wait_finish_task(StepWorkload(), NONLOCAL)
        nb.gpu_launch_kernel(sw, LOCAL)
        nb.gpu_launch_kernel(sw, NONLOCAL)
        # ... and it also writes forces of non-local atoms, which a DD local list never does: the non-local
        # copy-back must not overtake it (the library only orders local-after-non-local, like the reference)
        torch.cuda.synchronize()
        nb.gpu_launch_cpyback(nbat, sw, NONLOCAL)
        nb.gpu_launch_cpyback(nbat, sw, LOCAL)
        nb.gpu_wait_finish_task(sw, NONLOCAL)
        e_lj, e_el = nb.gpu_wait_finish_task(sw, LOCAL)
        check_forces(nbat.f.astype(np.float64), f_ref)
        assert abs(e_el - e_ref[1]) <= E_REL * abs(e_ref[1])
    finally:
        nb.gpu_free()
