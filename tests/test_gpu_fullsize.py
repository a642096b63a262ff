"""GPU tests at BASELINE.json's benchmark sizes: the 96k-atom box against the double-precision oracle on the
whole list, and the 1.5M-atom box through size-independent properties (Newton's third law, pruning does not
change forces, packed and scalar kernels agree, a sub-list against the oracle)."""
import numpy as np
import pytest

from util import relrms

pytestmark = pytest.mark.gpu


def orc_params(oracle, wl):
    p = oracle.OrcParams()
    for name, _ in wl.params._fields_:
        if hasattr(p, name):
            setattr(p, name, getattr(wl.params, name))
    p.ntypes = wl.nbat.numTypes
    return p


def gpu_forces(wl, plist, energy, dynamic_pruning=None, rolling_steps=0):
    from gromacs_b200 import LOCAL, NbnxmGpu, StepWorkload
    import copy
    params = copy.copy(wl.params)
    if dynamic_pruning is not None:
        params.use_dynamic_pruning = int(dynamic_pruning)
    nbat = wl.nbat
    nb = NbnxmGpu(params, nbat)
    try:
        sw = StepWorkload(computeEnergy=energy, computeVirial=energy)
        nb.gpu_init_atomdata(nbat)
        nb.gpu_init_pairlist(plist, LOCAL)
        nb.setupGpuShortRangeWork(LOCAL)
        nb.gpu_upload_shiftvec(nbat)
        nb.gpu_copy_xq_to_gpu(nbat, LOCAL)
        e = (0.0, 0.0)
        for step in range(1 + rolling_steps):
            nb.gpu_clear_outputs(True)
            nb.gpu_launch_kernel(sw, LOCAL)
            if step > 0 and params.use_dynamic_pruning:
                nb.gpu_launch_kernel_pruneonly(LOCAL, 3)
            nb.gpu_launch_cpyback(nbat, sw, LOCAL)
            e = nb.gpu_wait_finish_task(sw, LOCAL)
        return nbat.f.astype(np.float64).copy(), e
    finally:
        nb.gpu_free()


def test_96k_force_switch_box_matches_oracle(oracle):
    """BASELINE configs[1]: 96 000 atoms, rc 1.0, Ewald + LJ force-switch, F+E; whole list through the oracle."""
    from gromacs_b200.workload import make_workload
    wl = make_workload("water96k_fswitch")
    plist = wl.pairlist(min_sci=4000)
    g = wl.nbat
    f_ref, _, e_ref, _ = oracle.forces(orc_params(oracle, wl), plist.sci, plist.cjPacked, plist.excl, g.xq, g.type,
                                       np.zeros((g.numAtoms(), 2), np.float32), g.nbfp, g.nbfp_comb, g.shift_vec)
    f, (e_lj, e_el) = gpu_forces(wl, plist, energy=True, rolling_steps=6)
    assert relrms(f, f_ref) <= 5e-6
    assert np.abs(f - f_ref).max() <= 1e-4 * np.abs(f_ref).max()
    assert abs(e_el - e_ref[1]) <= 1e-6 * abs(e_ref[1]), (e_el, e_ref[1])
    assert abs(e_lj - e_ref[0]) <= 1e-6 * abs(e_ref[0]), (e_lj, e_ref[0])


def test_1536k_box_properties(oracle, monkeypatch):
    """BASELINE configs[3] at full size."""
    from gromacs_b200.workload import make_workload
    from gromacs_b200 import PairlistGpu
    wl = make_workload("water1536k")
    plist = wl.pairlist(min_sci=18944)
    # dynamic pruning + a full rolling cycle vs. the unpruned outer list: pruning only drops pairs beyond rc
    f_pruned, _ = gpu_forces(wl, plist, energy=False, dynamic_pruning=True, rolling_steps=6)
    f_outer, _ = gpu_forces(wl, plist, energy=False, dynamic_pruning=False)
    assert relrms(f_pruned, f_outer) <= 2e-6
    # Newton's third law: the forces sum to zero up to float32 accumulation noise
    scale = np.abs(f_pruned).sum(0)
    assert np.all(np.abs(f_pruned.sum(0)) <= 1e-6 * scale), (f_pruned.sum(0), scale)
    # packed (FFMA2) and scalar kernels are two evaluations of the same pairs
    monkeypatch.setenv("NBNXM_B200_SCALAR_KERNEL", "1")
    f_scalar, _ = gpu_forces(wl, plist, energy=False, dynamic_pruning=True, rolling_steps=2)
    monkeypatch.delenv("NBNXM_B200_SCALAR_KERNEL")
    assert relrms(f_pruned, f_scalar) <= 2e-6
    # a sub-list (every 40th sci entry) against the float32 oracle port on the same entries
    sub = PairlistGpu(sci=plist.sci[::40], cjPacked=plist.cjPacked, excl=plist.excl)
    g = wl.nbat
    f32, _, _ = oracle.forces_f32_omp(orc_params(oracle, wl), sub.sci, sub.cjPacked, sub.excl, g.xq, g.type, g.lj_comb, g.nbfp,
                                      g.nbfp_comb, g.shift_vec, nthreads=8)
    f_sub, _ = gpu_forces(wl, sub, energy=False, dynamic_pruning=False)
    assert relrms(f_sub, f32.astype(np.float64)) <= 5e-6


@pytest.mark.parametrize("energy", [False, True])
def test_long_sci_entries_span_descriptor_chunks(oracle, energy):
    """Unsplit sci entries with more than 64 cjPacked groups: the packed kernel stages the groups of an entry in
    shared memory 64 at a time (c_descChunk), so this walks more than one chunk per entry."""
    import copy
    from gromacs_b200.workload import make_workload
    wl = make_workload("water48k_test", energy=energy)
    wl.params = copy.copy(wl.params)
    plist = wl.grid.pairlist(1.7, wl.box.excl_index, wl.box.excl_atoms, min_sci=0)
    assert (plist.sci[:, 3] - plist.sci[:, 2]).max() > 64
    g = wl.nbat
    f_ref, _, e_ref, _ = oracle.forces(orc_params(oracle, wl), plist.sci, plist.cjPacked, plist.excl, g.xq, g.type,
                                       g.lj_comb, g.nbfp, g.nbfp_comb, g.shift_vec)
    f, (e_lj, e_el) = gpu_forces(wl, plist, energy=energy, dynamic_pruning=False)
    assert relrms(f, f_ref) <= 5e-6
    if energy:
        assert abs(e_el - e_ref[1]) <= 1e-6 * abs(e_ref[1]), (e_el, e_ref[1])
        assert abs(e_lj - e_ref[0]) <= 1e-6 * abs(e_ref[0]), (e_lj, e_ref[0])


def test_do_force_step_matches_call_sequence():
    """nbnxm_b200_do_force_step issues the same launches as the individual gpu_* calls: bit-identical masks, forces within
    the atomics' summation-order noise."""
    from gromacs_b200 import LOCAL, NbnxmGpu, StepWorkload
    from gromacs_b200.workload import make_workload
    wl = make_workload("water48k_test", energy=True)
    import copy
    params = copy.copy(wl.params)
    params.use_dynamic_pruning = 1
    params.rlist_inner_sq = np.float32(0.905 ** 2)
    plist = wl.pairlist(min_sci=2000)
    out = []
    for composite in (False, True):
        nbat = wl.nbat
        nb = NbnxmGpu(params, nbat)
        try:
            sw = StepWorkload(computeEnergy=True, computeVirial=True)
            nb.gpu_init_atomdata(nbat)
            nb.gpu_init_pairlist(plist, LOCAL)
            nb.setupGpuShortRangeWork(LOCAL)
            nb.gpu_upload_shiftvec(nbat)
            nb.gpu_copy_xq_to_gpu(nbat, LOCAL)
            for step in range(4):
                if composite:
                    sw.useGpuFBufferOps = False
                    nb.do_force_step(step, sw, have_halo=False, dynamic_pruning=True, num_parts=3, xq_host=nbat.xq, f_host=nbat.f)
                else:
                    nb.gpu_copy_xq_to_gpu(nbat, LOCAL)
                    nb.gpu_clear_outputs(True)
                    nb.gpu_launch_kernel(sw, LOCAL)
                    if step % 2 == 1:
                        nb.gpu_launch_kernel_pruneonly(LOCAL, 3)
                    nb.gpu_launch_cpyback(nbat, sw, LOCAL)
                e = nb.gpu_wait_finish_task(sw, LOCAL)
            out.append((nbat.f.astype(np.float64).copy(), e))
        finally:
            nb.gpu_free()
    (f0, e0), (f1, e1) = out
    assert relrms(f1, f0) <= 1e-6
    assert abs(e1[0] - e0[0]) <= 1e-6 * abs(e0[0]) and abs(e1[1] - e0[1]) <= 1e-6 * abs(e0[1])


@pytest.mark.parametrize("energy", [False, True])
def test_pipelined_step_matches_plain_step(energy):
    """nbnxm_b200_do_force_step_pipelined (chunked H2D / kernel / D2H) against the plain sequence on the same list."""
    import copy
    from gromacs_b200 import LOCAL, NbnxmGpu, StepWorkload
    from gromacs_b200.pipeline import make_chunk_plan
    from gromacs_b200.workload import make_workload
    wl = make_workload("water48k_test", energy=energy)
    params = copy.copy(wl.params)
    params.use_dynamic_pruning = 1
    params.rlist_inner_sq = np.float32(0.905 ** 2)
    plan = make_chunk_plan(wl.grid, wl.pairlist(min_sci=3000), 4)
    assert plan.first_sci[-1] == plan.plist.sci.shape[0] and plan.first_atom[-1] == wl.nbat.numAtoms()
    out = []
    for pipelined in (False, True):
        nbat = wl.nbat
        nbat.f[:] = 0
        nb = NbnxmGpu(params, nbat)
        try:
            sw = StepWorkload(computeEnergy=energy, computeVirial=energy, useGpuFBufferOps=False)
            nb.gpu_init_atomdata(nbat)
            nb.gpu_init_pairlist(plan.plist, LOCAL)
            nb.setupGpuShortRangeWork(LOCAL)
            nb.gpu_upload_shiftvec(nbat)
            for step in range(4):
                if pipelined:
                    nb.set_pipeline_timeline(step == 3)
                    nb.do_force_step_pipelined(step, sw, plan, nbat.xq, nbat.f, dynamic_pruning=True, num_parts=3)
                else:
                    nb.do_force_step(step, sw, dynamic_pruning=True, num_parts=3, xq_host=nbat.xq, f_host=nbat.f)
                e = nb.gpu_wait_finish_task(sw, LOCAL)
            if pipelined:
                # the recorded timeline of the last step: every chunk's kernel starts after the chunks it reads have arrived
                # and its forces come down after its kernel
                t = nb.pipeline_timeline()
                assert t.shape == (plan.nchunks, 4) and np.all(t >= 0)
                for k in range(plan.nchunks):
                    needed = [c for c in range(plan.nchunks) if int(plan.needs[k]) >> c & 1]
                    assert t[k, 1] >= max(t[c, 0] for c in needed) - 1e-3
                    assert t[k, 2] >= t[k, 1] and t[k, 3] >= t[k, 2] - 1e-3
            out.append((nbat.f.astype(np.float64).copy(), e))
        finally:
            nb.gpu_free()
    (f0, e0), (f1, e1) = out
    assert relrms(f1, f0) <= 1e-6
    if energy:
        assert abs(e1[0] - e0[0]) <= 1e-6 * abs(e0[0]) and abs(e1[1] - e0[1]) <= 1e-6 * abs(e0[1])
