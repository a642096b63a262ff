"""GPU parity on the configurations bench.py measures (BASELINE.json configs[2] and configs[4]), run the way bench.py runs
them: nbnxm_b200_do_force_step with dynamic pruning — first-pass prune on the list step, rolling prune on the reference's
schedule (every second step, numParts = nstlistPrune / 2, pairlist_tuning.cpp:685) — and the packed production kernel.

  water384k_ljpme, water384k_pswitch : forces of the production sequence against the double-precision oracle on the WHOLE
      list (pruning only drops pairs beyond the cut-off, so the oracle runs on the outer list), energies of an F+E step,
      and the first-pass prune masks bit for bit against the oracle's prune on every entry.
  water96k_fswitch : first-pass AND rolling prune masks bit for bit on the whole list with moved coordinates.
  water12m (the headline box, 12 288 000 atoms): the oracle cannot walk 8.5e9 pairs in a test, so
      (i)   every 400th sci entry of the PRUNED production list (downloaded after a full rolling cycle) against the double
            oracle on exactly those entries — strict tolerances;
      (ii)  the first-pass prune masks of every 400th entry bit for bit;
      (iii) a whole-system check no sample can miss: the box is 16 x 16 x 16 translated copies of the 3000-atom unit box
            (bench_system.cpp:95-158), so the force on every one of the 12.3 M atoms must equal the oracle's force on its
            image in the 24 000-atom 2 x 2 x 2 box (replicated_reference below).  The copies are translated in float32,
            which rounds coordinates near 49 nm to 3.8e-6 nm, so this comparison carries that input noise: the tolerance
            is the one the double oracle itself shows between two such inputs (tests/test_replication_property.py, CPU);
            a single missing or doubled cluster pair shows up as O(1) on its atoms and fails the max-component bound.
Tolerances are the north star's: forces rel. RMS 5e-6, max component 1e-4 of the largest force, energies 1e-6."""
import copy

import numpy as np
import pytest

from util import relrms

pytestmark = pytest.mark.gpu

NUM_PARTS = 3


def orc_params(oracle, wl):
    p = oracle.OrcParams()
    for name, _ in wl.params._fields_:
        if hasattr(p, name):
            setattr(p, name, getattr(wl.params, name))
    p.ntypes = wl.nbat.numTypes
    return p


def oracle_forces(oracle, wl, sci, cjp, excl, energy=True):
    g = wl.nbat
    return oracle.forces(orc_params(oracle, wl), sci, cjp, excl, g.xq, g.type, g.lj_comb, g.nbfp, g.nbfp_comb, g.shift_vec,
                         calc_energy=energy)


class Production:
    """The bench's step loop on one handle."""

    def __init__(self, wl, plist, dynamic_pruning=True):
        from gromacs_b200 import LOCAL, NbnxmGpu
        self.wl, self.nbat = wl, wl.nbat
        params = copy.copy(wl.params)
        params.use_dynamic_pruning = int(dynamic_pruning)
        self.dyn = dynamic_pruning
        self.nb = NbnxmGpu(params, self.nbat)
        self.nb.gpu_init_atomdata(self.nbat)
        self.nb.gpu_init_pairlist(plist, LOCAL)
        self.nb.setupGpuShortRangeWork(LOCAL)
        self.nb.gpu_upload_shiftvec(self.nbat)
        self.nb.gpu_copy_xq_to_gpu(self.nbat, LOCAL)

    def step(self, i, energy=False):
        from gromacs_b200 import LOCAL, StepWorkload
        sw = StepWorkload(computeEnergy=energy, computeVirial=energy, useGpuFBufferOps=False)
        self.nb.do_force_step(i, sw, have_halo=False, dynamic_pruning=self.dyn, num_parts=NUM_PARTS, xq_host=self.nbat.xq,
                              f_host=self.nbat.f)
        e = self.nb.gpu_wait_finish_task(sw, LOCAL)
        return self.nbat.f.astype(np.float64).copy(), e

    def close(self):
        self.nb.gpu_free()


def check_first_pass_masks(oracle, wl, plist, cj_gpu, outer_gpu, sci_sample):
    """the oracle's prune (fresh list) on the sampled entries of the ORIGINAL list against what the GPU left"""
    cj_o = plist.cjPacked.copy()
    outer_o = np.zeros(2 * cj_o.shape[0], np.uint32)
    oracle.prune(orc_params(oracle, wl), sci_sample, cj_o, outer_o, wl.nbat.xq, wl.nbat.shift_vec, fresh=True)
    nbits = 0
    for s in sci_sample:
        a, b = int(s[2]), int(s[3])
        assert np.array_equal(cj_gpu[a:b], cj_o[a:b]), "inner masks of sci entry %d" % s[0]
        assert np.array_equal(outer_gpu[2 * a:2 * b], outer_o[2 * a:2 * b]), "outer masks of sci entry %d" % s[0]
        nbits += b - a
    return nbits


@pytest.mark.parametrize("name", ["water384k_ljpme", "water384k_pswitch"])
def test_384k_production_sequence_matches_oracle(oracle, name):
    """BASELINE configs[2]: 384 000 atoms, LJ-PME (geometric grid) / potential switch, nstlist-100 style dynamic + rolling
    pruning; whole list through the double oracle."""
    from gromacs_b200.workload import make_workload
    wl = make_workload(name)
    plist = wl.pairlist(min_sci=9472)
    f_ref, _, e_ref, _ = oracle_forces(oracle, wl, plist.sci, plist.cjPacked, plist.excl)
    run = Production(wl, plist)
    try:
        f, _ = run.step(0)                      # list step: first-pass prune fused in front of the force kernel
        assert relrms(f, f_ref) <= 5e-6
        cj_gpu, outer_gpu, _, _, _ = run.nb.download_pairlist()
        assert check_first_pass_masks(oracle, wl, plist, cj_gpu, outer_gpu, plist.sci) == plist.cjPacked.shape[0]
        for i in range(1, 2 * NUM_PARTS + 1):   # a full rolling cycle
            f, _ = run.step(i)
        assert relrms(f, f_ref) <= 5e-6, relrms(f, f_ref)
        assert np.abs(f - f_ref).max() <= 1e-4 * np.abs(f_ref).max()
        f, (e_lj, e_el) = run.step(2 * NUM_PARTS + 1, energy=True)
        assert relrms(f, f_ref) <= 5e-6
        assert abs(e_el - e_ref[1]) <= 1e-6 * abs(e_ref[1]), (e_el, e_ref[1])
        assert abs(e_lj - e_ref[0]) <= 1e-6 * abs(e_ref[0]), (e_lj, e_ref[0])
    finally:
        run.close()


def test_96k_prune_masks_bit_exact_whole_list(oracle):
    """first-pass and rolling prune on a 96 000-atom list (5 400 sci entries, 67 620 cjPacked groups): every mask word"""
    from gromacs_b200 import LOCAL
    from gromacs_b200.workload import make_workload
    wl = make_workload("water96k_fswitch")
    plist = wl.pairlist(min_sci=4000)
    po = orc_params(oracle, wl)
    run = Production(wl, plist)
    try:
        run.step(0)
        cj_gpu, outer_gpu, sci_sorted, sci_count, _ = run.nb.download_pairlist()
        cj_o = plist.cjPacked.copy()
        outer_o = np.zeros(2 * cj_o.shape[0], np.uint32)
        cnt_o = oracle.prune(po, plist.sci, cj_o, outer_o, wl.nbat.xq, wl.nbat.shift_vec, fresh=True)
        assert np.array_equal(cj_gpu, cj_o) and np.array_equal(outer_gpu, outer_o) and np.array_equal(sci_count, cnt_o)
        assert (cj_o[:, 4] != plist.cjPacked[:, 4]).any(), "the first pass pruned nothing"
        rng = np.random.default_rng(11)
        xq = wl.nbat.xq.copy()
        for step in range(2 * NUM_PARTS):
            xq[:, :3] += rng.normal(0, 0.004, (xq.shape[0], 3)).astype(np.float32)
            wl.nbat.xq[:] = xq
            run.nb.gpu_copy_xq_to_gpu(wl.nbat, LOCAL)
            run.nb.gpu_launch_kernel_pruneonly(LOCAL, NUM_PARTS)
            before = cj_o.copy()
            oracle.prune(po, sci_sorted, cj_o, outer_o, xq, wl.nbat.shift_vec, fresh=False, part=step % NUM_PARTS, nparts=NUM_PARTS)
            cj_gpu, outer_gpu, _, _, _ = run.nb.download_pairlist()
            assert np.array_equal(cj_gpu, cj_o), "rolling prune step %d" % step
            assert np.array_equal(outer_gpu, outer_o)
            assert (cj_o != before).any(), "rolling step %d changed nothing: the test moves too little" % step
    finally:
        run.close()


def replicated_reference(oracle, wl):
    """Double-oracle forces for every atom of a benchmark box from the 2 x 2 x 2 box (24 000 atoms) with the same interaction
    settings.  BenchmarkSystem stacks translated copies of a 3000-atom unit box whose molecules were wrapped into the
    box (bench_system.cpp:95-158): a molecule cut by a unit-box face meets the other half of its neighbouring copy's
    molecule there (not excluded, as in the reference's own benchmark), the same way at every face as soon as there are
    two or more copies along a dimension.  So every box with >= 2 copies per dimension repeats the 2 x 2 x 2 box; the unit
    box itself (1 copy: the halves are the same molecule, excluded) does not."""
    import gromacs_b200.workload as W
    W.CONFIGS["_box222"] = dict(wl.cfg, k=8, energy=True, dynamic_pruning=False, rlist_inner=wl.cfg["rlist_outer"])
    try:
        w8 = W.make_workload("_box222")
    finally:
        del W.CONFIGS["_box222"]
    pl = w8.pairlist(min_sci=0)
    f8, _, e8, _ = oracle_forces(oracle, w8, pl.sci, pl.cjPacked, pl.excl)
    f8 = oracle.nbat_to_atom_order(f8, w8.grid.atom_index, w8.box.natoms)
    n0 = w8.box.natoms // 8
    fx, fy, fz = wl.box.factors
    assert min(fx, fy, fz) >= 2 and w8.box.factors == (2, 2, 2)
    ix, iy, iz = np.meshgrid(np.arange(fx), np.arange(fy), np.arange(fz), indexing="ij")
    copy8 = (((ix % 2) * 2 + (iy % 2)) * 2 + (iz % 2)).reshape(-1)          # copies are stacked x-major, z fastest
    ref = f8.reshape(8, n0, 3)[copy8].reshape(-1, 3)
    return ref, e8 * (fx * fy * fz / 8.0)


def test_water12m_production_sequence(oracle):
    """BASELINE configs[4], the box every bench line is quoted on."""
    from gromacs_b200 import PairlistGpu
    from gromacs_b200.workload import make_workload
    wl = make_workload("water12m")
    plist = wl.pairlist(min_sci=18944)
    run = Production(wl, plist)
    try:
        f0, _ = run.step(0)
        # (ii) first-pass prune masks of every 400th entry, bit for bit
        cj_gpu, outer_gpu, _, _, _ = run.nb.download_pairlist()
        ngroups = check_first_pass_masks(oracle, wl, plist, cj_gpu, outer_gpu, plist.sci[::400])
        assert ngroups > 20000
        for i in range(1, 2 * NUM_PARTS + 1):
            f, _ = run.step(i)
        # rolling pruning only drops pairs beyond the cut-off
        assert relrms(f, f0) <= 2e-6
        _, (e_lj_all, e_el_all) = run.step(2 * NUM_PARTS + 1, energy=True)
        cj_pruned, _, _, _, _ = run.nb.download_pairlist()
    finally:
        run.close()
    assert (cj_pruned[:, 4::2] != plist.cjPacked[:, 4::2]).any()

    # (i) every 400th sci entry of the pruned production list: packed kernel vs double oracle on the same entries
    sub = PairlistGpu(sci=plist.sci[::400], cjPacked=cj_pruned, excl=plist.excl)
    f_ref, _, e_ref, npairs = oracle_forces(oracle, wl, sub.sci, sub.cjPacked, sub.excl)
    assert npairs > 10_000_000
    run = Production(wl, sub, dynamic_pruning=False)
    try:
        f_sub, _ = run.step(0)
        f_sub_e, (e_lj, e_el) = run.step(1, energy=True)
    finally:
        run.close()
    assert relrms(f_sub, f_ref) <= 5e-6, relrms(f_sub, f_ref)
    assert np.abs(f_sub - f_ref).max() <= 1e-4 * np.abs(f_ref).max()
    assert relrms(f_sub_e, f_ref) <= 5e-6
    assert abs(e_el - e_ref[1]) <= 1e-6 * abs(e_ref[1]), (e_el, e_ref[1])
    assert abs(e_lj - e_ref[0]) <= 1e-6 * abs(e_ref[0]), (e_lj, e_ref[0])

    # (iii) every atom of the 12.3 M against its image in the 2 x 2 x 2 box
    ref, e_all_ref = replicated_reference(oracle, wl)
    # energies of the WHOLE 12.3 M-atom system: 512 times the oracle's energies of the 2 x 2 x 2 box
    assert abs(e_el_all - e_all_ref[1]) <= 1e-6 * abs(e_all_ref[1]), (e_el_all, e_all_ref[1])
    assert abs(e_lj_all - e_all_ref[0]) <= 1e-6 * abs(e_all_ref[0]), (e_lj_all, e_all_ref[0])
    f_atoms = oracle.nbat_to_atom_order(f, wl.grid.atom_index, wl.box.natoms)
    assert relrms(f_atoms, ref) <= 2e-4, relrms(f_atoms, ref)
    assert np.abs(f_atoms - ref).max() <= 5e-3 * np.abs(ref).max(), np.abs(f_atoms - ref).max() / np.abs(ref).max()
