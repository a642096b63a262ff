"""GPU tests of the pair-list builder on the device (nbnxm_b200_gpu_search_*, gromacs_b200/csrc/nbnxm_gpusearch.cu):
the list built from the coordinates resident in the handle must equal the host builder's (one thread) entry for
entry — integer / mask work, so bit-exact — and the force kernel run on the list installed from device memory must
give the forces it gives with the host builder's list."""
import json
import os
import time

import numpy as np
import pytest

from test_gpusearch_emu import assert_same_list
from util import load_golden, product_params, relrms

pytestmark = pytest.mark.gpu


def golden_system(case):
    from gromacs_b200.pairsearch import Grid
    d = load_golden(case)
    grid = Grid(d["sys_box"], d["sys_x"], nthreads=1)
    nt = int(d["nbat_ntypes"][0])
    nbat = grid.atomdata(d["sys_x"], d["sys_q"], d["sys_type"], d["nbat_nbfp"], nt, nbfp_comb=d["nbat_nbfp_comb"])
    return d, grid, nbat


def run_step(nb, nbat, energy=True):
    from gromacs_b200 import LOCAL, StepWorkload
    sw = StepWorkload(computeEnergy=energy, computeVirial=energy)
    nb.gpu_clear_outputs(True)
    nb.gpu_launch_kernel(sw, LOCAL)
    nb.gpu_launch_cpyback(nbat, sw, LOCAL)
    e = nb.gpu_wait_finish_task(sw, LOCAL)
    return nbat.f.astype(np.float64).copy(), e


@pytest.mark.parametrize("case,rlist,min_sci", [("test243_ewald_cutnone", 0.9, 0), ("bench1_ewald_cutnone", 1.0, 0),
                                                ("bench1_ewald_cutnone", 1.05, 500)])
def test_device_list_equals_host_list_and_gives_the_same_forces(case, rlist, min_sci):
    from gromacs_b200 import LOCAL, NbnxmGpu
    from gromacs_b200.pairsearch import GpuPairSearch
    d, grid, nbat = golden_system(case)
    ref = grid.pairlist(rlist, d["sys_excl_index"], d["sys_excl_atoms"], min_sci=min_sci)
    nb = NbnxmGpu(product_params(d, vdw="Cut"), nbat)
    try:
        nb.gpu_init_atomdata(nbat)
        nb.gpu_upload_shiftvec(nbat)
        nb.gpu_copy_xq_to_gpu(nbat, LOCAL)
        search = GpuPairSearch(nb, grid, d["sys_excl_index"], d["sys_excl_atoms"])
        sizes = search.build(rlist, LOCAL, min_sci=min_sci)
        got = search.download()
        assert sizes == (ref.sci.shape[0], ref.cjPacked.shape[0], ref.excl.shape[0])
        assert_same_list((got.sci, got.cjPacked, got.excl), (ref.sci, ref.cjPacked, ref.excl))
        assert search.nci_tot == ref.nci_tot
        nb.setupGpuShortRangeWork(LOCAL)
        f_dev, e_dev = run_step(nb, nbat)
        # a second build reuses every buffer
        assert search.build(rlist, LOCAL, min_sci=min_sci) == sizes
        f_dev2, _ = run_step(nb, nbat)
        nb.gpu_init_pairlist(ref, LOCAL)
        f_host, e_host = run_step(nb, nbat)
        search.free()
    finally:
        nb.gpu_free()
    # same entries in the same order: only the order of the floating-point atomics differs
    assert relrms(f_dev, f_host) < 1e-6 and relrms(f_dev2, f_host) < 1e-6
    assert abs(e_dev[0] - e_host[0]) <= 1e-6 * abs(e_host[0]) and abs(e_dev[1] - e_host[1]) <= 1e-6 * abs(e_host[1])


def test_device_list_of_the_96k_box(oracle):
    """BASELINE configs[1] (96 000 atoms, rlist 1.18): two-level scans, 67 k cjPacked groups, 14 k exclusion entries;
    records the device time of the build next to the host builder's wall time."""
    from gromacs_b200 import LOCAL, NbnxmGpu
    from gromacs_b200.pairsearch import GpuPairSearch
    from gromacs_b200.workload import make_workload
    wl = make_workload("water96k_fswitch", nthreads=1)
    t0 = time.time()
    ref = wl.pairlist(min_sci=9000)
    host_s = time.time() - t0
    nb = NbnxmGpu(wl.params, wl.nbat)
    try:
        nb.gpu_init_atomdata(wl.nbat)
        nb.gpu_upload_shiftvec(wl.nbat)
        nb.gpu_copy_xq_to_gpu(wl.nbat, LOCAL)
        search = GpuPairSearch(nb, wl.grid, wl.box.excl_index, wl.box.excl_atoms)
        search.build(wl.cfg["rlist_outer"], LOCAL, min_sci=9000)          # warm-up: allocations
        search.build(wl.cfg["rlist_outer"], LOCAL, min_sci=9000)
        got = search.download()
        assert_same_list((got.sci, got.cjPacked, got.excl), (ref.sci, ref.cjPacked, ref.excl))
        nb.setupGpuShortRangeWork(LOCAL)
        f_dev, _ = run_step(nb, wl.nbat)
        nb.gpu_init_pairlist(ref, LOCAL)
        f_host, _ = run_step(nb, wl.nbat)
        ms = search.build_ms
        search.free()
    finally:
        nb.gpu_free()
    assert relrms(f_dev, f_host) < 1e-6
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "gpu_search_96k.json"), "w") as fh:
            json.dump({"workload": "water96k_fswitch", "natoms": wl.box.natoms, "rlist": wl.cfg["rlist_outer"],
                       "ncj_packed": int(ref.cjPacked.shape[0]), "nsci": int(ref.sci.shape[0]),
                       "gpu_build_ms": ms, "host_build_1thread_s": host_s}, fh)


def test_device_list_of_the_1536k_box():
    """BASELINE configs[3] (1 536 000 atoms): 1.0 M cjPacked groups, 0.22 M exclusion entries, two-level scans of 2 M
    items; equality with the host builder's list (all threads: same entries, other exclusion numbering) and timing."""
    from gromacs_b200 import LOCAL, NbnxmGpu
    from gromacs_b200.pairsearch import GpuPairSearch
    from gromacs_b200.workload import make_workload
    wl = make_workload("water1536k")
    t0 = time.time()
    ref = wl.pairlist(min_sci=18944)
    host_s = time.time() - t0
    nb = NbnxmGpu(wl.params, wl.nbat)
    try:
        nb.gpu_init_atomdata(wl.nbat)
        nb.gpu_upload_shiftvec(wl.nbat)
        nb.gpu_copy_xq_to_gpu(wl.nbat, LOCAL)
        search = GpuPairSearch(nb, wl.grid, wl.box.excl_index, wl.box.excl_atoms)
        search.build(wl.cfg["rlist_outer"], LOCAL, min_sci=18944)
        search.build(wl.cfg["rlist_outer"], LOCAL, min_sci=18944)
        ms = search.build_ms
        got = search.download()
        search.free()
    finally:
        nb.gpu_free()
    assert_same_list((got.sci, got.cjPacked, got.excl), (ref.sci, ref.cjPacked, ref.excl))
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "gpu_search_1536k.json"), "w") as fh:
            json.dump({"workload": "water1536k", "natoms": wl.box.natoms, "rlist": wl.cfg["rlist_outer"],
                       "ncj_packed": int(ref.cjPacked.shape[0]), "nsci": int(ref.sci.shape[0]), "gpu_build_ms": ms,
                       "host_build_s": host_s, "host_threads": wl.grid.nthreads}, fh)


def test_force_reduction_gathers_nbat_forces_into_atom_order():
    """nbnxm_b200_reduce_f (reduceKernel, mdlib/gpuforcereduction_impl_internal.cu:61-118): set / accumulate, with and
    without an extra rvec force, on an atom sub-range, against the copied-back nbat forces."""
    import torch
    from gromacs_b200 import LOCAL, NbnxmGpu, StepWorkload
    d, grid, nbat = golden_system("bench1_ewald_cutnone")
    plist = grid.pairlist(0.9, d["sys_excl_index"], d["sys_excl_atoms"])
    n = d["sys_x"].shape[0]
    cell = np.zeros(n, np.int32)
    slots = np.nonzero(grid.atom_index >= 0)[0]
    cell[grid.atom_index[slots]] = slots
    nb = NbnxmGpu(product_params(d, vdw="Cut"), nbat)
    try:
        nb.gpu_init_atomdata(nbat)
        nb.gpu_upload_shiftvec(nbat)
        nb.gpu_copy_xq_to_gpu(nbat, LOCAL)
        nb.gpu_init_pairlist(plist, LOCAL)
        nb.setupGpuShortRangeWork(LOCAL)
        f_nbat, _ = run_step(nb, nbat, energy=False)
        f_atoms = f_nbat[cell].astype(np.float32)
        nb.gpu_force_reduction_reinit(cell)
        rng = np.random.default_rng(7)
        extra = rng.standard_normal((n, 3)).astype(np.float32)
        base = rng.standard_normal((n, 3)).astype(np.float32)
        d_extra = torch.from_numpy(extra).cuda()
        stream = nb.streams()[0]
        for accumulate in (False, True):
            for with_extra in (False, True):
                d_total = torch.from_numpy(base).cuda()
                torch.cuda.synchronize()
                a0, na = 64, n - 100
                nb.gpu_force_reduction_execute(d_total.data_ptr(), d_extra.data_ptr() if with_extra else None, a0, na,
                                               accumulate, stream)
                nb.gpu_wait_finish_task(StepWorkload(), LOCAL)
                torch.cuda.synchronize()
                want = base.copy()
                sl = slice(a0, a0 + na)
                want[sl] = (base[sl] if accumulate else 0) + f_atoms[sl] + (extra[sl] if with_extra else 0)
                got = d_total.cpu().numpy()
                assert np.array_equal(got[:a0], base[:a0]) and np.array_equal(got[a0 + na:], base[a0 + na:])
                assert np.allclose(got[sl], want[sl], rtol=1e-6, atol=1e-4 * np.abs(f_atoms).max() * 1e-3)
    finally:
        nb.gpu_free()


def device_search_step(nb, box, x, q, atom_type, ntypes, ei, ea, rlist, min_sci, lj_comb_per_type=None):
    """putAtomsOnGrid + constructPairlist on the device from coordinates in device memory (atom order)"""
    import torch
    from gromacs_b200 import LOCAL
    from gromacs_b200.pairsearch import GpuPairSearch
    x_dev = torch.from_numpy(np.ascontiguousarray(x, np.float32)).cuda()
    torch.cuda.synchronize()
    search = GpuPairSearch(nb)
    search.set_atoms(q, atom_type, ntypes, lj_comb_per_type, ei, ea)
    dims = search.put_atoms_on_grid(box, x_dev.data_ptr())
    search.put_atoms_on_grid(box, x_dev.data_ptr())        # again: every buffer is reused
    atom_index, first_bin, grid_ms = search.get_order()
    sizes = search.build(rlist, LOCAL, min_sci=min_sci)
    sizes = search.build(rlist, LOCAL, min_sci=min_sci)
    return search, x_dev, dims, atom_index, first_bin, grid_ms, sizes


@pytest.mark.parametrize("case,rlist,min_sci", [("bench1_ewald_cutnone", 1.0, 0), ("test243_ewald_cutnone", 0.9, 50)])
def test_search_step_entirely_on_the_device(case, rlist, min_sci):
    """grid order, atom data, list and forces from device-resident atom-order coordinates equal the host path's; the
    force reduction maps the forces back to atom order with the device-built atom -> slot map"""
    import torch
    from gromacs_b200 import LOCAL, NbnxmGpu, StepWorkload
    d, grid, nbat = golden_system(case)
    n = d["sys_x"].shape[0]
    ref = grid.pairlist(rlist, d["sys_excl_index"], d["sys_excl_atoms"], min_sci=min_sci)
    nb = NbnxmGpu(product_params(d, vdw="Cut"), nbat)
    try:
        nb.gpu_upload_shiftvec(nbat)
        search, x_dev, dims, atom_index, first_bin, _, sizes = device_search_step(
            nb, d["sys_box"], d["sys_x"], d["sys_q"], d["sys_type"], int(d["nbat_ntypes"][0]), d["sys_excl_index"],
            d["sys_excl_atoms"], rlist, min_sci)
        assert dims == (grid.natoms_nbat, grid.nbins, grid.ncx, grid.ncy)
        assert np.array_equal(atom_index, grid.atom_index) and np.array_equal(first_bin, grid.first_bin_of_column)
        got = search.download()
        assert_same_list((got.sci, got.cjPacked, got.excl), (ref.sci, ref.cjPacked, ref.excl))
        nb.setupGpuShortRangeWork(LOCAL)
        f_dev, e_dev = run_step(nb, nbat)
        # forces back in atom order through the device-built cell map
        d_total = torch.zeros((n, 3), dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        nb.gpu_force_reduction_execute(d_total.data_ptr(), None, 0, n, False, nb.streams()[0])
        nb.gpu_wait_finish_task(StepWorkload(), LOCAL)
        torch.cuda.synchronize()
        slots = np.nonzero(grid.atom_index >= 0)[0]
        f_atoms = np.zeros((n, 3))
        f_atoms[grid.atom_index[slots]] = f_dev[slots]
        assert np.array_equal(d_total.cpu().numpy(), f_atoms.astype(np.float32))
        search.free()
        # the host path on the same handle
        nb.gpu_init_atomdata(nbat)
        nb.gpu_copy_xq_to_gpu(nbat, LOCAL)
        nb.gpu_init_pairlist(ref, LOCAL)
        f_host, e_host = run_step(nb, nbat)
    finally:
        nb.gpu_free()
    assert relrms(f_dev, f_host) < 1e-6
    assert abs(e_dev[0] - e_host[0]) <= 1e-6 * abs(e_host[0]) and abs(e_dev[1] - e_host[1]) <= 1e-6 * abs(e_host[1])


def test_search_step_entirely_on_the_device_1536k():
    """BASELINE configs[3]: 784 columns of up to 2176 atoms (4096-element sorting networks), grid and list equal to the
    host's; records the device times of gridding and list construction next to the host's wall times."""
    from gromacs_b200 import LOCAL, NbnxmGpu
    from gromacs_b200.pairsearch import Grid
    from gromacs_b200.workload import make_workload
    wl = make_workload("water1536k")
    t0 = time.time()
    Grid(wl.box.box, wl.box.x)
    host_grid_s = time.time() - t0
    t0 = time.time()
    ref = wl.pairlist(min_sci=18944)
    host_list_s = time.time() - t0
    nb = NbnxmGpu(wl.params, wl.nbat)
    try:
        search, x_dev, dims, atom_index, first_bin, grid_ms, sizes = device_search_step(
            nb, wl.box.box, wl.box.x, wl.box.q, wl.box.type, wl.nbat.numTypes, wl.box.excl_index, wl.box.excl_atoms,
            wl.cfg["rlist_outer"], 18944, lj_comb_per_type=wl.nbat.nbfp_comb)      # LJ cut with the geometric rule
        build_ms = search.build_ms
        got = search.download()
        # forces with everything built on the device against the host path on the same handle
        nb.gpu_upload_shiftvec(wl.nbat)
        nb.setupGpuShortRangeWork(LOCAL)
        f_dev, _ = run_step(nb, wl.nbat, energy=False)
        search.free()
        nb.gpu_init_atomdata(wl.nbat)
        nb.gpu_copy_xq_to_gpu(wl.nbat, LOCAL)
        nb.gpu_init_pairlist(ref, LOCAL)
        f_host, _ = run_step(nb, wl.nbat, energy=False)
    finally:
        nb.gpu_free()
    assert dims == (wl.grid.natoms_nbat, wl.grid.nbins, wl.grid.ncx, wl.grid.ncy)
    assert np.array_equal(atom_index, wl.grid.atom_index) and np.array_equal(first_bin, wl.grid.first_bin_of_column)
    assert_same_list((got.sci, got.cjPacked, got.excl), (ref.sci, ref.cjPacked, ref.excl))
    assert relrms(f_dev, f_host) < 1e-6
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "gpu_search_step_1536k.json"), "w") as fh:
            json.dump({"workload": "water1536k", "natoms": wl.box.natoms, "gpu_grid_ms": grid_ms, "gpu_list_ms": build_ms,
                       "host_grid_s": host_grid_s, "host_list_s": host_list_s, "host_threads": wl.grid.nthreads}, fh)


def test_cooperative_mask_pass_gives_the_same_list(monkeypatch):
    """pass 3 as one warp per bin pair (the default) and as one thread per j-cluster (NBNXM_B200_SEARCH_COOP=0): same list; the build times go to gpurun_out"""
    from gromacs_b200 import LOCAL, NbnxmGpu
    from gromacs_b200.pairsearch import GpuPairSearch
    from gromacs_b200.workload import make_workload
    wl = make_workload("water1536k")
    ref = wl.pairlist(min_sci=18944)
    times = {}
    for coop in ("0", "1"):
        monkeypatch.setenv("NBNXM_B200_SEARCH_COOP", coop)
        nb = NbnxmGpu(wl.params, wl.nbat)
        try:
            nb.gpu_init_atomdata(wl.nbat)
            nb.gpu_copy_xq_to_gpu(wl.nbat, LOCAL)
            search = GpuPairSearch(nb, wl.grid, wl.box.excl_index, wl.box.excl_atoms)
            for _ in range(3):
                search.build(wl.cfg["rlist_outer"], LOCAL, min_sci=18944)
            times[coop] = search.build_ms
            got = search.download()
            search.free()
        finally:
            nb.gpu_free()
        assert_same_list((got.sci, got.cjPacked, got.excl), (ref.sci, ref.cjPacked, ref.excl))
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "gpu_search_coop_1536k.json"), "w") as fh:
            json.dump({"workload": "water1536k", "build_ms_thread_per_j_cluster": times["0"], "build_ms_warp_per_bin_pair": times["1"]}, fh)


@pytest.mark.parametrize("name,rlist", [("bench3k", 0.9), ("water24k_test", 0.95)])
def test_device_list_against_brute_force(oracle, name, rlist):
    """Independent of every product-side builder: the list built on the device, downloaded and walked by the double
    oracle, must give the forces and energies of the oracle's O(N^2) brute force over all atom pairs with the minimum
    image convention (no list at all) - a missing cluster pair, a wrong shift or a wrong exclusion bit shows up here."""
    from gromacs_b200 import LOCAL, NbnxmGpu
    from gromacs_b200.pairsearch import GpuPairSearch
    from gromacs_b200.workload import make_workload
    wl = make_workload(name, nthreads=4)
    g, b = wl.nbat, wl.box
    nb = NbnxmGpu(wl.params, g)
    try:
        nb.gpu_init_atomdata(g)
        nb.gpu_upload_shiftvec(g)
        nb.gpu_copy_xq_to_gpu(g, LOCAL)
        search = GpuPairSearch(nb, wl.grid, b.excl_index, b.excl_atoms)
        search.build(rlist, LOCAL, min_sci=0)
        got = search.download()
        search.free()
    finally:
        nb.gpu_free()
    p = oracle.OrcParams()
    for field, _ in wl.params._fields_:
        if hasattr(p, field):
            setattr(p, field, getattr(wl.params, field))
    p.ntypes = g.numTypes
    f_list, _, e_list, _ = oracle.forces(p, got.sci, got.cjPacked, got.excl, g.xq, g.type, g.lj_comb, g.nbfp, g.nbfp_comb, g.shift_vec)
    f_list = oracle.nbat_to_atom_order(f_list, wl.grid.atom_index, b.natoms)
    f_bf, e_bf = oracle.brute_force(p, b.x, b.q, b.type, g.nbfp, g.nbfp_comb, b.box, b.excl_index, b.excl_atoms)
    assert relrms(f_list, f_bf) <= 1e-6, relrms(f_list, f_bf)
    assert abs(e_list[0] - e_bf[0]) <= 1e-6 * abs(e_bf[0]) and abs(e_list[1] - e_bf[1]) <= 1e-6 * abs(e_bf[1])


@pytest.mark.parametrize("world", [2, 4])
def test_device_slab_lists_equal_the_host_plan(world):
    """The search step of a multi-GPU run without the host: the whole system gridded on the device, then for every rank its
    atom data gathered and its local (home x home) and non-local (home x halo) lists built, re-indexed and installed on the
    device (nbnxm_b200_gpu_search_gather_slab / _build_slab) - against gromacs_b200.multigpu.make_slab_plan, which does the
    same with the host gridder, the host builder and nbnxm_b200_pairlist_reindex.  Lists entry for entry; forces of both
    kernels on the device-built data against the host-built data (halo coordinates resident: one GPU plays every rank)."""
    import torch
    from gromacs_b200 import LOCAL, NONLOCAL, NbnxmGpu, StepWorkload
    from gromacs_b200.multigpu import make_slab_plan
    from gromacs_b200.pairsearch import GpuPairSearch
    from gromacs_b200.slabs import slab_bin_ranges_from_columns
    from gromacs_b200.workload import make_workload
    wl = make_workload("water48k_test", nthreads=4, nslabs=world)
    rlist, b = wl.cfg["rlist_outer"], wl.box
    whole = NbnxmGpu(wl.params, wl.nbat)            # holds the grid of the whole system on the device
    search = GpuPairSearch(whole)
    search.set_atoms(b.q, b.type, wl.nbat.numTypes, wl.nbat.nbfp_comb, b.excl_index, b.excl_atoms)
    x_dev = torch.from_numpy(np.ascontiguousarray(b.x, np.float32)).cuda()
    torch.cuda.synchronize()
    _, nbins, ncx, ncy = search.put_atoms_on_grid(b.box, x_dev.data_ptr(), nslabs=world)
    _, first_bin, _ = search.get_order()
    assert (ncx, ncy, nbins) == (wl.grid.ncx, wl.grid.ncy, wl.grid.nbins)
    sw = StepWorkload()

    def forces(nb, nbat):
        nb.gpu_upload_shiftvec(nbat)
        nb.setupGpuShortRangeWork(LOCAL)
        nb.setupGpuShortRangeWork(NONLOCAL)
        nb.gpu_clear_outputs(True)
        nb.gpu_launch_kernel(sw, LOCAL)
        nb.gpu_launch_kernel(sw, NONLOCAL)
        nb.gpu_launch_cpyback(nbat, StepWorkload(useGpuFBufferOps=True), NONLOCAL)
        nb.gpu_launch_cpyback(nbat, sw, LOCAL)
        nb.gpu_wait_finish_task(sw, NONLOCAL)
        nb.gpu_wait_finish_task(sw, LOCAL)
        return nbat.f[:nbat.numLocalAtoms].astype(np.float64).copy()

    try:
        for rank in range(world):
            plan = make_slab_plan(wl, rank, world, min_sci=2000)
            home, halo, tx = slab_bin_ranges_from_columns(b.box[0], ncx, ncy, first_bin, world, rank, rlist)
            assert (home, halo) == (plan.home_bins, plan.halo_bins)
            # host path
            nb_h = NbnxmGpu(wl.params, plan.nbat, bLocalAndNonlocal=True)
            nb_h.gpu_init_atomdata(plan.nbat)
            nb_h.gpu_init_pairlist(plan.local, LOCAL)
            nb_h.gpu_init_pairlist(plan.nonlocal_, NONLOCAL)
            nb_h.gpu_copy_xq_to_gpu(plan.nbat, LOCAL)
            nb_h.gpu_copy_xq_to_gpu(plan.nbat, NONLOCAL)
            f_host = forces(nb_h, plan.nbat)
            nb_h.gpu_free()
            # device path
            nb_d = NbnxmGpu(wl.params, plan.nbat, bLocalAndNonlocal=True)
            search.gather_slab(nb_d, home, halo)
            sizes = search.build_slab(nb_d, LOCAL, rlist, home, halo, min_sci=2000)
            assert sizes == (plan.local.sci.shape[0], plan.local.cjPacked.shape[0], plan.local.excl.shape[0])
            got = search.download()
            assert_same_list((got.sci, got.cjPacked, got.excl), (plan.local.sci, plan.local.cjPacked, plan.local.excl))
            search.build_slab(nb_d, NONLOCAL, rlist, home, halo, required_tx=tx, min_sci=1000)
            got = search.download()
            assert_same_list((got.sci, got.cjPacked, got.excl), (plan.nonlocal_.sci, plan.nonlocal_.cjPacked, plan.nonlocal_.excl))
            f_dev = forces(nb_d, plan.nbat)
            nb_d.gpu_free()
            assert relrms(f_dev, f_host) < 1e-6, (rank, relrms(f_dev, f_host))
    finally:
        search.free()
        whole.gpu_free()


@pytest.mark.parametrize("kind", ["tall uniform", "all z equal", "clustered z", "lattice z"])
def test_column_sort_forms_on_stressed_columns(monkeypatch, kind):
    """the gridding's column sort on the device in both forms - buckets + ranks (default) and bitonic networks
    (NBNXM_B200_SEARCH_BITONIC_SORT=1) - on columns of ~7000 atoms: uniform z (bucket count near its cap), every atom at the same z
    (one bucket holds the column), z crowded into 1 % of the range, z on a coarse lattice (ties): the host gridder's order in
    every case (the emulation checks the same systems on the CPU, tests/test_gpusearch_emu.py)"""
    import torch
    from gromacs_b200 import NbnxmGpu
    from gromacs_b200.pairsearch import Grid, GpuPairSearch
    d, _, nbat = golden_system("bench1_ewald_cutnone")
    rng = np.random.default_rng(7)
    box = np.array([1.2, 1.2, 60.0], np.float32)
    n = 7000
    x = rng.random((n, 3)) * box
    if kind == "all z equal":
        x[:, 2] = 30.0
    elif kind == "clustered z":
        x[:, 2] = np.where(rng.random(n) < 0.9, 5.0 + rng.random(n) * 0.01, x[:, 2])
    elif kind == "lattice z":
        x[:, 2] = np.minimum(np.round(x[:, 2] * 4) / 4, 59.75)
    x = np.ascontiguousarray(x, np.float32)
    grid = Grid(box, x, nthreads=2)
    q = np.zeros(n, np.float32)
    t = np.zeros(n, np.int32)
    for bitonic in ("0", "1"):
        monkeypatch.setenv("NBNXM_B200_SEARCH_BITONIC_SORT", bitonic)
        nb = NbnxmGpu(product_params(d, vdw="Cut"), nbat)
        try:
            x_dev = torch.from_numpy(x).cuda()
            torch.cuda.synchronize()
            search = GpuPairSearch(nb)
            search.set_atoms(q, t, int(d["nbat_ntypes"][0]), None, None, None)
            dims = search.put_atoms_on_grid(box, x_dev.data_ptr())
            atom_index, first_bin, _ = search.get_order()
            search.free()
        finally:
            nb.gpu_free()
        assert dims == (grid.natoms_nbat, grid.nbins, grid.ncx, grid.ncy)
        assert np.array_equal(first_bin, grid.first_bin_of_column) and np.array_equal(atom_index, grid.atom_index)
