"""CPU tests of the GPU pair-list builder's passes (gromacs_b200/csrc/gpusearch_bodies.h + gpusearch_driver.h), run
through the host-loop backend of tests/kernel_emu (one loop iteration per CUDA thread, items in reversed order, work
buffers poisoned): the list must equal, entry for entry, the one the independent host builder
(gromacs_b200/csrc/pairsearch.cpp, one thread) makes from the same grid — sci and cjPacked arrays bit for bit, the
exclusion masks after expansion per (cjPacked, half), since the two allocate nbnxm_excl_t entries in a different
order.  The host builder itself is checked against brute force and the reference's lists in test_pairsearch.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from util import load_golden

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module", params=["thread-per-j-cluster masks, bitonic column sort", "warp-per-bin-pair masks, bucket column sort"])
def emu(request):
    """both forms of pass 3 (cluster-pair masks): one thread per (bin pair, j-cluster), and the warp-cooperative one
    (the library's default; NBNXM_B200_SEARCH_COOP=0 selects the other), run lane by lane; and both forms of the column
    sort of the gridding: bitonic networks (NBNXM_B200_SEARCH_BITONIC_SORT=1 in the library) and buckets + ranks (default)"""
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "kernel_emu")], check=True)
    lib = C.CDLL(os.path.join(HERE, "kernel_emu", "libsearch_emu.so"))
    lib.search_emu_set_cooperative_masks(int(request.param.startswith("warp")))
    lib.search_emu_set_bitonic_column_sort(int("bitonic" in request.param))
    return lib


def _p(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


def emu_pairlist(emu, grid, xq, rlist, ei, ea, min_sci=0, bins=None, j_bins=None, inter_zone=False, required_tx=0):
    box = np.ascontiguousarray(grid.box, np.float32)
    ei = None if ei is None else np.ascontiguousarray(ei, np.int32)
    ea = None if ea is None else np.ascontiguousarray(ea, np.int32)
    assert emu.search_emu_set_grid(_p(box, C.c_float), grid.ncx, grid.ncy, _p(grid.first_bin_of_column, C.c_int),
                                   _p(grid.atom_index, C.c_int), grid.nbins, grid.natoms, _p(ei, C.c_int), _p(ea, C.c_int)) == 0
    b0, b1 = bins if bins is not None else (0, grid.nbins)
    j0, j1 = j_bins if j_bins is not None else (0, grid.nbins)
    sizes = (C.c_int * 4)()
    ncp = C.c_longlong()
    xq = np.ascontiguousarray(xq, np.float32)
    assert emu.search_emu_build(_p(xq, C.c_float), C.c_float(rlist), min_sci, b0, b1, j0, j1, int(inter_zone), required_tx,
                                sizes, C.byref(ncp)) == 0
    sci = np.zeros((sizes[0], 4), np.int32)
    cjp = np.zeros((sizes[1], 8), np.uint32)
    excl = np.zeros((sizes[2], 32), np.uint32)
    emu.search_emu_copy(_p(sci, C.c_int), _p(cjp, C.c_uint32), _p(excl, C.c_uint32))
    return sci, cjp, excl, ncp.value


def expanded_excl(cjp, excl):
    """exclusion words per (cjPacked, half): independent of how excl entries were numbered"""
    cjp = np.ascontiguousarray(cjp).view(np.uint32).reshape(-1, 8)
    return excl[cjp[:, 5].astype(np.int64)], excl[cjp[:, 7].astype(np.int64)]


def assert_same_list(a, b):
    sci_a, cjp_a, excl_a = a
    sci_b, cjp_b, excl_b = b
    sci_a = np.ascontiguousarray(sci_a).view(np.int32).reshape(-1, 4)
    sci_b = np.ascontiguousarray(sci_b).view(np.int32).reshape(-1, 4)
    cjp_a = np.ascontiguousarray(cjp_a).view(np.uint32).reshape(-1, 8)
    cjp_b = np.ascontiguousarray(cjp_b).view(np.uint32).reshape(-1, 8)
    assert sci_a.shape == sci_b.shape and np.array_equal(sci_a, sci_b)
    assert cjp_a.shape == cjp_b.shape
    assert np.array_equal(cjp_a[:, [0, 1, 2, 3, 4, 6]], cjp_b[:, [0, 1, 2, 3, 4, 6]])       # cj[4], imask of both halves
    assert np.array_equal(cjp_a[:, [5, 7]] != 0, cjp_b[:, [5, 7]] != 0)                      # who has exclusions
    assert excl_a.shape == excl_b.shape and np.all(excl_a[0] == 0xffffffff)
    for xa, xb in zip(expanded_excl(cjp_a, excl_a), expanded_excl(cjp_b, excl_b)):
        assert np.array_equal(xa, xb)
    # every exclusion entry but the shared entry 0 is used exactly once
    used = np.sort(cjp_a[:, [5, 7]][cjp_a[:, [5, 7]] != 0])
    assert np.array_equal(used, np.arange(1, excl_a.shape[0]))


def host_grid(d, nthreads=1):
    from gromacs_b200.pairsearch import Grid
    grid = Grid(d["sys_box"], d["sys_x"], nthreads=nthreads)
    nt = int(d["nbat_ntypes"][0])
    nbat = grid.atomdata(d["sys_x"], d["sys_q"], d["sys_type"], d["nbat_nbfp"], nt, nbfp_comb=d["nbat_nbfp_comb"])
    return grid, nbat


@pytest.mark.parametrize("case,rlist,min_sci", [("test243_ewald_cutnone", 0.9, 0), ("test243_ewald_cutnone", 0.93, 60),
                                                ("bench1_ewald_cutnone", 1.0, 0), ("bench1_ewald_cutnone", 1.05, 500),
                                                ("bench1_ewald_cutnone", 0.7, 3000)])
def test_passes_give_the_host_builders_list(emu, case, rlist, min_sci):
    d = load_golden(case)
    grid, nbat = host_grid(d)
    ref = grid.pairlist(rlist, d["sys_excl_index"], d["sys_excl_atoms"], min_sci=min_sci)
    sci, cjp, excl, ncp = emu_pairlist(emu, grid, nbat.xq, rlist, d["sys_excl_index"], d["sys_excl_atoms"], min_sci=min_sci)
    assert_same_list((sci, cjp, excl), (ref.sci, ref.cjPacked, ref.excl))
    assert ncp == ref.nci_tot
    assert cjp.shape[0] > 0 and excl.shape[0] > 1


def test_passes_without_topology_exclusions(emu):
    d = load_golden("bench1_ewald_cutnone")
    grid, nbat = host_grid(d)
    ref = grid.pairlist(0.9)
    got = emu_pairlist(emu, grid, nbat.xq, 0.9, None, None)
    assert_same_list(got[:3], (ref.sci, ref.cjPacked, ref.excl))


@pytest.mark.parametrize("nslabs", [2, 3])
def test_passes_build_slab_and_halo_lists(emu, nslabs):
    """bin ranges and the inter-zone mode (home x-slab against its +x neighbour's halo, one x shift)"""
    from gromacs_b200.pairsearch import Grid
    from gromacs_b200.slabs import slab_bin_ranges
    d = load_golden("bench1_ewald_cutnone")
    x = np.concatenate([d["sys_x"] + np.array([i * d["sys_box"][0], 0, 0], np.float32) for i in range(4)])
    n0 = d["sys_x"].shape[0]
    box = d["sys_box"] * np.array([4, 1, 1], np.float32)
    ei = np.concatenate([d["sys_excl_index"][:-1] + i * d["sys_excl_atoms"].shape[0] for i in range(4)]
                        + [[4 * d["sys_excl_atoms"].shape[0]]]).astype(np.int32)
    ea = np.concatenate([d["sys_excl_atoms"] + i * n0 for i in range(4)]).astype(np.int32)
    grid = Grid(box, x, nthreads=1)
    nt = int(d["nbat_ntypes"][0])
    nbat = grid.atomdata(x, np.tile(d["sys_q"], 4), np.tile(d["sys_type"], 4), d["nbat_nbfp"], nt, nbfp_comb=d["nbat_nbfp_comb"])
    rlist = 1.0
    for r in range(nslabs):
        home, halo, tx = slab_bin_ranges(grid, nslabs, r, rlist)
        ref = grid.pairlist(rlist, ei, ea, bins=home, j_bins=home, min_sci=200)
        got = emu_pairlist(emu, grid, nbat.xq, rlist, ei, ea, bins=home, j_bins=home, min_sci=200)
        assert_same_list(got[:3], (ref.sci, ref.cjPacked, ref.excl))
        ref = grid.pairlist(rlist, ei, ea, bins=home, j_bins=halo, inter_zone=True, required_tx=tx)
        got = emu_pairlist(emu, grid, nbat.xq, rlist, ei, ea, bins=home, j_bins=halo, inter_zone=True, required_tx=tx)
        assert_same_list(got[:3], (ref.sci, ref.cjPacked, ref.excl))
        assert got[0].shape[0] > 0


def test_passes_on_an_empty_range(emu):
    d = load_golden("test243_ewald_cutnone")
    grid, nbat = host_grid(d)
    sci, cjp, excl, ncp = emu_pairlist(emu, grid, nbat.xq, 0.9, None, None, bins=(0, 0))
    assert sci.shape[0] == 0 and cjp.shape[0] == 0 and excl.shape[0] == 1 and np.all(excl == 0xffffffff) and ncp == 0
    sci, cjp, excl, ncp = emu_pairlist(emu, grid, nbat.xq, 0.9, None, None, j_bins=(0, 0))
    assert sci.shape[0] == 0 and cjp.shape[0] == 0 and excl.shape[0] == 1


def emu_grid(emu, box, x, ncx, ncy, ei=None, ea=None):
    box = np.ascontiguousarray(box, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    ei = None if ei is None else np.ascontiguousarray(ei, np.int32)
    ea = None if ea is None else np.ascontiguousarray(ea, np.int32)
    nbins = C.c_int()
    assert emu.search_emu_put_atoms_on_grid(_p(box, C.c_float), ncx, ncy, x.shape[0], _p(x, C.c_float), _p(ei, C.c_int),
                                            _p(ea, C.c_int), C.byref(nbins)) == 0
    atom_index = np.zeros(nbins.value * 64, np.int32)
    first_bin = np.zeros(ncx * ncy + 1, np.int32)
    emu.search_emu_get_order(_p(atom_index, C.c_int), _p(first_bin, C.c_int))
    return nbins.value, atom_index, first_bin


@pytest.mark.parametrize("case", ["test243_ewald_cutnone", "bench1_ewald_cutnone"])
def test_gridding_passes_give_the_host_gridders_order(emu, case):
    """putAtomsOnGrid as passes (column histogram, scans, scatter, one bitonic-network block per column): the same
    columns, the same atom order and the same atom data as nbnxm_b200_grid_create / _grid_fill_atomdata."""
    d = load_golden(case)
    grid, nbat = host_grid(d)
    nbins, atom_index, first_bin = emu_grid(emu, d["sys_box"], d["sys_x"], grid.ncx, grid.ncy)
    assert nbins == grid.nbins
    assert np.array_equal(first_bin, grid.first_bin_of_column)
    assert np.array_equal(atom_index, grid.atom_index)
    n = d["sys_x"].shape[0]
    nt = int(d["nbat_ntypes"][0])
    xq = np.zeros((nbins * 64, 4), np.float32)
    tn = np.zeros(nbins * 64, np.int32)
    x = np.ascontiguousarray(d["sys_x"], np.float32)
    q = np.ascontiguousarray(d["sys_q"], np.float32)
    t = np.ascontiguousarray(d["sys_type"], np.int32)
    assert emu.search_emu_fill_atomdata(n, _p(x, C.c_float), _p(q, C.c_float), _p(t, C.c_int), nt, None, _p(xq, C.c_float),
                                        _p(tn, C.c_int), None) == 0
    assert np.array_equal(xq, nbat.xq) and np.array_equal(tn, nbat.type)


def test_gridding_and_list_from_raw_coordinates(emu):
    """the whole search step as passes, on a box with ragged columns (a water box with a slab of atoms removed, so
    that column heights differ and some columns are empty): grid, then the list, equal to the host's"""
    from gromacs_b200.pairsearch import Grid
    d = load_golden("bench1_ewald_cutnone")
    x, box = d["sys_x"], d["sys_box"]
    mol = np.arange(x.shape[0]) // 3
    ox = x[mol * 3]                                     # oxygen of each molecule decides
    keep = ~((ox[:, 0] < 0.9) & (ox[:, 1] < 0.9)) & ~((ox[:, 2] > 1.0) & (ox[:, 2] < 1.7) & (ox[:, 0] > 2.0))
    x = np.ascontiguousarray(x[keep])
    n = x.shape[0]
    assert n % 3 == 0 and n < d["sys_x"].shape[0]
    ei = (np.arange(n + 1) * 3).astype(np.int32)       # each atom excludes the three atoms of its molecule
    ea = (np.repeat(np.arange(n) // 3 * 3, 3) + np.tile(np.arange(3), n)).astype(np.int32)
    grid = Grid(box, x, nthreads=1)
    nbins, atom_index, first_bin = emu_grid(emu, box, x, grid.ncx, grid.ncy, ei, ea)
    assert np.array_equal(first_bin, grid.first_bin_of_column) and np.array_equal(atom_index, grid.atom_index)
    assert (np.diff(first_bin) == 0).any() or len(set(np.diff(first_bin).tolist())) > 1
    nbat = grid.atomdata(x, np.zeros(n, np.float32), np.zeros(n, np.int32), d["nbat_nbfp"], int(d["nbat_ntypes"][0]))
    ref = grid.pairlist(1.0, ei, ea, min_sci=300)
    xq = np.ascontiguousarray(nbat.xq, np.float32)
    sizes = (C.c_int * 4)()
    ncp = C.c_longlong()
    assert emu.search_emu_build(_p(xq, C.c_float), C.c_float(1.0), 300, 0, nbins, 0, nbins, 0, 0, sizes, C.byref(ncp)) == 0
    sci = np.zeros((sizes[0], 4), np.int32)
    cjp = np.zeros((sizes[1], 8), np.uint32)
    excl = np.zeros((sizes[2], 32), np.uint32)
    emu.search_emu_copy(_p(sci, C.c_int), _p(cjp, C.c_uint32), _p(excl, C.c_uint32))
    assert_same_list((sci, cjp, excl), (ref.sci, ref.cjPacked, ref.excl))


def chain_system(n, box, seed, reach=2):
    """random coordinates, charges and types; every atom excludes its neighbours within `reach` along the chain, so
    that exclusions cross clusters, bins and columns (a water box only has exclusions inside a cluster or two)"""
    rng = np.random.default_rng(seed)
    x = (rng.random((n, 3)) * box).astype(np.float32)
    x[:5] = 0.0                                      # atoms on the box corner / coinciding columns edge
    x[5, :] = box * (1 - 1e-7)
    q = rng.uniform(-0.5, 0.5, n).astype(np.float32)
    t = rng.integers(0, 2, n).astype(np.int32)
    ei, ea = [0], []
    for a in range(n):
        ea += [b for b in range(max(0, a - reach), min(n, a + reach + 1))]
        ei.append(len(ea))
    return x, q, t, np.array(ei, np.int32), np.array(ea, np.int32)


@pytest.mark.parametrize("n,seed,rlist,min_sci", [(1500, 1, 0.8, 0), (2500, 2, 1.1, 400), (700, 3, 1.2, 0)])
def test_passes_on_a_random_chain_system_against_host_builder_and_brute_force(emu, oracle, n, seed, rlist, min_sci):
    from gromacs_b200.pairsearch import Grid
    from util import oracle_params, relrms
    d = load_golden("bench1_rf_cutnone_split")
    box = np.array([2.6, 3.1, 3.7], np.float32)
    x, q, t, ei, ea = chain_system(n, box, seed)
    # separate the coinciding atoms a little: brute force and kernels clamp r^2 differently at r = 0
    x[:5] += np.arange(5, dtype=np.float32)[:, None] * 0.11
    grid = Grid(box, x, nthreads=2)
    nt = int(d["nbat_ntypes"][0])
    nbat = grid.atomdata(x, q, t, d["nbat_nbfp"], nt, nbfp_comb=d["nbat_nbfp_comb"])
    ref = grid.pairlist(rlist, ei, ea, min_sci=min_sci)
    # the device builder's passes: grid, then list, from the raw coordinates
    nbins, atom_index, first_bin = emu_grid(emu, box, x, grid.ncx, grid.ncy, ei, ea)
    assert np.array_equal(atom_index, grid.atom_index) and np.array_equal(first_bin, grid.first_bin_of_column)
    sizes = (C.c_int * 4)()
    ncp = C.c_longlong()
    xq = np.ascontiguousarray(nbat.xq, np.float32)
    assert emu.search_emu_build(_p(xq, C.c_float), C.c_float(rlist), min_sci, 0, nbins, 0, nbins, 0, 0, sizes, C.byref(ncp)) == 0
    sci = np.zeros((sizes[0], 4), np.int32)
    cjp = np.zeros((sizes[1], 8), np.uint32)
    excl = np.zeros((sizes[2], 32), np.uint32)
    emu.search_emu_copy(_p(sci, C.c_int), _p(cjp, C.c_uint32), _p(excl, C.c_uint32))
    assert_same_list((sci, cjp, excl), (ref.sci, ref.cjPacked, ref.excl))
    # and the list is right: forces through the oracle's list walk = brute force over all pairs
    p = oracle_params(oracle, d)
    rc = min(rlist, 0.5 * float(box.min()) - 0.01) - 0.05
    p.rcoulomb_sq = p.rvdw_sq = rc * rc
    f, _, e, _ = oracle.forces(p, sci, cjp, excl, nbat.xq, nbat.type, np.zeros((nbat.numAtoms(), 2), np.float32), nbat.nbfp,
                               nbat.nbfp_comb, nbat.shift_vec)
    f = oracle.nbat_to_atom_order(f, grid.atom_index, n)
    fb, eb = oracle.brute_force(p, x, q, t, d["nbat_nbfp"], d["nbat_nbfp_comb"], box, ei, ea)
    assert relrms(f, fb) < 1e-6
    assert abs(e[0] - eb[0]) <= 1e-6 * abs(eb[0]) + 1e-6 and abs(e[1] - eb[1]) <= 1e-6 * abs(eb[1]) + 1e-6


def test_gridding_passes_refuse_a_column_taller_than_one_block(emu):
    """8192 atoms per column is what one block sorts; the passes report it instead of mis-sorting"""
    rng = np.random.default_rng(5)
    box = np.array([1.0, 1.0, 400.0], np.float32)
    x = (rng.random((9000, 3)) * box).astype(np.float32)
    nbins = C.c_int()
    assert emu.search_emu_put_atoms_on_grid(_p(box, C.c_float), 1, 1, x.shape[0], _p(x, C.c_float), None, None, C.byref(nbins)) == 1


def test_gridding_passes_with_ties_everywhere(emu):
    """coordinates on a coarse lattice (many equal z, y, x values, atoms on cell boundaries and on the upper box face):
    the column sort is a total order (ties broken by atom index), so the device order still equals the host order"""
    from gromacs_b200.pairsearch import Grid
    rng = np.random.default_rng(3)
    box = np.array([3.0, 3.0, 3.0], np.float32)
    x = (rng.integers(0, 13, (4000, 3)) * 0.25).astype(np.float32)          # lattice 0, 0.25, ... 3.0 (= the box face)
    x[:50] = x[50:100]                                                     # exact duplicates
    grid = Grid(box, x, nthreads=3)
    nbins, atom_index, first_bin = emu_grid(emu, box, x, grid.ncx, grid.ncy)
    assert nbins == grid.nbins and np.array_equal(first_bin, grid.first_bin_of_column)
    assert np.array_equal(atom_index, grid.atom_index)


@pytest.mark.parametrize("kind", ["tall uniform", "all z equal", "clustered z", "two atoms", "lattice z", "signed zeros"])
def test_column_sort_corner_cases(emu, kind):
    """the column sort where its bucket form is stressed: a column of ~7000 atoms (bucket count near its cap), every atom at
    the same z (one bucket holds the column), z crowded into 1 % of the range, a two-atom system, z on a coarse lattice
    and +-0 keys: the order is the host gridder's in every case"""
    from gromacs_b200.pairsearch import Grid
    rng = np.random.default_rng(7)
    box = np.array([1.2, 1.2, 60.0], np.float32)
    n = 7000
    x = rng.random((n, 3)) * box
    if kind == "all z equal":
        x[:, 2] = 30.0
    elif kind == "clustered z":
        x[:, 2] = np.where(rng.random(n) < 0.9, 5.0 + rng.random(n) * 0.01, x[:, 2])
    elif kind == "two atoms":
        box = np.array([3.0, 3.0, 3.0], np.float32)
        x = rng.random((2, 3)) * box
    elif kind == "lattice z":
        x[:, 2] = np.minimum(np.round(x[:, 2] * 4) / 4, 59.75)
    elif kind == "signed zeros":
        box = np.array([3.0, 3.0, 3.0], np.float32)
        x = rng.random((5000, 3)) * box
        x[::7, 2] = -0.0
        x[1::7, 2] = 0.0
    x = np.ascontiguousarray(x, np.float32)
    grid = Grid(box, x, nthreads=2)
    nbins, atom_index, first_bin = emu_grid(emu, box, x, grid.ncx, grid.ncy)
    assert nbins == grid.nbins and np.array_equal(first_bin, grid.first_bin_of_column)
    assert np.array_equal(atom_index, grid.atom_index)


def emu_fep_list(emu):
    ni, npairs = C.c_int(), C.c_int()
    emu.search_emu_fep_sizes(C.byref(ni), C.byref(npairs))
    iinr = np.zeros(ni.value, np.int32)
    shift = np.zeros(ni.value, np.int32)
    pair_entry = np.zeros(npairs.value, np.int32)
    jjnr = np.zeros(npairs.value, np.int32)
    inter = np.zeros(npairs.value, np.uint8)
    emu.search_emu_fep_copy(_p(iinr, C.c_int), _p(shift, C.c_int), _p(pair_entry, C.c_int), _p(jjnr, C.c_int), _p(inter, C.c_ubyte))
    return iinr, shift, pair_entry, jjnr, inter


def assert_same_fep_list(dev, host):
    """device: (iinr, shift, pair_entry, jjnr, interacts); host: dict of split_fep_pairlist"""
    iinr, shift, pair_entry, jjnr, inter = dev
    assert np.array_equal(iinr, host["iinr"]) and np.array_equal(shift, host["shift"])
    assert np.array_equal(jjnr, host["jjnr"]) and np.array_equal(inter != 0, host["excl_fep"] != 0)
    # pair_entry is jindex expanded
    want = np.repeat(np.arange(host["iinr"].shape[0], dtype=np.int32), np.diff(host["jindex"]))
    assert np.array_equal(pair_entry, want)


@pytest.mark.parametrize("case,rlist,min_sci,seed", [("bench1_ewald_cutnone", 1.0, 0, 4), ("bench1_ewald_cutnone", 1.05, 500, 5),
                                                     ("test243_ewald_cutnone", 0.9, 60, 6), ("bench1_rf_cutnone_split", 0.8, 200, 7)])
def test_perturbed_pair_split_equals_the_host_builders(emu, case, rlist, min_sci, seed):
    """pass 8 of the device builder (perturbed pairs leave the cluster list for an atom-pair list) against
    nbnxm_b200_pairlist_split_fep: the same perturbed list entry for entry (i-entries, j-atoms, interacting / excluded), the
    same cluster list afterwards (masks bit for bit, exclusion words after expansion)"""
    from gromacs_b200.pairsearch import split_fep_pairlist
    d = load_golden(case)
    grid, nbat = host_grid(d)
    n = d["sys_x"].shape[0]
    rng = np.random.default_rng(seed)
    perturbed = np.zeros(n, np.uint8)
    if n % 3 == 0 and d["sys_excl_index"][1] == 3:          # water: whole molecules
        for m in rng.choice(n // 3, size=max(3, n // 200), replace=False):
            perturbed[3 * m:3 * m + 3] = 1
    else:
        perturbed[rng.choice(n, size=max(5, n // 40), replace=False)] = 1
    grid.pairlist(rlist, d["sys_excl_index"], d["sys_excl_atoms"], min_sci=min_sci)
    ref, fep = split_fep_pairlist(grid, perturbed)
    assert fep["jjnr"].size > 0 and (fep["excl_fep"] == 0).any() and (fep["excl_fep"] != 0).any()
    assert emu.search_emu_set_perturbed(n, _p(perturbed, C.c_ubyte)) == 0
    try:
        got = emu_pairlist(emu, grid, nbat.xq, rlist, d["sys_excl_index"], d["sys_excl_atoms"], min_sci=min_sci)
        assert_same_list(got[:3], (ref.sci, ref.cjPacked, ref.excl))
        assert_same_fep_list(emu_fep_list(emu), fep)
    finally:
        emu.search_emu_set_perturbed(0, None)
    # and without perturbed atoms the passes give the plain list again
    plain = grid.pairlist(rlist, d["sys_excl_index"], d["sys_excl_atoms"], min_sci=min_sci)
    got = emu_pairlist(emu, grid, nbat.xq, rlist, d["sys_excl_index"], d["sys_excl_atoms"], min_sci=min_sci)
    assert_same_list(got[:3], (plain.sci, plain.cjPacked, plain.excl))
    ni, npairs = C.c_int(), C.c_int()
    emu.search_emu_fep_sizes(C.byref(ni), C.byref(npairs))
    assert ni.value == 0 and npairs.value == 0
