"""CPU tests of the host-side grid + pair-list builder (gromacs_b200/csrc/pairsearch.cpp): the lists it
produces, walked by the oracle, must give the brute-force forces (every pair within the cut-off exactly
once, exclusion and diagonal masks right), also when i-entries are split and when the system is cut
into x-slabs with one-sided halos."""
import numpy as np
import pytest

from util import load_golden, oracle_params, relrms


def build(d, rlist, **kw):
    from gromacs_b200.pairsearch import Grid
    grid = Grid(d["sys_box"], d["sys_x"], nthreads=2)
    nt = int(d["nbat_ntypes"][0])
    nbat = grid.atomdata(d["sys_x"], d["sys_q"], d["sys_type"], d["nbat_nbfp"], nt, nbfp_comb=d["nbat_nbfp_comb"])
    return grid, nbat


def walk(oracle, p, nbat, plists, natoms, atom_index):
    f = np.zeros((nbat.numAtoms(), 3))
    e = np.zeros(2)
    npairs = 0
    for pl in plists:
        fi, _, ei, n = oracle.forces(p, pl.sci, pl.cjPacked, pl.excl, nbat.xq, nbat.type,
                                     np.zeros((nbat.numAtoms(), 2), np.float32), nbat.nbfp, nbat.nbfp_comb, nbat.shift_vec)
        f += fi
        e += ei
        npairs += n
    return oracle.nbat_to_atom_order(f, atom_index, natoms), e, npairs


@pytest.mark.parametrize("case,rlist,min_sci", [("test243_ewald_cutnone", 0.9, 0), ("test243_ewald_cutnone", 0.93, 60),
                                                ("bench1_ewald_cutnone", 1.0, 0), ("bench1_ewald_cutnone", 1.05, 500),
                                                ("bench1_ewald_ljpmegeom", 0.95, 0)])
def test_own_pairlist_covers_all_pairs_once(oracle, case, rlist, min_sci):
    d = load_golden(case)
    p = oracle_params(oracle, d)
    grid, nbat = build(d, rlist)
    pl = grid.pairlist(rlist, d["sys_excl_index"], d["sys_excl_atoms"], min_sci=min_sci)
    n = d["sys_x"].shape[0]
    assert grid.natoms_nbat % 64 == 0 and sorted(grid.atom_index[grid.atom_index >= 0].tolist()) == list(range(n))
    # format invariants the kernels rely on
    assert np.array_equal(pl.cjPacked[:, 4], pl.cjPacked[:, 6])            # both halves start identical
    assert np.all(pl.excl[0] == 0xffffffff)
    assert pl.sci[:, 2].min() >= 0 and pl.sci[:, 3].max() <= pl.cjPacked.shape[0]
    if min_sci:
        assert pl.sci.shape[0] >= min_sci // 2
    f, e, npairs = walk(oracle, p, nbat, [pl], n, grid.atom_index)
    fb, eb = oracle.brute_force(p, d["sys_x"], d["sys_q"], d["sys_type"], d["nbat_nbfp"], d["nbat_nbfp_comb"],
                                d["sys_box"], d["sys_excl_index"], d["sys_excl_atoms"])
    assert relrms(f, fb) < 1e-6
    assert abs(e[0] - eb[0]) < 1e-6 * abs(eb[0]) + 1e-6 and abs(e[1] - eb[1]) < 1e-6 * abs(eb[1])
    # and the same forces as the reference's own list for these coordinates
    fr, _, er, npairs_ref = oracle.forces(p, d["pl_sci"], d["pl_cjPacked"], d["pl_excl"], d["nbat_xq"], d["nbat_type"],
                                          d["nbat_lj_comb"], d["nbat_nbfp"], d["nbat_nbfp_comb"], d["shift_vec"])
    fr = oracle.nbat_to_atom_order(fr, d["nbat_atom_index"], n)
    assert relrms(f, fr) < 1e-6
    if abs(rlist - float(d["rlist"][0])) < 1e-6:
        # list quality: not more than 15 % larger than the reference's list at the same rlist
        assert npairs < 1.15 * npairs_ref, (npairs, npairs_ref)


@pytest.mark.parametrize("nslabs", [2, 3])
def test_x_slab_lists_partition_the_pairs(oracle, nslabs):
    """local (home x home) + non-local (home x +x-neighbour halo) lists over all slabs = the whole system."""
    from gromacs_b200.slabs import slab_bin_ranges
    d = load_golden("bench1_ewald_cutnone")
    # a wider box so that 3 slabs are each wider than rlist: stack the unit box 4x along x
    x = np.concatenate([d["sys_x"] + np.array([i * d["sys_box"][0], 0, 0], np.float32) for i in range(4)])
    n0 = d["sys_x"].shape[0]
    box = d["sys_box"] * np.array([4, 1, 1], np.float32)
    q = np.tile(d["sys_q"], 4)
    t = np.tile(d["sys_type"], 4)
    ei = np.concatenate([d["sys_excl_index"][:-1] + i * d["sys_excl_atoms"].shape[0] for i in range(4)]
                        + [[4 * d["sys_excl_atoms"].shape[0]]]).astype(np.int32)
    ea = np.concatenate([d["sys_excl_atoms"] + i * n0 for i in range(4)]).astype(np.int32)
    dd = dict(d)
    dd.update(sys_x=x, sys_box=box, sys_q=q, sys_type=t)
    rlist = 1.0
    p = oracle_params(oracle, d)
    grid, nbat = build(dd, rlist)
    whole = grid.pairlist(rlist, ei, ea)
    f_ref, e_ref, n_ref = walk(oracle, p, nbat, [whole], x.shape[0], grid.atom_index)
    lists = []
    for r in range(nslabs):
        home, halo, tx = slab_bin_ranges(grid, nslabs, r, rlist)
        lists.append(grid.pairlist(rlist, ei, ea, bins=home, j_bins=home))
        lists.append(grid.pairlist(rlist, ei, ea, bins=home, j_bins=halo, inter_zone=True, required_tx=tx))
        # the halo is a contiguous range of whole columns of the next slab
        assert halo[0] < halo[1]
    f, e, n = walk(oracle, p, nbat, lists, x.shape[0], grid.atom_index)
    # (x + shift) is rounded to float per image, so a pair seen from the other side differs by ~1e-7
    assert relrms(f, f_ref) < 1e-6
    assert abs(e[1] - e_ref[1]) < 1e-7 * abs(e_ref[1])


@pytest.mark.parametrize("nslabs", [2, 4, 8])
def test_slab_grid_gives_equal_slabs(nslabs):
    """nbnxm_b200_grid_create_slabs: a whole number of x columns per slab, so that every slab of the (uniform) water box
    holds the same number of bins, and the halo stays inside the next slab."""
    from gromacs_b200.slabs import slab_bin_ranges
    from gromacs_b200.workload import make_workload
    wl = make_workload("water48k_test", nthreads=4, nslabs=nslabs)
    assert wl.grid.ncx % nslabs == 0
    sizes = []
    for r in range(nslabs):
        home, halo, tx = slab_bin_ranges(wl.grid, nslabs, r, 0.95)
        sizes.append(home[1] - home[0])
        assert 0 < halo[1] - halo[0] <= 1.1 * (home[1] - home[0])
        assert tx == (-1 if r == nslabs - 1 else 0)
    assert max(sizes) - min(sizes) <= 0.05 * max(sizes), sizes
    assert sum(sizes) == wl.grid.nbins


@pytest.mark.parametrize("nchunks", [2, 5])
def test_chunk_plan_covers_every_atom_an_entry_touches(nchunks):
    """gromacs_b200/pipeline.py: the sci array grouped by chunk is a permutation of the list's entries, chunk ranges
    tile atoms and entries, and chunk_needs[k] contains the chunk of every i- and j-atom the entries of sci chunk k
    touch (checked entry by entry)."""
    from gromacs_b200.pipeline import make_chunk_plan
    from gromacs_b200.workload import make_workload
    wl = make_workload("water48k_test", nthreads=4)
    plist = wl.pairlist(min_sci=1500)
    plan = make_chunk_plan(wl.grid, plist, nchunks)
    k = plan.nchunks
    assert plan.first_atom[0] == 0 and plan.first_atom[-1] == wl.nbat.numAtoms()
    assert plan.first_sci[0] == 0 and plan.first_sci[-1] == plist.sci.shape[0]
    assert np.all(np.diff(plan.first_atom) > 0) and np.all(np.diff(plan.first_sci) >= 0)
    a = np.ascontiguousarray(plist.sci).reshape(-1, 4)
    b = np.ascontiguousarray(plan.plist.sci).reshape(-1, 4)
    assert sorted(map(tuple, a.tolist())) == sorted(map(tuple, b.tolist()))
    cjp = np.ascontiguousarray(plist.cjPacked).view(np.uint32).reshape(-1, 8)
    chunk_of_atom = lambda atom: int(np.searchsorted(plan.first_atom, atom, side="right") - 1)
    for c in range(k):
        for e in b[plan.first_sci[c]:plan.first_sci[c + 1]:7]:       # every 7th entry keeps the test short
            assert chunk_of_atom(e[0] * 64) == c
            for g in cjp[e[2]:e[3]]:
                for jm in range(4):
                    if ((int(g[4]) | int(g[6])) >> (8 * jm)) & 0xff:
                        assert plan.needs[c] & (1 << chunk_of_atom(int(g[jm]) * 8)), (c, jm, g)


def test_search_abi_rejects_bad_arguments():
    """the host-side search entry points fail with a status (never crash) on caller mistakes"""
    import ctypes as C
    from gromacs_b200 import load_library
    from gromacs_b200.pairsearch import Grid
    from gromacs_b200.slabs import slab_bin_ranges
    from gromacs_b200.nbnxm import NbnxmError
    lib = load_library()
    d = load_golden("test243_ewald_cutnone")
    g = C.c_void_p()
    box = np.ascontiguousarray(d["sys_box"], np.float32)
    x = np.ascontiguousarray(d["sys_x"], np.float32)
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    assert lib.nbnxm_b200_grid_create(C.byref(g), fp(box), C.c_int(0), fp(x), C.c_int(1)) != 0          # no atoms
    assert lib.nbnxm_b200_grid_create(C.byref(g), None, C.c_int(243), fp(x), C.c_int(1)) != 0            # no box
    ncx, ncy = C.c_int(), C.c_int()
    assert lib.nbnxm_b200_grid_dims(fp(box), C.c_int(243), C.c_int(0), C.byref(ncx), C.byref(ncy)) != 0  # no slabs
    grid = Grid(d["sys_box"], d["sys_x"], nthreads=1)
    with pytest.raises(NbnxmError):
        grid.pairlist(0.6 * float(d["sys_box"].min()))                                                    # rlist > half the box
    with pytest.raises(NbnxmError):
        grid.pairlist(0.9, bins=(0, grid.nbins + 1))                                                      # bin range
    with pytest.raises(ValueError):
        slab_bin_ranges(grid, 2, 0, 0.9)                                                                   # slabs thinner than the halo
    with pytest.raises(ValueError):
        slab_bin_ranges(grid, 2, 2, 0.9)                                                                   # rank outside the slabs
