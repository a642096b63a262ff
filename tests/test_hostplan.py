"""CPU tests of the C++ host-side planning (gromacs_b200/csrc/hostplan.cpp) against the independent numpy restatements
in tests/host_plan_reference.py: x-slab bin ranges, re-indexing of lists to rank order, chunk plans — integer work,
compared bit for bit."""
import numpy as np
import pytest

import host_plan_reference as R


@pytest.fixture(scope="module")
def wl():
    from gromacs_b200.workload import make_workload
    return make_workload("water48k_test", nthreads=4, nslabs=4)


@pytest.mark.parametrize("nslabs", [1, 2, 3, 4, 8])
def test_slab_bin_ranges(wl, nslabs):
    from gromacs_b200.slabs import slab_bin_ranges
    for r in range(nslabs):
        for rlist in (0.6, 0.95, 1.3):
            try:
                want = R.slab_bin_ranges(wl.grid, nslabs, r, rlist)
            except ValueError:
                with pytest.raises(ValueError):
                    slab_bin_ranges(wl.grid, nslabs, r, rlist)
                continue
            assert slab_bin_ranges(wl.grid, nslabs, r, rlist) == want


@pytest.mark.parametrize("nslabs", [2, 4])
def test_reindex_to_rank_order(wl, nslabs):
    from gromacs_b200.multigpu import _reindex
    from gromacs_b200.slabs import slab_bin_ranges
    rlist = 0.95
    for r in range(nslabs):
        home, halo, tx = slab_bin_ranges(wl.grid, nslabs, r, rlist)
        nhome, nhalo = home[1] - home[0], halo[1] - halo[0]
        ncl = (nhome + nhalo) * 8
        loc = wl.grid.pairlist(rlist, wl.box.excl_index, wl.box.excl_atoms, min_sci=300, bins=home, j_bins=home)
        nloc = wl.grid.pairlist(rlist, wl.box.excl_index, wl.box.excl_atoms, min_sci=150, bins=home, j_bins=halo,
                                inter_zone=True, required_tx=tx)
        for pl, is_halo in ((loc, False), (nloc, True)):
            got = _reindex(pl, home[0], halo[0], nhome, ncl, halo=is_halo)
            sci, cjp = R.reindex(pl, home[0], halo[0], nhome, ncl, halo=is_halo)
            assert np.array_equal(got.sci, sci) and np.array_equal(got.cjPacked, cjp)
            assert got.sci.shape[0] > 0 and got.cjPacked[:, :4].max() < ncl


@pytest.mark.parametrize("nchunks", [1, 2, 5, 24, 40])
def test_chunk_plan(wl, nchunks):
    from gromacs_b200.pipeline import make_chunk_plan
    plist = wl.pairlist(min_sci=1500)
    plan = make_chunk_plan(wl.grid, plist, nchunks)
    k, first_atom, first_sci, needs, sci_sorted = R.make_chunk_plan(wl.grid, plist, nchunks)
    assert plan.nchunks == k
    assert np.array_equal(plan.first_atom, first_atom) and np.array_equal(plan.first_sci, first_sci)
    assert np.array_equal(plan.needs, needs)
    assert np.array_equal(plan.plist.sci, sci_sorted)
    assert plan.plist.cjPacked is plist.cjPacked or np.array_equal(plan.plist.cjPacked, plist.cjPacked)
