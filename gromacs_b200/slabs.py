"""x-slab decomposition of one pair-search grid over G GPUs (one-sided halo, one process per GPU).

Grid columns are x-major (src/gromacs/nbnxm/grid.h:100), so the columns of slab r form a contiguous range
of bins and therefore of nbat slots: the halo rank r imports from rank (r+1) % G is a contiguous slice of
the coordinate array, and the forces it returns a contiguous slice of the force array.  The ownership
rule (the lower-x slab computes every home x halo pair, all y/z shifts allowed, like the reference's
eighth-shell zones, src/gromacs/domdec/domdec_zones.cpp:55-83) is implemented by the pair-list builder's
inter-zone mode (include/nbnxm_b200_search.h)."""
import ctypes as C

from .nbnxm import NbnxmError, load_library


def slab_columns(ncx, nslabs, r):
    return (ncx * r) // nslabs, (ncx * (r + 1)) // nslabs


def slab_bin_ranges(grid, nslabs, r, rlist):
    """Returns (home_bins, halo_bins, required_tx) for slab r of nslabs (nbnxm_b200_slab_bin_ranges, hostplan.cpp).
    home_bins / halo_bins: (begin, end) bin ranges; halo = the first columns of slab (r+1) % nslabs within
    rlist (+ one column of slack for atoms binned by their cluster's lower corner) of the slab boundary;
    required_tx: x shift index of i-atoms for home x halo pairs (-1 across the periodic boundary)."""
    v = [C.c_int() for _ in range(5)]
    if load_library().nbnxm_b200_slab_bin_ranges(grid._g, C.c_int(nslabs), C.c_int(r), C.c_float(rlist), *[C.byref(x) for x in v]):
        raise ValueError("slabs too thin for a one-sided halo (or bad arguments): %d slabs, rlist %g" % (nslabs, rlist))
    h0, h1, l0, l1, tx = [x.value for x in v]
    return (h0, h1), (l0, l1), tx


def slab_bin_ranges_from_columns(box_x, ncx, ncy, first_bin_of_column, nslabs, r, rlist):
    """slab_bin_ranges from the column table alone (e.g. GpuPairSearch.get_order() after gridding on the device): the same
    arithmetic as nbnxm_b200_slab_bin_ranges (hostplan.cpp) without a host grid object."""
    import math
    fb = first_bin_of_column
    if nslabs < 2:
        return (0, int(fb[ncx * ncy])), (0, 0), 0
    cx0, cx1 = slab_columns(ncx, nslabs, r)
    nx0, nx1 = slab_columns(ncx, nslabs, (r + 1) % nslabs)
    cell = float(box_x) / ncx
    ncol_halo = min(nx1 - nx0, int(math.ceil(float(rlist) / cell)) + 1)
    if nslabs == 2 and (cx1 - cx0) < 2 * ncol_halo:
        raise ValueError("slabs too thin for a one-sided halo: %d slabs, rlist %g" % (nslabs, rlist))
    home = (int(fb[cx0 * ncy]), int(fb[cx1 * ncy]))
    halo = (int(fb[nx0 * ncy]), int(fb[(nx0 + ncol_halo) * ncy]))
    return home, halo, (-1 if r == nslabs - 1 else 0)
