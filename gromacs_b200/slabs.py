"""x-slab decomposition of one pair-search grid over G GPUs (one-sided halo, one process per GPU).

Grid columns are x-major (src/gromacs/nbnxm/grid.h:100), so the columns of slab r form a contiguous range
of bins and therefore of nbat slots: the halo rank r imports from rank (r+1) % G is a contiguous slice of
the coordinate array, and the forces it returns a contiguous slice of the force array.  The ownership
rule (the lower-x slab computes every home x halo pair, all y/z shifts allowed, like the reference's
eighth-shell zones, src/gromacs/domdec/domdec_zones.cpp:55-83) is implemented by the pair-list builder's
inter-zone mode (include/nbnxm_b200_search.h)."""
import math


def slab_columns(ncx, nslabs, r):
    return (ncx * r) // nslabs, (ncx * (r + 1)) // nslabs


def slab_bin_ranges(grid, nslabs, r, rlist):
    """Returns (home_bins, halo_bins, required_tx) for slab r of nslabs.
    home_bins / halo_bins: (begin, end) bin ranges; halo = the first columns of slab (r+1) % nslabs within
    rlist (+ one column of slack for atoms binned by their cluster's lower corner) of the slab boundary;
    required_tx: x shift index of i-atoms for home x halo pairs (-1 across the periodic boundary)."""
    if nslabs < 2:
        return (0, grid.nbins), (0, 0), 0
    cx0, cx1 = slab_columns(grid.ncx, nslabs, r)
    nx0, nx1 = slab_columns(grid.ncx, nslabs, (r + 1) % nslabs)
    cell = float(grid.box[0]) / grid.ncx
    ncol_halo = min(nx1 - nx0, int(math.ceil(rlist / cell)) + 1)
    if nslabs == 2 and (cx1 - cx0) < 2 * ncol_halo:
        raise ValueError("slabs too thin for a one-sided halo: %d columns, halo %d" % (cx1 - cx0, ncol_halo))
    fb = grid.first_bin_of_column
    home = (int(fb[cx0 * grid.ncy]), int(fb[cx1 * grid.ncy]))
    halo = (int(fb[nx0 * grid.ncy]), int(fb[(nx0 + ncol_halo) * grid.ncy]))
    tx = -1 if r == nslabs - 1 else 0
    return home, halo, tx
