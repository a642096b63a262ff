"""TEST INFRASTRUCTURE: numpy restatements of the host-side planning functions (gromacs_b200/csrc/hostplan.cpp), kept
as the independent implementation tests/test_hostplan.py compares the C++ against."""
import math

import numpy as np


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


def reindex(pl, first_home_bin, first_halo_bin, num_home_bins, nclusters_total, halo):
    """Global bin / cluster indices -> rank order (home bins first, then halo bins)."""
    sci = pl.sci.copy()
    cjp = pl.cjPacked.copy()
    sci[:, 0] -= first_home_bin
    cj = cjp[:, :4].astype(np.int64)
    if halo:
        cj = cj - first_halo_bin * 8 + num_home_bins * 8
    else:
        cj = cj - first_home_bin * 8
    # unused slots of partially filled j-groups carry no mask bits and an unspecified index: keep them loadable
    cjp[:, :4] = np.clip(cj, 0, nclusters_total - 1).astype(np.uint32)
    return sci, cjp



def make_chunk_plan(grid, plist, nchunks):
    nchunks = int(max(1, min(nchunks, 32, grid.ncx)))
    fb = np.asarray(grid.first_bin_of_column)
    # tapered widths: weights 1, 1, 2, 3, 4, ..., 4, 3, 2, 1, 1 with eight chunks or more (hostplan.cpp)
    wsum = [0]
    for c in range(nchunks):
        d = min(c, nchunks - 1 - c)
        wsum.append(wsum[-1] + (4 if nchunks < 8 else (1 if d < 2 else (2 if d == 2 else (3 if d == 3 else 4)))))
    cols = [(grid.ncx * wsum[c]) // wsum[nchunks] for c in range(nchunks + 1)]
    first_bin = np.array([int(fb[cx * grid.ncy]) for cx in cols], dtype=np.int64)
    first_atom = (first_bin * 64).astype(np.int32)
    sci = np.ascontiguousarray(plist.sci).reshape(-1, 4)
    chunk_of_sci = np.searchsorted(first_bin, sci[:, 0], side="right") - 1
    order = np.argsort(chunk_of_sci, kind="stable")
    sci_sorted = np.ascontiguousarray(sci[order])
    chunk_sorted = chunk_of_sci[order]
    first_sci = np.searchsorted(chunk_sorted, np.arange(nchunks + 1), side="left").astype(np.int32)
    # atom chunks touched by the j-clusters of every cjPacked group (outer list masks: a superset of what is evaluated)
    cjp = np.ascontiguousarray(plist.cjPacked).view(np.uint32).reshape(-1, 8)
    counts = (sci[:, 3] - sci[:, 2]).astype(np.int64)
    owner_chunk = np.zeros(cjp.shape[0], np.int64)
    starts = sci[:, 2].astype(np.int64)
    # groups of an entry are contiguous [begin, end); entries do not overlap
    idx = np.repeat(np.arange(sci.shape[0]), counts)
    group_index = np.repeat(starts, counts) + (np.arange(counts.sum()) - np.repeat(np.cumsum(counts) - counts, counts))
    owner_chunk[group_index] = chunk_of_sci[idx]
    any_mask = cjp[:, 4] | cjp[:, 6]
    needs = np.zeros(nchunks, np.uint32)
    for k in range(nchunks):
        needs[k] |= np.uint32(1 << k)
    for jm in range(4):
        valid = ((any_mask >> np.uint32(8 * jm)) & np.uint32(0xff)) != 0
        cj_bin = (cjp[valid, jm].astype(np.int64) * 8) // 64
        cj_chunk = np.searchsorted(first_bin, cj_bin, side="right") - 1
        pairs = np.unique(owner_chunk[valid] * nchunks + cj_chunk)
        for pr in pairs:
            needs[int(pr) // nchunks] |= np.uint32(1 << (int(pr) % nchunks))
    return nchunks, first_atom, first_sci, needs, sci_sorted
