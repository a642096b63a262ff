"""Chunk plan for nbnxm_b200_do_force_step_pipelined: atoms cut into contiguous ranges of whole grid columns along x,
the sci array grouped by the chunk of its i-atoms, and for every sci chunk the set of atom chunks its entries touch."""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from .nbnxm import NbnxmError, PairlistGpu, load_library


@dataclass
class ChunkPlan:
    nchunks: int
    first_atom: np.ndarray      # int32[nchunks + 1]
    first_sci: np.ndarray       # int32[nchunks + 1]
    needs: np.ndarray           # uint32[nchunks]: bit c set = entries of sci chunk k read / write atoms of chunk c
    plist: PairlistGpu          # the list with sci grouped by chunk (cjPacked / excl unchanged)


def make_chunk_plan(grid, plist: PairlistGpu, nchunks: int) -> ChunkPlan:
    """nbnxm_b200_chunk_plan (gromacs_b200/csrc/hostplan.cpp)"""
    lib = load_library()
    sci = np.ascontiguousarray(plist.sci, np.int32).reshape(-1, 4).copy()
    cjp = np.ascontiguousarray(plist.cjPacked, np.uint32).reshape(-1, 8)
    first_atom = np.zeros(33, np.int32)
    first_sci = np.zeros(33, np.int32)
    needs = np.zeros(32, np.uint32)
    n = C.c_int()
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    if lib.nbnxm_b200_chunk_plan(grid._g, p(sci, C.c_int), C.c_int(sci.shape[0]), p(cjp, C.c_uint32), C.c_int(cjp.shape[0]),
                                 C.c_int(int(nchunks)), C.byref(n), p(first_atom, C.c_int), p(first_sci, C.c_int),
                                 p(needs, C.c_uint32)):
        raise NbnxmError("nbnxm_b200_chunk_plan failed")
    k = n.value
    out = PairlistGpu(sci=sci, cjPacked=plist.cjPacked, excl=plist.excl, na_ci=plist.na_ci, rlist=plist.rlist)
    return ChunkPlan(k, first_atom[:k + 1].copy(), first_sci[:k + 1].copy(), needs[:k].copy(), out)


def make_slab_chunk_plan(box_x, ncx, ncy, first_bin_of_column, nranks, rank, rlist, sci_local, nchunks) -> ChunkPlan:
    """Chunk plan of one x-slab for the pipelined step with a halo: the home atoms cut at whole grid columns (tapered widths as
    in nbnxm_b200_chunk_plan), the local list - whose entries are in grid order, i-bins in the rank's numbering - cut at the same
    places, and for every sci chunk the atom chunks within reach of it: a geometric superset of what its entries touch (the
    chunks whose columns lie within rlist + one column of the chunk's own), so that the plan needs nothing of the list but the
    i-bin of every entry.  sci_local: int32[nsci, 4] of the local list."""
    import math
    from .slabs import slab_columns
    cx0, cx1 = slab_columns(ncx, nranks, rank)
    ncol = cx1 - cx0
    n = int(max(1, min(nchunks, 32, ncol)))
    wsum = [0]
    for c in range(n):
        d = min(c, n - 1 - c)
        wsum.append(wsum[-1] + (4 if n < 8 else (1 if d < 2 else (2 if d == 2 else (3 if d == 3 else 4)))))
    cols = [cx0 + (ncol * wsum[c]) // wsum[n] for c in range(n + 1)]
    fb = np.asarray(first_bin_of_column)
    home0 = int(fb[cx0 * ncy])
    first_bin = np.array([int(fb[c * ncy]) - home0 for c in cols], np.int64)
    first_atom = (first_bin * 64).astype(np.int32)
    sci_bins = np.ascontiguousarray(sci_local, np.int32).reshape(-1, 4)[:, 0]
    if sci_bins.size and np.any(np.diff(sci_bins) < 0):
        raise NbnxmError("make_slab_chunk_plan: the local list is not in grid order")
    first_sci = np.searchsorted(sci_bins, first_bin, side="left").astype(np.int32)
    first_sci[-1] = sci_bins.size
    reach = int(math.ceil(float(rlist) / (float(box_x) / ncx))) + 1
    needs = np.zeros(n, np.uint32)
    for k in range(n):
        for c in range(n):
            if cols[c] < cols[k + 1] + reach and cols[c + 1] > cols[k] - reach:
                needs[k] |= np.uint32(1 << c)
    return ChunkPlan(n, first_atom, first_sci, needs, None)
