"""Chunk plan for nbnxm_b200_do_force_step_pipelined: atoms cut into contiguous ranges of whole grid columns along x,
the sci array grouped by the chunk of its i-atoms, and for every sci chunk the set of atom chunks its entries touch."""
from dataclasses import dataclass

import numpy as np

from .nbnxm import PairlistGpu


@dataclass
class ChunkPlan:
    nchunks: int
    first_atom: np.ndarray      # int32[nchunks + 1]
    first_sci: np.ndarray       # int32[nchunks + 1]
    needs: np.ndarray           # uint32[nchunks]: bit c set = entries of sci chunk k read / write atoms of chunk c
    plist: PairlistGpu          # the list with sci grouped by chunk (cjPacked / excl unchanged)


def make_chunk_plan(grid, plist: PairlistGpu, nchunks: int) -> ChunkPlan:
    nchunks = int(max(1, min(nchunks, 32, grid.ncx)))
    fb = np.asarray(grid.first_bin_of_column)
    cols = [(grid.ncx * c) // nchunks for c in range(nchunks + 1)]
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
    out = PairlistGpu(sci=sci_sorted, cjPacked=plist.cjPacked, excl=plist.excl, na_ci=plist.na_ci, rlist=plist.rlist)
    return ChunkPlan(nchunks, first_atom, first_sci, needs, out)
