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
