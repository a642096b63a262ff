"""Grid + GPU-layout pair-list construction (include/nbnxm_b200_search.h), the caller side of the force path:
mirrors nonbonded_verlet_t::putAtomsOnGrid / setAtomProperties / constructPairlist
(src/gromacs/nbnxm/nbnxm.cpp:78, atomdata.cpp:1107, pairlist.cpp:4056).  `Grid` is the host gridder / builder
(C++/OpenMP in the library), `GpuPairSearch` the same search step on the device."""
import ctypes as C
import os

import numpy as np

from .nbnxm import AtomData, NbnxmError, PairlistGpu, load_library

SEARCH_SYMBOLS = [
    "nbnxm_b200_pairlist_split_fep", "nbnxm_b200_pairlist_fep_sizes", "nbnxm_b200_pairlist_fep_copy",
    "nbnxm_b200_grid_dims", "nbnxm_b200_grid_box", "nbnxm_b200_slab_bin_ranges", "nbnxm_b200_pairlist_reindex", "nbnxm_b200_chunk_plan",
    "nbnxm_b200_grid_create", "nbnxm_b200_grid_create_slabs", "nbnxm_b200_grid_free", "nbnxm_b200_grid_info", "nbnxm_b200_grid_get_order",
    "nbnxm_b200_grid_fill_atomdata", "nbnxm_b200_pairlist_build", "nbnxm_b200_pairlist_sizes",
    "nbnxm_b200_pairlist_copy",
    "nbnxm_b200_gpu_search_create", "nbnxm_b200_gpu_search_free", "nbnxm_b200_gpu_search_set_grid",
    "nbnxm_b200_gpu_search_build", "nbnxm_b200_gpu_search_sizes", "nbnxm_b200_gpu_search_download",
    "nbnxm_b200_gpu_search_set_atoms", "nbnxm_b200_gpu_search_put_atoms_on_grid", "nbnxm_b200_gpu_search_get_order",
    "nbnxm_b200_gpu_search_set_perturbed", "nbnxm_b200_gpu_search_fep_sizes", "nbnxm_b200_gpu_search_fep_download",
    "nbnxm_b200_gpu_search_gather_slab", "nbnxm_b200_gpu_search_build_slab",
]


def _p(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


class Grid:
    """One pair-search grid over a rectangular periodic box (Grid / GridSet of the reference)."""

    def __init__(self, box, x, nthreads=None, nslabs=1):
        self._lib = load_library()
        self._g = C.c_void_p()
        self.nthreads = nthreads or min(os.cpu_count() or 1, 64)
        self.box = np.ascontiguousarray(box, np.float32)
        x = np.ascontiguousarray(x, np.float32)
        self.natoms = x.shape[0]
        if self._lib.nbnxm_b200_grid_create_slabs(C.byref(self._g), _p(self.box, C.c_float), C.c_int(self.natoms),
                                                  _p(x, C.c_float), C.c_int(self.nthreads), C.c_int(nslabs)):
            raise NbnxmError("nbnxm_b200_grid_create_slabs failed")
        n, nb, ncx, ncy = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._lib.nbnxm_b200_grid_info(self._g, C.byref(n), C.byref(nb), C.byref(ncx), C.byref(ncy))
        self.natoms_nbat, self.nbins, self.ncx, self.ncy = n.value, nb.value, ncx.value, ncy.value
        self.atom_index = np.zeros(self.natoms_nbat, np.int32)
        self.first_bin_of_column = np.zeros(self.ncx * self.ncy + 1, np.int32)
        self._lib.nbnxm_b200_grid_get_order(self._g, _p(self.atom_index, C.c_int), _p(self.first_bin_of_column, C.c_int))

    def __del__(self):
        try:
            if self._g:
                self._lib.nbnxm_b200_grid_free(self._g)
                self._g = C.c_void_p()
        except Exception:
            pass

    def atomdata(self, x, q, atom_type, nbfp, ntypes, nbfp_comb=None, lj_comb_per_type=None, shift_vec=None):
        """Build the nbnxm_atomdata_t contents in grid order (fillers: x = -1e6, q = 0, type = ntypes-1)."""
        x = np.ascontiguousarray(x, np.float32)
        q = np.ascontiguousarray(q, np.float32)
        t = np.ascontiguousarray(atom_type, np.int32)
        xq = np.zeros((self.natoms_nbat, 4), np.float32)
        tn = np.zeros(self.natoms_nbat, np.int32)
        ljt = None if lj_comb_per_type is None else np.ascontiguousarray(lj_comb_per_type, np.float32)
        lj = None if ljt is None else np.zeros((self.natoms_nbat, 2), np.float32)
        if self._lib.nbnxm_b200_grid_fill_atomdata(self._g, _p(x, C.c_float), _p(q, C.c_float), _p(t, C.c_int),
                                                   C.c_int(ntypes), _p(ljt, C.c_float), _p(xq, C.c_float),
                                                   _p(tn, C.c_int), _p(lj, C.c_float)):
            raise NbnxmError("nbnxm_b200_grid_fill_atomdata failed")
        return AtomData(xq=xq, type=tn, lj_comb=lj, nbfp=nbfp, nbfp_comb=nbfp_comb, numTypes=ntypes,
                        shift_vec=shift_vec if shift_vec is not None else shift_vectors(self.box))

    def pairlist(self, rlist, excl_index=None, excl_atoms=None, min_sci=0, bins=None, j_bins=None,
                 inter_zone=False, required_tx=0):
        """constructPairlist for the GPU layout; returns a PairlistGpu. bins / j_bins: (begin, end) bin ranges."""
        b0, b1 = bins if bins is not None else (0, self.nbins)
        j0, j1 = j_bins if j_bins is not None else (0, self.nbins)
        ei = None if excl_index is None else np.ascontiguousarray(excl_index, np.int32)
        ea = None if excl_atoms is None else np.ascontiguousarray(excl_atoms, np.int32)
        if self._lib.nbnxm_b200_pairlist_build(self._g, C.c_float(rlist), _p(ei, C.c_int), _p(ea, C.c_int),
                                               C.c_int(min_sci), C.c_int(b0), C.c_int(b1), C.c_int(j0), C.c_int(j1),
                                               C.c_int(int(inter_zone)), C.c_int(required_tx), C.c_int(self.nthreads)):
            raise NbnxmError("nbnxm_b200_pairlist_build failed (rlist must be < half the box)")
        nsci, ncj, nex, ncp = C.c_int(), C.c_int(), C.c_int(), C.c_longlong()
        self._lib.nbnxm_b200_pairlist_sizes(self._g, C.byref(nsci), C.byref(ncj), C.byref(nex), C.byref(ncp))
        sci = np.zeros((nsci.value, 4), np.int32)
        cjp = np.zeros((ncj.value, 8), np.uint32)
        excl = np.zeros((nex.value, 32), np.uint32)
        self._lib.nbnxm_b200_pairlist_copy(self._g, _p(sci, C.c_int), _p(cjp, C.c_uint32), _p(excl, C.c_uint32))
        pl = PairlistGpu(sci=sci, cjPacked=cjp, excl=excl, na_ci=8, rlist=rlist)
        pl.nci_tot = ncp.value
        return pl


def split_fep_pairlist(grid: Grid, perturbed):
    """make_fep_list for the list `grid.pairlist()` built last: returns (cluster list with the perturbed pairs' bits
    cleared, dict(iinr, jindex, jjnr, shift, excl_fep) in nbat indices for gpu_init_feppairlist)."""
    lib = load_library()
    pert = np.ascontiguousarray(perturbed, np.uint8)
    if lib.nbnxm_b200_pairlist_split_fep(grid._g, _p(pert, C.c_ubyte)):
        raise NbnxmError("nbnxm_b200_pairlist_split_fep failed")
    ni, nj = C.c_int(), C.c_int()
    lib.nbnxm_b200_pairlist_fep_sizes(grid._g, C.byref(ni), C.byref(nj))
    iinr, jindex = np.zeros(ni.value, np.int32), np.zeros(ni.value + 1, np.int32)
    jjnr, shift, inter = np.zeros(nj.value, np.int32), np.zeros(ni.value, np.int32), np.zeros(nj.value, np.uint8)
    lib.nbnxm_b200_pairlist_fep_copy(grid._g, _p(iinr, C.c_int), _p(jindex, C.c_int), _p(jjnr, C.c_int), _p(shift, C.c_int),
                                     _p(inter, C.c_ubyte))
    nsci, ncj, nex, ncp = C.c_int(), C.c_int(), C.c_int(), C.c_longlong()
    lib.nbnxm_b200_pairlist_sizes(grid._g, C.byref(nsci), C.byref(ncj), C.byref(nex), C.byref(ncp))
    sci = np.zeros((nsci.value, 4), np.int32)
    cjp = np.zeros((ncj.value, 8), np.uint32)
    excl = np.zeros((nex.value, 32), np.uint32)
    lib.nbnxm_b200_pairlist_copy(grid._g, _p(sci, C.c_int), _p(cjp, C.c_uint32), _p(excl, C.c_uint32))
    return (PairlistGpu(sci=sci, cjPacked=cjp, excl=excl, na_ci=8),
            dict(iinr=iinr, jindex=jindex, jjnr=jjnr, shift=shift, excl_fep=inter))


class GpuPairSearch:
    """constructPairlist + gpu_init_pairlist on the device (nbnxm_b200_gpu_search_*): the list is built from the
    coordinates resident in the NbnxmGpu handle and becomes its list for `iloc`; nothing is staged on the host."""

    def __init__(self, nb, grid: Grid = None, excl_index=None, excl_atoms=None):
        """With a host Grid: its columns / atom order / exclusions are uploaded (nbnxm_b200_gpu_search_set_grid).
        Without: call set_atoms() once and put_atoms_on_grid() at every search step (gridding on the device)."""
        self._lib = load_library()
        self._nb = nb
        self._s = C.c_void_p()
        self.grid = grid
        self.nbins = 0
        self.build_ms = 0.0
        self.nci_tot = 0
        nb._check(self._lib.nbnxm_b200_gpu_search_create(C.byref(self._s), nb._h))
        if grid is not None:
            ei = None if excl_index is None else np.ascontiguousarray(excl_index, np.int32)
            ea = None if excl_atoms is None else np.ascontiguousarray(excl_atoms, np.int32)
            nb._check(self._lib.nbnxm_b200_gpu_search_set_grid(
                self._s, _p(grid.box, C.c_float), C.c_int(grid.ncx), C.c_int(grid.ncy), _p(grid.first_bin_of_column, C.c_int),
                _p(grid.atom_index, C.c_int), C.c_int(grid.nbins), C.c_int(grid.natoms), _p(ei, C.c_int), _p(ea, C.c_int)))
            self.nbins = grid.nbins

    def set_atoms(self, q, atom_type, ntypes, lj_comb_per_type=None, excl_index=None, excl_atoms=None):
        """Static per-atom properties and topology exclusions, atom order (nbnxm_b200_gpu_search_set_atoms)."""
        q = None if q is None else np.ascontiguousarray(q, np.float32)
        t = None if atom_type is None else np.ascontiguousarray(atom_type, np.int32)
        lj = None if lj_comb_per_type is None else np.ascontiguousarray(lj_comb_per_type, np.float32)
        ei = None if excl_index is None else np.ascontiguousarray(excl_index, np.int32)
        ea = None if excl_atoms is None else np.ascontiguousarray(excl_atoms, np.int32)
        self.natoms = (q if q is not None else t).shape[0]
        self._nb._check(self._lib.nbnxm_b200_gpu_search_set_atoms(
            self._s, C.c_int(self.natoms), _p(q, C.c_float), _p(t, C.c_int), C.c_int(ntypes), _p(lj, C.c_float),
            _p(ei, C.c_int), _p(ea, C.c_int)))

    def set_perturbed(self, perturbed):
        """make_fep_list on the device (nbnxm_b200_gpu_search_set_perturbed): perturbed[natoms] flags in atom order, or None to
        switch the split off.  The builds that follow move the pairs of these atoms to the handle's perturbed list."""
        if perturbed is None:
            self._nb._check(self._lib.nbnxm_b200_gpu_search_set_perturbed(self._s, C.c_int(0), None))
            return
        pert = np.ascontiguousarray(perturbed, np.uint8)
        self._nb._check(self._lib.nbnxm_b200_gpu_search_set_perturbed(self._s, C.c_int(pert.shape[0]), _p(pert, C.c_ubyte)))

    def fep_download(self):
        """the perturbed list built last, in the form of split_fep_pairlist: dict(iinr, jindex, jjnr, shift, excl_fep)"""
        ni, nj = C.c_int(), C.c_int()
        self._nb._check(self._lib.nbnxm_b200_gpu_search_fep_sizes(self._s, C.byref(ni), C.byref(nj)))
        iinr, shift = np.zeros(ni.value, np.int32), np.zeros(ni.value, np.int32)
        jindex = np.zeros(ni.value + 1, np.int32)
        jjnr, inter = np.zeros(nj.value, np.int32), np.zeros(nj.value, np.uint8)
        self._nb._check(self._lib.nbnxm_b200_gpu_search_fep_download(self._s, _p(iinr, C.c_int), _p(jindex, C.c_int), _p(jjnr, C.c_int),
                                                                     _p(shift, C.c_int), _p(inter, C.c_ubyte)))
        return dict(iinr=iinr, jindex=jindex, jjnr=jjnr, shift=shift, excl_fep=inter)

    def put_atoms_on_grid(self, box, d_x_ptr, nslabs=1, x_ready_event=None):
        """putAtomsOnGrid + setAtomProperties + gpu_init_atomdata on the device from natoms rvecs in device memory;
        returns (natoms_nbat, nbins, ncx, ncy)."""
        box = np.ascontiguousarray(box, np.float32)
        n, nb_, ncx, ncy = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._nb._check(self._lib.nbnxm_b200_gpu_search_put_atoms_on_grid(
            self._s, _p(box, C.c_float), C.c_int(nslabs), C.c_void_p(d_x_ptr), C.c_void_p(x_ready_event), C.byref(n),
            C.byref(nb_), C.byref(ncx), C.byref(ncy)))
        self.nbins, self.ncx, self.ncy = nb_.value, ncx.value, ncy.value
        self._nb._natoms = n.value
        return n.value, nb_.value, ncx.value, ncy.value

    def get_order(self):
        """(atom_index[natoms_nbat], first_bin_of_column[ncx*ncy+1], grid_ms) of the grid in use."""
        ncol = (self.grid.ncx * self.grid.ncy) if self.grid is not None else self.ncx * self.ncy
        ai = np.zeros(self.nbins * 64, np.int32)
        fb = np.zeros(ncol + 1, np.int32)
        ms = C.c_float()
        self._nb._check(self._lib.nbnxm_b200_gpu_search_get_order(self._s, _p(ai, C.c_int), _p(fb, C.c_int), C.byref(ms)))
        return ai, fb, ms.value

    def free(self):
        if self._s:
            self._lib.nbnxm_b200_gpu_search_free(self._s)
            self._s = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def build(self, rlist, iloc=0, min_sci=0, bins=None, j_bins=None, inter_zone=False, required_tx=0):
        """Builds and installs the list; returns (nsci, ncj_packed, nexcl)."""
        b0, b1 = bins if bins is not None else (0, self.nbins)
        j0, j1 = j_bins if j_bins is not None else (0, self.nbins)
        self._nb._check(self._lib.nbnxm_b200_gpu_search_build(
            self._s, C.c_int(iloc), C.c_float(rlist), C.c_int(min_sci), C.c_int(b0), C.c_int(b1), C.c_int(j0), C.c_int(j1),
            C.c_int(int(inter_zone)), C.c_int(required_tx)))
        nsci, ncj, nex, ncp, ms = C.c_int(), C.c_int(), C.c_int(), C.c_longlong(), C.c_float()
        self._nb._check(self._lib.nbnxm_b200_gpu_search_sizes(self._s, C.byref(nsci), C.byref(ncj), C.byref(nex),
                                                             C.byref(ncp), C.byref(ms)))
        self.build_ms, self.nci_tot, self.rlist = ms.value, ncp.value, rlist
        self._sizes = (nsci.value, ncj.value, nex.value)
        self._nb._set_list_sizes(iloc, nsci.value, ncj.value)
        return self._sizes

    def gather_slab(self, target, home, halo):
        """Atom data of one x-slab (home bins, then halo bins) from this search object's whole-system grid into the rank's
        handle `target`, device to device (nbnxm_b200_gpu_search_gather_slab)."""
        self._nb._check(self._lib.nbnxm_b200_gpu_search_gather_slab(
            self._s, target._h, C.c_int(home[0]), C.c_int(home[1]), C.c_int(halo[0]), C.c_int(halo[1])))
        target._natoms = (home[1] - home[0] + halo[1] - halo[0]) * 64

    def build_slab(self, target, iloc, rlist, home, halo, required_tx=0, min_sci=0):
        """Local (iloc 0: home x home) or non-local (iloc 1: home x halo) list of one x-slab, built on the device from the
        whole-system grid, re-indexed to the rank's order and installed in `target` (nbnxm_b200_gpu_search_build_slab);
        returns (nsci, ncj_packed, nexcl)."""
        self._nb._check(self._lib.nbnxm_b200_gpu_search_build_slab(
            self._s, target._h, C.c_int(iloc), C.c_float(rlist), C.c_int(min_sci), C.c_int(home[0]), C.c_int(home[1]),
            C.c_int(halo[0]), C.c_int(halo[1]), C.c_int(required_tx)))
        nsci, ncj, nex, ncp, ms = C.c_int(), C.c_int(), C.c_int(), C.c_longlong(), C.c_float()
        self._nb._check(self._lib.nbnxm_b200_gpu_search_sizes(self._s, C.byref(nsci), C.byref(ncj), C.byref(nex),
                                                             C.byref(ncp), C.byref(ms)))
        self.build_ms, self.nci_tot, self.rlist = ms.value, ncp.value, rlist
        self._sizes = (nsci.value, ncj.value, nex.value)
        target._set_list_sizes(iloc, nsci.value, ncj.value)
        return self._sizes

    def download_sci(self):
        """int32[nsci, 4]: the sci entries of the list built last (16 bytes each; the cjPacked groups stay on the device)"""
        sci = np.zeros((self._sizes[0], 4), np.int32)
        self._nb._check(self._lib.nbnxm_b200_gpu_search_download(self._s, _p(sci, C.c_int), None, None))
        return sci

    def download(self) -> PairlistGpu:
        """Host copy of the list built last."""
        nsci, ncj, nex = self._sizes
        sci = np.zeros((nsci, 4), np.int32)
        cjp = np.zeros((ncj, 8), np.uint32)
        excl = np.zeros((nex, 32), np.uint32)
        self._nb._check(self._lib.nbnxm_b200_gpu_search_download(self._s, _p(sci, C.c_int), _p(cjp, C.c_uint32),
                                                                _p(excl, C.c_uint32)))
        pl = PairlistGpu(sci=sci, cjPacked=cjp, excl=excl, na_ci=8, rlist=self.rlist)
        pl.nci_tot = self.nci_tot
        return pl


def shift_vectors(box):
    """calc_shifts for a rectangular box (src/gromacs/pbcutil/pbc.cpp): 45 vectors, index
    ((z+1)*3 + (y+1))*5 + (x+2), central = 22 (pbcutil/ishift.h:43-56)."""
    sv = np.zeros((45, 3), np.float32)
    for z in (-1, 0, 1):
        for y in (-1, 0, 1):
            for x in (-2, -1, 0, 1, 2):
                sv[((z + 1) * 3 + (y + 1)) * 5 + (x + 2)] = (x * box[0], y * box[1], z * box[2])
    return sv
