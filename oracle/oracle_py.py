"""TEST INFRASTRUCTURE: ctypes binding of oracle/liboracle.so (the CPU oracle).

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package gromacs_b200 never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

ELEC = {"cut": 0, "rf": 1, "ewaldtab": 2, "ewaldtabtwin": 3, "ewald": 4, "ewaldtwin": 5}
VDW = {"cutnone": 0, "cutgeom": 1, "cutlb": 2, "fswitch": 3, "pswitch": 4, "ljpmegeom": 5, "ljpmelb": 6}


class OrcParams(C.Structure):
    _fields_ = [("elec_type", C.c_int), ("vdw_type", C.c_int)] + [
        (n, C.c_float) for n in (
            "epsfac", "c_rf", "two_k_rf", "ewald_beta", "sh_ewald", "sh_lj_ewald", "ewaldcoeff_lj",
            "rcoulomb_sq", "rvdw_sq", "rvdw_switch", "rlist_outer_sq", "rlist_inner_sq",
            "disp_c2", "disp_c3", "disp_cpot", "rep_c2", "rep_c3", "rep_cpot",
            "sw_c3", "sw_c4", "sw_c5", "coulomb_tab_scale")] + [("ntypes", C.c_int)]


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _lib = C.CDLL(path)
        _lib.orc_forces.restype = C.c_int64
        _lib.orc_forces_f32_omp.restype = C.c_int64
    return _lib


def _p(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


def params_from_golden(d, rlist_inner=None):
    """Build OrcParams from a tests/golden npz, as set_cutoff_parameters does
    (src/gromacs/nbnxm/nbnxm_gpu_data_mgmt.cpp:218-240)."""
    g = lambda k: float(d[k][0])
    p = OrcParams()
    p.elec_type = ELEC[str(d["case_coulomb"])]
    p.vdw_type = VDW[str(d["case_vdw"])]
    p.epsfac = g("ic_epsfac")
    p.c_rf = g("ic_c_rf")
    p.two_k_rf = 2.0 * g("ic_k_rf")
    p.ewald_beta = g("ic_ewald_beta")
    p.sh_ewald = g("ic_sh_ewald")
    p.sh_lj_ewald = g("ic_sh_lj_ewald")
    p.ewaldcoeff_lj = g("ic_ewaldcoeff_lj")
    p.rcoulomb_sq = g("ic_rcoulomb") ** 2
    p.rvdw_sq = g("ic_rvdw") ** 2
    p.rvdw_switch = g("ic_rvdw_switch")
    p.rlist_outer_sq = g("rlist") ** 2
    p.rlist_inner_sq = (rlist_inner if rlist_inner is not None else g("rlist")) ** 2
    for k in ("disp_c2", "disp_c3", "disp_cpot", "rep_c2", "rep_c3", "rep_cpot", "sw_c3", "sw_c4", "sw_c5"):
        setattr(p, k, g("ic_" + k))
    p.coulomb_tab_scale = g("ic_coulomb_tab_scale") if "ic_coulomb_tab_scale" in d else 0.0
    p.ntypes = int(d["nbat_ntypes"][0])
    return p


def forces(p, sci, cjp, excl, xq, atype, lj_comb, nbfp, nbfp_comb, shift_vec, tab=None,
           calc_energy=True, calc_fshift=True):
    """Double-precision forces on the GPU list layout. Returns (f[n,3], fshift[45,3], e[2], npairs)."""
    n = xq.shape[0]
    f = np.zeros((n, 3), np.float64)
    fsh = np.zeros((45, 3), np.float64)
    e = np.zeros(2, np.float64)
    sci = np.ascontiguousarray(sci, np.int32)
    cjp = np.ascontiguousarray(cjp, np.uint32)
    excl = np.ascontiguousarray(excl, np.uint32)
    xq = np.ascontiguousarray(xq, np.float32)
    atype = np.ascontiguousarray(atype, np.int32)
    lj_comb = np.ascontiguousarray(lj_comb, np.float32)
    if lj_comb.size < 2 * n:
        lj_comb = np.zeros((n, 2), np.float32)
    nbfp = np.ascontiguousarray(nbfp, np.float32)
    nbfp_comb = np.ascontiguousarray(nbfp_comb, np.float32)
    if nbfp_comb.size < 2 * p.ntypes:
        nbfp_comb = np.zeros((p.ntypes, 2), np.float32)
    shift_vec = np.ascontiguousarray(shift_vec, np.float32)
    tabp = None if tab is None else np.ascontiguousarray(tab, np.float32)
    np_ = lib().orc_forces(C.byref(p), C.c_int(sci.shape[0]), _p(sci, C.c_int), _p(cjp, C.c_uint32),
                           _p(excl, C.c_uint32), _p(xq, C.c_float), _p(atype, C.c_int),
                           _p(lj_comb, C.c_float), _p(nbfp, C.c_float), _p(nbfp_comb, C.c_float),
                           _p(tabp, C.c_float), _p(shift_vec, C.c_float), C.c_int(int(calc_energy)),
                           C.c_int(int(calc_fshift)), _p(f, C.c_double), _p(fsh, C.c_double),
                           _p(e, C.c_double))
    return f, fsh, e, int(np_)


def forces_f32_omp(p, sci, cjp, excl, xq, atype, lj_comb, nbfp, nbfp_comb, shift_vec,
                   calc_energy=False, nthreads=1):
    n = xq.shape[0]
    f = np.zeros((n, 3), np.float32)
    e = np.zeros(2, np.float64)
    lj_comb = np.ascontiguousarray(lj_comb, np.float32)
    if lj_comb.size < 2 * n:
        lj_comb = np.zeros((n, 2), np.float32)
    nbfp_comb = np.ascontiguousarray(nbfp_comb, np.float32)
    if nbfp_comb.size < 2 * p.ntypes:
        nbfp_comb = np.zeros((p.ntypes, 2), np.float32)
    np_ = lib().orc_forces_f32_omp(
        C.byref(p), C.c_int(sci.shape[0]), _p(np.ascontiguousarray(sci, np.int32), C.c_int),
        _p(np.ascontiguousarray(cjp, np.uint32), C.c_uint32),
        _p(np.ascontiguousarray(excl, np.uint32), C.c_uint32),
        _p(np.ascontiguousarray(xq, np.float32), C.c_float),
        _p(np.ascontiguousarray(atype, np.int32), C.c_int), _p(lj_comb, C.c_float),
        _p(np.ascontiguousarray(nbfp, np.float32), C.c_float), _p(nbfp_comb, C.c_float),
        _p(np.ascontiguousarray(shift_vec, np.float32), C.c_float), C.c_int(n),
        C.c_int(int(calc_energy)), _p(f, C.c_float), _p(e, C.c_double), C.c_int(nthreads))
    return f, e, int(np_)


def prune(p, sci_order, cjp, imask_outer, xq, shift_vec, fresh, part=0, nparts=1):
    """In-place prune of cjp (uint32 [ncjp,8]) and imask_outer (uint32 [2*ncjp]); returns sci_count."""
    assert cjp.dtype == np.uint32 and cjp.flags.c_contiguous
    assert imask_outer.dtype == np.uint32 and imask_outer.flags.c_contiguous
    sci_order = np.ascontiguousarray(sci_order, np.int32)
    cnt = np.zeros(sci_order.shape[0], np.int32)
    lib().orc_prune(C.byref(p), C.c_int(sci_order.shape[0]), _p(sci_order, C.c_int), _p(cjp, C.c_uint32),
                    _p(imask_outer, C.c_uint32), _p(np.ascontiguousarray(xq, np.float32), C.c_float),
                    _p(np.ascontiguousarray(shift_vec, np.float32), C.c_float), C.c_int(int(fresh)),
                    C.c_int(part), C.c_int(nparts), _p(cnt, C.c_int))
    return cnt


def brute_force(p, x, q, atype, nbfp, nbfp_comb, box, excl_index, excl_atoms, calc_energy=True):
    n = x.shape[0]
    f = np.zeros((n, 3), np.float64)
    e = np.zeros(2, np.float64)
    nbfp_comb = np.ascontiguousarray(nbfp_comb, np.float32)
    if nbfp_comb.size < 2 * p.ntypes:
        nbfp_comb = np.zeros((p.ntypes, 2), np.float32)
    lib().orc_brute_force(
        C.byref(p), C.c_int(n), _p(np.ascontiguousarray(x, np.float32), C.c_float),
        _p(np.ascontiguousarray(q, np.float32), C.c_float),
        _p(np.ascontiguousarray(atype, np.int32), C.c_int),
        _p(np.ascontiguousarray(nbfp, np.float32), C.c_float), _p(nbfp_comb, C.c_float),
        _p(np.ascontiguousarray(box, np.float32), C.c_float),
        _p(np.ascontiguousarray(excl_index, np.int32), C.c_int),
        _p(np.ascontiguousarray(excl_atoms, np.int32), C.c_int), C.c_int(int(calc_energy)),
        _p(f, C.c_double), _p(e, C.c_double))
    return f, e


def nbat_to_atom_order(f_nbat, atom_index, natoms):
    """Scatter nbat-ordered forces to atom order (fillers have index -1), as
    nbnxm_atomdata_t::reduceForces does (src/gromacs/nbnxm/atomdata.cpp:1523-1590)."""
    out = np.zeros((natoms, 3), f_nbat.dtype)
    m = atom_index >= 0
    out[atom_index[m]] = f_nbat[: atom_index.shape[0]][m]
    return out
