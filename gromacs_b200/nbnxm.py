"""Host-side mirror of the reference's NBNXM GPU-backend interface, over the C ABI.

`NbnxmGpu` stands for the reference's `NbnxmGpu` object and its methods carry the names of the free
functions in src/gromacs/nbnxm/nbnxm_gpu.h:68-313 and src/gromacs/nbnxm/gpu_data_mgmt.h:66-181
(`gpu_init_pairlist`, `gpu_copy_xq_to_gpu`, `gpu_launch_kernel`, ...), with the same argument meaning
and the same error behaviour (a failed call raises, it never falls back to a CPU path).

The library (gromacs_b200/libnbnxm_b200.so) is plain C ABI (include/nbnxm_b200.h); nothing here
touches torch.  The GPU work is entirely inside the library.
"""
import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libnbnxm_b200.so")

LOCAL, NONLOCAL, ALL = 0, 1, 2

ELEC_TYPES = {"Cut": 0, "RF": 1, "EwaldTab": 2, "EwaldTabTwin": 3, "EwaldAna": 4, "EwaldAnaTwin": 5, "None": 6}
VDW_TYPES = {"Cut": 0, "CutCombGeom": 1, "CutCombLB": 2, "FSwitch": 3, "PSwitch": 4, "EwaldGeom": 5, "EwaldLB": 6}

# every symbol include/nbnxm_b200.h declares
EXPORTED_SYMBOLS = [
    "nbnxm_b200_last_error", "nbnxm_b200_init", "nbnxm_b200_free", "nbnxm_b200_update_params", "nbnxm_b200_do_force_step", "nbnxm_b200_do_force_step_pipelined", "nbnxm_b200_set_pipeline_timeline", "nbnxm_b200_get_pipeline_timeline", "nbnxm_b200_peer_blob_size", "nbnxm_b200_peer_export", "nbnxm_b200_peer_import",
    "nbnxm_b200_peer_close", "nbnxm_b200_peer_error",
    "nbnxm_b200_init_pairlist", "nbnxm_b200_init_pairlist_device", "nbnxm_b200_init_atomdata", "nbnxm_b200_init_atomdata_device", "nbnxm_b200_upload_shiftvec",
    "nbnxm_b200_copy_xq_to_gpu", "nbnxm_b200_init_x_to_nbat_x", "nbnxm_b200_x_to_nbat_x",
    "nbnxm_b200_launch_kernel", "nbnxm_b200_launch_kernel_pruneonly", "nbnxm_b200_launch_cpyback",
    "nbnxm_b200_try_finish_task", "nbnxm_b200_wait_finish_task", "nbnxm_b200_clear_outputs",
    "nbnxm_b200_insert_nonlocal_dependency", "nbnxm_b200_setup_short_range_work",
    "nbnxm_b200_have_short_range_work", "nbnxm_b200_min_ci_balanced",
    "nbnxm_b200_is_kernel_ewald_analytical", "nbnxm_b200_get_timings", "nbnxm_b200_reset_timings",
    "nbnxm_b200_copy_fepparams", "nbnxm_b200_init_fep_atomdata", "nbnxm_b200_init_feppairlist", "nbnxm_b200_init_feppairlist_device",
    "nbnxm_b200_launch_free_energy_kernel", "nbnxm_b200_get_fep_dvdl", "nbnxm_b200_launch_foreign_energy_kernel",
    "nbnxm_b200_get_fep_foreign",
    "nbnxm_b200_set_timing", "nbnxm_b200_init_reduce_f", "nbnxm_b200_reduce_f", "nbnxm_b200_get_device_buffers", "nbnxm_b200_get_shared_outputs", "nbnxm_b200_get_streams",
    "nbnxm_b200_download_pairlist", "nbnxm_b200_set_pair_counting", "nbnxm_b200_get_pair_count",
    "nbnxm_b200_launch_count", "nbnxm_b200_pack_xq", "nbnxm_b200_unpack_xq", "nbnxm_b200_pack_f",
    "nbnxm_b200_unpack_add_f", "nbnxm_b200_measure_fp32_peak",
    "nbnxm_b200_halo_get_unique_id", "nbnxm_b200_halo_init", "nbnxm_b200_halo_free", "nbnxm_b200_halo_set_ranges",
    "nbnxm_b200_halo_exchange_x", "nbnxm_b200_halo_exchange_f", "nbnxm_b200_halo_set_timing",
    "nbnxm_b200_halo_get_timings",
]


class NbnxmError(RuntimeError):
    """Raised where the reference would gmx_fatal / GMX_THROW."""


class Params(C.Structure):
    """nbnxm_b200_params_t: the scalar part of NBParamGpu (gpu_types_common.h:222-262)."""
    _fields_ = [("elec_type", C.c_int), ("vdw_type", C.c_int)] + [
        (n, C.c_float) for n in (
            "epsfac", "c_rf", "two_k_rf", "ewald_beta", "sh_ewald", "sh_lj_ewald", "ewaldcoeff_lj",
            "rcoulomb_sq", "rvdw_sq", "rvdw_switch", "rlist_outer_sq", "rlist_inner_sq",
            "disp_c2", "disp_c3", "disp_cpot", "rep_c2", "rep_c3", "rep_cpot",
            "sw_c3", "sw_c4", "sw_c5", "coulomb_tab_scale")] + [("use_dynamic_pruning", C.c_int)]


class StepFlags(C.Structure):
    """nbnxm_b200_step_flags_t"""
    _fields_ = [("compute_energy", C.c_int), ("compute_virial", C.c_int), ("have_halo", C.c_int),
                ("dynamic_pruning", C.c_int), ("rolling_prune_parts", C.c_int)]


class Timings(C.Structure):
    _fields_ = [("force_ms", (C.c_double * 2) * 2), ("force_count", (C.c_int * 2) * 2),
                ("prune_ms", C.c_double), ("rolling_prune_ms", C.c_double),
                ("prune_count", C.c_int), ("rolling_prune_count", C.c_int),
                ("xq_h2d_ms", C.c_double), ("f_d2h_ms", C.c_double), ("pairlist_h2d_ms", C.c_double),
                ("force_nonlocal_ms", C.c_double), ("force_nonlocal_count", C.c_int)]


_lib = None


def load_library():
    """Load libnbnxm_b200.so; fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise NbnxmError(
                "%s is missing: build it with `make -C gromacs_b200/csrc` (there is no CPU fallback)" % _LIB_PATH)
        lib = C.CDLL(_LIB_PATH)
        lib.nbnxm_b200_last_error.restype = C.c_char_p
        lib.nbnxm_b200_launch_count.restype = C.c_longlong
        lib.nbnxm_b200_launch_count.argtypes = [C.c_void_p]
        _lib = lib
    return _lib


def measure_fp32_peak(device=0):
    """Pure-FFMA throughput of the device in TFLOP/s (roofline denominator of the force kernel)."""
    v = C.c_double(0)
    if load_library().nbnxm_b200_measure_fp32_peak(C.c_int(device), C.byref(v)):
        raise NbnxmError("nbnxm_b200_measure_fp32_peak failed (no CUDA device?)")
    return v.value


def _ptr(a, ctype):
    return None if a is None else a.ctypes.data_as(C.POINTER(ctype))


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


@dataclass
class StepWorkload:
    """The fields of gmx::StepWorkload this path reads (mdtypes/simulation_workload.h:63-111)."""
    computeForces: bool = True
    computeEnergy: bool = False
    computeVirial: bool = False
    useGpuFBufferOps: bool = False


@dataclass
class PairlistGpu:
    """NbnxmPairlistGpu as plain arrays (pairlist.h:335): sci int32[n,4], cjPacked uint32[n,8], excl uint32[n,32]."""
    sci: np.ndarray
    cjPacked: np.ndarray
    excl: np.ndarray
    na_ci: int = 8
    rlist: float = 0.0

    def __post_init__(self):
        self.sci = np.ascontiguousarray(self.sci, np.int32).reshape(-1, 4)
        self.cjPacked = np.ascontiguousarray(self.cjPacked, np.uint32).reshape(-1, 8)
        self.excl = np.ascontiguousarray(self.excl, np.uint32).reshape(-1, 32)


@dataclass
class AtomData:
    """The parts of nbnxm_atomdata_t this path reads/writes (atomdata.h:184-375)."""
    xq: np.ndarray                      # x(): float32[natoms,4], XYZQ
    type: np.ndarray = None             # params().type
    lj_comb: np.ndarray = None          # params().lj_comb float32[natoms,2]
    nbfp: np.ndarray = None             # params().nbfp float32[ntypes*ntypes,2] (6*C6, 12*C12)
    nbfp_comb: np.ndarray = None        # params().nbfp_comb float32[ntypes,2]
    numTypes: int = 0
    shift_vec: np.ndarray = None        # float32[45,3]
    numLocalAtoms: int = -1
    f: np.ndarray = field(default=None)  # outputBuffer(0).f float32[natoms,3]

    def numAtoms(self):
        return int(self.xq.shape[0])

    def __post_init__(self):
        self.xq = _f32(self.xq).reshape(-1, 4)
        if self.numLocalAtoms < 0:
            self.numLocalAtoms = self.numAtoms()
        if self.f is None:
            self.f = np.zeros((self.numAtoms(), 3), np.float32)


class NbnxmGpu:
    """One per rank and device, like the reference's NbnxmGpu (cuda/nbnxm_cuda_types.h:69)."""

    def __init__(self, params: Params, nbat: AtomData, device=0, bLocalAndNonlocal=False,
                 coulomb_tab=None, local_stream=None, nonlocal_stream=None):
        """gpu_init (nbnxm_gpu_data_mgmt.cpp:637)."""
        self._lib = load_library()
        self._h = C.c_void_p()
        self.params = params
        nbfp = _f32(nbat.nbfp)
        nbfp_comb = _f32(nbat.nbfp_comb) if nbat.nbfp_comb is not None and np.size(nbat.nbfp_comb) else None
        tab = _f32(coulomb_tab)
        self._check(self._lib.nbnxm_b200_init(
            C.byref(self._h), C.c_int(device), C.byref(params), C.c_int(nbat.numTypes), _ptr(nbfp, C.c_float),
            _ptr(nbfp_comb, C.c_float), _ptr(tab, C.c_float), C.c_int(0 if tab is None else tab.size),
            C.c_int(int(bLocalAndNonlocal)), C.c_void_p(local_stream), C.c_void_p(nonlocal_stream)))
        self._keep = []   # host buffers with async copies in flight

    # ---- plumbing -------------------------------------------------------------------------
    def _check(self, status):
        if status != 0:
            raise NbnxmError(self._lib.nbnxm_b200_last_error().decode())

    def gpu_free(self):
        if self._h:
            self._lib.nbnxm_b200_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.gpu_free()
        except Exception:
            pass

    # ---- search-step functions --------------------------------------------------------------
    def gpu_init_pairlist(self, h_plist: PairlistGpu, iloc=LOCAL):
        self._check(self._lib.nbnxm_b200_init_pairlist(
            self._h, C.c_int(iloc), _ptr(h_plist.sci, C.c_int), C.c_int(h_plist.sci.shape[0]),
            _ptr(h_plist.cjPacked, C.c_uint32), C.c_int(h_plist.cjPacked.shape[0]),
            _ptr(h_plist.excl, C.c_uint32), C.c_int(h_plist.excl.shape[0]), C.c_int(h_plist.na_ci)))
        self._set_list_sizes(iloc, h_plist.sci.shape[0], h_plist.cjPacked.shape[0])

    def _set_list_sizes(self, iloc, nsci, ncj):
        self._numSci = getattr(self, "_numSci", {})
        self._numSci[iloc] = nsci
        self._ncj = getattr(self, "_ncj", {})
        self._ncj[iloc] = ncj

    def gpu_init_atomdata(self, nbat: AtomData):
        t = _i32(nbat.type)
        lj = _f32(nbat.lj_comb) if nbat.lj_comb is not None and np.size(nbat.lj_comb) >= 2 * nbat.numAtoms() else None
        self._natoms = nbat.numAtoms()
        self._check(self._lib.nbnxm_b200_init_atomdata(
            self._h, C.c_int(nbat.numAtoms()), C.c_int(nbat.numLocalAtoms), _ptr(t, C.c_int), _ptr(lj, C.c_float)))

    def gpu_upload_shiftvec(self, nbat: AtomData, bDynamicBox=True):
        sv = _f32(nbat.shift_vec)
        self._check(self._lib.nbnxm_b200_upload_shiftvec(self._h, _ptr(sv, C.c_float), C.c_int(int(bDynamicBox))))

    def setupGpuShortRangeWork(self, iloc=LOCAL, haveBondedWork=False):
        self._check(self._lib.nbnxm_b200_setup_short_range_work(self._h, C.c_int(iloc), C.c_int(int(haveBondedWork))))

    def haveGpuShortRangeWork(self, iloc=LOCAL):
        return bool(self._lib.nbnxm_b200_have_short_range_work(self._h, C.c_int(iloc)))

    def nbnxm_gpu_init_x_to_nbat_x(self, atom_index, atom_offset=0, grid=0, ngrids=1):
        ai = _i32(atom_index)
        self._check(self._lib.nbnxm_b200_init_x_to_nbat_x(
            self._h, C.c_int(grid), C.c_int(ngrids), _ptr(ai, C.c_int), C.c_int(ai.shape[0]), None, None,
            C.c_int(0), C.c_int(64), C.c_int(atom_offset)))

    # ---- per-step functions -------------------------------------------------------------------
    def gpu_copy_xq_to_gpu(self, nbat: AtomData, aloc=LOCAL):
        self._keep.append(nbat.xq)
        self._check(self._lib.nbnxm_b200_copy_xq_to_gpu(self._h, C.c_int(aloc), _ptr(nbat.xq, C.c_float)))

    def nbnxm_gpu_x_to_nbat_x(self, d_x_ptr, xReadyOnDevice=None, aloc=LOCAL):
        self._check(self._lib.nbnxm_b200_x_to_nbat_x(self._h, C.c_void_p(d_x_ptr), C.c_void_p(xReadyOnDevice), C.c_int(aloc)))

    # ---- perturbed (free-energy) pair kernels ------------------------------------------------------
    def copy_gpu_fepparams(self, bFepGpuNonBonded, alphaCoul, alphaVdw, lambdaPower, sigma6WithInvalidSigma, sigma6Minimum,
                           lambdaCoul, lambdaVdw):
        """copy_gpu_fepparams (gpu_data_mgmt.h:75)"""
        self._check(self._lib.nbnxm_b200_copy_fepparams(
            self._h, C.c_int(int(bFepGpuNonBonded)), C.c_float(alphaCoul), C.c_float(alphaVdw), C.c_int(lambdaPower),
            C.c_float(sigma6WithInvalidSigma), C.c_float(sigma6Minimum), C.c_float(lambdaCoul), C.c_float(lambdaVdw)))

    def gpu_init_fep_atomdata(self, qA, qB, typeA=None, typeB=None, ljCombA=None, ljCombB=None):
        """the end-state part of gpu_init_atomdata (q4 / atomTypes4 / ljComb4), nbat order"""
        a = [_f32(qA), _f32(qB), _i32(typeA), _i32(typeB), _f32(ljCombA), _f32(ljCombB)]
        self._check(self._lib.nbnxm_b200_init_fep_atomdata(
            self._h, _ptr(a[0], C.c_float), _ptr(a[1], C.c_float), _ptr(a[2], C.c_int), _ptr(a[3], C.c_int),
            _ptr(a[4], C.c_float), _ptr(a[5], C.c_float)))

    def gpu_init_feppairlist(self, iinr, jindex, jjnr, shift, excl_fep=None, iloc=LOCAL):
        """gpu_init_feppairlist (nbnxm_gpu_data_mgmt.cpp:880) with nbat atom indices"""
        a = [_i32(iinr), _i32(jindex), _i32(jjnr), _i32(shift)]
        ex = None if excl_fep is None else np.ascontiguousarray(excl_fep, np.uint8)
        self._check(self._lib.nbnxm_b200_init_feppairlist(
            self._h, C.c_int(iloc), C.c_int(a[0].shape[0]), _ptr(a[0], C.c_int), _ptr(a[1], C.c_int), _ptr(a[2], C.c_int),
            _ptr(a[3], C.c_int), _ptr(ex, C.c_ubyte)))

    def gpu_launch_free_energy_kernel(self, stepWork: StepWorkload, iloc=LOCAL):
        self._check(self._lib.nbnxm_b200_launch_free_energy_kernel(
            self._h, C.c_int(iloc), C.c_int(int(stepWork.computeEnergy)), C.c_int(int(stepWork.computeVirial))))

    def gpu_launch_foreign_energy_kernel(self, lambda_coul, lambda_vdw, iloc=LOCAL):
        """energies and dV/dlambda of the perturbed pairs at foreign lambdas; returns float64[nlambda, 4]:
        E_lj, E_el, dvdl_lj, dvdl_el"""
        lc, lv = _f32(lambda_coul), _f32(lambda_vdw)
        self._check(self._lib.nbnxm_b200_launch_foreign_energy_kernel(
            self._h, C.c_int(iloc), C.c_int(lc.shape[0]), _ptr(lc, C.c_float), _ptr(lv, C.c_float)))
        out = np.zeros((lc.shape[0], 4), np.float64)
        self._check(self._lib.nbnxm_b200_get_fep_foreign(self._h, C.c_int(lc.shape[0]), _ptr(out, C.c_double)))
        return out

    def gpu_get_fep_dvdl(self, clear=False):
        """(dvdl_lj, dvdl_el) accumulated by the perturbed kernels' energy launches"""
        a, b = C.c_float(0), C.c_float(0)
        self._check(self._lib.nbnxm_b200_get_fep_dvdl(self._h, C.byref(a), C.byref(b), C.c_int(int(clear))))
        return a.value, b.value

    def gpu_force_reduction_reinit(self, cell):
        """GpuForceReduction::reinit: cell[natoms] maps atoms to nbat slots (GridSet::cells())."""
        c = _i32(cell)
        self._check(self._lib.nbnxm_b200_init_reduce_f(self._h, _ptr(c, C.c_int), C.c_int(c.shape[0])))

    def gpu_force_reduction_execute(self, d_f_total_ptr, d_rvec_to_add_ptr=None, atom_start=0, num_atoms=None, accumulate=False,
                                    stream=None):
        """GpuForceReduction::execute: f_total[a] (+)= f_nbat[cell[a]] (+ rvec[a]) on the device."""
        self._check(self._lib.nbnxm_b200_reduce_f(
            self._h, C.c_void_p(d_f_total_ptr), C.c_void_p(d_rvec_to_add_ptr), C.c_int(atom_start), C.c_int(num_atoms),
            C.c_int(int(accumulate)), C.c_void_p(stream)))

    def gpu_launch_kernel(self, stepWork: StepWorkload, iloc=LOCAL):
        self._check(self._lib.nbnxm_b200_launch_kernel(
            self._h, C.c_int(iloc), C.c_int(int(stepWork.computeEnergy)), C.c_int(int(stepWork.computeVirial))))

    def gpu_launch_kernel_pruneonly(self, iloc=LOCAL, numParts=1):
        self._check(self._lib.nbnxm_b200_launch_kernel_pruneonly(self._h, C.c_int(iloc), C.c_int(numParts)))

    def gpu_launch_cpyback(self, nbat: AtomData, stepWork: StepWorkload, aloc=LOCAL):
        self._keep.append(nbat.f)
        self._check(self._lib.nbnxm_b200_launch_cpyback(
            self._h, C.c_int(aloc), _ptr(nbat.f, C.c_float), C.c_int(int(stepWork.computeEnergy)),
            C.c_int(int(stepWork.computeVirial)), C.c_int(int(stepWork.useGpuFBufferOps))))

    def gpu_wait_finish_task(self, stepWork: StepWorkload, aloc=LOCAL, shiftForces=None):
        """Returns (e_lj, e_el) added by this task; shift forces are added into shiftForces[45,3]."""
        e_lj, e_el = C.c_float(0), C.c_float(0)
        fs = np.zeros((45, 3), np.float32)
        self._check(self._lib.nbnxm_b200_wait_finish_task(
            self._h, C.c_int(aloc), C.c_int(int(stepWork.computeEnergy)), C.c_int(int(stepWork.computeVirial)),
            C.byref(e_lj), C.byref(e_el), _ptr(fs, C.c_float)))
        if shiftForces is not None:
            shiftForces += fs
        self._keep.clear()
        return e_lj.value, e_el.value

    def gpu_try_finish_task(self, stepWork: StepWorkload, aloc=LOCAL, shiftForces=None):
        e_lj, e_el, done = C.c_float(0), C.c_float(0), C.c_int(0)
        fs = np.zeros((45, 3), np.float32)
        self._check(self._lib.nbnxm_b200_try_finish_task(
            self._h, C.c_int(aloc), C.c_int(int(stepWork.computeEnergy)), C.c_int(int(stepWork.computeVirial)),
            C.byref(e_lj), C.byref(e_el), _ptr(fs, C.c_float), C.byref(done)))
        if done.value and shiftForces is not None:
            shiftForces += fs
        return bool(done.value), e_lj.value, e_el.value

    def do_force_step(self, step, stepWork: StepWorkload, have_halo=False, dynamic_pruning=False, num_parts=1,
                      xq_host=None, f_host=None):
        """The nonbonded part of one do_force step in one foreign call (nbnxm_b200_do_force_step): the same sequence
        as the individual gpu_* calls, for callers that pay per call."""
        fl = StepFlags(int(stepWork.computeEnergy), int(stepWork.computeVirial), int(have_halo), int(dynamic_pruning),
                       int(num_parts))
        xq = _ptr(xq_host, C.c_float) if xq_host is not None else None
        f = _ptr(f_host, C.c_float) if f_host is not None else None
        self._check(self._lib.nbnxm_b200_do_force_step(self._h, C.c_int(step), C.byref(fl), xq, f))

    def do_force_step_pipelined(self, step, stepWork: StepWorkload, plan, xq_host, f_host, dynamic_pruning=False, num_parts=1,
                                have_halo=0):
        """nbnxm_b200_do_force_step_pipelined with a ChunkPlan (gromacs_b200/pipeline.py); the pair list uploaded with
        gpu_init_pairlist must be plan.plist.  have_halo=3: a slab of a multi-GPU run with the peer-memory halo (the chunks
        cover the home atoms and the local list)."""
        fl = StepFlags(int(stepWork.computeEnergy), int(stepWork.computeVirial), int(have_halo), int(dynamic_pruning), int(num_parts))
        self._check(self._lib.nbnxm_b200_do_force_step_pipelined(
            self._h, C.c_int(step), C.byref(fl), _ptr(xq_host, C.c_float), _ptr(f_host, C.c_float), C.c_int(plan.nchunks),
            _ptr(plan.first_atom, C.c_int), _ptr(plan.first_sci, C.c_int), _ptr(plan.needs, C.c_uint32)))

    def set_pipeline_timeline(self, enable=True):
        self._check(self._lib.nbnxm_b200_set_pipeline_timeline(self._h, C.c_int(int(enable))))

    def pipeline_timeline(self):
        """float32[nchunks, 4]: H2D end, kernel start, kernel end, D2H end of each chunk of the last pipelined step, in ms"""
        ms = np.zeros((32, 4), np.float32)
        n = C.c_int(0)
        self._check(self._lib.nbnxm_b200_get_pipeline_timeline(self._h, C.c_int(32), C.byref(n), _ptr(ms, C.c_float)))
        return ms[:n.value].copy()

    def gpu_clear_outputs(self, computeVirial=True):
        self._check(self._lib.nbnxm_b200_clear_outputs(self._h, C.c_int(int(computeVirial))))

    def nbnxmInsertNonlocalGpuDependency(self, iloc):
        self._check(self._lib.nbnxm_b200_insert_nonlocal_dependency(self._h, C.c_int(iloc)))

    # ---- queries --------------------------------------------------------------------------------
    def gpu_min_ci_balanced(self):
        return int(self._lib.nbnxm_b200_min_ci_balanced(self._h))

    def gpu_is_kernel_ewald_analytical(self):
        return bool(self._lib.nbnxm_b200_is_kernel_ewald_analytical(self._h))

    def gpu_pme_loadbal_update_param(self, params: Params, coulomb_tab=None):
        tab = _f32(coulomb_tab)
        self._check(self._lib.nbnxm_b200_update_params(
            self._h, C.byref(params), _ptr(tab, C.c_float), C.c_int(0 if tab is None else tab.size)))
        self.params = params

    def set_timing(self, enable=True):
        self._check(self._lib.nbnxm_b200_set_timing(self._h, C.c_int(int(enable))))

    def gpu_get_timings(self):
        t = Timings()
        self._check(self._lib.nbnxm_b200_get_timings(self._h, C.byref(t)))
        return t

    def gpu_reset_timings(self):
        self._check(self._lib.nbnxm_b200_reset_timings(self._h))

    def device_buffers(self):
        """(d_xq, d_f, natoms): raw device addresses of the float4 coordinates and float3 forces."""
        xq, f, n = C.c_void_p(), C.c_void_p(), C.c_int()
        self._check(self._lib.nbnxm_b200_get_device_buffers(self._h, C.byref(xq), C.byref(f), C.byref(n)))
        return xq.value, f.value, n.value

    def gpuGetNBAtomData(self):
        """(d_f, d_fshift): the outputs other device code may add into (gpuGetNBAtomData, nbnxm_gpu_data_mgmt.cpp:1823); from
        this call on gpu_clear_outputs zeroes them and the copy-back adds the kernels' forces on top."""
        f, fs = C.c_void_p(), C.c_void_p()
        self._check(self._lib.nbnxm_b200_get_shared_outputs(self._h, C.byref(f), C.byref(fs)))
        return f.value, fs.value

    def streams(self):
        a, b = C.c_void_p(), C.c_void_p()
        self._check(self._lib.nbnxm_b200_get_streams(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def download_pairlist(self, iloc=LOCAL):
        """(cjPacked uint32[n,8], imask_outer uint32[2n], sciSorted int32[nsci,4], sciCount int32[nsci], rollingPart)."""
        nsci, ncj = self._numSci[iloc], self._ncj[iloc]
        cj = np.zeros((ncj, 8), np.uint32)
        im = np.zeros(2 * ncj, np.uint32)
        ss = np.zeros((nsci, 4), np.int32)
        sc = np.zeros(nsci, np.int32)
        rp = np.zeros(nsci, np.int32)
        self._check(self._lib.nbnxm_b200_download_pairlist(
            self._h, C.c_int(iloc), _ptr(cj, C.c_uint32), _ptr(im, C.c_uint32), _ptr(ss, C.c_int),
            _ptr(sc, C.c_int), _ptr(rp, C.c_int)))
        return cj, im, ss, sc, rp

    def set_pair_counting(self, enable=True):
        self._check(self._lib.nbnxm_b200_set_pair_counting(self._h, C.c_int(int(enable))))

    def get_pair_count(self, iloc=LOCAL):
        v = C.c_longlong(0)
        self._check(self._lib.nbnxm_b200_get_pair_count(self._h, C.c_int(iloc), C.byref(v)))
        return v.value

    def launch_count(self):
        return int(self._lib.nbnxm_b200_launch_count(self._h))

    # ---- halo helpers (x-slab decomposition) -------------------------------------------------------
    def pack_xq(self, d_index_ptr, n, shift3, d_send_ptr, stream=None):
        sh = _f32(shift3)
        self._check(self._lib.nbnxm_b200_pack_xq(self._h, C.c_void_p(d_index_ptr), C.c_int(n), _ptr(sh, C.c_float),
                                                C.c_void_p(d_send_ptr), C.c_void_p(stream)))

    def unpack_xq(self, first, n, d_recv_ptr, stream=None):
        self._check(self._lib.nbnxm_b200_unpack_xq(self._h, C.c_int(first), C.c_int(n), C.c_void_p(d_recv_ptr), C.c_void_p(stream)))

    def pack_f(self, first, n, d_send_ptr, stream=None):
        self._check(self._lib.nbnxm_b200_pack_f(self._h, C.c_int(first), C.c_int(n), C.c_void_p(d_send_ptr), C.c_void_p(stream)))

    def unpack_add_f(self, d_index_ptr, n, d_recv_ptr, stream=None):
        self._check(self._lib.nbnxm_b200_unpack_add_f(self._h, C.c_void_p(d_index_ptr), C.c_int(n), C.c_void_p(d_recv_ptr), C.c_void_p(stream)))
