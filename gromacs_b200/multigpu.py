"""x-slab decomposition of the NBNXM force step over the GPUs of one box, one process per GPU.

Mirrors the reference's domain-decomposed step (do_force, src/gromacs/mdlib/sim_util.cpp:1815-1922,
2424-2442): every rank holds its home atoms followed by a halo imported from the +x neighbour
(eighth-shell zones restricted to one dimension, src/gromacs/domdec/domdec_zones.cpp:55-83), runs a
local list (home x home) on the local stream and a non-local list (home x halo) on the non-local
stream, and returns the halo forces to their owner.  The transport is the library's NCCL send/recv
halo exchange (include/nbnxm_b200.h, nbnxm_b200_halo_*); torch.distributed is only used to launch the
ranks, broadcast the NCCL id, and take barriers / max-over-ranks timings.

Host-side pieces (slab ranges, list re-indexing) are pure numpy and are tested on CPU with gloo.
"""
import ctypes as C
import json
import os
import sys
import time
from dataclasses import dataclass

import numpy as np

from .nbnxm import LOCAL, NONLOCAL, AtomData, NbnxmGpu, PairlistGpu, StepWorkload
from .slabs import slab_bin_ranges


@dataclass
class SlabPlan:
    """What one rank needs at a search step: its atoms (home then halo), its two lists re-indexed to that
    order, and the contiguous send / receive ranges of the halo exchange."""
    rank: int
    nranks: int
    home_bins: tuple
    halo_bins: tuple
    nbat: AtomData              # home atoms then halo atoms
    local: PairlistGpu
    nonlocal_: PairlistGpu
    send_first: int
    send_count: int
    recv_first: int
    recv_count: int
    home_slice: slice           # where the home atoms live in the global nbat order
    halo_slice: slice


def _reindex(pl, first_home_bin, first_halo_bin, num_home_bins, nclusters_total, halo):
    """Global bin / cluster indices -> rank order (home bins first, then halo bins): nbnxm_b200_pairlist_reindex."""
    import ctypes as C
    from .nbnxm import load_library
    sci = np.ascontiguousarray(pl.sci, np.int32).reshape(-1, 4).copy()
    cjp = np.ascontiguousarray(pl.cjPacked, np.uint32).reshape(-1, 8).copy()
    if load_library().nbnxm_b200_pairlist_reindex(
            sci.ctypes.data_as(C.POINTER(C.c_int)), C.c_int(sci.shape[0]), cjp.ctypes.data_as(C.POINTER(C.c_uint32)),
            C.c_int(cjp.shape[0]), C.c_int(first_home_bin), C.c_int(first_halo_bin), C.c_int(num_home_bins),
            C.c_int(nclusters_total), C.c_int(int(halo))):
        raise RuntimeError("nbnxm_b200_pairlist_reindex failed")
    return PairlistGpu(sci=sci, cjPacked=cjp, excl=pl.excl, na_ci=pl.na_ci, rlist=pl.rlist)


def make_slab_plan(wl, rank, nranks, min_sci=0):
    """constructPairlist(Local) + constructPairlist(NonLocal) for slab `rank` of `nranks`
    (sim_util.cpp:1458, :1503), from one global grid."""
    grid, cfg, g = wl.grid, wl.cfg, wl.nbat
    rlist = cfg["rlist_outer"]
    home, halo, tx = slab_bin_ranges(grid, nranks, rank, rlist)
    ex_i, ex_a = wl.box.excl_index, wl.box.excl_atoms
    if nranks == 1:
        local = grid.pairlist(rlist, ex_i, ex_a, min_sci=min_sci)
        empty = PairlistGpu(sci=np.zeros((0, 4), np.int32), cjPacked=np.zeros((0, 8), np.uint32),
                            excl=np.full((1, 32), 0xffffffff, np.uint32))
        return SlabPlan(rank, 1, home, halo, g, local, empty, 0, 0, g.numAtoms(), 0, slice(0, g.numAtoms()),
                        slice(0, 0))
    # non-local work is about rlist / slab width of the local work: share the balance target accordingly
    loc_g = grid.pairlist(rlist, ex_i, ex_a, min_sci=min_sci, bins=home, j_bins=home)
    nloc_g = grid.pairlist(rlist, ex_i, ex_a, min_sci=min_sci // 2, bins=home, j_bins=halo, inter_zone=True,
                           required_tx=tx)
    nhome, nhalo = home[1] - home[0], halo[1] - halo[0]
    ncl = (nhome + nhalo) * 8
    local = _reindex(loc_g, home[0], halo[0], nhome, ncl, halo=False)
    nonloc = _reindex(nloc_g, home[0], halo[0], nhome, ncl, halo=True)
    hs, ls = slice(home[0] * 64, home[1] * 64), slice(halo[0] * 64, halo[1] * 64)
    cat = lambda a: None if a is None else np.concatenate([a[hs], a[ls]])
    nbat = AtomData(xq=cat(g.xq), type=cat(g.type), lj_comb=cat(g.lj_comb), nbfp=g.nbfp, nbfp_comb=g.nbfp_comb,
                    numTypes=g.numTypes, shift_vec=g.shift_vec, numLocalAtoms=nhome * 64)
    # what the -x neighbour imports from us: the first columns of our slab = its halo range
    _, halo_prev, _ = slab_bin_ranges(grid, nranks, (rank - 1) % nranks, rlist)
    assert halo_prev[0] == home[0], "the -x neighbour's halo must start at our first column"
    return SlabPlan(rank, nranks, home, halo, nbat, local, nonloc, 0, (halo_prev[1] - halo_prev[0]) * 64,
                    nhome * 64, nhalo * 64, hs, ls)


class _DeviceArray:
    """__cuda_array_interface__ view of a raw float32 device buffer of the library"""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "strides": None,
                                         "version": 3}


class DeviceSlabSearch:
    """The search step of one rank without the host (the multi-GPU form of SURVEY 8f #1): the whole system is put on the grid
    on this rank's GPU (nbnxm_b200_gpu_search_put_atoms_on_grid, 15 ms for 12.3 M atoms against 0.3 s on 16 host threads),
    the rank's atoms (home bins, then the halo bins of the +x neighbour) are gathered into its handle and its local and
    non-local lists are built, re-indexed and installed on the device (nbnxm_b200_gpu_search_gather_slab / _build_slab).
    No rank grids or searches anything on the host; make_slab_plan (host gridder + host builder) stays as the reference the
    tests compare this with (tests/test_gpu_search.py::test_device_slab_lists_equal_the_host_plan)."""

    def __init__(self, wl, device):
        import torch
        from .pairsearch import GpuPairSearch
        self.wl = wl
        b = wl.box
        self.whole = NbnxmGpu(wl.params, wl.nbat, device=device)
        self.search = GpuPairSearch(self.whole)
        self.search.set_atoms(b.q, b.type, wl.nbat.numTypes, wl.nbat.nbfp_comb, b.excl_index, b.excl_atoms)
        self.x_dev = torch.from_numpy(np.ascontiguousarray(b.x, np.float32)).cuda()
        torch.cuda.synchronize()

    def search_step(self, nb, rank, nranks, min_sci=0):
        """putAtomsOnGrid + constructPairlist(Local) + constructPairlist(NonLocal) + gpu_init_* of rank `rank` into handle `nb`;
        returns the SlabPlan (ranges and host buffers; the lists stay on the device)."""
        import torch
        from .slabs import slab_bin_ranges_from_columns
        wl, b = self.wl, self.wl.box
        rlist = wl.cfg["rlist_outer"]
        _, nbins, ncx, ncy = self.search.put_atoms_on_grid(b.box, self.x_dev.data_ptr(), nslabs=nranks)
        _, first_bin, self.grid_ms = self.search.get_order()
        home, halo, tx = slab_bin_ranges_from_columns(b.box[0], ncx, ncy, first_bin, nranks, rank, rlist)
        _, halo_prev, _ = slab_bin_ranges_from_columns(b.box[0], ncx, ncy, first_bin, nranks, (rank - 1) % nranks, rlist)
        if nranks > 1:
            assert halo_prev[0] == home[0], "the -x neighbour's halo must start at our first column"
        self.search.gather_slab(nb, home, halo)
        self.list_ms = 0.0
        self.search.build_slab(nb, LOCAL, rlist, home, halo, min_sci=min_sci)
        self.list_ms += self.search.build_ms
        # what the chunk plan of the pipelined end-to-end step needs of the local list: the i-bin of every entry
        self.local_sci = self.search.download_sci()
        self.columns = (float(b.box[0]), ncx, ncy, first_bin)
        self.search.build_slab(nb, NONLOCAL, rlist, home, halo, required_tx=tx, min_sci=min_sci // 2)
        self.list_ms += self.search.build_ms
        nhome, nhalo = home[1] - home[0], halo[1] - halo[0]
        # the host's copy of the rank's coordinates in grid order (nbat->x()), for callers that keep x on the host
        d_xq, _, n = nb.device_buffers()
        xq = torch.as_tensor(_DeviceArray(d_xq, (n, 4)), device="cuda").cpu().numpy().copy()
        g = wl.nbat
        nbat = AtomData(xq=xq, type=None, lj_comb=None, nbfp=g.nbfp, nbfp_comb=g.nbfp_comb, numTypes=g.numTypes,
                        shift_vec=g.shift_vec, numLocalAtoms=nhome * 64)
        hs, ls = slice(home[0] * 64, home[1] * 64), slice(halo[0] * 64, halo[1] * 64)
        return SlabPlan(rank, nranks, home, halo, nbat, None, None, 0, (halo_prev[1] - halo_prev[0]) * 64 if nranks > 1 else 0,
                        nhome * 64, nhalo * 64, hs, ls)

    def whole_system_forces(self, min_sci=0):
        """single-domain F+E step of the whole system on this GPU (list built on the device, unpruned): forces in the global
        grid order and (e_lj, e_el) - the reference of the multi-GPU parity figure"""
        import copy
        wl = self.wl
        whole = self.whole
        n = self.search.nbins * 64
        f = np.zeros((n, 3), np.float32)
        stub = AtomData(xq=np.zeros((n, 4), np.float32), nbfp=wl.nbat.nbfp, nbfp_comb=wl.nbat.nbfp_comb, numTypes=wl.nbat.numTypes,
                        shift_vec=wl.nbat.shift_vec)
        stub.f = f
        p1 = copy.copy(wl.params)
        p1.use_dynamic_pruning = 0
        whole.gpu_pme_loadbal_update_param(p1)
        self.search.build(wl.cfg["rlist_outer"], LOCAL, min_sci=min_sci)
        sw = StepWorkload(computeEnergy=True, computeVirial=True, useGpuFBufferOps=False)
        whole.setupGpuShortRangeWork(LOCAL)
        whole.gpu_upload_shiftvec(stub)
        whole.gpu_clear_outputs(True)
        whole.gpu_launch_kernel(sw, LOCAL)
        whole.gpu_launch_cpyback(stub, sw, LOCAL)
        e = whole.gpu_wait_finish_task(sw, LOCAL)
        return f, e

    def free(self):
        self.search.free()
        self.whole.gpu_free()
        self.x_dev = None


# ---- the halo part of the C ABI -------------------------------------------------------------------

class HaloExchange:
    """nbnxm_b200_halo_*: GpuHaloExchange for x-slabs (domdec/gpuhaloexchange.h:80-130)."""

    def __init__(self, nb: NbnxmGpu, unique_id: bytes, rank, nranks):
        self.nb, self._lib = nb, nb._lib
        self.peer = False
        buf = C.create_string_buffer(unique_id, 128)
        nb._check(self._lib.nbnxm_b200_halo_init(nb._h, buf, C.c_int(rank), C.c_int(nranks)))

    @staticmethod
    def unique_id(lib):
        buf = C.create_string_buffer(128)
        if lib.nbnxm_b200_halo_get_unique_id(buf, C.c_int(128)):
            raise RuntimeError(lib.nbnxm_b200_last_error().decode())
        return buf.raw

    def reinitHalo(self, plan: SlabPlan):
        self.nb._check(self._lib.nbnxm_b200_halo_set_ranges(
            self.nb._h, C.c_int(plan.send_first), C.c_int(plan.send_count), C.c_int(plan.recv_first),
            C.c_int(plan.recv_count)))

    def enable_peer_memory(self, all_gather_bytes, rank, nranks):
        """Peer-memory halo (nbnxm_b200_peer_*): maps the +x neighbour's xq / force accumulator; afterwards the step
        needs no transport.  `all_gather_bytes(b)` returns the list of every rank's bytes `b`.  Call after
        gpu_init_atomdata and reinitHalo of a search step."""
        n = self._lib.nbnxm_b200_peer_blob_size()
        buf = C.create_string_buffer(n)
        self.nb._check(self._lib.nbnxm_b200_peer_export(self.nb._h, buf, C.c_int(n)))
        blobs = all_gather_bytes(buf.raw)
        up = C.create_string_buffer(blobs[(rank + 1) % nranks], n)
        self.nb._check(self._lib.nbnxm_b200_peer_import(self.nb._h, up, C.c_int(n)))
        self.peer = True

    def peer_error(self):
        e = C.c_int(0)
        self.nb._check(self._lib.nbnxm_b200_peer_error(self.nb._h, C.byref(e)))
        return e.value

    def communicateGpuHaloCoordinates(self):
        self.nb._check(self._lib.nbnxm_b200_halo_exchange_x(self.nb._h))

    def communicateGpuHaloForces(self):
        self.nb._check(self._lib.nbnxm_b200_halo_exchange_f(self.nb._h))

    def set_timing(self, enable=True):
        self.nb._check(self._lib.nbnxm_b200_halo_set_timing(self.nb._h, C.c_int(int(enable))))

    def timings(self, reset=False):
        x, f = C.c_double(0), C.c_double(0)
        self.nb._check(self._lib.nbnxm_b200_halo_get_timings(self.nb._h, C.byref(x), C.byref(f), C.c_int(int(reset))))
        return x.value, f.value


class SlabStep:
    """One rank's force step: the do_force sequence around the two kernels (sim_util.cpp:1639-2442)."""

    def __init__(self, nb, halo, plan, energy, dynamic_pruning, num_parts=3):
        self.nb, self.halo, self.plan = nb, halo, plan
        self.energy, self.dynamic_pruning, self.num_parts = energy, dynamic_pruning, num_parts
        self.sw = StepWorkload(computeEnergy=energy, computeVirial=energy, useGpuFBufferOps=True)
        self.multi = plan.nranks > 1
        self.chunk_plan = None      # set to a pipeline.ChunkPlan: steps with host buffers run chunk-pipelined (peer halo only)

    def search_step(self):
        nb, plan = self.nb, self.plan
        if plan.local is None:
            # atom data and both lists were built and installed on the device (DeviceSlabSearch.search_step)
            nb.setupGpuShortRangeWork(LOCAL)
            nb.setupGpuShortRangeWork(NONLOCAL)
            nb.gpu_upload_shiftvec(plan.nbat)
            if self.multi and self.halo is not None:
                self.halo.reinitHalo(plan)
            return
        nb.gpu_init_atomdata(plan.nbat)
        nb.gpu_init_pairlist(plan.local, LOCAL)
        nb.gpu_init_pairlist(plan.nonlocal_, NONLOCAL)
        nb.setupGpuShortRangeWork(LOCAL)
        nb.setupGpuShortRangeWork(NONLOCAL)
        nb.gpu_upload_shiftvec(plan.nbat)
        if self.multi and self.halo is not None:
            self.halo.reinitHalo(plan)
        nb.gpu_copy_xq_to_gpu(plan.nbat, LOCAL)

    def __call__(self, step, host_io=False):
        """The do_force sequence (sim_util.cpp:1639-2442) through nbnxm_b200_do_force_step, which issues the same
        gpu_* calls as `sequence()` below in one foreign call."""
        nb, sw = self.nb, self.sw
        sw.useGpuFBufferOps = not host_io
        if host_io and self.multi and self.chunk_plan is not None and self.halo is not None and self.halo.peer:
            nb.do_force_step_pipelined(step, sw, self.chunk_plan, self.plan.nbat.xq, self.plan.nbat.f,
                                       dynamic_pruning=self.dynamic_pruning, num_parts=self.num_parts, have_halo=3)
            nb.gpu_wait_finish_task(sw, NONLOCAL)
            return nb.gpu_wait_finish_task(sw, LOCAL)
        nb.do_force_step(step, sw, have_halo=(2 if self.halo is None else (3 if self.halo.peer else 1)) if self.multi else 0, dynamic_pruning=self.dynamic_pruning, num_parts=self.num_parts,
                         xq_host=self.plan.nbat.xq if host_io else None, f_host=self.plan.nbat.f if host_io else None)
        if host_io:
            if self.multi:
                nb.gpu_wait_finish_task(sw, NONLOCAL)
            return nb.gpu_wait_finish_task(sw, LOCAL)
        return None

    def sequence(self, step, host_io=False):
        """The same step spelled out call by call, as the reference's do_force issues it (used by the tests)."""
        nb, sw = self.nb, self.sw
        if host_io:
            nb.gpu_copy_xq_to_gpu(self.plan.nbat, LOCAL)
        nb.gpu_clear_outputs(computeVirial=self.energy)
        if self.multi:
            nb.nbnxmInsertNonlocalGpuDependency(LOCAL)         # clear + H2D done -> non-local stream may start
            self.halo.communicateGpuHaloCoordinates()
        nb.gpu_launch_kernel(sw, LOCAL)
        if self.multi:
            nb.gpu_launch_kernel(sw, NONLOCAL)
            self.halo.communicateGpuHaloForces()
        if self.dynamic_pruning:
            # with DD the rolling prune alternates local (even) / non-local (odd), prunekerneldispatch.cpp:123-128
            if not self.multi:
                if step % 2 == 1:
                    nb.gpu_launch_kernel_pruneonly(LOCAL, self.num_parts)
            elif step % 2 == 0:
                nb.gpu_launch_kernel_pruneonly(LOCAL, self.num_parts)
            else:
                nb.gpu_launch_kernel_pruneonly(NONLOCAL, self.num_parts)
        sw.useGpuFBufferOps = not host_io
        if self.multi:
            nb.gpu_launch_cpyback(self.plan.nbat, StepWorkload(useGpuFBufferOps=True), NONLOCAL)
        nb.gpu_launch_cpyback(self.plan.nbat, sw, LOCAL)
        if host_io:
            if self.multi:
                nb.gpu_wait_finish_task(sw, NONLOCAL)
            return nb.gpu_wait_finish_task(sw, LOCAL)
        return None

    def finish(self):
        if self.multi:
            self.nb.gpu_wait_finish_task(self.sw, NONLOCAL)
        return self.nb.gpu_wait_finish_task(self.sw, LOCAL)


def parity_vs_single_gpu(wl, plan, nb, halo, step, cfg, num_parts, rank, world, local_rank, args, dsearch=None):
    """Correctness figure of the N-GPU step on the bench line: the forces every rank holds for its home atoms after an
    end-to-end step (dynamically pruned local + non-local lists, halo transport as benchmarked) and the energies of one
    extra F+E step, against the single-GPU path on the whole system (rank 0 runs it on its GPU through the same library:
    unpruned outer list, one domain) - product against product; the single-GPU path is checked against the oracle by
    tests/test_gpu_benched_configs.py and by the N = 1 bench line's own parity figure."""
    import copy
    import torch
    import torch.distributed as dist
    n_all = torch.tensor([dsearch.search.nbins * 64 if dsearch is not None else (wl.nbat.numAtoms() if wl.grid is not None else 0)],
                         dtype=torch.int64, device="cuda")
    dist.broadcast(n_all, 0)
    n_all = int(n_all.item())
    f_full = torch.zeros((n_all, 3), dtype=torch.float32, device="cuda")
    e_full = torch.zeros(2, dtype=torch.float64, device="cuda")
    # N ranks: forces of the last end-to-end step are in plan.nbat.f; one F+E step for the energies
    estep = SlabStep(nb, halo, plan, True, cfg["dynamic_pruning"], num_parts)
    estep.chunk_plan = step.chunk_plan      # the pipelined end-to-end step, where the bench used it
    e_lj, e_el = estep(1, host_io=True)
    f_mine = np.array(plan.nbat.f[:plan.nbat.numLocalAtoms], np.float64)
    e_n = torch.tensor([e_lj, e_el], dtype=torch.float64, device="cuda")
    dist.all_reduce(e_n)
    if rank == 0 and dsearch is not None:
        f1, e1 = dsearch.whole_system_forces(min_sci=args.min_sci or nb.gpu_min_ci_balanced())
        f_full.copy_(torch.from_numpy(f1))
        e_full.copy_(torch.tensor(e1, dtype=torch.float64))
    elif rank == 0:
        p1 = copy.copy(wl.params)
        p1.use_dynamic_pruning = 0
        g = wl.nbat
        if g.f is None or g.f.shape[0] != n_all:
            g.f = np.zeros((n_all, 3), np.float32)
        nb1 = NbnxmGpu(p1, g, device=local_rank)
        try:
            pl = wl.grid.pairlist(cfg["rlist_outer"], wl.box.excl_index, wl.box.excl_atoms, min_sci=args.min_sci or nb1.gpu_min_ci_balanced())
            sw = StepWorkload(computeEnergy=True, computeVirial=True, useGpuFBufferOps=False)
            nb1.gpu_init_atomdata(g)
            nb1.gpu_init_pairlist(pl, LOCAL)
            nb1.setupGpuShortRangeWork(LOCAL)
            nb1.gpu_upload_shiftvec(g)
            nb1.do_force_step(0, sw, have_halo=False, dynamic_pruning=False, num_parts=1, xq_host=g.xq, f_host=g.f)
            e1 = nb1.gpu_wait_finish_task(sw, LOCAL)
            f_full.copy_(torch.from_numpy(np.ascontiguousarray(g.f)))
            e_full.copy_(torch.tensor(e1, dtype=torch.float64))
        finally:
            nb1.gpu_free()
    dist.broadcast(f_full, 0)
    dist.broadcast(e_full, 0)
    ref = f_full[plan.home_slice].cpu().numpy().astype(np.float64)
    acc = torch.tensor([((f_mine - ref) ** 2).sum(), (ref ** 2).sum()], dtype=torch.float64, device="cuda")
    mx = torch.tensor([np.abs(f_mine - ref).max(), np.abs(ref).max()], dtype=torch.float64, device="cuda")
    dist.all_reduce(acc)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    e1 = e_full.cpu().numpy()
    en = e_n.cpu().numpy()
    return {"f_relrms_vs_n1": float(np.sqrt(acc[0].item() / acc[1].item())), "f_maxcomp_rel_vs_n1": float(mx[0].item() / mx[1].item()),
            "e_lj_rel_vs_n1": float(abs(en[0] - e1[0]) / abs(e1[0])), "e_el_rel_vs_n1": float(abs(en[1] - e1[1]) / abs(e1[1])),
            "tolerance": {"f_relrms": 5e-6, "f_maxcomp_rel": 1e-4, "e_rel": 1e-6},
            "what": "home-atom forces of all ranks after an end-to-end step on the dynamically pruned slab lists, energies of one "
                    "F+E step summed over ranks, against the single-GPU path on the whole system (rank 0, same library)"}


def bench_multi_gpu(args, rank, world, local_rank):
    """bench.py's N > 1 arm: strong scaling of one water box over `world` x-slabs."""
    import torch
    import torch.distributed as dist
    from .nbnxm import load_library, measure_fp32_peak
    from .workload import make_workload, rolling_prune_parts

    torch.cuda.set_device(local_rank)
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    lib = load_library()
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(HaloExchange.unique_id(lib)), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    uid = bytes(idt.cpu().numpy().tobytes())

    ncores = len(os.sched_getaffinity(0))
    device_lists = getattr(args, "slab_lists", "device") == "device"
    # device_lists: nothing of the search step runs on the host - every rank grids the system and builds its two lists on
    # its own GPU (DeviceSlabSearch); "host": one host grid, the host builder per slab (make_slab_plan)
    wl = make_workload(args.workload, nthreads=max(1, ncores // world), nslabs=world, host_grid=not device_lists)
    cfg = wl.cfg
    energy = cfg["energy"]
    nb = NbnxmGpu(wl.params, wl.nbat, device=local_rank, bLocalAndNonlocal=True)
    halo = HaloExchange(nb, uid, rank, world)
    dsearch = None
    if device_lists:
        dsearch = DeviceSlabSearch(wl, local_rank)
        plan = dsearch.search_step(nb, rank, world, min_sci=args.min_sci or nb.gpu_min_ci_balanced())
        first = (dsearch.grid_ms, dsearch.list_ms)
        slab_local_sci, slab_columns_info = dsearch.local_sci, dsearch.columns
        # once more with every buffer in place: what a search step costs from the second list on
        plan = dsearch.search_step(nb, rank, world, min_sci=args.min_sci or nb.gpu_min_ci_balanced())
        search_rec = {"where": "device, every rank", "gpu_grid_ms_whole_system": dsearch.grid_ms, "gpu_lists_ms_this_rank": dsearch.list_ms,
                      "first_call_with_allocations_ms": {"grid": first[0], "lists": first[1]}}
        if rank != 0:
            dsearch.free()      # rank 0 keeps the whole-system grid for the parity figure
            dsearch = None
    else:
        plan = make_slab_plan(wl, rank, world, min_sci=args.min_sci or nb.gpu_min_ci_balanced())
        search_rec = {"where": "host: one grid of the whole system, host builder per slab"}
    nbat = plan.nbat
    xq_pin = torch.empty((nbat.numAtoms(), 4), dtype=torch.float32).pin_memory()
    xq_pin.numpy()[:] = nbat.xq
    nbat.xq = xq_pin.numpy()
    f_pin = torch.zeros((nbat.numAtoms(), 3), dtype=torch.float32).pin_memory()
    nbat.f = f_pin.numpy()

    num_parts = rolling_prune_parts(cfg)
    step = SlabStep(nb, halo, plan, energy, cfg["dynamic_pruning"], num_parts)
    step.search_step()
    if getattr(args, "halo", "peer") == "peer":
        def gather(b):
            t = torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda()
            parts = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(parts, t)
            return [bytes(p.cpu().numpy().tobytes()) for p in parts]
        # all ranks use the same transport: fall back to NCCL if mapping the neighbour's memory fails anywhere
        ok = torch.ones(1, dtype=torch.int32, device="cuda")
        try:
            halo.enable_peer_memory(gather, rank, world)
        except Exception as exc:    # e.g. CUDA IPC not permitted in this container
            print("rank %d: peer-memory halo unavailable (%s), using NCCL" % (rank, exc), file=sys.stderr)
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0 and halo.peer:
            nb._check(lib.nbnxm_b200_peer_close(nb._h))
            halo.peer = False
        dist.barrier()
    local_stream = torch.cuda.ExternalStream(nb.streams()[0])
    flush = None if args.no_l2_flush else torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    nb.set_pair_counting(True)
    step(0)
    step.finish()
    nb.get_pair_count(LOCAL), nb.get_pair_count(NONLOCAL)
    for i in range(max(args.warmup, 4 * num_parts)):
        step(i)
    step.finish()
    nb.get_pair_count(LOCAL), nb.get_pair_count(NONLOCAL)
    step(0)
    step.finish()
    pairs = torch.tensor([nb.get_pair_count(LOCAL), nb.get_pair_count(NONLOCAL)], dtype=torch.float64, device="cuda")
    nb.set_pair_counting(False)
    dist.all_reduce(pairs)
    computed_pairs = float(pairs.sum().item())

    def timed_run(host_io):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        dist.barrier()
        torch.cuda.synchronize()
        l0 = nb.launch_count()
        for i in range(args.steps):
            if flush is not None:
                with torch.cuda.stream(local_stream):
                    flush.zero_()
            ev[i][0].record(local_stream)
            step(i, host_io)
            ev[i][1].record(local_stream)
        step.finish()
        torch.cuda.synchronize()
        dist.barrier()
        ms = sum(a.elapsed_time(b) for a, b in ev) / len(ev)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        n = torch.tensor([nb.launch_count() - l0], dtype=torch.float64, device="cuda")
        dist.all_reduce(n)
        return float(t.item()), int(n.item())

    from bench import ClockSampler, METRIC, UNIT
    for i in range(args.warmup):
        step(i)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ms_step, launches = timed_run(False)
    clock_rec = clocks.stop() if rank == 0 else None

    nb.gpu_reset_timings()
    nb.set_timing(True)
    halo.set_timing(True)
    for i in range(4):
        step(i)
        step.finish()
        halo.timings()
    t = nb.gpu_get_timings()
    hx, hf = halo.timings(reset=True)
    if halo.peer and halo.peer_error():
        raise RuntimeError("peer-memory halo: a wait on a neighbour's step counter timed out")
    nb.set_timing(False)
    halo.set_timing(False)
    e = 1 if energy else 0
    # the two force launches of a step run on two streams at the same time: report them per stream, never their sum
    k_all, n_all = t.force_ms[0][e], t.force_count[0][e]
    k_nl_ms = t.force_nonlocal_ms / max(1, t.force_nonlocal_count)
    k_loc_ms = (k_all - t.force_nonlocal_ms) / max(1, n_all - t.force_nonlocal_count)
    kt = torch.tensor([k_loc_ms, k_nl_ms, hx, hf], dtype=torch.float64, device="cuda")
    dist.all_reduce(kt, op=dist.ReduceOp.MAX)
    k_loc_ms, k_nl_ms, hx, hf = [float(v) for v in kt.tolist()]
    fp32_peak = measure_fp32_peak(local_rank)

    # end to end: coordinates up, forces down every step; with the peer-memory halo the step is chunk-pipelined like the
    # single-rank one (the plain copy - compute - copy sequence first, for the record)
    for i in range(args.warmup):
        step(i, True)
    ms_e2e_plain, _ = timed_run(True)
    e2e_pipeline = "none: copy, compute, copy"
    ms_e2e = ms_e2e_plain
    # one chunk per 250 k home atoms as in the single-rank step; below four chunks the plain sequence is as fast (measured:
    # 768 k atoms per rank, 3 chunks: 1.070 against 1.062 ms; 6.1 M atoms per rank, 16 chunks: 7.74 against 9.89 ms)
    nchunks = getattr(args, "e2e_chunks", 0) or max(1, min(24, plan.nbat.numLocalAtoms // 250000))
    if halo.peer and nchunks >= 4:
        from .pipeline import make_slab_chunk_plan
        if device_lists:
            box_x, ncx, ncy, first_bin = slab_columns_info
            sci_local = slab_local_sci
        else:
            box_x, ncx, ncy, first_bin = float(wl.grid.box[0]), wl.grid.ncx, wl.grid.ncy, wl.grid.first_bin_of_column
            sci_local = plan.local.sci
        step.chunk_plan = make_slab_chunk_plan(box_x, ncx, ncy, first_bin, world, rank, cfg["rlist_outer"], sci_local, nchunks)
        for i in range(args.warmup):
            step(i, True)
        ms_e2e, _ = timed_run(True)
        e2e_pipeline = ("%d chunks of grid columns per rank: H2D, local force kernels and D2H of different chunks overlap, the "
                        "non-local kernel runs against the neighbour's memory beside them (nbnxm_b200_do_force_step_pipelined)"
                        % step.chunk_plan.nchunks)
    sizes = torch.tensor([plan.nbat.numLocalAtoms, plan.recv_count, plan.send_count], dtype=torch.float64, device="cuda")
    dist.all_reduce(sizes)
    n_home, n_halo, _ = [int(v) for v in sizes.tolist()]

    parity = parity_vs_single_gpu(wl, plan, nb, halo, step, cfg, num_parts, rank, world, local_rank, args, dsearch)
    if dsearch is not None:
        dsearch.free()

    if rank == 0:
        # fraction of the N-GPU FP32 peak the whole step reaches (kernels of both streams, halo waits and launch gaps included)
        achieved = computed_pairs * wl.flops_per_pair / (ms_step * 1e-3) * 1e-12
        line = {
            "metric": METRIC, "value": wl.useful_pairs / (ms_step * 1e-3) * 1e-9, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "natoms": wl.box.natoms, "rc_nm": cfg["rc"], "vdw": cfg["vdw"],
                       "elec": "ewald_analytical", "energy_every_step": energy, "rlist_outer_nm": cfg["rlist_outer"],
                       "rlist_inner_nm": cfg["rlist_inner"], "rolling_prune_parts": num_parts,
                       "decomposition": "%d x-slabs, one-sided halo from the +x neighbour" % world,
                       "halo_atoms_total": n_halo, "home_atoms_total": n_home,
                       "l2": "256 MiB flush between steps, outside the per-step CUDA-event intervals" if flush is not None else "no flush",
                       "timing": "mean of per-step CUDA-event intervals on each rank's local stream, max over ranks"},
            "us_per_force_step": ms_step * 1e3,
            "search_step": search_rec,
            "computed_pairs_per_step": computed_pairs,
            "computed_gpairs_per_s": computed_pairs / (ms_step * 1e-3) * 1e-9,
            "gpu_launches": launches,
            "clocks": clock_rec,
            "halo": ({"transport": "peer memory over NVLink: the non-local kernel reads the +x neighbour's xq and reduces "
                                   "forces into its accumulator (CUDA IPC mapping, step counters in device memory), no transport calls",
                      "bytes_per_step_total": int(n_halo * 32)} if halo.peer else
                     {"x_exchange_wait_and_transfer_us": hx * 1e3, "f_exchange_wait_and_transfer_us": hf * 1e3,
                      "note": "CUDA events around the ncclGroup{Send, Recv} of the slowest rank: they include the wait for the "
                              "neighbour to reach its matching call (rank skew), not only the transfer",
                      "transport": "ncclSend/ncclRecv",
                      "bytes_per_step_total": int(n_halo * 32)}),
            "e2e": {"value": wl.useful_pairs / (ms_e2e * 1e-3) * 1e-9, "unit": UNIT, "ms_per_step": ms_e2e,
                    "pipeline": e2e_pipeline, "ms_per_step_copy_compute_copy": ms_e2e_plain,
                    "h2d_bytes_per_step": int(n_home * 16),
                    "d2h_bytes_per_step": int(n_home * 12 + (16 + 45 * 24 if energy else 0) * world)},
            "parity": parity,
            "roofline": {"bound": "fp32_fma", "achieved": achieved, "peak": fp32_peak * world, "unit": "TFLOP/s",
                         "frac": achieved / (fp32_peak * world), "traffic": None, "kernel": "nbnxm_force_kernel",
                         "local_kernel_us": k_loc_ms * 1e3, "nonlocal_kernel_us": k_nl_ms * 1e3, "flops_per_pair": wl.flops_per_pair,
                         "note": "N > 1: pairs of all ranks x flops per pair / ms_per_step against the measured FFMA peak x n_gpus, "
                                 "i.e. the whole step, not one kernel (the local and non-local launches overlap on two streams; "
                                 "their durations are the slowest rank's, per stream)"},
        }
        import __main__ as main_module
        emit_line = getattr(main_module, "emit", None)       # bench.py keeps the real stdout for this line
        if emit_line is None:
            from bench import emit as emit_line
        emit_line(line)
    torch.cuda.synchronize()
    dist.barrier()          # nobody unmaps or frees memory a neighbour's kernels may still touch
    halo_free = getattr(lib, "nbnxm_b200_halo_free")
    halo_free(nb._h)
    dist.barrier()
    nb.gpu_free()
    dist.barrier()
    dist.destroy_process_group()
