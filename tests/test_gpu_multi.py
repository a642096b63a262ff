"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): the x-slab step as bench.py runs it - dynamic pruning
with the rolling prune alternating between the local and the non-local list (prunekerneldispatch.cpp:123-128), the
non-local prune reading halo coordinates - with the library's NCCL halo exchange and with its peer-memory halo (no
transport), one process per GPU, 2 / 4 / 8 ranks, over 8 steps with moving atoms, against the oracle on the whole system."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _workload_name(world):
    """slabs must be at least one halo wide: 12.4 nm of x for up to 4 ranks, 24.9 nm for 8"""
    return "water48k_test" if world <= 4 else "water384k_test"


def _worker(rank, world, port, out, peer):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from gromacs_b200 import NbnxmGpu
        from gromacs_b200.multigpu import HaloExchange, SlabStep, make_slab_plan
        from gromacs_b200.nbnxm import load_library
        from gromacs_b200.workload import make_workload
        lib = load_library()
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(HaloExchange.unique_id(lib)), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        wl = make_workload(_workload_name(world), nthreads=4, nslabs=world)
        _dynamic_pruning_params(wl)
        nb = NbnxmGpu(wl.params, wl.nbat, device=rank, bLocalAndNonlocal=True)
        halo = HaloExchange(nb, idt.cpu().numpy().tobytes(), rank, world)
        plan = make_slab_plan(wl, rank, world, min_sci=2000)
        # the halo part of the host coordinates is never uploaded: it has to arrive through the exchange
        plan.nbat.xq[plan.recv_first:] = 0
        step = SlabStep(nb, halo, plan, energy=True, dynamic_pruning=True, num_parts=3)
        step.search_step()
        if peer:
            # no transport: the non-local kernel reads the neighbour's coordinates / adds to its forces over NVLink
            def gather(b):
                parts = [None] * world
                dist.all_gather_object(parts, b)
                return parts
            halo.enable_peer_memory(gather, rank, world)
            dist.barrier()
        if peer == "pipelined":
            # coordinates up and forces down in chunks of grid columns, overlapped with the kernels (the bench's e2e path)
            from gromacs_b200.pipeline import make_slab_chunk_plan
            g = wl.grid
            step.chunk_plan = make_slab_chunk_plan(g.box[0], g.ncx, g.ncy, g.first_bin_of_column, world, rank, wl.cfg["rlist_outer"],
                                                   plan.local.sci, 6)
            assert step.chunk_plan.nchunks > 1
        results = []
        nloc = plan.nbat.numLocalAtoms
        for i, disp in enumerate(_displacements(wl.nbat.numAtoms())):
            # every rank moves its home atoms; halo coordinates only ever arrive through the exchange
            plan.nbat.xq[:nloc, :3] += disp[plan.home_slice]
            e_lj, e_el = step(i, host_io=True)
            results.append((plan.nbat.f[:nloc].astype(np.float64).copy(), e_lj, e_el))
        assert halo.peer_error() == 0
        parts = [None] * world
        dist.all_gather_object(parts, (plan.home_slice.start, results))
        dist.barrier()
        lib.nbnxm_b200_halo_free(nb._h)
        nb.gpu_free()
        if rank == 0:
            out.put(parts)
    finally:
        dist.destroy_process_group()


NSTEPS = 8


def _displacements(natoms):
    """the same small random walk on every rank and in the checker, in the global grid order"""
    rng = np.random.default_rng(2027)
    for _ in range(NSTEPS):
        yield rng.normal(0.0, 3e-4, (natoms, 3)).astype(np.float32)


def _dynamic_pruning_params(wl):
    """rlistInner 0.905 inside the 0.95 nm outer list, as the single-GPU tests use it"""
    import copy
    wl.params = copy.copy(wl.params)
    wl.params.use_dynamic_pruning = 1
    wl.params.rlist_inner_sq = np.float32(0.905 ** 2)


@pytest.mark.parametrize("peer", [False, True, "pipelined"], ids=["nccl", "peer", "peer-pipelined"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_x_slab_step_matches_oracle(oracle, world, peer):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    from gromacs_b200.workload import make_workload
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out, peer)) for r in range(world)]
    for pr in procs:
        pr.start()
    parts = out.get(timeout=600)
    for pr in procs:
        pr.join(120)
        assert pr.exitcode == 0
    wl = make_workload(_workload_name(world), nthreads=4, nslabs=world)
    p = oracle.OrcParams()
    for name, _ in wl.params._fields_:
        if hasattr(p, name):
            setattr(p, name, getattr(wl.params, name))
    p.ntypes = wl.nbat.numTypes
    whole, g = wl.pairlist(), wl.nbat
    xq = g.xq.copy()
    for i, disp in enumerate(_displacements(g.numAtoms())):
        xq[:, :3] += disp
        # list step, mid-cycle, and after local and non-local lists were both rolled over (the big box: last step only)
        if i not in ((0, 3, NSTEPS - 1) if world <= 4 else (NSTEPS - 1,)):
            continue
        f_ref, _, e_ref, _ = oracle.forces(p, whole.sci, whole.cjPacked, whole.excl, xq, g.type, g.lj_comb, g.nbfp,
                                           g.nbfp_comb, g.shift_vec)
        f = np.zeros_like(f_ref)
        e = np.zeros(2)
        for start, results in parts:
            fpart, e_lj, e_el = results[i]
            f[start:start + fpart.shape[0]] += fpart
            e += (e_lj, e_el)
        assert np.sqrt(((f - f_ref) ** 2).sum() / (f_ref ** 2).sum()) <= 5e-6, "step %d" % i
        assert np.abs(f - f_ref).max() <= 1e-4 * np.abs(f_ref).max()
        assert abs(e[1] - e_ref[1]) <= 1e-6 * abs(e_ref[1]), (i, e, e_ref)
        assert abs(e[0] - e_ref[0]) <= 1e-6 * abs(e_ref[0]), (i, e, e_ref)
