"""CPU test of the multi-GPU host logic (gromacs_b200/multigpu.py) with torch.distributed / gloo, world size 2:
each rank builds its slab plan (home + halo atoms, re-indexed local / non-local lists, contiguous send and
receive ranges), the halo exchange is emulated with gloo send/recv on the same ranges the NCCL path uses,
and the oracle walks each rank's lists.  Forces gathered over the ranks must equal the single-rank forces."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gromacs_b200.multigpu import make_slab_plan
        from gromacs_b200.workload import make_workload
        from oracle import oracle_py as O
        wl = make_workload("water48k_test", nthreads=2)
        plan = make_slab_plan(wl, rank, world, min_sci=200)
        nbat = plan.nbat
        p = O.OrcParams()
        for name, _ in wl.params._fields_:
            if hasattr(p, name):
                setattr(p, name, getattr(wl.params, name))
        p.ntypes = nbat.numTypes
        up, down = (rank + 1) % world, (rank - 1) % world
        # --- coordinate halo: our first columns to the -x neighbour, the +x neighbour's into our halo range
        xq = nbat.xq.copy()
        xq[plan.recv_first:plan.recv_first + plan.recv_count] = np.nan
        send = torch.from_numpy(xq[plan.send_first:plan.send_first + plan.send_count].copy())
        recv = torch.empty((plan.recv_count, 4), dtype=torch.float32)
        ops = [dist.P2POp(dist.isend, send, down), dist.P2POp(dist.irecv, recv, up)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        xq[plan.recv_first:plan.recv_first + plan.recv_count] = recv.numpy()
        assert np.array_equal(xq, nbat.xq), "halo coordinates must be the +x neighbour's first columns"
        # --- forces on the rank's atoms from its local and non-local lists
        f = np.zeros((nbat.numAtoms(), 3))
        e = np.zeros(2)
        npairs = 0
        for pl in (plan.local, plan.nonlocal_):
            fi, _, ei, n = O.forces(p, pl.sci, pl.cjPacked, pl.excl, xq, nbat.type, nbat.lj_comb, nbat.nbfp,
                                    nbat.nbfp_comb, nbat.shift_vec)
            f += fi
            e += ei
            npairs += n
        # --- force halo: halo forces back to their owner (+x), ours from -x added to our first columns
        fsend = torch.from_numpy(f[plan.recv_first:plan.recv_first + plan.recv_count].copy())
        frecv = torch.empty((plan.send_count, 3), dtype=torch.float64)
        ops = [dist.P2POp(dist.isend, fsend, up), dist.P2POp(dist.irecv, frecv, down)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        f[plan.send_first:plan.send_first + plan.send_count] += frecv.numpy()
        fh = torch.from_numpy(f[:nbat.numLocalAtoms].copy())
        parts = [None] * world
        dist.all_gather_object(parts, (plan.home_slice.start, fh.numpy(), e, npairs))
        if rank == 0:
            whole = wl.pairlist()
            g = wl.nbat
            f_ref, _, e_ref, n_ref = O.forces(p, whole.sci, whole.cjPacked, whole.excl, g.xq, g.type, g.lj_comb, g.nbfp,
                                              g.nbfp_comb, g.shift_vec)
            f_all = np.zeros_like(f_ref)
            e_all = np.zeros(2)
            for start, fpart, epart, _ in parts:
                f_all[start:start + fpart.shape[0]] += fpart
                e_all += epart
            err = float(np.sqrt(((f_all - f_ref) ** 2).sum() / (f_ref ** 2).sum()))
            out.put((err, float(abs(e_all[1] - e_ref[1]) / abs(e_ref[1])), int(plan.send_count), int(plan.recv_count)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_slab_plans_and_halo_ranges_gloo(oracle, world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(300)
        assert pr.exitcode == 0
    err, e_err, nsend, nrecv = out.get(timeout=10)
    assert nsend > 0 and nrecv > 0
    # (x + shift) is rounded to float per periodic image, so a pair seen from the other slab differs by ~1e-7
    assert err < 1e-6, err
    assert e_err < 1e-6, e_err


def test_reindexed_lists_stay_inside_the_rank_atoms(oracle):
    from gromacs_b200.multigpu import make_slab_plan
    from gromacs_b200.workload import make_workload
    wl = make_workload("water48k_test", nthreads=2)
    total_home = 0
    for r in range(3):
        plan = make_slab_plan(wl, r, 3)
        ncl = plan.nbat.numAtoms() // 8
        for pl in (plan.local, plan.nonlocal_):
            assert pl.sci[:, 0].min() >= 0 and pl.sci[:, 0].max() < plan.nbat.numLocalAtoms // 64
            assert pl.cjPacked[:, :4].max() < ncl
        # local lists only touch home atoms, non-local j-clusters are all in the halo
        used = (plan.nonlocal_.cjPacked[:, 4] | plan.nonlocal_.cjPacked[:, 6]) != 0
        m = plan.nonlocal_.cjPacked[used]
        for jm in range(4):
            sel = ((m[:, 4] | m[:, 6]) >> (8 * jm)) & 0xff != 0
            assert (m[sel, jm] >= plan.nbat.numLocalAtoms // 8).all()
        assert plan.send_count <= plan.nbat.numLocalAtoms and plan.recv_first == plan.nbat.numLocalAtoms
        total_home += plan.nbat.numLocalAtoms
    assert total_home == wl.nbat.numAtoms()
