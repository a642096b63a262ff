#!/usr/bin/env python
"""bench.py — NBNXM short-range nonbonded force step on B200 (driver contract in the task prompt).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

A "step" is one pass of the hot path over the synthetic water box named in config.workload, the nonbonded part of
do_force in the reference's order (nbnxm_b200_do_force_step): clear outputs -> force(+energy) kernel on the
dynamically pruned list -> rolling prune on the reference's schedule (every 2nd step, numParts = nstlistPrune/2)
-> pack forces (float4 -> float3), with coordinates already resident in HBM.  `value` = useful pair interactions
per second, the quantity `gmx nonbonded-benchmark` prints (N/2 (rho 4/3 pi rc^3 + 1) pairs per step,
src/gromacs/nbnxm/benchmark/bench_setup.cpp:332-337).  `e2e` is the same metric through the public API with host
buffers (H2D of xq and D2H of forces + energies inside the timed region).  `roofline` reports the force kernel
alone against the measured FP32-FMA peak using the reference's flop model (src/gromacs/gmxlib/nrnb.cpp:92-112)
on the pairs the kernel actually evaluates.

The default workload is the same for every N, so that the driver's 1/2/4/8-GPU lines are one strong-scaling
series: water12m, BASELINE.json's configs[4] (12.3 M atoms, rc 1.2 nm, dynamic + rolling pruning), the box the
north-star targets (kernel fraction of FP32 peak at 1 GPU, parallel efficiency at 8) are stated on; it fits one
GPU.  The other configurations (--workload water96k_fswitch = configs[1], water384k_*, water1536k, bench3k) are
recorded in profiles/ and DESIGN.md.  N > 1: x-slabs, one process per GPU, peer-memory halo over NVLink by
default (--halo nccl for the ncclSend/ncclRecv exchange).

--impl reference times the UNMODIFIED reference's CPU SIMD kernel (oracle/_ref/bench_ref, linked against
the reference's own libgromacs) on the host cores for the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "useful_pair_interactions_per_s"
UNIT = "Gpairs/s"
DEFAULT_WORKLOAD = "water12m"
REF_VDW = {"cut": "cut", "fswitch": "fswitch", "pswitch": "pswitch", "ljpme": "ljpme"}


_JSON_FD = None


def claim_stdout():
    """Keep stdout for the ONE JSON line: everything else that libraries write to file descriptor 1 (NCCL prints its
    version there) goes to stderr."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def host_has_avx512():
    try:
        with open("/proc/cpuinfo") as fh:
            flags = fh.read()
        return all(f in flags for f in (" avx512f", " avx512bw", " avx512vl", " avx512dq", " avx512cd"))
    except Exception:
        return False


def run_reference_cpu(cfg, natoms_k, budget_s=20.0, threads=None):
    """Time the reference's SIMD kernel on a bounded number of iterations (about budget_s of CPU work).  Two builds of
    the unmodified reference may be present (oracle/ref_harness/build_ref.sh): AVX2_256 and, where the host has
    AVX-512, AVX_512; a short probe picks the faster one, which is then timed."""
    threads = threads or host_cores()
    exes = [("AVX2_256", os.path.join(ROOT, "oracle", "_ref", "bench_ref"))]
    if host_has_avx512():
        exes.append(("AVX_512", os.path.join(ROOT, "oracle", "_ref", "bench_ref_avx512")))
    exes = [(n, e) for n, e in exes if os.path.exists(e)]
    if not exes:
        return None
    args = ["--size", str(natoms_k), "--rc", str(cfg["rc"]), "--vdw", REF_VDW[cfg["vdw"]],
            "--energy", "1" if cfg["energy"] else "0", "--nt", str(threads)]
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="close", OMP_PLACES="cores")

    def call(exe, kernel, iters, warm, size=natoms_k):
        a = list(args)
        a[1] = str(size)
        out = subprocess.run([exe] + a + ["--kernel", kernel, "--iter", str(iters), "--warmup", str(warm)], env=env,
                             check=True, capture_output=True, text=True).stdout.strip().splitlines()[-1]
        return json.loads(out)
    # the reference's two SIMD kernel layouts (Cpu4xN_Simd_4xN, Cpu4xN_Simd_2xNN) in every build present, ranked on a
    # 96 k-atom box of the same flavor (seconds), so that the full-size system is set up only for the one that is timed
    probes = []
    for name, exe in exes:
        for kernel in ("4xm", "2xmm"):
            try:
                probes.append((call(exe, kernel, 12, 3, size=min(natoms_k, 32))["sec_per_iter"], name + " " + kernel, exe, kernel))
            except Exception:      # a layout this SIMD width does not have, an instruction set the host lacks after all
                pass
    if not probes:
        return None
    _, name, exe, kernel = min(probes)
    sec = call(exe, kernel, 2, 1)["sec_per_iter"]
    iters = int(max(3, min(2000, budget_s / max(sec, 1e-6))))
    res = call(exe, kernel, iters, 2)
    res["threads"] = threads
    res["simd"] = name
    res["simd_probed"] = {n: round(s, 6) for s, n, _, _ in probes}
    return res


def run_port_cpu(wl, plist, budget_s=15.0, threads=None):
    """Fallback CPU baseline: the oracle's float32 OpenMP port on a slice of the sci entries."""
    import numpy as np
    from oracle import oracle_py as O
    threads = threads or host_cores()
    p = O.OrcParams()
    for name, _ in wl.params._fields_:
        if hasattr(p, name):
            setattr(p, name, getattr(wl.params, name))
    p.ntypes = wl.nbat.numTypes
    nsci = plist.sci.shape[0]
    sample = max(1, min(nsci, 64 * threads))
    while True:
        t0 = time.perf_counter()
        _, _, n = O.forces_f32_omp(p, plist.sci[:sample], plist.cjPacked, plist.excl, wl.nbat.xq, wl.nbat.type,
                                   wl.nbat.lj_comb, wl.nbat.nbfp, wl.nbat.nbfp_comb, wl.nbat.shift_vec,
                                   calc_energy=wl.cfg["energy"], nthreads=threads)
        dt = time.perf_counter() - t0
        if dt > 0.3 * budget_s or sample == nsci:
            break
        sample = min(nsci, sample * 4)
    return dict(sec=dt, computed_pairs=n, sample_sci=sample, nsci=nsci, threads=threads)


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML every few ms DURING the timed region (the same
    counters `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*` prints, B200_PROFILING.md)."""

    def __init__(self, index, period_s=0.002):
        self.index, self.period, self.rows, self.thread, self._stop = index, period_s, [], None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            uuid_index = index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                uuid_index = int(vis.split(",")[index]) if vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(uuid_index)
        except Exception:
            self.nv = None

    def start(self):
        if self.nv is None:
            return
        self._stop = False
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        nv = self.nv
        while not self._stop:
            try:
                self.rows.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM),
                                  nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                                  if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons")
                                  else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h),
                                  nv.nvmlDeviceGetPowerUsage(self.h) * 1e-3))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self._stop = True
        if self.thread:
            self.thread.join(timeout=2)
        nv = self.nv
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        sm = sorted(r[0] for r in self.rows)
        reasons = sorted(n for n, b in bits.items() if any(r[1] & b for r in self.rows))
        try:
            smax = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:
            smax = None
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": reasons,
                "samples": len(sm), "power_w_max": max((r[2] for r in self.rows), default=None)}


def reference_arm(args):
    """--impl reference: the reference's own CPU kernels on the host cores, rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from gromacs_b200.workload import CONFIGS
    import math
    cfg = CONFIGS[args.workload]
    res = run_reference_cpu(cfg, cfg["k"], budget_s=max(5.0, min(60.0, 6.0 * args.steps)))
    if res is None:
        # oracle/_ref was not built in this container (it needs the reference sources and its CMake build): time the
        # oracle's float32 OpenMP restatement of the same kernel on a bounded slice of the same list instead
        from gromacs_b200.workload import make_workload
        wl = make_workload(args.workload)
        plist = wl.pairlist(min_sci=0)
        import numpy as np
        cjp = np.ascontiguousarray(plist.cjPacked).view(np.uint32).reshape(-1, 8)
        list_pairs = 32 * int(np.unpackbits(np.ascontiguousarray(cjp[:, [4, 6]]).view(np.uint8)).sum())
        r = run_port_cpu(wl, plist, budget_s=max(5.0, min(60.0, 6.0 * args.steps)))
        value = wl.useful_pairs * (r["computed_pairs"] / max(1, list_pairs)) / r["sec"] * 1e-9
        emit({
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": 1, "warmup": 0,
            "ms_per_step": r["sec"] * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": args.workload, "natoms": wl.box.natoms, "rc_nm": cfg["rc"], "vdw": cfg["vdw"],
                       "elec": "ewald_analytical", "energy_every_step": cfg["energy"],
                       "note": "oracle port (oracle/nbnxm_oracle.c, float32, OpenMP) on the GPU-layout list: oracle/_ref is not built here"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["threads"], "kind": "port",
                             "sample": "%d of %d sci entries of the outer list" % (r["sample_sci"], r["nsci"])},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return
    value = res["useful_pairs"] / res["sec_per_iter"] * 1e-9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": res["iters"],
        "warmup": 2, "ms_per_step": res["sec_per_iter"] * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "natoms": int(res["natoms"]), "rc_nm": cfg["rc"], "vdw": cfg["vdw"],
                   "elec": "ewald_analytical", "energy_every_step": cfg["energy"],
                   "note": "reference SIMD kernel (%s; s/iteration of its builds / layouts on a 96 k-atom probe: %s), its own CPU pair list with rlist = rc"
                           % (res["simd"], json.dumps(res["simd_probed"]))},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": res["threads"], "kind": "reference",
                         "sample": "%d iterations of the full %d-atom system" % (res["iters"], int(res["natoms"]))},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def parity_sample(nb, wl, plist, sw):
    """Part of the cpu_baseline leg (the one place bench.py uses oracle/, as the checker): a sample of the sci entries of
    the PRUNED production list - every n-th entry, about 600 of them - through the double-precision oracle and through
    the CUDA path (same handle, same kernels), forces and energies compared.  tests/test_gpu_benched_configs.py holds
    the full-size checks; this puts a correctness figure on the bench line itself."""
    import numpy as np
    from oracle import oracle_py as O
    from gromacs_b200 import LOCAL, PairlistGpu, StepWorkload
    O.build()
    cj_pruned, _, _, _, _ = nb.download_pairlist()
    stride = max(1, plist.sci.shape[0] // 600)
    sub = PairlistGpu(sci=plist.sci[::stride], cjPacked=cj_pruned, excl=plist.excl)
    g = wl.nbat
    po = O.OrcParams()
    for name, _ in wl.params._fields_:
        if hasattr(po, name):
            setattr(po, name, getattr(wl.params, name))
    po.ntypes = g.numTypes
    f_ref, _, e_ref, npairs = O.forces(po, sub.sci, sub.cjPacked, sub.excl, g.xq, g.type, g.lj_comb, g.nbfp, g.nbfp_comb,
                                       g.shift_vec)
    swe = StepWorkload(computeEnergy=True, computeVirial=True, useGpuFBufferOps=False)
    nb.gpu_init_pairlist(sub, LOCAL)
    nb.do_force_step(0, swe, have_halo=False, dynamic_pruning=bool(wl.cfg["dynamic_pruning"]), num_parts=1, xq_host=g.xq, f_host=g.f)
    e_lj, e_el = nb.gpu_wait_finish_task(swe, LOCAL)
    f = np.asarray(g.f, np.float64)
    return {"vs_oracle_sample": {"sci_entries": int(sub.sci.shape[0]), "of": int(plist.sci.shape[0]), "pairs_in_range": int(npairs),
                                 "f_relrms": float(np.sqrt(((f - f_ref) ** 2).sum() / (f_ref ** 2).sum())),
                                 "f_maxcomp_rel": float(np.abs(f - f_ref).max() / np.abs(f_ref).max()),
                                 "e_lj_rel": float(abs(e_lj - e_ref[0]) / abs(e_ref[0])),
                                 "e_el_rel": float(abs(e_el - e_ref[1]) / abs(e_ref[1]))},
            "tolerance": {"f_relrms": 5e-6, "f_maxcomp_rel": 1e-4, "e_rel": 1e-6}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--energy", type=int, default=None, choices=[0, 1],
                    help="N = 1: force + energy (+ virial) kernel on every step (1) or never (0) instead of what the workload says")
    ap.add_argument("--no-l2-flush", action="store_true")
    ap.add_argument("--min-sci", type=int, default=0, help="override gpu_min_ci_balanced (list splitting target)")
    ap.add_argument("--e2e-chunks", type=int, default=0,
                    help="N = 1: chunks of the pipelined end-to-end step (1 = plain copy-compute-copy sequence; "
                         "0 = one chunk per 250k atoms, at most 32: measured 22.0 / 17.9 / 15.9 / 14.6 / 14.4 ms per step "
                         "with 1 / 8 / 16 / 24 / 32 chunks on the 12.3 M-atom box)")
    ap.add_argument("--slab-lists", default="device", choices=["device", "host"],
                    help="N > 1: where the search step of every rank runs (device: gridding and both lists on the rank's GPU)")
    ap.add_argument("--halo", default="peer", choices=["peer", "nccl"],
                    help="N > 1: peer-memory halo over NVLink (no transport calls) or ncclSend/ncclRecv")
    args = ap.parse_args()
    claim_stdout()
    if args.workload is None:
        args.workload = DEFAULT_WORKLOAD
    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
    import torch
    from gromacs_b200 import LOCAL, NbnxmGpu, StepWorkload
    from gromacs_b200.nbnxm import measure_fp32_peak
    from gromacs_b200.workload import make_workload, rolling_prune_parts

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        from gromacs_b200.multigpu import bench_multi_gpu
        return bench_multi_gpu(args, rank, world, local_rank)
    if args.gpus != 1:
        sys.exit("launch multi-GPU runs with torch.distributed.run (one rank per GPU)")

    torch.cuda.set_device(local_rank)
    wl = make_workload(args.workload, energy=None if args.energy is None else bool(args.energy))
    cfg, nbat = wl.cfg, wl.nbat
    energy = cfg["energy"]
    sw = StepWorkload(computeEnergy=energy, computeVirial=energy, useGpuFBufferOps=True)
    # pinned host buffers, like the reference's HostAllocationPolicy
    xq_pin = torch.empty((nbat.numAtoms(), 4), dtype=torch.float32).pin_memory()
    xq_pin.numpy()[:] = nbat.xq
    nbat.xq = xq_pin.numpy()
    f_pin = torch.zeros((nbat.numAtoms(), 3), dtype=torch.float32).pin_memory()
    nbat.f = f_pin.numpy()

    nb = NbnxmGpu(wl.params, nbat, device=local_rank)
    min_sci = args.min_sci or nb.gpu_min_ci_balanced()
    plist = wl.pairlist(min_sci=min_sci)
    # end-to-end path: coordinates go up and forces come down in chunks of grid columns, pipelined against the kernel
    # (nbnxm_b200_do_force_step_pipelined); the list is the same, with its sci entries grouped by chunk
    from gromacs_b200.pipeline import make_chunk_plan
    nchunks = args.e2e_chunks if args.e2e_chunks > 0 else max(1, min(32, wl.box.natoms // 250000))
    chunks = make_chunk_plan(wl.grid, plist, nchunks) if nchunks > 1 else None
    if chunks is not None:
        plist = chunks.plist
    nb.gpu_init_atomdata(nbat)
    nb.gpu_init_pairlist(plist, LOCAL)
    nb.setupGpuShortRangeWork(LOCAL)
    nb.gpu_upload_shiftvec(nbat)
    nb.gpu_copy_xq_to_gpu(nbat, LOCAL)
    local_stream = torch.cuda.ExternalStream(nb.streams()[0])
    num_parts = rolling_prune_parts(cfg)   # numRollingPruningParts = nstlistPrune / 2 (pairlist_tuning.cpp:685)
    flush = None if args.no_l2_flush else torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def step(i, host_io):
        # the do_force sequence: [H2D xq] -> clear outputs -> force(+energy) kernel -> rolling prune on odd steps
        # (isDynamicPruningStepGpu, pairlistsets.h:108-115) -> f4 -> f3 [-> D2H f, energies], one foreign call
        sw.useGpuFBufferOps = not host_io
        if host_io and chunks is not None:
            nb.do_force_step_pipelined(i, sw, chunks, nbat.xq, nbat.f, dynamic_pruning=cfg["dynamic_pruning"], num_parts=num_parts)
        else:
            nb.do_force_step(i, sw, have_halo=False, dynamic_pruning=cfg["dynamic_pruning"], num_parts=num_parts,
                             xq_host=nbat.xq if host_io else None, f_host=nbat.f if host_io else None)
        if host_io:
            return nb.gpu_wait_finish_task(sw, LOCAL)

    # search step + first-pass prune + a full rolling cycle, untimed (the first-pass prune is timed by the library for
    # `value_with_search` below)
    outer_list_pairs = 32 * int(np.unpackbits(np.ascontiguousarray(plist.cjPacked[:, [4, 6]]).view(np.uint8)).sum(dtype=np.int64))
    nb.set_pair_counting(True)
    nb.gpu_reset_timings()
    nb.set_timing(True)
    step(0, False)
    nb.gpu_wait_finish_task(sw, LOCAL)
    t0 = nb.gpu_get_timings()
    first_prune_ms = t0.prune_ms / max(1, t0.prune_count)
    nb.set_timing(False)
    nb.gpu_reset_timings()
    pairs_first = nb.get_pair_count(LOCAL)
    for i in range(max(args.warmup, 2 * num_parts)):
        step(i, False)
    nb.gpu_wait_finish_task(sw, LOCAL)
    nb.get_pair_count(LOCAL)
    step(0, False)
    nb.gpu_wait_finish_task(sw, LOCAL)
    computed_pairs = nb.get_pair_count(LOCAL)       # pairs one force launch evaluates on the pruned list
    nb.set_pair_counting(False)

    def timed_run(host_io):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        torch.cuda.synchronize()
        l0 = nb.launch_count()
        t0 = time.perf_counter()
        for i in range(args.steps):
            if flush is not None:
                with torch.cuda.stream(local_stream):
                    flush.zero_()
            ev[i][0].record(local_stream)
            step(i, host_io)
            ev[i][1].record(local_stream)
        nb.gpu_wait_finish_task(sw, LOCAL)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = [a.elapsed_time(b) for a, b in ev]
        return sum(ms) / len(ms), nb.launch_count() - l0, wall

    for i in range(args.warmup):
        step(i, False)
    clocks = ClockSampler(local_rank)
    clocks.start()
    ms_step, launches, _ = timed_run(False)
    clock_rec = clocks.stop()

    # force kernel alone (dominant kernel): library-side CUDA events around each launch
    nb.gpu_reset_timings()
    nb.set_timing(True)
    timed_run(False)
    t = nb.gpu_get_timings()
    nb.set_timing(False)
    k_ms = t.force_ms[0][1 if energy else 0] / max(1, t.force_count[0][1 if energy else 0])
    prune_ms = t.rolling_prune_ms / max(1, t.rolling_prune_count)
    fp32_peak = measure_fp32_peak(local_rank)
    achieved = computed_pairs * wl.flops_per_pair / (k_ms * 1e-3) * 1e-12
    # dram__bytes_read.sum + dram__bytes_write.sum of one force-kernel launch from the committed ncu --set full captures
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "force_kernel_dram_traffic.json")) as fh:
            traffic = json.load(fh).get(args.workload, {}).get("bytes_per_launch")
    except Exception:
        pass

    # end to end through the public API with host buffers
    for i in range(args.warmup):
        step(i, True)
    ms_e2e, _, _ = timed_run(True)

    # the search step on either side of the path, untimed above: grid and list built on the device
    # (nbnxm_b200_gpu_search_put_atoms_on_grid / _build) next to the host gridder / builder + upload they replace.  Run in
    # a process of its own (profiles/tools/search_profile.py, the script behind the numbers in DESIGN.md 4.4), so that
    # nothing it does can touch the force-step measurement above.
    try:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "tools", "search_profile.py"), args.workload, "3",
                              str(min_sci)], capture_output=True, text=True, timeout=600)
        recs = [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]
        lst = next(r for r in recs if "gpu_build_ms" in r)
        dev = next(r for r in recs if r.get("device_search_step"))
        search_rec = {"gpu_grid_ms": min(dev["gpu_grid_ms"][1:]), "gpu_list_ms": min(dev["gpu_list_ms"][1:]),
                      "host_grid_s": dev["host_grid_s"], "host_list_s": lst["host_build_s"], "host_threads": lst["host_threads"],
                      "host_list_upload_s": lst["host_list_upload_s"], "list_bytes_not_uploaded": lst["list_bytes"],
                      "same_grid_order_as_host": dev["same_order_as_host"],
                      "same_list_sizes_as_host": bool(lst["same_sizes_as_host"] and dev["same_sizes_as_host"]),
                      "same_list_entries_as_host": lst.get("same_entries_as_host")}
    except Exception as e:      # reported, never fatal for the force-step measurement
        search_rec = {"error": str(e)[:200]}

    value = wl.useful_pairs / (ms_step * 1e-3) * 1e-9
    # the search step amortised over the list's lifetime (nstlist 100): device gridding + device list build + first-pass
    # prune of the fresh list once per 100 force steps
    nstlist = 100
    if "gpu_grid_ms" in search_rec:
        search_ms = search_rec["gpu_grid_ms"] + search_rec["gpu_list_ms"] + first_prune_ms
        search_rec["first_pass_prune_ms"] = first_prune_ms
        search_rec["nstlist"] = nstlist
        search_rec["ms_per_step_with_search"] = (nstlist * ms_step + search_ms) / nstlist
        value_with_search = wl.useful_pairs / (search_rec["ms_per_step_with_search"] * 1e-3) * 1e-9
    else:
        value_with_search = None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": args.workload, "natoms": wl.box.natoms, "rc_nm": cfg["rc"], "vdw": cfg["vdw"],
                   "elec": "ewald_analytical", "energy_every_step": energy, "rlist_outer_nm": cfg["rlist_outer"],
                   "rlist_inner_nm": cfg["rlist_inner"], "rolling_prune_parts": num_parts,
                   "nsci": int(plist.sci.shape[0]), "ncj_packed": int(plist.cjPacked.shape[0]),
                   "l2": "256 MiB flush between steps, outside the per-step CUDA-event intervals" if flush is not None else "no flush",
                   "timing": "mean of per-step CUDA-event intervals on the library's local stream"},
        "us_per_force_step": ms_step * 1e3,
        "value_with_search": value_with_search,
        "computed_pairs_per_step": computed_pairs,
        "computed_gpairs_per_s": computed_pairs / (ms_step * 1e-3) * 1e-9,
        "outer_list_pairs": outer_list_pairs,
        "pairs_after_first_pass_prune": pairs_first,
        "gpu_launches": launches,
        "search_step": search_rec,
        "clocks": clock_rec,
        "e2e": {"value": wl.useful_pairs / (ms_e2e * 1e-3) * 1e-9, "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(nbat.numAtoms() * 16),
                "d2h_bytes_per_step": int(nbat.numAtoms() * 12 + (16 + 45 * 24 if energy else 0)),
                "pipeline": ("%d chunks of grid columns: H2D, force kernel and D2H of different chunks overlap "
                             "(nbnxm_b200_do_force_step_pipelined)" % chunks.nchunks) if chunks is not None else "none"},
        "roofline": {"bound": "fp32_fma", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                     "frac": achieved / fp32_peak, "traffic": traffic,
                     "kernel": "nbnxm_force_kernel", "kernel_us": k_ms * 1e3, "rolling_prune_us": prune_ms * 1e3,
                     "flops_per_pair": wl.flops_per_pair,
                     "peak_source": "measured live: pure-FFMA kernel (nbnxm_b200_measure_fp32_peak); nominal 148 SM x 128 x 2 x 1.965 GHz = 74.45"},
    }
    if not args.no_cpu_baseline:
        line["parity"] = parity_sample(nb, wl, plist, sw)
        res = run_reference_cpu(cfg, cfg["k"], budget_s=15.0)
        if res is not None:
            line["cpu_baseline"] = {"value": res["useful_pairs"] / res["sec_per_iter"] * 1e-9, "unit": UNIT,
                                    "cores": res["threads"], "kind": "reference",
                                    "sample": "%d iterations of the reference SIMD kernel (%s, the fastest of its builds / layouts here) on the full system" % (res["iters"], res["simd"])}
        else:
            r = run_port_cpu(wl, plist)
            frac = r["computed_pairs"] / max(1, pairs_first)
            line["cpu_baseline"] = {"value": wl.useful_pairs * frac / r["sec"] * 1e-9, "unit": UNIT, "cores": r["threads"],
                                    "kind": "port", "sample": "%d of %d sci entries" % (r["sample_sci"], r["nsci"])}
    nb.gpu_free()
    emit(line)


if __name__ == "__main__":
    main()
