#!/usr/bin/env python
"""Profiling aid: run ONE slab of an N-slab decomposition alone on one GPU (halo coordinates resident, no transport),
to see what the local and non-local kernels of a rank cost without the exchange.
  python profiles/tools/slab_alone.py --workload water1536k --world 8 --rank 3 --steps 20"""
import argparse, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from gromacs_b200 import LOCAL, NONLOCAL, NbnxmGpu
from gromacs_b200.multigpu import SlabStep, make_slab_plan
from gromacs_b200.workload import make_workload

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="water1536k")
ap.add_argument("--world", type=int, default=8)
ap.add_argument("--rank", type=int, default=3)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--min-sci", type=int, default=0)
a = ap.parse_args()
wl = make_workload(a.workload, nslabs=a.world)
nb = NbnxmGpu(wl.params, wl.nbat, device=0, bLocalAndNonlocal=True)
plan = make_slab_plan(wl, a.rank, a.world, min_sci=a.min_sci or nb.gpu_min_ci_balanced())
step = SlabStep(nb, None, plan, wl.cfg["energy"], wl.cfg["dynamic_pruning"], 3)
step.search_step()
for i in range(12):
    step(i)
step.finish()
st = torch.cuda.ExternalStream(nb.streams()[0])
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
torch.cuda.synchronize()
for i in range(a.steps):
    ev[i][0].record(st)
    step(i)
    ev[i][1].record(st)
step.finish()
torch.cuda.synchronize()
ms = sorted(x.elapsed_time(y) for x, y in ev)
nb.gpu_reset_timings(); nb.set_timing(True)
for i in range(8):
    step(i); step.finish()
t = nb.gpu_get_timings(); nb.set_timing(False)
e = 1 if wl.cfg["energy"] else 0
print(json.dumps({"workload": a.workload, "world": a.world, "rank": a.rank, "step_us_median": ms[len(ms) // 2] * 1e3,
                  "step_us_min": ms[0] * 1e3, "nsci_local": int(plan.local.sci.shape[0]), "nsci_nonlocal": int(plan.nonlocal_.sci.shape[0]),
                  "force_us_avg_per_launch": t.force_ms[0][e] / max(1, t.force_count[0][e]) * 1e3,
                  "home_atoms": int(plan.nbat.numLocalAtoms), "halo_atoms": int(plan.recv_count)}))
