"""Timeline of the chunk-pipelined end-to-end step (nbnxm_b200_do_force_step_pipelined) of a workload: when each chunk's
coordinates are up, when its force kernel starts and ends, when its forces are down - CUDA timing events recorded by the
library (nbnxm_b200_set_pipeline_timeline).  usage: python profiles/tools/pipeline_timeline.py [workload] [nchunks]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gromacs_b200 import LOCAL, NbnxmGpu, StepWorkload  # noqa: E402
from gromacs_b200.pipeline import make_chunk_plan  # noqa: E402
from gromacs_b200.workload import make_workload, rolling_prune_parts  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "water12m"
nchunks = int(sys.argv[2]) if len(sys.argv) > 2 else 24
wl = make_workload(name)
cfg, nbat = wl.cfg, wl.nbat
xq_pin = torch.empty((nbat.numAtoms(), 4), dtype=torch.float32).pin_memory()
xq_pin.numpy()[:] = nbat.xq
nbat.xq = xq_pin.numpy()
f_pin = torch.zeros((nbat.numAtoms(), 3), dtype=torch.float32).pin_memory()
nbat.f = f_pin.numpy()
nb = NbnxmGpu(wl.params, nbat)
plan = make_chunk_plan(wl.grid, wl.pairlist(min_sci=nb.gpu_min_ci_balanced()), nchunks)
nb.gpu_init_atomdata(nbat)
nb.gpu_init_pairlist(plan.plist, LOCAL)
nb.setupGpuShortRangeWork(LOCAL)
nb.gpu_upload_shiftvec(nbat)
sw = StepWorkload(useGpuFBufferOps=False)
parts = rolling_prune_parts(cfg)
for i in range(2 * parts + 2):
    nb.do_force_step_pipelined(i, sw, plan, nbat.xq, nbat.f, dynamic_pruning=cfg["dynamic_pruning"], num_parts=parts)
    nb.gpu_wait_finish_task(sw, LOCAL)
nb.set_pipeline_timeline(True)
rows = []
for i in (2 * parts + 2, 2 * parts + 3):          # an even step (no rolling prune) and an odd one
    ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    stream = torch.cuda.ExternalStream(nb.streams()[0])
    ev[0].record(stream)
    nb.do_force_step_pipelined(i, sw, plan, nbat.xq, nbat.f, dynamic_pruning=cfg["dynamic_pruning"], num_parts=parts)
    ev[1].record(stream)
    nb.gpu_wait_finish_task(sw, LOCAL)
    torch.cuda.synchronize()
    t = nb.pipeline_timeline()
    k = t[:, 2] - t[:, 1]
    rows.append({"step": i, "rolling_prune": bool(i % 2), "step_ms": ev[0].elapsed_time(ev[1]), "chunks": int(t.shape[0]),
                 "h2d_done_ms": [round(float(v), 3) for v in t[:, 0]], "kernel_start_ms": [round(float(v), 3) for v in t[:, 1]],
                 "kernel_end_ms": [round(float(v), 3) for v in t[:, 2]], "d2h_done_ms": [round(float(v), 3) for v in t[:, 3]],
                 "sum_of_kernel_intervals_ms": float(k.sum()), "last_kernel_end_ms": float(t[:, 2].max()),
                 "last_d2h_end_ms": float(t[:, 3].max()), "first_kernel_start_ms": float(t[:, 1].min())})
for r in rows:
    print(json.dumps(r))
nb.gpu_free()
