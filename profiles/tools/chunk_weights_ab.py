"""End-to-end step (nbnxm_b200_do_force_step_pipelined) of a workload under different chunk width patterns
(NBNXM_B200_CHUNK_WEIGHTS), one process: mean ms per step over `steps` steps per pattern, even (no rolling prune) and odd.
usage: python profiles/tools/chunk_weights_ab.py [workload] [steps]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gromacs_b200 import LOCAL, NbnxmGpu, StepWorkload  # noqa: E402
from gromacs_b200.pipeline import make_chunk_plan  # noqa: E402
from gromacs_b200.workload import make_workload, rolling_prune_parts  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "water12m"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
PATTERNS = {
    "default32": None,
    "default24": 24,
    "w16": "1,1,2,3,6,8,8,8,8,8,8,6,3,2,1,1",
    "w20": "1,1,2,3,4,6,8,8,8,8,8,8,8,6,4,3,2,1,1,1",
    "w12": "1,1,2,4,8,16,16,8,4,2,1,1",
    "w14": "1,1,2,4,8,12,12,12,8,6,4,2,1,1",
    "w24": "1,1,2,3,4,4,6,6,6,6,6,6,6,6,6,6,4,4,3,3,2,2,1,1",
}
wl = make_workload(name)
cfg, nbat = wl.cfg, wl.nbat
xq_pin = torch.empty((nbat.numAtoms(), 4), dtype=torch.float32).pin_memory()
xq_pin.numpy()[:] = nbat.xq
nbat.xq = xq_pin.numpy()
f_pin = torch.zeros((nbat.numAtoms(), 3), dtype=torch.float32).pin_memory()
nbat.f = f_pin.numpy()
nb = NbnxmGpu(wl.params, nbat)
base = wl.pairlist(min_sci=nb.gpu_min_ci_balanced())
nb.gpu_init_atomdata(nbat)
nb.gpu_upload_shiftvec(nbat)
sw = StepWorkload(useGpuFBufferOps=False)
parts = rolling_prune_parts(cfg)
stream = torch.cuda.ExternalStream(nb.streams()[0])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for tag, pat in PATTERNS.items():
    os.environ.pop("NBNXM_B200_CHUNK_WEIGHTS", None)
    n = 32
    if isinstance(pat, str):
        os.environ["NBNXM_B200_CHUNK_WEIGHTS"] = pat
    elif isinstance(pat, int):
        n = pat
    plan = make_chunk_plan(wl.grid, base, n)
    nb.gpu_init_pairlist(plan.plist, LOCAL)
    nb.setupGpuShortRangeWork(LOCAL)
    for i in range(2 * parts + 2):
        nb.do_force_step_pipelined(i, sw, plan, nbat.xq, nbat.f, dynamic_pruning=cfg["dynamic_pruning"], num_parts=parts)
        nb.gpu_wait_finish_task(sw, LOCAL)
    ms = {0: [], 1: []}
    for i in range(2 * parts + 2, 2 * parts + 2 + steps):
        with torch.cuda.stream(stream):
            flush.zero_()
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record(stream)
        nb.do_force_step_pipelined(i, sw, plan, nbat.xq, nbat.f, dynamic_pruning=cfg["dynamic_pruning"], num_parts=parts)
        ev[1].record(stream)
        nb.gpu_wait_finish_task(sw, LOCAL)
        torch.cuda.synchronize()
        ms[i % 2].append(ev[0].elapsed_time(ev[1]))
    print(json.dumps({"pattern": tag, "weights": pat, "chunks": int(plan.nchunks), "ms_even_steps": sum(ms[0]) / len(ms[0]),
                      "ms_odd_steps_with_rolling_prune": sum(ms[1]) / len(ms[1]),
                      "ms_per_step": (sum(ms[0]) + sum(ms[1])) / (len(ms[0]) + len(ms[1]))}), flush=True)
nb.gpu_free()
