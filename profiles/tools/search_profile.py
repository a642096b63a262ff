"""Builds the pair list of a workload on the GPU (nbnxm_b200_gpu_search_build) a few times and prints the device time
of each build; run under `ncu --metrics gpu__time_duration.sum` for the per-pass launch list.
usage: python profiles/tools/search_profile.py [workload] [builds] [min_sci]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gromacs_b200 import LOCAL, NbnxmGpu  # noqa: E402
from gromacs_b200.pairsearch import GpuPairSearch  # noqa: E402
from gromacs_b200.workload import make_workload  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "water1536k"
builds = int(sys.argv[2]) if len(sys.argv) > 2 else 3
min_sci = int(sys.argv[3]) if len(sys.argv) > 3 else 18944
wl = make_workload(name)
t0 = time.time()
ref = wl.pairlist(min_sci=min_sci)
host_s = time.time() - t0
nb = NbnxmGpu(wl.params, wl.nbat)
nb.gpu_init_atomdata(wl.nbat)
nb.gpu_upload_shiftvec(wl.nbat)
nb.gpu_copy_xq_to_gpu(wl.nbat, LOCAL)
search = GpuPairSearch(nb, wl.grid, wl.box.excl_index, wl.box.excl_atoms)
ms = []
for _ in range(builds):
    sizes = search.build(wl.cfg["rlist_outer"], LOCAL, min_sci=min_sci)
    ms.append(search.build_ms)
# entry-for-entry comparison with the host builder's list (all threads: same sci / cjPacked bits, the exclusion entries
# are numbered in another order, so they are compared through the groups that point at them)
import numpy as np  # noqa: E402
got = search.download()


def same_entries(a, b):
    sa, sb = np.asarray(a.sci).reshape(-1, 4), np.asarray(b.sci).reshape(-1, 4)
    ca, cb = np.asarray(a.cjPacked).reshape(-1, 8), np.asarray(b.cjPacked).reshape(-1, 8)
    if sa.shape != sb.shape or ca.shape != cb.shape or not np.array_equal(sa, sb):
        return False
    if not np.array_equal(ca[:, [0, 1, 2, 3, 4, 6]], cb[:, [0, 1, 2, 3, 4, 6]]):
        return False
    ea, eb = np.asarray(a.excl).reshape(-1, 32), np.asarray(b.excl).reshape(-1, 32)
    for h in (5, 7):
        ia, ib = ca[:, h], cb[:, h]
        m = ia != 0
        if not np.array_equal(m, ib != 0) or not np.array_equal(ea[ia[m]], eb[ib[m]]):
            return False
    return True


entries_equal = bool(same_entries(got, ref))
del got
# the host path for comparison: upload of the host builder's list (gpu_init_pairlist from pageable numpy arrays)
from gromacs_b200 import StepWorkload  # noqa: E402
nb.gpu_wait_finish_task(StepWorkload(), LOCAL)
t0 = time.time()
nb.gpu_init_pairlist(ref, LOCAL)
nb.gpu_wait_finish_task(StepWorkload(), LOCAL)
h2d_s = time.time() - t0
print(json.dumps({"host_list_upload_s": h2d_s, "workload": name, "natoms": wl.box.natoms, "rlist": wl.cfg["rlist_outer"], "nsci": sizes[0],
                  "ncj_packed": sizes[1], "nexcl": sizes[2], "same_sizes_as_host": sizes == (ref.sci.shape[0], ref.cjPacked.shape[0], ref.excl.shape[0]),
                  "same_entries_as_host": entries_equal, "gpu_build_ms": ms, "host_build_s": host_s, "host_threads": wl.grid.nthreads,
                  "list_bytes": int(ref.sci.nbytes + ref.cjPacked.nbytes + ref.excl.nbytes)}))
search.free()
# the whole search step on the device: gridding from atom-order coordinates in device memory, then the list
import torch  # noqa: E402
from gromacs_b200.pairsearch import Grid  # noqa: E402
import numpy as np  # noqa: E402
t0 = time.time()
Grid(wl.box.box, wl.box.x)
host_grid_s = time.time() - t0
x_dev = torch.from_numpy(np.ascontiguousarray(wl.box.x, np.float32)).cuda()
torch.cuda.synchronize()
s2 = GpuPairSearch(nb)
s2.set_atoms(wl.box.q, wl.box.type, wl.nbat.numTypes, wl.nbat.nbfp_comb, wl.box.excl_index, wl.box.excl_atoms)
grid_ms, list_ms = [], []
for _ in range(builds):
    dims = s2.put_atoms_on_grid(wl.box.box, x_dev.data_ptr())
    ai, fb, gms = s2.get_order()
    sizes2 = s2.build(wl.cfg["rlist_outer"], LOCAL, min_sci=min_sci)
    grid_ms.append(gms)
    list_ms.append(s2.build_ms)
print(json.dumps({"device_search_step": True, "workload": name, "gpu_grid_ms": grid_ms, "gpu_list_ms": list_ms,
                  "host_grid_s": host_grid_s, "same_order_as_host": bool(np.array_equal(ai, wl.grid.atom_index)),
                  "same_sizes_as_host": sizes2 == sizes, "max_column_atoms": int(np.diff(fb).max() * 64)}))
s2.free()
nb.gpu_free()
