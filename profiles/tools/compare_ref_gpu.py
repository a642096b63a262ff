"""Compares the outputs of two runs of oracle/_ref/cuda/bench_ref_gpu (stock CUDA backend of the reference vs the same
libgromacs with the nbnxm_b200 backend behind gmx::gpu_*): atom-order forces and shift forces from the --dump files,
energies and timings from the JSON lines.  Prints one JSON line per workload."""
import json
import sys

import numpy as np


def load(dump, natoms):
    a = np.fromfile(dump, dtype=np.float32)
    return a[:3 * natoms].reshape(natoms, 3).astype(np.float64), a[3 * natoms:3 * natoms + 135].reshape(45, 3).astype(np.float64)


def main():
    name, js, jd, ds, dd = sys.argv[1:6]
    s = json.loads(open(js).read().strip().splitlines()[-1])
    d = json.loads(open(jd).read().strip().splitlines()[-1])
    n = int(s["natoms"])
    fs, shs = load(ds, n)
    fd, shd = load(dd, n)
    rms = np.sqrt(((fd - fs) ** 2).sum() / (fs ** 2).sum())
    maxc = np.abs(fd - fs).max() / np.abs(fs).max()
    out = {"workload": name, "natoms": n,
           "f_relrms_shim_vs_stock": rms, "f_maxcomp_rel": maxc,
           "e_lj_rel": abs(d["e_lj"] - s["e_lj"]) / max(abs(s["e_lj"]), 1e-30) if s["energy"] else None,
           "e_el_rel": abs(d["e_el"] - s["e_el"]) / max(abs(s["e_el"]), 1e-30) if s["energy"] else None,
           "virial_rel": abs(d["virial_shift_part"] - s["virial_shift_part"]) / max(abs(s["virial_shift_part"]), 1e-30) if s["energy"] else None,
           "stock": {k: s[k] for k in ("force_kernel_ms", "first_prune_ms", "rolling_prune_ms", "sec_per_step_host_buffers", "h2d_ms", "d2h_ms")},
           "shim": {k: d[k] for k in ("force_kernel_ms", "first_prune_ms", "rolling_prune_ms", "sec_per_step_host_buffers", "h2d_ms", "d2h_ms")},
           "force_kernel_speedup": s["force_kernel_ms"] / d["force_kernel_ms"] if d["force_kernel_ms"] else None,
           "rolling_prune_speedup": s["rolling_prune_ms"] / d["rolling_prune_ms"] if d["rolling_prune_ms"] else None,
           "step_speedup_host_buffers": s["sec_per_step_host_buffers"] / d["sec_per_step_host_buffers"]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
