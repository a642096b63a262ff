#!/bin/bash
# r02x: packed f32x2 distances on top of the two-deep prefetch in the prune kernel (-DNBNXM_PRUNE_PACKED) against the default
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
touch gromacs_b200/csrc/nbnxm_prune.cu
make -s -j32 -C gromacs_b200/csrc EXTRA="-DNBNXM_PRUNE_PACKED" > gpurun_out/r02x_build.log 2>&1 || { echo "build failed"; tail -5 gpurun_out/r02x_build.log; }
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k prune 2>&1 | tail -n 2
for wl in water12m water1536k; do
    timeout 900 python bench.py --workload $wl --steps 20 --warmup 12 --no-cpu-baseline > gpurun_out/r02x_bench_$wl.json 2> gpurun_out/r02x_bench_$wl.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02x_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f[22:-5], "packed+two-deep: rolling_prune_us %.1f first_pass_prune_ms %.3f" % (d["roofline"]["rolling_prune_us"], d["search_step"].get("first_pass_prune_ms", -1)))
    except Exception as e:
        print(f, "failed", e)
PY
