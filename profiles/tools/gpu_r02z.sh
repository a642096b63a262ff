#!/bin/bash
# r02z: A/B of packed j-force accumulators in the force-only both-halves body (-DNBNXM_PACKED_FJ2: 3 FFMA2 instead of 6 FFMA)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
bench() { timeout 900 python bench.py --workload $2 --steps 20 --warmup 12 --no-cpu-baseline > gpurun_out/r02z_bench_$2_$1.json 2> gpurun_out/r02z_bench_$2_$1.err; }
bench base water12m; bench base water1536k; bench base water384k_ljpme; bench base water384k_pswitch
touch gromacs_b200/csrc/*.cuh
make -s -j32 -C gromacs_b200/csrc EXTRA="-DNBNXM_PACKED_FJ2" > gpurun_out/r02z_build.log 2>&1 || { echo "build failed"; tail -5 gpurun_out/r02z_build.log; }
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -n 1
bench fj2 water12m; bench fj2 water1536k; bench fj2 water384k_ljpme; bench fj2 water384k_pswitch
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02z_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f[22:-5], "ms/step %.4f kernel_us %.1f frac %.4f" % (d["ms_per_step"], d["roofline"]["kernel_us"], d["roofline"]["frac"]))
    except Exception as e:
        print(f, "failed", e)
PY
