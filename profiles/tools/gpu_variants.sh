#!/bin/bash
# usage: gpu_variants.sh "<name>:<EXTRA flags>" ...   Builds each variant, runs the parity tests once (first variant) and both benches.
mkdir -p gpurun_out
first=1
for v in "$@"; do
  name=${v%%:*}; extra=${v#*:}
  touch gromacs_b200/csrc/*.cuh
  make -s -j32 -C gromacs_b200/csrc EXTRA="$extra" > gpurun_out/build_$name.log 2>&1 || { echo "build failed $name"; tail -5 gpurun_out/build_$name.log; continue; }
  if [ $first = 1 ] && [ -z "$SKIP_TESTS" ]; then
    timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
  fi
  first=0
  timeout 600 python bench.py --workload water1536k --steps 20 --no-cpu-baseline > gpurun_out/bench_1536k_$name.json 2> gpurun_out/bench_1536k_$name.err
  timeout 600 python bench.py --workload water96k_fswitch --no-cpu-baseline > gpurun_out/bench_96k_$name.json 2> gpurun_out/bench_96k_$name.err
  python - <<PY
import json
for w in ("1536k","96k"):
    try:
        d=json.load(open("gpurun_out/bench_%s_$name.json"%w))
        print("$name",w,"kernel_us %.1f frac %.3f step_us %.1f e2e %.1f"%(d["roofline"]["kernel_us"],d["roofline"]["frac"],d["us_per_force_step"],d["e2e"]["value"]))
    except Exception as e:
        print("$name",w,"failed",e)
PY
done
