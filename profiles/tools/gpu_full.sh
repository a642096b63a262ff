#!/bin/bash
# usage: gpu_full.sh <tag>  — GPU parity tests, bench lines (96k default + 1536k) and ncu --set full captures of the force kernels
tag=${1:-run}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_$tag.log
tail -4 gpurun_out/pytest_gpu_$tag.log
timeout 600 python bench.py > gpurun_out/bench_96k_$tag.json 2> gpurun_out/bench_96k_$tag.err
timeout 600 python bench.py --workload water1536k --steps 20 --no-cpu-baseline > gpurun_out/bench_1536k_$tag.json 2> gpurun_out/bench_1536k_$tag.err
python - <<PY
import json
for w in ("1536k","96k"):
    try:
        d=json.load(open("gpurun_out/bench_%s_$tag.json"%w))
        print("$tag",w,"kernel_us %.1f frac %.3f step_us %.1f value %.1f e2e %.1f prune_us %.1f"%(d["roofline"]["kernel_us"],d["roofline"]["frac"],d["us_per_force_step"],d["value"],d["e2e"]["value"],d["roofline"]["rolling_prune_us"]))
    except Exception as e:
        print("$tag",w,"failed",e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nbnxm_force_kernel -s 4 -c 1 -f -o gpurun_out/prof_1536k_$tag \
    python bench.py --workload water1536k --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_1536k_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nbnxm_force_kernel -s 4 -c 1 -f -o gpurun_out/prof_96k_$tag \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_96k_$tag.log 2>&1
