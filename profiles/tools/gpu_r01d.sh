#!/bin/bash
# One gpurun call: GPU tests, bench lines (96k default, 1536k), launch list and ncu --set full captures of the force kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_96k.json 2> gpurun_out/bench_96k.err; tail -c 600 gpurun_out/bench_96k.json
timeout 600 python bench.py --workload water1536k --steps 20 --no-cpu-baseline > gpurun_out/bench_1536k.json 2> gpurun_out/bench_1536k.err; tail -c 600 gpurun_out/bench_1536k.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nbnxm_force_kernel -s 4 -c 1 -f -o gpurun_out/prof_1536k \
    python bench.py --workload water1536k --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_1536k.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nbnxm_force_kernel -s 4 -c 1 -f -o gpurun_out/prof_96k \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_96k.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches_96k.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/launches_96k.log 2>&1
ls -la gpurun_out
