#!/bin/bash
tag=${1:-last}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -2 gpurun_out/pytest_gpu_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_default_$tag.json 2> gpurun_out/bench_default_$tag.err; tail -c 1200 gpurun_out/bench_default_$tag.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 80 --csv --log-file gpurun_out/launches_default_$tag.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
