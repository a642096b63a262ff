#!/bin/bash
# usage: gpu_final.sh <tag> : GPU tests, default bench line + reference arm, secondary workloads, ncu captures, launch list, microbenchmarks
tag=${1:-final}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_$tag.log
tail -3 gpurun_out/pytest_gpu_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; tail -1 gpurun_out/smoke_$tag.log
timeout 900 python bench.py > gpurun_out/bench_default_$tag.json 2> gpurun_out/bench_default_$tag.err
timeout 900 python bench.py --impl reference --steps 10 > gpurun_out/bench_reference_$tag.json 2> gpurun_out/bench_reference_$tag.err
for w in water96k_fswitch water1536k water384k_ljpme water384k_pswitch; do
  timeout 600 python bench.py --workload $w --steps 20 --no-cpu-baseline > gpurun_out/bench_${w}_$tag.json 2> gpurun_out/bench_${w}_$tag.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_*_$tag.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get("roofline",{})
        print(f.split("/")[-1], "value %.1f ms/step %.4f e2e %.1f kernel_us %s frac %s cpu %s"%(d["value"], d["ms_per_step"], d["e2e"]["value"], r.get("kernel_us"), r.get("frac"), d.get("cpu_baseline",{}).get("value")))
    except Exception as e:
        print(f, "failed", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nbnxm_force_kernel -s 4 -c 1 -f -o gpurun_out/prof_12m_$tag \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_12m_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nbnxm_prune_kernel -s 8 -c 1 -f -o gpurun_out/prof_prune_12m_$tag \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_prune_12m_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nbnxm_force_kernel -s 4 -c 1 -f -o gpurun_out/prof_96k_$tag \
    python bench.py --workload water96k_fswitch --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_96k_$tag.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/launches_default_$tag.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/launches_default_$tag.log 2>&1
./profiles/microbench/ffma2_probe > gpurun_out/ffma2_probe_$tag.txt 2>&1; tail -30 gpurun_out/ffma2_probe_$tag.txt
