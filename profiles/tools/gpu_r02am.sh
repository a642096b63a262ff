#!/bin/bash
# r02am: the code as the round ends (device perturbed-pair split, gridding with block counters on top of r02ah): all single-GPU
# tests, smoke, the default bench line and the 96 k / 1.5 M lines, launch list of the default command
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -n 60 > gpurun_out/r02am_pytest_gpu.log; tail -n 3 gpurun_out/r02am_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2 | tee gpurun_out/r02am_smoke.log
timeout 900 python bench.py > gpurun_out/r02am_bench_default.json 2> gpurun_out/r02am_bench_default.err
for wl in water96k_fswitch water1536k; do
    timeout 600 python bench.py --workload $wl --steps 40 --warmup 12 > gpurun_out/r02am_bench_$wl.json 2> gpurun_out/r02am_bench_$wl.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 80 --csv --log-file gpurun_out/r02am_launches_default.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import json
for n in ("default", "water1536k", "water96k_fswitch"):
    try:
        d = json.loads(open("gpurun_out/r02am_bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, "ms/step %.4f value %.1f vws %.1f kernel_us %.1f frac %.4f e2e_ms %.3f e2e %.1f cpu %.2f" % (d["ms_per_step"], d["value"], d["value_with_search"], d["roofline"]["kernel_us"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["cpu_baseline"]["value"]), d["parity"]["vs_oracle_sample"]["f_relrms"], d["search_step"].get("gpu_grid_ms"), d["search_step"].get("gpu_list_ms"), d["search_step"].get("same_list_entries_as_host"))
    except Exception as e:
        print(n, "failed", e)
PY
