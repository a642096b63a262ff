#!/bin/bash
# r02ah: the state the round ends with - all single-GPU tests, smoke, the bench lines of the five workloads (default command first),
# the reference arm, the launch list of the default command, ncu --set full of the dominant kernels, compute-sanitizer over the
# new column sort of the device gridding.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -n 60 > gpurun_out/r02ah_pytest_gpu.log; tail -n 3 gpurun_out/r02ah_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2 | tee gpurun_out/r02ah_smoke.log
timeout 900 python bench.py > gpurun_out/r02ah_bench_default.json 2> gpurun_out/r02ah_bench_default.err
for wl in water96k_fswitch water384k_ljpme water384k_pswitch water1536k; do
    timeout 600 python bench.py --workload $wl --steps 40 --warmup 12 > gpurun_out/r02ah_bench_$wl.json 2> gpurun_out/r02ah_bench_$wl.err
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02ah_bench_reference.json 2> gpurun_out/r02ah_bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 80 --csv --log-file gpurun_out/r02ah_launches_default.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
cap() { # tag regex skip bench-args
    tag=$1; k=$2; skip=$3; shift 3
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/r02ah_prof_$tag \
        python bench.py "$@" --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02ah_ncu_$tag.log 2>&1
    ncu -i gpurun_out/r02ah_prof_$tag.ncu-rep --page raw --csv > gpurun_out/r02ah_prof_$tag.csv 2>/dev/null
    python profiles/tools/ncu_summary.py gpurun_out/r02ah_prof_$tag.csv > gpurun_out/r02ah_prof_$tag.txt 2>&1
    rm -f gpurun_out/r02ah_prof_$tag.ncu-rep gpurun_out/r02ah_prof_$tag.csv
}
cap force12m nbnxm_force_kernel 4
cap force1536k nbnxm_force_kernel 4 --workload water1536k
cap energy96k nbnxm_force_kernel 4 --workload water96k_fswitch
for tool in memcheck racecheck; do
    timeout 600 compute-sanitizer --tool $tool --error-exitcode 1 python -m pytest tests/test_gpu_search.py -m gpu -q -x -k "column_sort_forms and (uniform or equal) or search_step_entirely_on_the_device and not 1536k" > gpurun_out/r02ah_sanitizer_$tool.log 2>&1
    echo "column sort $tool: exit $?" | tee -a gpurun_out/r02ah_sanitizer_summary.log
done
python - <<'PY'
import json
for n in ("default", "water1536k", "water96k_fswitch", "water384k_ljpme", "water384k_pswitch"):
    try:
        d = json.loads(open("gpurun_out/r02ah_bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, "ms/step %.4f value %.1f vws %.1f kernel_us %.1f frac %.4f e2e_ms %.3f e2e %.1f cpu %.2f" % (d["ms_per_step"], d["value"], d["value_with_search"], d["roofline"]["kernel_us"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["cpu_baseline"]["value"]), d["parity"]["vs_oracle_sample"]["f_relrms"], d["search_step"].get("gpu_grid_ms"), d["search_step"].get("gpu_list_ms"))
    except Exception as e:
        print(n, "failed", e)
print(open("gpurun_out/r02ah_bench_reference.json").read()[:300])
for t in ("force12m", "force1536k", "energy96k"):
    print("==", t)
    for l in open("gpurun_out/r02ah_prof_%s.txt" % t):
        if any(k in l for k in ("gpu__time_duration", "issue_active", "pipe_fma_cycles", "warps_active", "inst_executed.sum", "registers")):
            print(l.rstrip()[:110])
PY
