#!/bin/bash
# r02w: prune kernel with the masks requested two groups ahead: bit-exact masks, timings
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_benched_configs.py -m gpu -q 2>&1 | tail -n 4 > gpurun_out/r02w_pytest.log; tail -n 2 gpurun_out/r02w_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
for wl in water12m water1536k water384k_ljpme; do
    timeout 900 python bench.py --workload $wl --steps 20 --warmup 12 --no-cpu-baseline > gpurun_out/r02w_bench_$wl.json 2> gpurun_out/r02w_bench_$wl.err
done
timeout 600 ncu --set full --clock-control none -k regex:nbnxm_prune_kernel -s 3 -c 1 -f -o gpurun_out/r02w_prof_prune12m \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02w_ncu.log 2>&1
ncu -i gpurun_out/r02w_prof_prune12m.ncu-rep --page raw --csv > gpurun_out/r02w_prof_prune12m.csv 2>/dev/null
python profiles/tools/ncu_summary.py gpurun_out/r02w_prof_prune12m.csv > gpurun_out/r02w_prof_prune12m.txt 2>&1
rm -f gpurun_out/r02w_prof_prune12m.ncu-rep gpurun_out/r02w_prof_prune12m.csv
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02w_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f[22:-5], "ms/step %.4f kernel_us %.1f rolling_prune_us %.1f first_pass_prune_ms %.3f e2e_ms %.3f vws %.1f" % (d["ms_per_step"], d["roofline"]["kernel_us"], d["roofline"]["rolling_prune_us"], d["search_step"].get("first_pass_prune_ms", -1), d["e2e"]["ms_per_step"], d["value_with_search"] or -1))
    except Exception as e:
        print(f, "failed", e)
PY
grep -E "gpu__time_duration|issue_active|long_scoreboard|wait" gpurun_out/r02w_prof_prune12m.txt
