#!/bin/bash
# r02zz: all single-GPU tests, smoke, the pipelined step's timeline with the need-ordered upload, the bench lines of the five
# workloads and the reference arm - the state the round ends with.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -n 60 > gpurun_out/r02zz_pytest_gpu.log; tail -n 5 gpurun_out/r02zz_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
python profiles/tools/pipeline_timeline.py water12m 32 > gpurun_out/r02zz_timeline_12m_32.jsonl 2> gpurun_out/r02zz_timeline.err
timeout 900 python bench.py --steps 20 --warmup 12 > gpurun_out/r02zz_bench_water12m.json 2> gpurun_out/r02zz_bench_water12m.err
for wl in water96k_fswitch water384k_ljpme water384k_pswitch water1536k; do
    timeout 600 python bench.py --workload $wl --steps 40 --warmup 12 > gpurun_out/r02zz_bench_$wl.json 2> gpurun_out/r02zz_bench_$wl.err
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02zz_bench_reference.json 2> gpurun_out/r02zz_bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 80 --csv --log-file gpurun_out/r02zz_launches_default.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/r02zz_timeline_12m_32.jsonl"):
    if l.startswith("{"):
        d = json.loads(l)
        print("timeline step", d["step"], "prune", d["rolling_prune"], "step_ms %.3f first_k_start %.3f last_k_end %.3f last_d2h %.3f" % (d["step_ms"], d["first_kernel_start_ms"], d["last_kernel_end_ms"], d["last_d2h_end_ms"]))
for n in ("water12m", "water1536k", "water96k_fswitch", "water384k_ljpme", "water384k_pswitch"):
    try:
        d = json.loads(open("gpurun_out/r02zz_bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, "ms/step %.4f value %.1f vws %.1f frac %.4f e2e_ms %.3f e2e %.1f cpu %.2f" % (d["ms_per_step"], d["value"], d["value_with_search"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["cpu_baseline"]["value"]), d["parity"]["vs_oracle_sample"]["f_relrms"])
    except Exception as e:
        print(n, "failed", e)
print(open("gpurun_out/r02zz_bench_reference.json").read()[:300])
PY
