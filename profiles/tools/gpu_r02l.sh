#!/bin/bash
# r02l: all single-GPU tests with the bulk-copy descriptor staging as the default and the device-side slab lists;
# compute-sanitizer over the packed kernel's new staging (memcheck, racecheck, synccheck).
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -n 60 > gpurun_out/r02l_pytest_gpu.log; tail -n 6 gpurun_out/r02l_pytest_gpu.log
for tool in memcheck racecheck synccheck; do
    timeout 500 compute-sanitizer --tool $tool --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
        -k "packed and (test243_ewald_cutgeom or test243_ewald_pswitch or bench1_ewald_cutgeom or bench1_rf_cutnone_split)" > gpurun_out/r02l_sanitizer_$tool.log 2>&1
    echo "compute-sanitizer $tool: exit $?" >> gpurun_out/r02l_sanitizer_summary.log
    tail -n 3 gpurun_out/r02l_sanitizer_$tool.log
done
cat gpurun_out/r02l_sanitizer_summary.log
timeout 900 python bench.py --steps 20 --warmup 12 > gpurun_out/r02l_bench_12m.json 2> gpurun_out/r02l_bench_12m.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02l_bench_12m.json").read().strip().splitlines()[-1])
print("12m ms/step %.4f kernel_us %.1f frac %.4f e2e_ms %.3f vws %.1f parity %s" % (d["ms_per_step"], d["roofline"]["kernel_us"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["value_with_search"], d["parity"]["vs_oracle_sample"]))
print(d["search_step"])
PY
