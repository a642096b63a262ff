#!/bin/bash
# r02ab: column sort of the device gridding as buckets + ranks (default) against the bitonic networks
# (NBNXM_B200_SEARCH_BITONIC_SORT=1); j-force reduction with packed x/y sums.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -n 5 > gpurun_out/r02ab_pytest.log; tail -n 2 gpurun_out/r02ab_pytest.log
bench() { tag=$1; wl=$2; shift 2; env "$@" timeout 900 python bench.py --workload $wl --steps 20 --warmup 12 --no-cpu-baseline > gpurun_out/r02ab_bench_${wl}_$tag.json 2> gpurun_out/r02ab_bench_${wl}_$tag.err; }
bench bucket water12m X=1
bench bitonic water12m NBNXM_B200_SEARCH_BITONIC_SORT=1
bench bucket water1536k X=1
bench bitonic water1536k NBNXM_B200_SEARCH_BITONIC_SORT=1
bench bucket water96k_fswitch X=1
# launch list of the gridding passes at 1.5 M atoms with the new sort
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:search --csv --log-file gpurun_out/r02ab_search_launches_1536k.csv python profiles/tools/search_profile.py water1536k 2 > /dev/null 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02ab_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        s = d["search_step"]
        print(f[23:-5], "ms/step %.4f kernel_us %.1f frac %.4f vws %.2f grid_ms %.3f list_ms %.3f same order %s same entries %s" % (d["ms_per_step"], d["roofline"]["kernel_us"], d["roofline"]["frac"], d["value_with_search"], s["gpu_grid_ms"], s["gpu_list_ms"], s["same_grid_order_as_host"], s["same_list_entries_as_host"]))
    except Exception as e:
        print(f, "failed", e)
PY
