#!/bin/bash
# r02al: perturbed-pair split on the device (pass 8 of the builder) and the gridding passes with block-level column counters:
# GPU tests of the search step and the perturbed path, search times at 12.3 M and 1.5 M atoms against the per-atom global atomics
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zz_fep.py tests/test_gpu_search.py -m gpu -q -x 2>&1 | tail -n 12 > gpurun_out/r02al_pytest.log; tail -n 3 gpurun_out/r02al_pytest.log
for wl in water12m water1536k; do
    timeout 600 python profiles/tools/search_profile.py $wl 3 > gpurun_out/r02al_search_${wl}_block.json 2> gpurun_out/r02al_search_${wl}_block.err
    NBNXM_B200_SEARCH_GLOBAL_ATOMICS=1 timeout 600 python profiles/tools/search_profile.py $wl 3 > gpurun_out/r02al_search_${wl}_global.json 2> gpurun_out/r02al_search_${wl}_global.err
done
for tool in memcheck racecheck; do
    timeout 600 compute-sanitizer --tool $tool --error-exitcode 1 python -m pytest tests/test_gpu_zz_fep.py tests/test_gpu_search.py -m gpu -q -x -k "device_built_and_split or search_step_entirely_on_the_device and not 1536k" > gpurun_out/r02al_sanitizer_$tool.log 2>&1
    echo "fep split + gridding $tool: exit $?" | tee -a gpurun_out/r02al_sanitizer_summary.log
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02al_search_*.json")):
    try:
        recs = [json.loads(l) for l in open(f) if l.startswith("{")]
        dev = next(r for r in recs if r.get("device_search_step"))
        print(f[24:-5], "grid_ms", [round(v, 3) for v in dev["gpu_grid_ms"]], "list_ms", [round(v, 3) for v in dev["gpu_list_ms"]], "same order", dev["same_order_as_host"], "same sizes", dev["same_sizes_as_host"])
    except Exception as e:
        print(f, "failed", e)
PY
