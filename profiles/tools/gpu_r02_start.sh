#!/bin/bash
# First gpurun call of the next round: what round 1 left unmeasured on the GPU.
#   1. the whole -m gpu suite plus the opt-in test of the warp-cooperative mask pass (NBNXM_B200_SEARCH_COOP=1, so far only
#      run through the CPU emulation), which also writes its build time next to the default form's;
#   2. per-pass launch list of the device search step at 12.3 M atoms (gridding + list), default and cooperative;
#   3. the default bench line (now carrying "search_step");
#   4. compute-sanitizer (memcheck, racecheck, initcheck) over the 243-atom force / prune / search cases (SURVEY section 5:
#      the reference has no GPU race detection; ours is this).
# usage: gpurun --timeout 900 -- 'bash profiles/tools/gpu_r02_start.sh'
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
NBNXM_B200_TEST_UNVERIFIED=1 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu.log
for coop in 0 1; do
    NBNXM_B200_SEARCH_COOP=$coop ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
        --log-file gpurun_out/r02_search_launches_12m_coop$coop.csv \
        python profiles/tools/search_profile.py water12m 2 150000 > gpurun_out/r02_search_profile_ncu_coop$coop.log 2>&1
    NBNXM_B200_SEARCH_COOP=$coop python profiles/tools/search_profile.py water12m 3 150000 \
        > gpurun_out/r02_search_profile_12m_coop$coop.json 2> gpurun_out/r02_search_profile_12m_coop$coop.err
done
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
for tool in memcheck racecheck initcheck; do
    timeout 600 compute-sanitizer --tool $tool --error-exitcode 1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_search.py \
        -m gpu -q -x -k "test243 and not twin and not rf" > gpurun_out/r02_sanitizer_$tool.log 2>&1
    echo "compute-sanitizer $tool: exit $?" >> gpurun_out/r02_sanitizer_summary.log
done
cat gpurun_out/r02_sanitizer_summary.log
cat gpurun_out/r02_pytest_gpu.log gpurun_out/r02_search_profile_12m_coop*.json gpurun_out/r02_bench_default.json
