#!/bin/bash
# r02a: run everything round 1 left unexecuted on a GPU (FEP launch path, cooperative mask pass) + compute-sanitizer.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
NBNXM_B200_TEST_UNVERIFIED=1 timeout 600 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_zz_fep.py 2>&1 | tail -30 > gpurun_out/r02a_pytest_gpu.log
NBNXM_B200_TEST_UNVERIFIED=1 timeout 600 python -m pytest tests/test_gpu_zz_fep.py -m gpu -q 2>&1 | tail -80 > gpurun_out/r02a_pytest_fep.log
for tool in memcheck racecheck initcheck; do
    timeout 500 compute-sanitizer --tool $tool --error-exitcode 1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_search.py \
        -m gpu -q -x -k "test243 and not twin and not rf" > gpurun_out/r02a_sanitizer_$tool.log 2>&1
    echo "compute-sanitizer $tool: exit $?" >> gpurun_out/r02a_sanitizer_summary.log
done
cat gpurun_out/r02a_sanitizer_summary.log
tail -5 gpurun_out/r02a_pytest_gpu.log; tail -30 gpurun_out/r02a_pytest_fep.log
