#!/bin/bash
# r02s: compute-sanitizer over the kernels that are new in round 2 - the tiled sci sort (forced on small lists), the device slab
# builder and its re-index kernel, the pipelined step with the overlapped rolling prune, the bulk-copy staging under the prune path
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
rm -f gpurun_out/r02s_sanitizer_summary.log
san() { # tag tool env... -- pytest args
    tag=$1; tool=$2; shift 2
    timeout 800 env "$@" > gpurun_out/r02s_${tag}_$tool.log 2>&1
    echo "$tag $tool: exit $?" >> gpurun_out/r02s_sanitizer_summary.log
    tail -n 2 gpurun_out/r02s_${tag}_$tool.log
}
for tool in memcheck racecheck; do
    san tiledsort $tool NBNXM_B200_SCI_SORT=tiled compute-sanitizer --tool $tool --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "prune_masks and (test243_ewald_cutnone or bench1_ewald_cutgeom or bench1_rf_cutnone_split)"
    san slab $tool compute-sanitizer --tool $tool --error-exitcode 1 python -m pytest tests/test_gpu_search.py -m gpu -q -x -k "device_slab_lists and 2"
    san pipelined $tool compute-sanitizer --tool $tool --error-exitcode 1 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -k "pipelined_step"
done
cat gpurun_out/r02s_sanitizer_summary.log
