#!/bin/bash
# r02y: ncu --set full of the dominant kernels as the round ends (force F-only at 12.3 M and 1.5 M atoms, F+E at 96 k, rolling and
# first-pass prune at 12.3 M), the launch list of the default bench command, and the default bench line next to them.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
cap() { # tag regex skip bench-args
    tag=$1; k=$2; skip=$3; shift 3
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/r02y_prof_$tag \
        python bench.py "$@" --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02y_ncu_$tag.log 2>&1
    ncu -i gpurun_out/r02y_prof_$tag.ncu-rep --page raw --csv > gpurun_out/r02y_prof_$tag.csv 2>/dev/null
    python profiles/tools/ncu_summary.py gpurun_out/r02y_prof_$tag.csv > gpurun_out/r02y_prof_$tag.txt 2>&1
    rm -f gpurun_out/r02y_prof_$tag.ncu-rep gpurun_out/r02y_prof_$tag.csv
}
cap force12m nbnxm_force_kernel 4
cap force1536k nbnxm_force_kernel 4 --workload water1536k
cap energy96k nbnxm_force_kernel 4 --workload water96k_fswitch
cap prune_rolling12m "nbnxm_prune_kernel<\(bool\)0>|nbnxm_prune_kernelILb0" 2
cap prune_first12m "nbnxm_prune_kernel<\(bool\)1>|nbnxm_prune_kernelILb1" 0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02y_launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 900 python bench.py > gpurun_out/r02y_bench_default.json 2> gpurun_out/r02y_bench_default.err
for t in force12m force1536k energy96k prune_rolling12m prune_first12m; do echo "== $t"; grep -E "Kernel Name|gpu__time_duration|dram__bytes|issue_active|pipe_fma_cycles|warps_active" gpurun_out/r02y_prof_$t.txt | cut -c1-110; done
cut -c1-300 gpurun_out/r02y_bench_default.json
