#!/bin/bash
# r02n: source-level ncu capture of the packed force kernel (1.5 M atoms, F-only) - per-instruction stall samples
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nbnxm_force_kernel -s 4 -c 1 -f -o gpurun_out/r02n_prof_1536k \
    python bench.py --workload water1536k --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02n_ncu.log 2>&1
ncu -i gpurun_out/r02n_prof_1536k.ncu-rep --page source --csv > gpurun_out/r02n_source_1536k.csv 2>/dev/null
ncu -i gpurun_out/r02n_prof_1536k.ncu-rep --page raw --csv > gpurun_out/r02n_raw_1536k.csv 2>/dev/null
rm -f gpurun_out/r02n_prof_1536k.ncu-rep
ls -la gpurun_out/r02n_*
