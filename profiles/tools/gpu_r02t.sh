#!/bin/bash
# r02t: source-level ncu capture of the packed energy kernel (96 k atoms, force switch, F+E); FEP GPU tests after the split change
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zz_fep.py -m gpu -q 2>&1 | tail -n 3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nbnxm_force_kernel -s 4 -c 1 -f -o gpurun_out/r02t_prof_96k \
    python bench.py --workload water96k_fswitch --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02t_ncu.log 2>&1
ncu -i gpurun_out/r02t_prof_96k.ncu-rep --page source --csv > gpurun_out/r02t_source_96k.csv 2>/dev/null
rm -f gpurun_out/r02t_prof_96k.ncu-rep
ls -la gpurun_out/r02t_*
