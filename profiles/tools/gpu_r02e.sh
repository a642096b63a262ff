#!/bin/bash
# r02e: all GPU tests (no -x), bench lines of the five workloads with the reference-derived list radii, ncu --set full of the energy kernel (96 k) and the rolling prune kernel (12 M).
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -n 150 > gpurun_out/r02e_pytest_gpu.log; tail -n 8 gpurun_out/r02e_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
for wl in water96k_fswitch water384k_ljpme water384k_pswitch water1536k; do
    timeout 600 python bench.py --workload $wl --steps 40 --warmup 12 > gpurun_out/r02e_bench_$wl.json 2> gpurun_out/r02e_bench_$wl.err
done
timeout 900 python bench.py --steps 20 --warmup 12 > gpurun_out/r02e_bench_water12m.json 2> gpurun_out/r02e_bench_water12m.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02e_bench_reference.json 2> gpurun_out/r02e_bench_reference.err
cap() { # tag, kernel regex, skip, bench args...
    tag=$1; k=$2; skip=$3; shift 3
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/r02e_prof_$tag \
        python bench.py "$@" --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02e_ncu_$tag.log 2>&1
    ncu -i gpurun_out/r02e_prof_$tag.ncu-rep --page raw --csv > gpurun_out/r02e_prof_$tag.csv 2>/dev/null
    python profiles/tools/ncu_summary.py gpurun_out/r02e_prof_$tag.csv > gpurun_out/r02e_prof_$tag.txt 2>&1
}
cap energy96k nbnxm_force_kernel 4 --workload water96k_fswitch
cap prune12m nbnxm_prune_kernel 3
ls -la gpurun_out/*.ncu-rep
