#!/bin/bash
# r02c: (1) energy steps, stock CUDA backend of the reference vs the shim; (2) the sci order at 12.3 M atoms: count-sorted
# over the whole list against sorted within tiles of 4096 entries - bench line, ncu DRAM traffic; (3) GPU parity tests.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
H=oracle/_ref/cuda/bench_ref_gpu
NT=$(nproc)
run() {
    name=$1; shift
    for impl in stock shim; do
        lib=oracle/_ref/cuda/lib; [ $impl = shim ] && lib=oracle/_ref/cuda/lib_shim
        GMX_ENABLE_GPU_TIMING=1 LD_LIBRARY_PATH=$lib:$LD_LIBRARY_PATH timeout 900 $H "$@" --nt $NT --dump /tmp/f_${name}_$impl.bin \
            > gpurun_out/r02b_${name}_$impl.json 2> gpurun_out/r02b_${name}_$impl.err
        echo "$name $impl exit $?" >> gpurun_out/r02c_summary.log
    done
    python profiles/tools/compare_ref_gpu.py $name gpurun_out/r02b_${name}_stock.json gpurun_out/r02b_${name}_shim.json \
        /tmp/f_${name}_stock.bin /tmp/f_${name}_shim.bin >> gpurun_out/r02c_compare.jsonl 2>> gpurun_out/r02c_summary.log
}
rm -f gpurun_out/r02c_summary.log gpurun_out/r02c_compare.jsonl
run bench3k_energy    --size 1    --rc 0.9 --vdw cut     --energy 1 --iter 200 --warmup 10
run bench3k_cutlb     --size 1    --rc 0.9 --vdw cutlb   --energy 1 --iter 100 --warmup 10
run water96k_fswitch  --size 32   --rc 1.0 --vdw fswitch --energy 1 --rlist-outer 1.18 --rlist-inner 1.002 --dynamic-pruning 1 --iter 100 --warmup 12
run water1536k_energy --size 512  --rc 1.0 --vdw cut     --energy 1 --rlist-outer 1.18 --rlist-inner 1.002 --dynamic-pruning 1 --iter 40 --warmup 12
run water12m_energy   --size 4096 --rc 1.2 --vdw cut     --energy 1 --rlist-outer 1.35 --rlist-inner 1.202 --dynamic-pruning 1 --iter 12 --warmup 12
cat gpurun_out/r02c_summary.log; cat gpurun_out/r02c_compare.jsonl
# (2) sci order at 12.3 M atoms
for mode in global tiled; do
    NBNXM_B200_SCI_SORT=$mode timeout 900 python bench.py --steps 20 --warmup 12 --no-cpu-baseline > gpurun_out/r02c_bench_12m_$mode.json 2> gpurun_out/r02c_bench_12m_$mode.err
    NBNXM_B200_SCI_SORT=$mode timeout 600 ncu --set full --clock-control none --import-source on -k regex:nbnxm_force_kernel -s 4 -c 1 -f \
        -o gpurun_out/r02c_prof_12m_$mode python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02c_ncu_12m_$mode.log 2>&1
    ncu -i gpurun_out/r02c_prof_12m_$mode.ncu-rep --page raw --csv > gpurun_out/r02c_prof_12m_$mode.csv 2>/dev/null
    python profiles/tools/ncu_summary.py gpurun_out/r02c_prof_12m_$mode.csv > gpurun_out/r02c_prof_12m_$mode.txt 2>&1
    rm -f gpurun_out/r02c_prof_12m_$mode.ncu-rep
done
NBNXM_B200_SCI_SORT=tiled timeout 600 python bench.py --workload water1536k --steps 40 --warmup 12 --no-cpu-baseline > gpurun_out/r02c_bench_1536k_tiled.json 2>&1
grep -h -E "gpu__time_duration|dram__bytes|lts__t_bytes" gpurun_out/r02c_prof_12m_*.txt
# (3) parity under the default order
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -n 15 > gpurun_out/r02c_pytest_gpu.log; tail -n 6 gpurun_out/r02c_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3
