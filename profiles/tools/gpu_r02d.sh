#!/bin/bash
# r02d: all GPU tests (no -x), bench lines of the five workloads with the reference-derived list radii, stock vs shim with
# the same radii, ncu --set full of the energy kernel (96 k) and the rolling prune kernel (12 M).
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -n 40 > gpurun_out/r02d_pytest_gpu.log; tail -n 8 gpurun_out/r02d_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
for wl in water96k_fswitch water384k_ljpme water384k_pswitch water1536k; do
    timeout 600 python bench.py --workload $wl --steps 40 --warmup 12 > gpurun_out/r02d_bench_$wl.json 2> gpurun_out/r02d_bench_$wl.err
done
timeout 900 python bench.py --steps 20 --warmup 12 > gpurun_out/r02d_bench_water12m.json 2> gpurun_out/r02d_bench_water12m.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02d_bench_reference.json 2> gpurun_out/r02d_bench_reference.err
H=oracle/_ref/cuda/bench_ref_gpu
NT=$(nproc)
run() {
    name=$1; shift
    for impl in stock shim; do
        lib=oracle/_ref/cuda/lib; [ $impl = shim ] && lib=oracle/_ref/cuda/lib_shim
        GMX_ENABLE_GPU_TIMING=1 LD_LIBRARY_PATH=$lib:$LD_LIBRARY_PATH timeout 900 $H "$@" --nt $NT --dump /tmp/f_${name}_$impl.bin \
            > gpurun_out/r02d_${name}_$impl.json 2> gpurun_out/r02d_${name}_$impl.err
        echo "$name $impl exit $?" >> gpurun_out/r02d_summary.log
    done
    python profiles/tools/compare_ref_gpu.py $name gpurun_out/r02d_${name}_stock.json gpurun_out/r02d_${name}_shim.json \
        /tmp/f_${name}_stock.bin /tmp/f_${name}_shim.bin >> gpurun_out/r02d_compare.jsonl 2>> gpurun_out/r02d_summary.log
}
rm -f gpurun_out/r02d_summary.log gpurun_out/r02d_compare.jsonl
run bench3k           --size 1    --rc 0.9 --vdw cut     --iter 200 --warmup 10
run bench3k_energy    --size 1    --rc 0.9 --vdw cut     --energy 1 --iter 200 --warmup 10
run water96k_fswitch  --size 32   --rc 1.0 --vdw fswitch --energy 1 --rlist-outer 1.172 --rlist-inner 1.003 --nstlist-prune 10 --dynamic-pruning 1 --iter 100 --warmup 12
run water384k_ljpme   --size 128  --rc 1.0 --vdw ljpme   --rlist-outer 1.172 --rlist-inner 1.003 --nstlist-prune 10 --dynamic-pruning 1 --iter 100 --warmup 12
run water384k_pswitch --size 128  --rc 1.0 --vdw pswitch --rlist-outer 1.172 --rlist-inner 1.003 --nstlist-prune 10 --dynamic-pruning 1 --iter 100 --warmup 12
run water1536k        --size 512  --rc 1.0 --vdw cut     --rlist-outer 1.172 --rlist-inner 1.003 --nstlist-prune 10 --dynamic-pruning 1 --iter 60 --warmup 12
run water12m          --size 4096 --rc 1.2 --vdw cut     --rlist-outer 1.358 --rlist-inner 1.201 --nstlist-prune 10 --dynamic-pruning 1 --iter 24 --warmup 12
run water12m_energy   --size 4096 --rc 1.2 --vdw cut     --energy 1 --rlist-outer 1.358 --rlist-inner 1.201 --nstlist-prune 10 --dynamic-pruning 1 --iter 12 --warmup 12
cat gpurun_out/r02d_summary.log; cut -c1-400 gpurun_out/r02d_compare.jsonl
cap() { # tag, kernel regex, skip, bench args...
    tag=$1; k=$2; skip=$3; shift 3
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/r02d_prof_$tag \
        python bench.py "$@" --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02d_ncu_$tag.log 2>&1
    ncu -i gpurun_out/r02d_prof_$tag.ncu-rep --page raw --csv > gpurun_out/r02d_prof_$tag.csv 2>/dev/null
    python profiles/tools/ncu_summary.py gpurun_out/r02d_prof_$tag.csv > gpurun_out/r02d_prof_$tag.txt 2>&1
}
cap energy96k nbnxm_force_kernel 4 --workload water96k_fswitch
cap prune12m nbnxm_prune_kernel 3
ls -la gpurun_out/*.ncu-rep
