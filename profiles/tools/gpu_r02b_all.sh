#!/bin/bash
# r02b: the reference's own CUDA nbnxm backend (GMX_GPU=CUDA, sm_100, -use_fast_math) and the same libgromacs with the
# nbnxm_b200 shim linked in, driven by the same harness (oracle/ref_harness/bench_ref_gpu.cpp) on the same B200.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
H=oracle/_ref/cuda/bench_ref_gpu
NT=$(nproc)
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > gpurun_out/r02b_clocks.csv
run() {
    name=$1; shift
    for impl in stock shim; do
        lib=oracle/_ref/cuda/lib; [ $impl = shim ] && lib=oracle/_ref/cuda/lib_shim
        GMX_ENABLE_GPU_TIMING=1 LD_LIBRARY_PATH=$lib:$LD_LIBRARY_PATH timeout 900 $H "$@" --nt $NT --dump /tmp/f_${name}_$impl.bin \
            > gpurun_out/r02b_${name}_$impl.json 2> gpurun_out/r02b_${name}_$impl.err
        echo "$name $impl exit $?" >> gpurun_out/r02b_summary.log
        tail -n 3 gpurun_out/r02b_${name}_$impl.err >> gpurun_out/r02b_summary.log
    done
    python profiles/tools/compare_ref_gpu.py $name gpurun_out/r02b_${name}_stock.json gpurun_out/r02b_${name}_shim.json \
        /tmp/f_${name}_stock.bin /tmp/f_${name}_shim.bin >> gpurun_out/r02b_compare.jsonl 2>> gpurun_out/r02b_summary.log
}
rm -f gpurun_out/r02b_summary.log gpurun_out/r02b_compare.jsonl
run bench3k           --size 1    --rc 0.9 --vdw cut     --iter 200 --warmup 10
if ! grep -q natoms gpurun_out/r02b_bench3k_stock.json; then
    echo "stock harness failed on the smallest box; stopping"; cat gpurun_out/r02b_bench3k_stock.err | tail -40; cat gpurun_out/r02b_bench3k_shim.err | tail -40; exit 1
fi
run bench3k_energy    --size 1    --rc 0.9 --vdw cut     --energy 1 --iter 200 --warmup 10
run water96k_fswitch  --size 32   --rc 1.0 --vdw fswitch --energy 1 --rlist-outer 1.18 --rlist-inner 1.002 --dynamic-pruning 1 --iter 100 --warmup 12
run water384k_ljpme   --size 128  --rc 1.0 --vdw ljpme   --rlist-outer 1.18 --rlist-inner 1.002 --dynamic-pruning 1 --iter 100 --warmup 12
run water384k_pswitch --size 128  --rc 1.0 --vdw pswitch --rlist-outer 1.18 --rlist-inner 1.002 --dynamic-pruning 1 --iter 100 --warmup 12
run water1536k        --size 512  --rc 1.0 --vdw cut     --rlist-outer 1.18 --rlist-inner 1.002 --dynamic-pruning 1 --iter 60 --warmup 12
run water12m          --size 4096 --rc 1.2 --vdw cut     --rlist-outer 1.35 --rlist-inner 1.202 --dynamic-pruning 1 --iter 24 --warmup 12
# launch list of the stock backend (per-kernel durations without the reference's timers)
LD_LIBRARY_PATH=oracle/_ref/cuda/lib:$LD_LIBRARY_PATH timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/r02b_stock_launches_1536k.csv $H --size 512 --rc 1.0 --vdw cut --rlist-outer 1.18 --rlist-inner 1.002 \
    --dynamic-pruning 1 --iter 14 --warmup 2 --nt $NT > /dev/null 2>&1
cat gpurun_out/r02b_summary.log; cat gpurun_out/r02b_compare.jsonl
