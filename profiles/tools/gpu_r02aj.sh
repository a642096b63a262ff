#!/bin/bash
# r02aj: the reference's CUDA backend and this library (kernels as the round ends) behind the same callers of the reference
# (oracle/_ref/cuda/bench_ref_gpu, GMX_GPU=CUDA build of the unmodified tree; lib = stock backend, lib_shim = the shim + libnbnxm_b200.so)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
H=oracle/_ref/cuda/bench_ref_gpu
NT=$(nproc)
run() {
    name=$1; shift
    for impl in stock shim; do
        lib=oracle/_ref/cuda/lib; [ $impl = shim ] && lib=oracle/_ref/cuda/lib_shim
        GMX_ENABLE_GPU_TIMING=1 LD_LIBRARY_PATH=$lib:$LD_LIBRARY_PATH timeout 900 $H "$@" --nt $NT --dump /tmp/f_${name}_$impl.bin \
            > gpurun_out/r02aj_${name}_$impl.json 2> gpurun_out/r02aj_${name}_$impl.err
        echo "$name $impl exit $?" >> gpurun_out/r02aj_summary.log
    done
    python profiles/tools/compare_ref_gpu.py $name gpurun_out/r02aj_${name}_stock.json gpurun_out/r02aj_${name}_shim.json \
        /tmp/f_${name}_stock.bin /tmp/f_${name}_shim.bin >> gpurun_out/r02aj_compare.jsonl 2>> gpurun_out/r02aj_summary.log
}
rm -f gpurun_out/r02aj_summary.log gpurun_out/r02aj_compare.jsonl
run water96k_fswitch  --size 32   --rc 1.0 --vdw fswitch --energy 1 --rlist-outer 1.172 --rlist-inner 1.003 --nstlist-prune 10 --dynamic-pruning 1 --iter 100 --warmup 12
run water384k_ljpme   --size 128  --rc 1.0 --vdw ljpme   --rlist-outer 1.172 --rlist-inner 1.003 --nstlist-prune 10 --dynamic-pruning 1 --iter 100 --warmup 12
run water384k_pswitch --size 128  --rc 1.0 --vdw pswitch --rlist-outer 1.172 --rlist-inner 1.003 --nstlist-prune 10 --dynamic-pruning 1 --iter 100 --warmup 12
run water1536k        --size 512  --rc 1.0 --vdw cut     --rlist-outer 1.172 --rlist-inner 1.003 --nstlist-prune 10 --dynamic-pruning 1 --iter 60 --warmup 12
run water12m          --size 4096 --rc 1.2 --vdw cut     --rlist-outer 1.358 --rlist-inner 1.201 --nstlist-prune 10 --dynamic-pruning 1 --iter 24 --warmup 12
run water12m_energy   --size 4096 --rc 1.2 --vdw cut     --energy 1 --rlist-outer 1.358 --rlist-inner 1.201 --nstlist-prune 10 --dynamic-pruning 1 --iter 12 --warmup 12
cat gpurun_out/r02aj_summary.log; cut -c1-400 gpurun_out/r02aj_compare.jsonl
