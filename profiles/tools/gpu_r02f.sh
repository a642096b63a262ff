#!/bin/bash
# r02f (2 GPUs): multi-GPU parity tests at 2 ranks (dynamic pruning, moving atoms, both transports), 2-GPU bench lines with
# the parity figure, boundary-corner tests, and the stock-vs-shim comparison with the shift vectors in place.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_boundary_corners.py -m gpu -q 2>&1 | tail -n 60 > gpurun_out/r02f_pytest.log; tail -n 6 gpurun_out/r02f_pytest.log
for wl in water1536k water12m; do
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --workload $wl --steps 20 --warmup 5 \
        > gpurun_out/r02f_bench_${wl}_n2.json 2> gpurun_out/r02f_bench_${wl}_n2.err
    tail -c 400 gpurun_out/r02f_bench_${wl}_n2.err; cut -c1-300 gpurun_out/r02f_bench_${wl}_n2.json
done
H=oracle/_ref/cuda/bench_ref_gpu
NT=$(nproc)
run() {
    name=$1; shift
    for impl in stock shim; do
        lib=oracle/_ref/cuda/lib; [ $impl = shim ] && lib=oracle/_ref/cuda/lib_shim
        GMX_ENABLE_GPU_TIMING=1 LD_LIBRARY_PATH=$lib:$LD_LIBRARY_PATH timeout 900 $H "$@" --nt $NT --dump /tmp/f_${name}_$impl.bin \
            > gpurun_out/r02f_${name}_$impl.json 2> gpurun_out/r02f_${name}_$impl.err
        echo "$name $impl exit $?" >> gpurun_out/r02f_summary.log
    done
    python profiles/tools/compare_ref_gpu.py $name gpurun_out/r02f_${name}_stock.json gpurun_out/r02f_${name}_shim.json \
        /tmp/f_${name}_stock.bin /tmp/f_${name}_shim.bin >> gpurun_out/r02f_compare.jsonl 2>> gpurun_out/r02f_summary.log
}
rm -f gpurun_out/r02f_summary.log gpurun_out/r02f_compare.jsonl
run bench3k           --size 1    --rc 0.9 --vdw cut     --iter 200 --warmup 10
run bench3k_energy    --size 1    --rc 0.9 --vdw cut     --energy 1 --iter 200 --warmup 10
run bench3k_cutlb     --size 1    --rc 0.9 --vdw cutlb   --energy 1 --iter 100 --warmup 10
run water96k_fswitch  --size 32   --rc 1.0 --vdw fswitch --energy 1 --rlist-outer 1.172 --rlist-inner 1.003 --nstlist-prune 10 --dynamic-pruning 1 --iter 100 --warmup 12
run water384k_ljpme   --size 128  --rc 1.0 --vdw ljpme   --rlist-outer 1.172 --rlist-inner 1.003 --nstlist-prune 10 --dynamic-pruning 1 --iter 100 --warmup 12
run water384k_pswitch --size 128  --rc 1.0 --vdw pswitch --rlist-outer 1.172 --rlist-inner 1.003 --nstlist-prune 10 --dynamic-pruning 1 --iter 100 --warmup 12
run water1536k        --size 512  --rc 1.0 --vdw cut     --rlist-outer 1.172 --rlist-inner 1.003 --nstlist-prune 10 --dynamic-pruning 1 --iter 60 --warmup 12
run water12m          --size 4096 --rc 1.2 --vdw cut     --rlist-outer 1.358 --rlist-inner 1.201 --nstlist-prune 10 --dynamic-pruning 1 --iter 24 --warmup 12
run water12m_energy   --size 4096 --rc 1.2 --vdw cut     --energy 1 --rlist-outer 1.358 --rlist-inner 1.201 --nstlist-prune 10 --dynamic-pruning 1 --iter 12 --warmup 12
cat gpurun_out/r02f_summary.log; cut -c1-330 gpurun_out/r02f_compare.jsonl
