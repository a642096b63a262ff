#!/bin/bash
# r02h (8 GPUs): the multi-GPU parity tests at 2, 4 and 8 ranks (dynamic pruning, moving atoms, NCCL and peer-memory halo) and
# the 8-GPU bench lines with their parity figure.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/r02h_ngpus.txt
# 2 ranks on GPUs 0-1 and 4 ranks on GPUs 2-5 at the same time, then 8 ranks on the whole box
(CUDA_VISIBLE_DEVICES=0,1 timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "2-" 2>&1 | tail -n 30 > gpurun_out/r02h_pytest_world2.log) &
(CUDA_VISIBLE_DEVICES=2,3,4,5 timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "4-" 2>&1 | tail -n 30 > gpurun_out/r02h_pytest_world4.log) &
wait
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "8-" 2>&1 | tail -n 30 > gpurun_out/r02h_pytest_world8.log
tail -n 3 gpurun_out/r02h_pytest_world2.log gpurun_out/r02h_pytest_world4.log gpurun_out/r02h_pytest_world8.log
for wl in water12m water1536k; do
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus 8 --workload $wl --steps 20 --warmup 5 \
        > gpurun_out/r02h_bench_${wl}_n8.json 2> gpurun_out/r02h_bench_${wl}_n8.err
    tail -c 300 gpurun_out/r02h_bench_${wl}_n8.err; grep "^{" gpurun_out/r02h_bench_${wl}_n8.json | cut -c1-200
done
