#!/bin/bash
# 2-GPU sanity check of the default multi-GPU path exactly as the driver launches it, plus the multi-GPU parity tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -k "2" > gpurun_out/pytest_multi2.log 2>&1; tail -2 gpurun_out/pytest_multi2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/drv_n2.json 2> gpurun_out/drv_n2.err
echo "stdout lines: $(wc -l < gpurun_out/drv_n2.json)"; head -c 400 gpurun_out/drv_n2.json; echo; grep -v "^\*\|OMP_NUM" gpurun_out/drv_n2.err | tail -3 | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/drv_ref_n2.json 2> gpurun_out/drv_ref_n2.err
echo "reference stdout lines: $(wc -l < gpurun_out/drv_ref_n2.json)"; head -c 300 gpurun_out/drv_ref_n2.json; echo
