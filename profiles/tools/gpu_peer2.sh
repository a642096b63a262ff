#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -k "2" > gpurun_out/pytest_multi.log 2>&1; echo "exit $?" >> gpurun_out/pytest_multi.log; tail -5 gpurun_out/pytest_multi.log
for h in peer nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --halo $h > gpurun_out/n2_$h.json 2> gpurun_out/n2_$h.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/n2_$h.json").read().strip().splitlines()[-1]); print("$h", "ms/step %.4f e2e ms %.4f"%(d["ms_per_step"], d["e2e"]["ms_per_step"]))
except Exception as e:
    print("$h failed", e)
PY
grep -v "^\*\|OMP_NUM" gpurun_out/n2_$h.err | tail -4 | cut -c1-300
done
