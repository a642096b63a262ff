#!/bin/bash
# r02q (8 GPUs): multi-GPU parity at 8 ranks on the three step variants (NCCL, peer memory, peer memory chunk-pipelined) and at 4
# ranks pipelined; bench lines at 8 and 4 GPUs with the search step on every rank's GPU and the pipelined end-to-end step.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "8-" 2>&1 | tail -n 30 > gpurun_out/r02q_pytest_world8.log
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "4-peer-pipelined" 2>&1 | tail -n 30 > gpurun_out/r02q_pytest_world4.log
tail -n 3 gpurun_out/r02q_pytest_world8.log gpurun_out/r02q_pytest_world4.log
run() { # gpus workload
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus $1 --workload $2 --steps 20 --warmup 5 \
        > gpurun_out/r02q_bench_$2_n$1.json 2> gpurun_out/r02q_bench_$2_n$1.err
    tail -c 500 gpurun_out/r02q_bench_$2_n$1.err | grep -v "OMP_NUM_THREADS\|^\*\*\*\|NCCL version"
}
run 8 water12m
run 4 water12m
run 8 water1536k
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02q_bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f[22:-5], "ms/step %.4f value %.1f e2e_ms %.3f (plain %.3f) e2e %.1f" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["ms_per_step_copy_compute_copy"], d["e2e"]["value"]), {k: v for k, v in d["parity"].items() if k.endswith("n1")}, d["search_step"])
    except Exception as e:
        print(f, "failed", e)
PY
