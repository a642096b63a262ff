#!/bin/bash
# r02an (8 GPUs, what is left of the round's GPU budget): bench lines at 8 and 4 ranks with the kernels as the round ends,
# the chunk-pipelined slab step against the oracle at 8 ranks
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
run() { # gpus workload
    timeout 70 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus $1 --workload $2 --steps 20 --warmup 5 \
        > gpurun_out/r02an_bench_$2_n$1.json 2> gpurun_out/r02an_bench_$2_n$1.err
}
run 8 water12m
timeout 60 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "8-peer-pipelined" 2>&1 | tail -n 5 > gpurun_out/r02an_pytest_world8.log
run 8 water1536k
CUDA_VISIBLE_DEVICES=0,1,2,3 run 4 water12m
tail -n 2 gpurun_out/r02an_pytest_world8.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02an_bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f[23:-5], "ms/step %.4f value %.1f e2e_ms %.3f (plain %.3f) e2e %.1f" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["ms_per_step_copy_compute_copy"], d["e2e"]["value"]), {k: v for k, v in d["parity"].items() if k.endswith("n1")})
    except Exception as e:
        print(f, "failed", e)
PY
