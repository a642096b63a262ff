#!/bin/bash
# r02p (2 GPUs): the chunk-pipelined end-to-end step of a slab (peer-memory halo): parity test at 2 ranks, bench lines
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "2-" 2>&1 | tail -n 40 > gpurun_out/r02p_pytest.log; tail -n 5 gpurun_out/r02p_pytest.log
run() { # tag workload extra
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 2 --workload $2 --steps 20 --warmup 5 $3 \
        > gpurun_out/r02p_bench_$2_n2_$1.json 2> gpurun_out/r02p_bench_$2_n2_$1.err
    tail -c 600 gpurun_out/r02p_bench_$2_n2_$1.err | grep -v "OMP_NUM_THREADS\|^\*\*\*\|NCCL version"
}
run pipe water1536k ""
run pipe water12m ""
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02p_bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f[22:-5], "ms/step %.4f e2e_ms %.3f (plain %.3f)" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["ms_per_step_copy_compute_copy"]), {k: v for k, v in d["parity"].items() if k.endswith("n1")}, d["search_step"])
    except Exception as e:
        print(f, "failed", e)
PY
