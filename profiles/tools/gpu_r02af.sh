#!/bin/bash
# r02af (2 GPUs): multi-GPU paths with the kernels, stream priorities and the device column sort of this session: parity tests at 2 ranks (NCCL, peer memory, chunk-pipelined), bench lines incl. the reference arm under torchrun
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "2-" 2>&1 | tail -n 40 > gpurun_out/r02af_pytest.log; tail -n 5 gpurun_out/r02af_pytest.log
run() { # tag workload extra
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 2 --workload $2 --steps 20 --warmup 5 $3 \
        > gpurun_out/r02af_bench_$2_n2_$1.json 2> gpurun_out/r02af_bench_$2_n2_$1.err
    tail -c 600 gpurun_out/r02af_bench_$2_n2_$1.err | grep -v "OMP_NUM_THREADS\|^\*\*\*\|NCCL version"
}
run pipe water1536k ""
run pipe water12m ""
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29554 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02af_bench_reference_n2.json 2> gpurun_out/r02af_bench_reference_n2.err; tail -c 300 gpurun_out/r02af_bench_reference_n2.json
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02af_bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f[22:-5], "ms/step %.4f e2e_ms %.3f (plain %.3f)" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["ms_per_step_copy_compute_copy"]), {k: v for k, v in d["parity"].items() if k.endswith("n1")}, d["search_step"])
    except Exception as e:
        print(f, "failed", e)
PY
