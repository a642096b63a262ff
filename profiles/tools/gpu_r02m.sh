#!/bin/bash
# r02m (2 GPUs): the multi-GPU bench path with the search step of every rank on its GPU (device slab lists), next to the host plan
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
run() { # tag workload extra
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --workload $2 --steps 20 --warmup 5 $3 \
        > gpurun_out/r02m_bench_$2_n2_$1.json 2> gpurun_out/r02m_bench_$2_n2_$1.err
    tail -c 500 gpurun_out/r02m_bench_$2_n2_$1.err | grep -v "OMP_NUM_THREADS\|^\*\*\*\|NCCL version" ; grep "^{" gpurun_out/r02m_bench_$2_n2_$1.json | cut -c1-160
}
run device water1536k ""
run host water1536k "--slab-lists host"
run device water12m ""
run device_nccl water1536k "--halo nccl"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02m_bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f[22:-5], "ms/step %.4f e2e_ms %.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"]), {k: v for k, v in d["parity"].items() if k.endswith("n1")}, d.get("search_step"))
    except Exception as e:
        print(f, "failed", e)
PY
