#!/bin/bash
# r02ag (2 GPUs): end-to-end slab step at 12.3 M atoms, 16 against 24 chunks per rank, stream priorities on / off
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
run() { # tag chunks env
    env $3 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 2 --workload water12m --steps 20 --warmup 5 --e2e-chunks $2 \
        > gpurun_out/r02ag_bench_water12m_n2_$1.json 2> gpurun_out/r02ag_bench_water12m_n2_$1.err
}
run c16_prio 16 X=1
run c24_prio 24 X=1
run c16_flat 16 NBNXM_B200_STREAM_PRIORITIES=0
run c24_flat 24 NBNXM_B200_STREAM_PRIORITIES=0
run c12_prio 12 X=1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02ag_bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f[22:-5], "ms/step %.4f e2e_ms %.3f (plain %.3f)" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["ms_per_step_copy_compute_copy"]))
    except Exception as e:
        print(f, "failed", e)
PY
