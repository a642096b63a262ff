#!/bin/bash
# usage: gpu_scale2.sh <workload> <halo modes...> -- N...
wl=$1; shift
modes=()
while [ "$1" != "--" ]; do modes+=("$1"); shift; done; shift
mkdir -p gpurun_out
for n in "$@"; do for h in "${modes[@]}"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $n --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --halo $h > gpurun_out/scale_${wl}_${h}_n$n.json 2> gpurun_out/scale_${wl}_${h}_n$n.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_${wl}_${h}_n$n.json").read().strip().splitlines()[-1]); print("N=$n $h", "ms/step %.4f value %.1f e2e ms %.4f"%(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"]))
except Exception as e:
    print("N=$n $h failed", e)
PY
done; done
