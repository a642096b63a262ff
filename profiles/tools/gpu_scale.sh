#!/bin/bash
# usage: gpu_scale.sh <workload> N [N...]   (inside gpurun --gpus 8): strong-scaling bench lines
wl=$1; shift
mkdir -p gpurun_out
for n in "$@"; do
  if [ "$n" = 1 ]; then
    timeout 900 python bench.py --workload $wl --steps 20 --no-cpu-baseline > gpurun_out/scale_${wl}_n$n.json 2> gpurun_out/scale_${wl}_n$n.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --workload $wl --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_${wl}_n$n.json 2> gpurun_out/scale_${wl}_n$n.err
  fi
  tail -c 1500 gpurun_out/scale_${wl}_n$n.json | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=$n', 'ms_per_step %.4f value %.1f' % (d['ms_per_step'], d['value']), {k:v for k,v in d.items() if k in ('halo','roofline')})
except Exception as e: print('N=$n failed', e)
"
  tail -3 gpurun_out/scale_${wl}_n$n.err
done
