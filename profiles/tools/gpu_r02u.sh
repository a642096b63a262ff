#!/bin/bash
# r02u: list-splitting granularity (gpu_min_ci_balanced target) on the small boxes
cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -k pipelined 2>&1 | tail -n 3
mkdir -p gpurun_out
for spec in "water96k_fswitch 0" "water96k_fswitch 24000" "water96k_fswitch 36000" "water384k_pswitch 0" "water384k_pswitch 36000" "water384k_ljpme 36000"; do
    set -- $spec
    python bench.py --workload $1 --steps 40 --warmup 12 --no-cpu-baseline --min-sci $2 > gpurun_out/r02u_$1_$2.json 2>/dev/null
    python - "$1" "$2" <<'PY'
import json, sys
d = json.loads(open("gpurun_out/r02u_%s_%s.json" % (sys.argv[1], sys.argv[2])).read().strip().splitlines()[-1])
print(sys.argv[1], "min_sci", sys.argv[2], "nsci", d["config"]["nsci"], "ms/step %.4f kernel_us %.1f frac %.4f" % (d["ms_per_step"], d["roofline"]["kernel_us"], d["roofline"]["frac"]))
PY
done
