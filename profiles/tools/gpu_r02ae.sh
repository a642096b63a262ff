#!/bin/bash
# r02ae: energy both-halves body with scalar accumulators for j forces and energies (ptxas gave the packed accumulators two moves per
# update): all GPU tests, the benched workloads, F+E on the 1.5 M-atom box
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -n 5 > gpurun_out/r02ae_pytest.log; tail -n 2 gpurun_out/r02ae_pytest.log
bench() { tag=$1; wl=$2; shift 2; en=""; case $tag in *_energy) en="--energy 1";; esac; env "$@" timeout 900 python bench.py --workload $wl $en --steps 40 --warmup 12 > gpurun_out/r02ae_bench_${wl}_$tag.json 2> gpurun_out/r02ae_bench_${wl}_$tag.err; }
bench asboth water96k_fswitch X=1
bench asboth_energy water1536k X=1
bench asboth water384k_pswitch X=1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02ae_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f[23:-5], "ms/step %.4f kernel_us %.1f frac %.4f e2e_ms %.4f parity" % (d["ms_per_step"], d["roofline"]["kernel_us"], d["roofline"]["frac"], d["e2e"]["ms_per_step"]), d.get("parity", {}).get("vs_oracle_sample"))
    except Exception as e:
        print(f, "failed", e)
PY
