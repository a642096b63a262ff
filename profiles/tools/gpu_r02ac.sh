#!/bin/bash
# r02ac: energy kernels take the Ewald real-space force of pairs without exclusions from the erfc / exp of the energy instead
# of the rational correction (A/B: rebuilt with -DNBNXM_PACKED_ENERGY_PMECORR)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_benched_configs.py tests/test_gpu_fullsize.py tests/test_gpu_boundary_corners.py tests/test_gpu_zz_elec_none.py -m gpu -q -x 2>&1 | tail -n 5 > gpurun_out/r02ac_pytest.log; tail -n 2 gpurun_out/r02ac_pytest.log
cp gpurun_out/parity_errors.jsonl gpurun_out/r02ac_parity_errors.jsonl 2>/dev/null
bench() { tag=$1; wl=$2; shift 2; en=""; case $tag in *_energy) en="--energy 1";; esac; env "$@" timeout 900 python bench.py --workload $wl $en --steps 40 --warmup 12 > gpurun_out/r02ac_bench_${wl}_$tag.json 2> gpurun_out/r02ac_bench_${wl}_$tag.err; }
bench erfc water96k_fswitch X=1
bench erfc_energy water1536k X=1
touch gromacs_b200/csrc/*.cuh
make -s -j32 -C gromacs_b200/csrc EXTRA="-DNBNXM_PACKED_ENERGY_PMECORR" > gpurun_out/r02ac_build.log 2>&1 || { echo "build failed"; tail -5 gpurun_out/r02ac_build.log; }
bench pmecorr water96k_fswitch X=1
bench pmecorr_energy water1536k X=1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02ac_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f[23:-5], "ms/step %.4f kernel_us %.1f frac %.4f e2e_ms %.4f parity" % (d["ms_per_step"], d["roofline"]["kernel_us"], d["roofline"]["frac"], d["e2e"]["ms_per_step"]), d.get("parity", {}).get("vs_oracle_sample"))
    except Exception as e:
        print(f, "failed", e)
PY
