#!/bin/bash
# r02ak: energy kernels at 96 registers / 20 warps per SM (-DNBNXM_PACKED_MIN_BLOCKS_ENERGY=20: 94 ... 122 bytes of spills, all outside the
# i-cluster chain) against 120 ... 126 registers / 16 warps (as shipped)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
bench() { tag=$1; wl=$2; shift 2; en=""; case $tag in *_energy) en="--energy 1";; esac; timeout 900 python bench.py --workload $wl $en --steps 40 --warmup 12 --no-cpu-baseline > gpurun_out/r02ak_bench_${wl}_$tag.json 2> gpurun_out/r02ak_bench_${wl}_$tag.err; }
bench w16 water96k_fswitch
bench w16_energy water1536k
bench w16_energy water384k_ljpme
touch gromacs_b200/csrc/*.cuh
make -s -j32 -C gromacs_b200/csrc EXTRA="-DNBNXM_PACKED_MIN_BLOCKS_ENERGY=20" > gpurun_out/r02ak_build.log 2>&1 || { echo "build failed"; tail -5 gpurun_out/r02ak_build.log; }
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -n 1
bench w20 water96k_fswitch
bench w20_energy water1536k
bench w20_energy water384k_ljpme
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02ak_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f[23:-5], "ms/step %.4f kernel_us %.1f frac %.4f" % (d["ms_per_step"], d["roofline"]["kernel_us"], d["roofline"]["frac"]))
    except Exception as e:
        print(f, "failed", e)
PY
