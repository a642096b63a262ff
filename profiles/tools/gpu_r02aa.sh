#!/bin/bash
# r02aa: (1) the both-halves body with FSETP + FSEL cut-off masks and the pmeCorr polynomials divided by num[6] (3 instructions
# less per 64 pairs), (2) the rolling prune as background work on a lowest-priority stream beside the force kernel.
# A/B: new default | new kernels with the prune in line | kernels and prune as before (rebuilt with the old forms).
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -n 15 > gpurun_out/r02aa_pytest_gpu.log; tail -n 3 gpurun_out/r02aa_pytest_gpu.log
bench() { tag=$1; wl=$2; shift 2; env "$@" timeout 900 python bench.py --workload $wl --steps 20 --warmup 12 --no-cpu-baseline > gpurun_out/r02aa_bench_${wl}_$tag.json 2> gpurun_out/r02aa_bench_${wl}_$tag.err; }
for wl in water12m water1536k water96k_fswitch water384k_ljpme water384k_pswitch; do bench new $wl X=1; done
for wl in water12m water1536k water96k_fswitch; do bench new_inline $wl NBNXM_B200_BACKGROUND_PRUNE=0; done
touch gromacs_b200/csrc/*.cuh
make -s -j32 -C gromacs_b200/csrc EXTRA="-DNBNXM_PACKED_NO_SELP -DNBNXM_PACKED_PLAIN_PMECORR" > gpurun_out/r02aa_build.log 2>&1 || { echo "build failed"; tail -5 gpurun_out/r02aa_build.log; }
for wl in water12m water1536k water96k_fswitch water384k_ljpme water384k_pswitch; do bench old_inline $wl NBNXM_B200_BACKGROUND_PRUNE=0; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02aa_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f[23:-5], "ms/step %.4f kernel_us %.1f prune_us %.1f frac %.4f e2e_ms %.4f" % (d["ms_per_step"], d["roofline"]["kernel_us"], d["roofline"]["rolling_prune_us"], d["roofline"]["frac"], d["e2e"]["ms_per_step"]))
    except Exception as e:
        print(f, "failed", e)
PY
