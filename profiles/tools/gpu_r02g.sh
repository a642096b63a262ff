#!/bin/bash
# r02g: all single-GPU tests on the current library; A/B of the prune kernel's subtraction (ALU FADD vs FMA-pipe FFMA) and of
# the sort tile size at 12.3 M atoms.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -n 80 > gpurun_out/r02g_pytest_gpu.log; tail -n 5 gpurun_out/r02g_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 12 --no-cpu-baseline > gpurun_out/r02g_bench_12m_tile4096.json 2> gpurun_out/r02g_bench_12m_tile4096.err
NBNXM_B200_SCI_SORT_TILE=8192 timeout 900 python bench.py --steps 20 --warmup 12 --no-cpu-baseline > gpurun_out/r02g_bench_12m_tile8192.json 2> gpurun_out/r02g_bench_12m_tile8192.err
NBNXM_B200_SCI_SORT=global timeout 900 python bench.py --steps 20 --warmup 12 --no-cpu-baseline > gpurun_out/r02g_bench_12m_global.json 2> gpurun_out/r02g_bench_12m_global.err
timeout 600 python bench.py --workload water1536k --steps 40 --warmup 12 --no-cpu-baseline > gpurun_out/r02g_bench_1536k.json 2> gpurun_out/r02g_bench_1536k.err
python - <<'PY'
import json
for n in ("12m_tile4096", "12m_tile8192", "12m_global", "1536k"):
    try:
        d = json.loads(open("gpurun_out/r02g_bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, "ms/step %.4f kernel_us %.1f frac %.4f rolling_prune_us %.1f e2e_ms %.3f" % (d["ms_per_step"], d["roofline"]["kernel_us"], d["roofline"]["frac"], d["roofline"]["rolling_prune_us"], d["e2e"]["ms_per_step"]))
    except Exception as e:
        print(n, "failed", e)
PY
