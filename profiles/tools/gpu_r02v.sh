#!/bin/bash
# r02v: prune kernel with packed f32x2 distances (default) against the scalar form (-DNBNXM_PRUNE_SCALAR): bit-exact masks, timings, ncu
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_benched_configs.py tests/test_gpu_search.py -m gpu -q 2>&1 | tail -n 6 > gpurun_out/r02v_pytest.log; tail -n 2 gpurun_out/r02v_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
bench() { timeout 900 python bench.py --workload $2 --steps 20 --warmup 12 --no-cpu-baseline > gpurun_out/r02v_bench_$2_$1.json 2> gpurun_out/r02v_bench_$2_$1.err; }
bench packed water12m; bench packed water1536k
timeout 600 ncu --set full --clock-control none -k regex:nbnxm_prune_kernel -s 3 -c 1 -f -o gpurun_out/r02v_prof_prune12m \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02v_ncu.log 2>&1
ncu -i gpurun_out/r02v_prof_prune12m.ncu-rep --page raw --csv > gpurun_out/r02v_prof_prune12m.csv 2>/dev/null
python profiles/tools/ncu_summary.py gpurun_out/r02v_prof_prune12m.csv > gpurun_out/r02v_prof_prune12m_packed.txt 2>&1
rm -f gpurun_out/r02v_prof_prune12m.ncu-rep gpurun_out/r02v_prof_prune12m.csv
touch gromacs_b200/csrc/nbnxm_prune.cu
make -s -j32 -C gromacs_b200/csrc EXTRA="-DNBNXM_PRUNE_SCALAR" > gpurun_out/r02v_build_scalar.log 2>&1 || { echo "scalar build failed"; tail -5 gpurun_out/r02v_build_scalar.log; }
bench scalar water12m; bench scalar water1536k
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02v_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f[22:-5], "ms/step %.4f kernel_us %.1f rolling_prune_us %.1f first_pass_prune_ms %.3f e2e_ms %.3f vws %.1f" % (d["ms_per_step"], d["roofline"]["kernel_us"], d["roofline"]["rolling_prune_us"], d["search_step"].get("first_pass_prune_ms", -1), d["e2e"]["ms_per_step"], d["value_with_search"] or -1))
    except Exception as e:
        print(f, "failed", e)
PY
grep -E "gpu__time_duration|inst_executed.sum|issue_active|pipe_alu|pipe_fma_cycles" gpurun_out/r02v_prof_prune12m_packed.txt
