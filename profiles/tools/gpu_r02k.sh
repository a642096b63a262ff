#!/bin/bash
# r02k: A/B of the cjPacked descriptor staging in the packed force kernel - coalesced loads + st.shared by all lanes (default)
# against cp.async.bulk + mbarrier, two-stage ring (-DNBNXM_PACKED_TMA_DESC) - and of the number of streams the chunk kernels of
# the pipelined end-to-end step rotate over.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
bench() { # tag workload extra-args
    timeout 900 python bench.py --workload $2 --steps 20 --warmup 12 --no-cpu-baseline $3 > gpurun_out/r02k_bench_$2_$1.json 2> gpurun_out/r02k_bench_$2_$1.err
}
for ns in 2 3 4; do
    NBNXM_B200_PIPE_STREAMS=$ns bench streams$ns water12m
done
bench ldg water12m; bench ldg water1536k; bench ldg water96k_fswitch
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nbnxm_force_kernel -s 4 -c 1 -f -o gpurun_out/r02k_prof_1536k_ldg \
    python bench.py --workload water1536k --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02k_ncu_ldg.log 2>&1
touch gromacs_b200/csrc/*.cuh
make -s -j32 -C gromacs_b200/csrc EXTRA="-DNBNXM_PACKED_TMA_DESC" > gpurun_out/r02k_build_tma.log 2>&1 || { echo "TMA build failed"; tail -5 gpurun_out/r02k_build_tma.log; }
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -n 5 > gpurun_out/r02k_pytest_tma.log; tail -n 2 gpurun_out/r02k_pytest_tma.log
bench tma water12m; bench tma water1536k; bench tma water96k_fswitch
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nbnxm_force_kernel -s 4 -c 1 -f -o gpurun_out/r02k_prof_1536k_tma \
    python bench.py --workload water1536k --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02k_ncu_tma.log 2>&1
for v in ldg tma; do
    ncu -i gpurun_out/r02k_prof_1536k_$v.ncu-rep --page raw --csv > gpurun_out/r02k_prof_1536k_$v.csv 2>/dev/null
    python profiles/tools/ncu_summary.py gpurun_out/r02k_prof_1536k_$v.csv > gpurun_out/r02k_prof_1536k_$v.txt 2>&1
    rm -f gpurun_out/r02k_prof_1536k_$v.ncu-rep
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02k_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f[22:-5], "ms/step %.4f kernel_us %.1f frac %.4f e2e_ms %.3f" % (d["ms_per_step"], d["roofline"]["kernel_us"], d["roofline"]["frac"], d["e2e"]["ms_per_step"]))
    except Exception as e:
        print(f, "failed", e)
PY
