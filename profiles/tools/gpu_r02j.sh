#!/bin/bash
# r02j: the pipelined end-to-end step with tapered chunk widths and the rolling prune overlapped with the force kernels
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_benched_configs.py -m gpu -q 2>&1 | tail -n 15 > gpurun_out/r02j_pytest.log; tail -n 3 gpurun_out/r02j_pytest.log
python profiles/tools/pipeline_timeline.py water12m 24 > gpurun_out/r02j_timeline_12m_24.jsonl 2> gpurun_out/r02j_timeline.err
python profiles/tools/pipeline_timeline.py water12m 32 > gpurun_out/r02j_timeline_12m_32.jsonl 2>> gpurun_out/r02j_timeline.err
for ch in 24 32; do
    timeout 900 python bench.py --steps 20 --warmup 12 --no-cpu-baseline --e2e-chunks $ch > gpurun_out/r02j_bench_12m_ch$ch.json 2> gpurun_out/r02j_bench_12m_ch$ch.err
done
timeout 600 python bench.py --workload water1536k --steps 40 --warmup 12 --no-cpu-baseline > gpurun_out/r02j_bench_1536k.json 2> gpurun_out/r02j_bench_1536k.err
python - <<'PY'
import json
for f in ("r02j_timeline_12m_24", "r02j_timeline_12m_32"):
    for l in open("gpurun_out/%s.jsonl" % f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, "step", d["step"], "prune", d["rolling_prune"], "step_ms %.3f first_k_start %.3f last_k_end %.3f last_d2h %.3f" % (d["step_ms"], d["first_kernel_start_ms"], d["last_kernel_end_ms"], d["last_d2h_end_ms"]))
for n in ("12m_ch24", "12m_ch32", "1536k"):
    try:
        d = json.loads(open("gpurun_out/r02j_bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, "ms/step %.4f e2e_ms %.3f e2e %.1f value %.1f" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["value"]))
    except Exception as e:
        print(n, "failed", e)
PY
