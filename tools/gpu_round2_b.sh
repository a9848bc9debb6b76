#!/bin/bash
# GPU call B (round 2): fused-LayerNorm GEMM epilogue: selftest, PoseNet parity tests, bench with / without the fusion.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 tools/gemm_selftest > gpurun_out/r2b_selftest.log 2>&1; echo "selftest exit $?" >> gpurun_out/r2b_selftest.log
grep -E "LayerNorm|timing|FAIL|SELFTEST|exit" gpurun_out/r2b_selftest.log | head -40
timeout 900 python -m pytest tests/test_gpu_posenet.py tests/test_gpu_pipeline.py -x -q -m gpu -s > gpurun_out/r2b_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2b_tests.log
tail -5 gpurun_out/r2b_tests.log
timeout 600 python bench.py --config posenet --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_fused.json 2> gpurun_out/r2b_bench_fused.err; echo "bench fused exit $?"
ROHM_B200_FUSED_LN=0 timeout 600 python bench.py --config posenet --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_unfused.json 2> gpurun_out/r2b_bench_unfused.err; echo "bench unfused exit $?"
python - <<'PY'
import json
for n in ("fused", "unfused"):
    try:
        d = json.loads(open(f"gpurun_out/r2b_bench_{n}.json").read().strip().splitlines()[-1])
        r = d["roofline"]
        print(n, "clips/s %.2f" % d["value"], "ms/step %.1f" % d["ms_per_step"], "graph ms %.4f" % r["forward_graph_ms"], r["forward_ms_by_kernel_class"], r["launches_by_kernel_class"])
    except Exception as e:
        print(n, "ERR", e)
PY
