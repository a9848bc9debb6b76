#!/bin/bash
# quick GPU check: GEMM selftest (LayerNorm cases), PoseNet parity tests, one PoseNet bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-quick}
timeout 300 tools/gemm_selftest > gpurun_out/${TAG}_selftest.log 2>&1
grep -A2 "LayerNorm M" gpurun_out/${TAG}_selftest.log | cut -c1-400
grep -E "SELFTEST|FAIL" gpurun_out/${TAG}_selftest.log | head
timeout 600 python -m pytest tests/test_gpu_posenet.py -x -q -m gpu > gpurun_out/${TAG}_tests.log 2>&1; tail -3 gpurun_out/${TAG}_tests.log
timeout 600 python bench.py --config posenet --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("clips/s %.2f ms/step %.1f graph %.4f ms" % (d["value"], d["ms_per_step"], r["forward_graph_ms"]), r["forward_ms_by_kernel_class"], "frac %.3f graphfrac %.3f" % (r["frac"], r["frac_from_graph_share"]))
except Exception as e:
    print("bench ERR", e, open("gpurun_out/${TAG}_bench.err").read()[-1500:])
PY
