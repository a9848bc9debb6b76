#!/bin/bash
# scratch experiment runner: args = tag, pytest -k expression ("" = none), 1 = run the GEMM selftest, 1 = run the default bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-x}
if [ -n "$2" ]; then
  timeout 1200 python -m pytest tests -q -m gpu -s -x -k "$2" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/${TAG}_tests.log
fi
if [ "$3" = "1" ]; then
  timeout 300 tools/gemm_selftest > gpurun_out/${TAG}_selftest.log 2>&1; echo "selftest exit $?"; grep -c OK gpurun_out/${TAG}_selftest.log; grep "FAIL" gpurun_out/${TAG}_selftest.log | head
fi
if [ "$4" = "1" ]; then
  timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
  python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("clips/s", d["value"], "e2e", d["e2e"]["value"], "graph ms", d.get("step_graph_ms"), "frac", d["roofline"]["frac"], d.get("forward_ms_by_kernel_class"))
PY
fi
