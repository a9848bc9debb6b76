#!/bin/bash
# Quick GPU check used while iterating: args = tag, then a pytest -k expression (or "" for none)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-q}
KEXPR=${2:-}
if [ -n "$KEXPR" ]; then
  timeout 900 python -m pytest tests -q -m gpu -s -x -k "$KEXPR" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests exit $?"; tail -5 gpurun_out/${TAG}_tests.log
fi
ROHM_B200_ATTN_TS=1 timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("clips/s", d["value"], "e2e", d["e2e"]["value"], "graph ms", d.get("step_graph_ms"), "frac", d["roofline"]["frac"])
PY
grep -h "timeline" gpurun_out/${TAG}_bench.err | tail -2
