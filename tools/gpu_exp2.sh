#!/bin/bash
# scratch: args = tag, pytest -k expression, bench config list (space separated, may be empty)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-x}
if [ -n "$2" ]; then
  timeout 1500 python -m pytest tests -q -m gpu -s -x -k "$2" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/${TAG}_tests.log
fi
for c in $3; do
  timeout 900 python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err; echo "bench $c exit $?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_$c.json").read().strip().splitlines()[-1])
    print("$c", d["value"], d["unit"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], "achieved", d["roofline"]["achieved"])
except Exception as ex:
    print("no bench line", ex)
PY
done
