#!/bin/bash
# quick LBS check while iterating: vertex tests, CTA-0 timeline, bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-q}
timeout 900 python -m pytest tests -q -m gpu -x -k "vertices or lbs or dense or from_repr or edge_rot or fused" 2>&1 | tail -3
ROHM_B200_LBS_TS=1 timeout 300 python tools/profile_lbs.py 4 2>&1 | grep timeline
timeout 900 python bench.py --config lbs --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_lbs.json 2> gpurun_out/${TAG}_bench_lbs.err; echo "bench lbs exit $?"
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_lbs.json").read().strip().splitlines()[-1])
print("lbs", d["value"], "ms/step", d["ms_per_step"], "call_ms", d["roofline"]["call_ms"], "frac", d["roofline"]["frac"])
PY
