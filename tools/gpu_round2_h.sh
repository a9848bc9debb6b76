#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py --config posenet --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2h_bench.json").read().strip().splitlines()[-1]); r = d["roofline"]
print("clips/s %.2f ms/step %.1f fwd graph %.4f step graph %.4f host us %.1f" % (d["value"], d["ms_per_step"], r["forward_graph_ms"], r["step_graph_ms"], r["host_enqueue_us_per_step"]))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 140 -c 200 --csv --log-file gpurun_out/r2h_launches.csv python tools/profile_target.py 4 > gpurun_out/r2h_ncu.log 2>&1
python tools/launch_list_summary.py gpurun_out/r2h_launches.csv 2>/dev/null | head -40
