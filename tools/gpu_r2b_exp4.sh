#!/bin/bash
# Round-2 second session, call 5: pitched TMA vertex stores of the fused LBS launch (A/B against the dense 4-byte stores), the
# register-resident GroupNorm kernel.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r2b5}
timeout -k 10 1200 python -m pytest tests/test_gpu_body.py tests/test_gpu_glue.py tests/test_gpu_trajnet.py -q -m gpu -x > gpurun_out/${TAG}_tests.log 2>&1; echo "body+glue+trajnet tests exit $?"; tail -3 gpurun_out/${TAG}_tests.log
run_bench() {  # name, config, env...
  local name=$1 cfg=$2; shift 2
  env "$@" timeout -k 10 900 python bench.py --config $cfg --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_${name}.json 2> gpurun_out/${TAG}_bench_${name}.err
  echo "bench $name exit $?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_${name}.json").read().strip().splitlines()[-1])
    r = d.get("roofline", {})
    print("  ${name}: value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "ms/step", round(d["ms_per_step"], 4),
          "frac", round(r.get("frac") or 0, 4), "call/fwd ms", r.get("call_ms") or r.get("forward_graph_ms") or r.get("forward_ms"), "clocks", d.get("clocks", {}).get("sm_mhz"))
except Exception as e:
    print("  ${name}: no line:", e)
PY
}
run_bench lbs lbs ROHM_B200_X=0
run_bench lbs_dense lbs ROHM_B200_LBS_TMA_STORE=0
run_bench trajcontrol trajcontrol ROHM_B200_X=0
ROHM_B200_LBS_TS=1 timeout -k 10 300 python tools/profile_lbs.py 4 2> gpurun_out/${TAG}_lbs_timeline.txt > /dev/null; grep -h timeline gpurun_out/${TAG}_lbs_timeline.txt
du -sh gpurun_out
