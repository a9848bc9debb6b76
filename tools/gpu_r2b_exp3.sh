#!/bin/bash
# Round-2 second session, call 4: every GPU test on the tree with the GroupNorm statistics inside the GroupNorm kernel and the
# weight-tile prefetch ahead of the setup barrier; A/B bench lines.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r2b4}
timeout -k 10 300 tools/gemm_selftest > gpurun_out/${TAG}_selftest.log 2>&1; echo "selftest exit $?"; grep -c OK gpurun_out/${TAG}_selftest.log; grep -c -i "fail\|mismatch" gpurun_out/${TAG}_selftest.log
timeout -k 10 2400 python -m pytest tests -q -m gpu -x > gpurun_out/${TAG}_tests.log 2>&1; echo "gpu tests exit $?"; tail -3 gpurun_out/${TAG}_tests.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/${TAG}_smoke.log
run_bench() {  # name, config, env...
  local name=$1 cfg=$2; shift 2
  env "$@" timeout -k 10 900 python bench.py --config $cfg --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_${name}.json 2> gpurun_out/${TAG}_bench_${name}.err
  echo "bench $name exit $?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_${name}.json").read().strip().splitlines()[-1])
    r = d.get("roofline", {})
    print("  ${name}: value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "ms/step", round(d["ms_per_step"], 2),
          "frac", round(r.get("frac") or 0, 4), "fwd graph ms", r.get("forward_graph_ms") or r.get("forward_ms"), "clocks", d.get("clocks", {}).get("sm_mhz"))
except Exception as e:
    print("  ${name}: no line:", e)
PY
}
run_bench trajcontrol trajcontrol ROHM_B200_X=0
run_bench trajcontrol_gn_epilogue trajcontrol ROHM_B200_TRAJ_GN_EPILOGUE=1
run_bench posenet posenet ROHM_B200_X=0
timeout -k 10 300 python tools/profile_trajnet.py > gpurun_out/${TAG}_profile_trajnet.txt 2>&1; tail -2 gpurun_out/${TAG}_profile_trajnet.txt
ROHM_B200_GRAPH=0 ROHM_B200_TRAJ_TS=diff_enc1.c2,diff_enc2.c2,diff_mid_block1.c1,diff_dec1.c1 timeout -k 10 300 python tools/profile_target_trajnet.py 5 2>&1 | grep "timeline" > gpurun_out/${TAG}_conv_timelines.txt; cat gpurun_out/${TAG}_conv_timelines.txt
du -sh gpurun_out
