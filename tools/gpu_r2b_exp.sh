#!/bin/bash
# Round-2 second session: A/B run of the split-K TrajNet convolutions, the split Q/K load of the tcgen05 attention and the
# programmatic-dependent-launch edges of the small PoseNet kernels.  Outputs under gpurun_out/<tag>_*.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r2b}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout -k 10 900 python -m pytest tests/test_gpu_trajnet.py -q -m gpu -x > gpurun_out/${TAG}_tests_trajnet.log 2>&1; echo "trajnet tests exit $?"; tail -3 gpurun_out/${TAG}_tests_trajnet.log
timeout -k 10 900 python -m pytest tests/test_gpu_posenet.py -q -m gpu -x > gpurun_out/${TAG}_tests_posenet.log 2>&1; echo "posenet tests exit $?"; tail -3 gpurun_out/${TAG}_tests_posenet.log
timeout -k 10 1500 python -m pytest tests -q -m gpu --deselect tests/test_gpu_trajnet.py --deselect tests/test_gpu_posenet.py > gpurun_out/${TAG}_tests_rest.log 2>&1; echo "other gpu tests exit $?"; tail -3 gpurun_out/${TAG}_tests_rest.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/${TAG}_smoke.log
run_bench() {  # name, config, env...
  local name=$1 cfg=$2; shift 2
  env "$@" timeout -k 10 600 python bench.py --config $cfg --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_${name}.json 2> gpurun_out/${TAG}_bench_${name}.err
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
run_bench trajcontrol_nosplit trajcontrol ROHM_B200_TRAJ_SPLITK=0
run_bench posenet posenet ROHM_B200_X=0
run_bench posenet_nosplitload posenet ROHM_B200_ATTN_SPLIT_LOAD=0
ROHM_B200_ATTN_TS=1 timeout -k 10 300 python tools/profile_target.py 12 2> gpurun_out/${TAG}_timelines.txt > /dev/null
ROHM_B200_ATTN_TS=1 ROHM_B200_ATTN_SPLIT_LOAD=0 timeout -k 10 300 python tools/profile_target.py 12 2>> gpurun_out/${TAG}_timelines.txt > /dev/null
grep -h timeline gpurun_out/${TAG}_timelines.txt
du -sh gpurun_out
