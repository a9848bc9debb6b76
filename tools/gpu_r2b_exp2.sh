#!/bin/bash
# Round-2 second session, call 2: TrajNet tests + bench with the downsampling convolutions split as well, warm-cache launch
# lists (ncu --cache-control none) of the TrajNet forward and the PoseNet step, the other bench configs.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r2b2}
timeout -k 10 900 python -m pytest tests/test_gpu_trajnet.py tests/test_gpu_pipeline.py -q -m gpu -x > gpurun_out/${TAG}_tests_trajnet.log 2>&1; echo "trajnet+pipeline tests exit $?"; tail -3 gpurun_out/${TAG}_tests_trajnet.log
timeout -k 10 300 tools/gemm_selftest > gpurun_out/${TAG}_selftest.log 2>&1; echo "selftest exit $?"; grep -c OK gpurun_out/${TAG}_selftest.log; grep -c -i "fail\|mismatch" gpurun_out/${TAG}_selftest.log
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
run_bench trajcontrol_serial trajcontrol ROHM_B200_TRAJ_PARALLEL=0
timeout -k 10 300 python tools/profile_trajnet.py > gpurun_out/${TAG}_profile_trajnet.txt 2>&1; cat gpurun_out/${TAG}_profile_trajnet.txt | tail -2
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/${TAG}_launches_trajnet_forward_warm.csv python tools/profile_target_trajnet.py 3 > /dev/null 2>&1
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 140 -c 200 --csv --log-file gpurun_out/${TAG}_launches_posenet_step_warm.csv python tools/profile_target.py 4 > /dev/null 2>&1
python tools/launch_list_summary.py gpurun_out/${TAG}_launches_trajnet_forward_warm.csv pack_rows unpack_rows > gpurun_out/${TAG}_launches_trajnet_forward_warm_summary.txt 2>&1
python tools/launch_list_summary.py gpurun_out/${TAG}_launches_posenet_step_warm.csv > gpurun_out/${TAG}_launches_posenet_step_warm_summary.txt 2>&1
cat gpurun_out/${TAG}_launches_trajnet_forward_warm_summary.txt gpurun_out/${TAG}_launches_posenet_step_warm_summary.txt
run_bench respaced100 respaced100 ROHM_B200_X=0
run_bench pipeline pipeline ROHM_B200_X=0
du -sh gpurun_out
