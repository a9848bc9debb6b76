#!/bin/bash
# Full GPU pass of a round: every GPU test, one bench line per BASELINE config (+ the reference arm), smoke(), launch lists of
# the three engines and ncu --set full captures of the dominant kernels.  Outputs under gpurun_out/<tag>_*.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r2}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 2400 python -m pytest tests -q -m gpu -s > gpurun_out/${TAG}_tests.log 2>&1; echo "gpu tests exit $?"; tail -3 gpurun_out/${TAG}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/${TAG}_smoke.log
timeout 300 tools/gemm_selftest > gpurun_out/${TAG}_selftest.log 2>&1; echo "selftest exit $?"
for c in posenet trajcontrol lbs respaced100 pipeline; do
  timeout 900 python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err; echo "bench $c exit $?"
done
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "reference arm exit $?"
# launch lists (cold-cache, serialised: shares, not absolutes); the ncu --set full captures live in tools/gpu_round2_ncu.sh
# (a separate call: gpurun_out/ may not exceed 64 MiB per call)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 140 -c 200 --csv --log-file gpurun_out/${TAG}_launches_posenet_step.csv python tools/profile_target.py 4 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_trajnet_forward.csv python tools/profile_target_trajnet.py 3 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_lbs.csv python tools/profile_lbs.py 3 > /dev/null 2>&1
ROHM_B200_ATTN_TS=1 ROHM_B200_LBS_TS=1 timeout 300 python tools/profile_target.py 12 2> gpurun_out/${TAG}_timelines.txt > /dev/null
ROHM_B200_LBS_TS=1 timeout 300 python tools/profile_lbs.py 4 2>> gpurun_out/${TAG}_timelines.txt > /dev/null
python tools/launch_list_summary.py gpurun_out/${TAG}_launches_posenet_step.csv > gpurun_out/${TAG}_launches_posenet_step_summary.txt 2>&1
python tools/launch_list_summary.py gpurun_out/${TAG}_launches_trajnet_forward.csv pack_rows unpack_rows > gpurun_out/${TAG}_launches_trajnet_forward_summary.txt 2>&1
python tools/launch_list_summary.py gpurun_out/${TAG}_launches_lbs.csv repr_to_smplx gemm_tile > gpurun_out/${TAG}_launches_lbs_summary.txt 2>&1
du -sh gpurun_out
