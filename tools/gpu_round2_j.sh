#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_body.py tests/test_gpu_glue.py -x -q -m gpu > gpurun_out/r2j_tests.log 2>&1; tail -3 gpurun_out/r2j_tests.log
for cfg in "384 1" "384 0" "256 1" "512 1" "768 1" "4608 0"; do
  set -- $cfg
  ROHM_B200_LBS_CHUNK=$1 ROHM_B200_LBS_OVERLAP=$2 timeout 300 python bench.py --config lbs --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('chunk $1 overlap $2: call %.3f ms frac %.3f  bench ms/step %.3f' % (r['call_ms'], r['frac'], d['ms_per_step']))"
done
