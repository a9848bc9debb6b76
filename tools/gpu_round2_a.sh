#!/bin/bash
# GPU call A (round 2): the new parity tests first, then the whole GPU suite, then one bench line per config.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_glue.py tests/test_gpu_pipeline.py -x -q -m gpu -s > gpurun_out/r2a_tests_new.log 2>&1
echo "new tests exit $?" >> gpurun_out/r2a_tests_new.log
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_gpu_glue.py --deselect tests/test_gpu_pipeline.py > gpurun_out/r2a_tests_old.log 2>&1
echo "old tests exit $?" >> gpurun_out/r2a_tests_old.log
for c in posenet trajcontrol lbs respaced100 pipeline; do
  timeout 900 python bench.py --config $c --steps 2 --warmup 3 > gpurun_out/r2a_bench_$c.json 2> gpurun_out/r2a_bench_$c.err
  echo "bench $c exit $?" >> gpurun_out/r2a_tests_new.log
done
tail -5 gpurun_out/r2a_tests_new.log; tail -3 gpurun_out/r2a_tests_old.log
