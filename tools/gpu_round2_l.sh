#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 tools/gemm_selftest > gpurun_out/r2l_selftest.log 2>&1; echo "selftest exit $?"
grep -E "multicast|FAIL|SELFTEST" gpurun_out/r2l_selftest.log | cut -c1-200 | head -30
grep -A2 "\[A multicast\]" gpurun_out/r2l_selftest.log | grep timing | head
grep -B0 -A3 "f16x2 linear M4640 N1536 K512 bn128 act0 a_scale 1 \[TMA store fp16 pair\]" gpurun_out/r2l_selftest.log | cut -c1-300 | head -12
timeout 600 python -m pytest tests/test_gpu_posenet.py -x -q -m gpu > gpurun_out/r2l_tests.log 2>&1; tail -3 gpurun_out/r2l_tests.log
for mc in 1 0; do
ROHM_B200_MULTICAST=$mc timeout 600 python bench.py --config posenet --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('multicast $mc: clips/s %.2f ms/step %.1f fwd graph %.4f step graph %.4f' % (d['value'], d['ms_per_step'], r['forward_graph_ms'], r['step_graph_ms']), r['forward_ms_by_kernel_class'])"
done
