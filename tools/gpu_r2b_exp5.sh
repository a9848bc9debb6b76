#!/bin/bash
# Round-2 second session: the fused TrajNet sampler step (rohm_trajnet_sample_step): TrajNet / pipeline / PoseNet tests + bench line.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r2b7}
timeout -k 10 900 python -m pytest tests/test_gpu_trajnet.py tests/test_gpu_posenet.py tests/test_gpu_pipeline.py -q -m gpu -x > gpurun_out/${TAG}_tests.log 2>&1; echo "trajnet+posenet+pipeline tests exit $?"; tail -3 gpurun_out/${TAG}_tests.log
timeout -k 10 600 python bench.py --config trajcontrol --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_trajcontrol.json 2> gpurun_out/${TAG}_bench_trajcontrol.err; echo "bench exit $?"
python -c "
import json
d = json.loads(open('gpurun_out/${TAG}_bench_trajcontrol.json').read().strip().splitlines()[-1]); r = d['roofline']
print('trajcontrol: value', round(d['value'], 2), 'e2e', round(d['e2e']['value'], 2), 'ms/step', round(d['ms_per_step'], 3), 'fwd ms', r['forward_ms'], 'launches', d['gpu_launches'])"
