#!/bin/bash
# 2-GPU pass of the final tree (gpurun --gpus 2): sharding checks over NCCL and the weak-scaling headline bench line.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r2fn2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout -k 10 400 $TR --master-port 29511 tools/n2_sharded_check.py > gpurun_out/${TAG}_sharded_check.log 2>&1; echo "sharded check exit $?"; grep "n2 check" gpurun_out/${TAG}_sharded_check.log
for c in posenet; do
  timeout -k 10 600 $TR --master-port 29512 bench.py --config $c --gpus 2 --steps 2 --warmup 3 > gpurun_out/${TAG}_bench_${c}_n2.json 2> gpurun_out/${TAG}_bench_${c}_n2.err; echo "bench $c n2 exit $?"
  tail -1 gpurun_out/${TAG}_bench_${c}_n2.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['name'], d['n_gpus'], 'GPUs', d['value'], d['unit'], 'e2e', d['e2e']['value'], d['clocks'])"
done
timeout -k 10 300 $TR --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_reference_n2.json 2> gpurun_out/${TAG}_bench_reference_n2.err; echo "reference arm n2 exit $?"; tail -1 gpurun_out/${TAG}_bench_reference_n2.json | cut -c1-200
