#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_trajnet.py tests/test_gpu_body.py tests/test_gpu_glue.py -x -q -m gpu > gpurun_out/r2i_tests.log 2>&1; tail -4 gpurun_out/r2i_tests.log
for c in trajcontrol lbs; do
  timeout 600 python bench.py --config $c --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_bench_$c.json 2> gpurun_out/r2i_bench_$c.err
done
ROHM_B200_TRAJ_PARALLEL=0 timeout 600 python bench.py --config trajcontrol --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_bench_trajserial.json 2> gpurun_out/r2i_bench_trajserial.err
python - <<'PY'
import json
for n in ("trajcontrol", "trajserial", "lbs"):
    try:
        d = json.loads(open(f"gpurun_out/r2i_bench_{n}.json").read().strip().splitlines()[-1]); r = d["roofline"]
        print(n, "value %.2f ms/step %.1f frac %.3f" % (d["value"], d["ms_per_step"], r["frac"]), {k: r[k] for k in ("forward_ms", "call_ms") if k in r})
    except Exception as e:
        print(n, "ERR", e, open(f"gpurun_out/r2i_bench_{n}.err").read()[-1200:])
PY
