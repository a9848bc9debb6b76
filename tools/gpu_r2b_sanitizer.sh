#!/bin/bash
# compute-sanitizer over the second-session kernels (tools/sanitizer_target.py) and the smoke path; 2-GPU bench line.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r2s}
timeout -k 10 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitizer_target.py > gpurun_out/${TAG}_sanitizer_memcheck_target.log 2>&1; echo "memcheck target exit $?"
timeout -k 10 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitizer_target.py > gpurun_out/${TAG}_sanitizer_racecheck_target.log 2>&1; echo "racecheck target exit $?"
timeout -k 10 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_sanitizer_memcheck_smoke.log 2>&1; echo "memcheck smoke exit $?"
timeout -k 10 900 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_sanitizer_racecheck_smoke.log 2>&1; echo "racecheck smoke exit $?"
tail -4 gpurun_out/${TAG}_sanitizer_*.log
