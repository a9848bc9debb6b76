#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "vertices or lbs or dense or from_repr or edge_rot" 2>&1 | tail -3
for f in 0 3; do
ROHM_B200_LBS_DEBUG=$f ROHM_B200_LBS_TS=1 timeout 300 python tools/profile_lbs.py 4 2>&1 | grep timeline
done
ROHM_B200_LBS_MULTICAST=0 ROHM_B200_LBS_TS=1 timeout 300 python tools/profile_lbs.py 4 2>&1 | grep timeline
