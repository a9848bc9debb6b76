#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export ROHM_B200_LBS_CHUNK=4608 ROHM_B200_LBS_OVERLAP=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2k_lbs_launches.csv python tools/profile_lbs.py 3 > gpurun_out/r2k_ncu1.log 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r2k_lbs_launches.csv")))
i0 = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[i0]; kn, mv = h.index("Kernel Name"), h.index("Metric Value")
for r in rows[i0 + 1:][-8:]:
    print(r[kn].split('(')[0][-50:], r[mv])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"skin_kernel|gemm_tile|fk_full" -s 6 -c 3 -o gpurun_out/r2k_lbs python tools/profile_lbs.py 3 > gpurun_out/r2k_ncu2.log 2>&1
ls -la gpurun_out/r2k_lbs.ncu-rep
