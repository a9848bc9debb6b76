#!/bin/bash
# Final GPU pass of round 2 (second session): every GPU test, one bench line per BASELINE config (+ the reference arm), smoke(),
# launch lists, CTA timelines and ncu --set full captures of the dominant kernels of the three engines, summarised on the box.
# Outputs under gpurun_out/<tag>_*; tools/collect_profiles.sh <tag> copies the judged ones into profiles/.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r2f}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout -k 10 2400 python -m pytest tests -q -m gpu -s > gpurun_out/${TAG}_tests.log 2>&1; echo "gpu tests exit $?"; tail -3 gpurun_out/${TAG}_tests.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/${TAG}_smoke.log
timeout -k 10 300 tools/gemm_selftest > gpurun_out/${TAG}_selftest.log 2>&1; echo "selftest exit $?"
for c in posenet trajcontrol lbs respaced100 pipeline; do
  timeout -k 10 900 python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err; echo "bench $c exit $?"
  python -c "
import json
d = json.loads(open('gpurun_out/${TAG}_bench_$c.json').read().strip().splitlines()[-1]); r = d.get('roofline', {})
print('  $c: value', round(d['value'], 2), 'e2e', round(d['e2e']['value'], 2), 'ms/step', round(d['ms_per_step'], 3), 'frac', round(r.get('frac') or 0, 4), 'clocks', d.get('clocks', {}).get('sm_mhz'))" 2>&1 | tail -1
done
timeout -k 10 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "reference arm exit $?"
# launch lists (cold-cache, serialised: shares, not absolutes)
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 140 -c 200 --csv --log-file gpurun_out/${TAG}_launches_posenet_step.csv python tools/profile_target.py 4 > /dev/null 2>&1
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_trajnet_forward.csv python tools/profile_target_trajnet.py 3 > /dev/null 2>&1
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_lbs.csv python tools/profile_lbs.py 3 > /dev/null 2>&1
python tools/launch_list_summary.py gpurun_out/${TAG}_launches_posenet_step.csv > gpurun_out/${TAG}_launches_posenet_step_summary.txt 2>&1
python tools/launch_list_summary.py gpurun_out/${TAG}_launches_trajnet_forward.csv pack_rows unpack_rows > gpurun_out/${TAG}_launches_trajnet_forward_summary.txt 2>&1
python tools/launch_list_summary.py gpurun_out/${TAG}_launches_lbs.csv repr_to_smplx gemm_tile > gpurun_out/${TAG}_launches_lbs_summary.txt 2>&1
# CTA-0 %globaltimer timelines: attention, fused LBS launch, TrajNet convolutions
ROHM_B200_ATTN_TS=1 timeout -k 10 300 python tools/profile_target.py 12 2> gpurun_out/${TAG}_timelines.txt > /dev/null
ROHM_B200_LBS_TS=1 timeout -k 10 300 python tools/profile_lbs.py 4 2>> gpurun_out/${TAG}_timelines.txt > /dev/null
ROHM_B200_GRAPH=0 ROHM_B200_TRAJ_TS=diff_enc1.c2,diff_enc2.c2,diff_enc3.c1,diff_mid_block1.c1,diff_down4,diff_dec1.c1,up1e timeout -k 10 300 python tools/profile_target_trajnet.py 5 2>> gpurun_out/${TAG}_timelines.txt > /dev/null
grep -c timeline gpurun_out/${TAG}_timelines.txt
# ncu --set full of the dominant kernels (one PoseNet layer; the fused LBS launch; a TrajNet stretch)
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tile|attention_tc|ddpm_step" -s 52 -c 6 -o gpurun_out/${TAG}_prof_posenet python tools/profile_target.py 4 > gpurun_out/${TAG}_ncu_posenet.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tile|fk_full" -s 2 -c 2 -o gpurun_out/${TAG}_prof_lbs python tools/profile_lbs.py 3 > gpurun_out/${TAG}_ncu_lbs.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none -k regex:"gemm_tile|gn_mish|sum_split" -s 130 -c 8 -o gpurun_out/${TAG}_prof_trajnet python tools/profile_target_trajnet.py 3 > gpurun_out/${TAG}_ncu_trajnet.log 2>&1
for n in posenet lbs trajnet; do
  python tools/ncu_summary.py gpurun_out/${TAG}_prof_$n.ncu-rep > gpurun_out/${TAG}_ncu_${n}_summary.txt 2>&1
done
python tools/ncu_traffic.py gpurun_out/${TAG}_prof_posenet.ncu-rep gemm_tile_kernel gpurun_out/${TAG}_gemm_traffic.json > /dev/null 2>&1
python tools/ncu_traffic.py gpurun_out/${TAG}_prof_lbs.ncu-rep gemm_tile_kernel gpurun_out/${TAG}_lbs_traffic.json > /dev/null 2>&1
python tools/ncu_traffic.py gpurun_out/${TAG}_prof_trajnet.ncu-rep gemm_tile_kernel gpurun_out/${TAG}_trajnet_traffic.json > /dev/null 2>&1
# keep the reports only while the directory stays under the 64 MiB limit (largest first out)
while [ $(du -sm gpurun_out | cut -f1) -ge 56 ]; do
  big=$(ls -S gpurun_out/*.ncu-rep 2>/dev/null | head -1); [ -z "$big" ] && break; echo "dropping $big"; rm -f "$big"
done
du -sh gpurun_out
