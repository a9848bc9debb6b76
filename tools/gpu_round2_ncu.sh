#!/bin/bash
# ncu --set full captures of the dominant kernels (one layer of the PoseNet forward; the fused LBS launch; a TrajNet stretch),
# summarised ON THE BOX (tools/ncu_summary.py, tools/ncu_traffic.py) because gpurun_out/ may not exceed 64 MiB per call; the
# .ncu-rep files are kept only while they fit.  Also: compute-sanitizer racecheck / memcheck of the smoke path.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r2}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tile|attention_tc|ddpm_step" -s 52 -c 6 -o gpurun_out/${TAG}_prof_posenet python tools/profile_target.py 4 > gpurun_out/${TAG}_ncu_posenet.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tile|fk_full" -s 2 -c 2 -o gpurun_out/${TAG}_prof_lbs python tools/profile_lbs.py 3 > gpurun_out/${TAG}_ncu_lbs.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"gemm_tile|gn_mish" -s 130 -c 6 -o gpurun_out/${TAG}_prof_trajnet python tools/profile_target_trajnet.py 3 > gpurun_out/${TAG}_ncu_trajnet.log 2>&1
# the same PoseNet layer with caches left warm between ncu's replay passes (closer to the running loop, where operands are L2 hits)
timeout 900 ncu --set full --clock-control none --cache-control none -k regex:"gemm_tile|attention_tc" -s 52 -c 5 -o gpurun_out/${TAG}_prof_posenet_warm python tools/profile_target.py 4 > gpurun_out/${TAG}_ncu_posenet_warm.log 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_prof_posenet_warm.ncu-rep > gpurun_out/${TAG}_ncu_posenet_warm_summary.txt 2>&1
rm -f gpurun_out/${TAG}_prof_posenet_warm.ncu-rep
for n in posenet lbs trajnet; do
  python tools/ncu_summary.py gpurun_out/${TAG}_prof_$n.ncu-rep > gpurun_out/${TAG}_ncu_${n}_summary.txt 2>&1
done
python tools/ncu_traffic.py gpurun_out/${TAG}_prof_posenet.ncu-rep gemm_tile_kernel gpurun_out/${TAG}_gemm_traffic.json > /dev/null 2>&1
python tools/ncu_traffic.py gpurun_out/${TAG}_prof_lbs.ncu-rep gemm_tile_kernel gpurun_out/${TAG}_lbs_traffic.json > /dev/null 2>&1
python tools/ncu_traffic.py gpurun_out/${TAG}_prof_trajnet.ncu-rep gemm_tile_kernel gpurun_out/${TAG}_trajnet_traffic.json > /dev/null 2>&1
# source-level hot spots of the PoseNet GEMM (stall samples per line), for the record
ncu -i gpurun_out/${TAG}_prof_posenet.ncu-rep --page source --csv --kernel-name regex:gemm_tile 2>/dev/null | head -400 > gpurun_out/${TAG}_ncu_posenet_source_head.csv
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_sanitizer_racecheck_smoke.log 2>&1; echo "racecheck exit $?"
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_sanitizer_memcheck_smoke.log 2>&1; echo "memcheck exit $?"
tail -2 gpurun_out/${TAG}_sanitizer_racecheck_smoke.log gpurun_out/${TAG}_sanitizer_memcheck_smoke.log
ls -la gpurun_out/
# keep the reports only while the directory stays under the 64 MiB limit (largest first out)
while [ $(du -sm gpurun_out | cut -f1) -ge 60 ]; do
  big=$(ls -S gpurun_out/*.ncu-rep 2>/dev/null | head -1); [ -z "$big" ] && break; echo "dropping $big"; rm -f "$big"
done
du -sh gpurun_out
