#!/bin/bash
# Copies the judged artefacts of a full GPU run (tools/gpu_round2_full.sh <tag>, tools/gpu_round2_ncu.sh <tag2>) from the
# scratch directory gpurun_out/ into profiles/ under round-2 names.  usage: tools/collect_profiles.sh <tag> [<ncu tag>]
cd "$(dirname "$0")/.."
T=$1; N=${2:-$1}; G=gpurun_out; P=profiles
for c in posenet trajcontrol lbs respaced100 pipeline reference; do [ -s $G/${T}_bench_$c.json ] && tail -1 $G/${T}_bench_$c.json > $P/r2_bench_$c.json; done
for f in posenet_step trajnet_forward lbs; do
  [ -s $G/${T}_launches_$f.csv ] && grep -v "^==" $G/${T}_launches_$f.csv > $P/r2_launches_$f.csv
  [ -s $G/${T}_launches_${f}_summary.txt ] && cp $G/${T}_launches_${f}_summary.txt $P/r2_launches_${f}_summary.txt
done
[ -s $G/${T}_timelines.txt ] && grep timeline $G/${T}_timelines.txt > $P/r2_timelines.txt
[ -s $G/${T}_selftest.log ] && grep -B3 "debug_flags" $G/${T}_selftest.log | grep -v "^--" > $P/r2_gemm_selftest_timelines.txt
[ -s $G/${T}_tests.log ] && { grep -A60 "^guided tail 32x143" $G/${T}_tests.log | grep "^guided\|^t=" > $P/r2_guided_tail_32x143.txt; tail -3 $G/${T}_tests.log > $P/r2_gpu_tests_tail.txt; grep -h "max |cuda\|projection guidance\|free-run\|stagewise\|teacher" $G/${T}_tests.log | head -40 >> $P/r2_gpu_tests_tail.txt; }
[ -s $G/${T}_smoke.log ] && tail -1 $G/${T}_smoke.log >> $P/r2_gpu_tests_tail.txt
[ -s $G/${T}_smi.txt ] && cp $G/${T}_smi.txt $P/r2_nvidia_smi.txt
for n in posenet posenet_warm lbs trajnet; do [ -s $G/${N}_ncu_${n}_summary.txt ] && cp $G/${N}_ncu_${n}_summary.txt $P/r2_ncu_${n}_summary.txt; done
for n in gemm lbs trajnet; do [ -s $G/${N}_${n}_traffic.json ] && cp $G/${N}_${n}_traffic.json $P/r2_${n}_traffic.json; done
for n in racecheck memcheck; do [ -s $G/${N}_sanitizer_${n}_smoke.log ] && tail -15 $G/${N}_sanitizer_${n}_smoke.log > $P/r2_sanitizer_${n}_smoke.txt; done
ls -la $P | tail -40
