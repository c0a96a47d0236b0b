#!/bin/bash
# Round-1 profiling pass (run under gpurun): launch list of a short bench run, then one full capture of each hot kernel.
set -x
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$B > gpurun_out/plain_launches.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv $B > gpurun_out/ncu_launches.log 2>&1
S="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --contigs 296000"
$S > gpurun_out/plain_full.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:kmer_hist_kernel -s 1 -c 1 -f -o gpurun_out/prof_hist_r1 $S > gpurun_out/ncu_hist.log 2>&1
$S > gpurun_out/plain_full2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:score_tc_kernel -s 1 -c 1 -f -o gpurun_out/prof_score_r1 $S > gpurun_out/ncu_score.log 2>&1
