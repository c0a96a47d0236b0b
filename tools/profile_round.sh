#!/bin/bash
# Profiling pass of one round (run under gpurun):  tools/profile_round.sh r2
# Launch lists of short bench runs (shipped and enlarged references), then one full capture of each hot kernel.
# Every ncu command is preceded by the same command without ncu, and only runs if that exited 0.
TAG=${1:-r2}
set -x
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs"
$B > gpurun_out/plain_launches_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/ncu_launches_$TAG.log 2>&1
E="python bench.py --workload enlarged --contigs 200000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$E > gpurun_out/plain_enl_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_enl_$TAG.csv $E > gpurun_out/ncu_launches_enl_$TAG.log 2>&1
S="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --contigs 296000"
$S > gpurun_out/plain_full_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:kmer_hist_kernel -s 1 -c 1 -f -o gpurun_out/prof_hist_$TAG $S > gpurun_out/ncu_hist_$TAG.log 2>&1
# per step the scorer launches score_tc_kernel three times (first pass, list pass twice): skip the warm-up step's three
$S > gpurun_out/plain_full2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:score_tc_kernel -s 3 -c 1 -f -o gpurun_out/prof_score_$TAG $S > gpurun_out/ncu_score_$TAG.log 2>&1
$S > gpurun_out/plain_full3_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:score_decide_kernel -s 1 -c 1 -f -o gpurun_out/prof_decide_$TAG $S > gpurun_out/ncu_decide_$TAG.log 2>&1
true
