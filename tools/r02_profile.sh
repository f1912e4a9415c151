#!/bin/bash
# Round-2 ncu evidence for the default bench command (1 GPU).  Run on the GPU box:  bash tools/r02_profile.sh
# Writes into gpurun_out/; the summaries are copied to profiles/ afterwards.
set -x
CMD="python bench.py --steps 3 --warmup 3 --no-solve"
$CMD > gpurun_out/r02_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r02_plain.log; exit 1; }
tail -c 300 gpurun_out/r02_plain.log
# every launch with its device time (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_bench.csv \
    $CMD > gpurun_out/r02_ncu_launches.log 2>&1
# the dominant kernel (level-0 passes are among the first 16 matches: 2 untimed + timed applies)
ncu --set full --clock-control none --import-source on -k regex:k_batched_gemv$ -s 7 -c 7 -o gpurun_out/r02_gemv \
    $CMD > gpurun_out/r02_ncu_gemv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_refine -c 3 -o gpurun_out/r02_refine \
    $CMD > gpurun_out/r02_ncu_refine.log 2>&1
ls -la gpurun_out/*.ncu-rep
