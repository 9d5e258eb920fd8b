#!/bin/bash
# compute-sanitizer over every kernel of libevgsim (run on the GPU box): racecheck (shared-memory hazards — the combat
# phase has lanes reading and writing other matches' rows between __syncwarp()s), memcheck, and initcheck-free synccheck.
# Logs go to gpurun_out/ (copy the summaries into profiles/).
set -u
OUT=${1:-gpurun_out}
mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in racecheck memcheck synccheck; do
    extra=""
    [ $tool = racecheck ] && extra="--racecheck-report all"
    timeout 1500 $CS --tool $tool $extra --print-limit 50 --log-file $OUT/sanitizer_$tool.log python tools/sanitize_workload.py 90 > $OUT/sanitizer_$tool.out 2>&1
    echo "$tool exit $?" >> $OUT/sanitizer_$tool.out
    tail -3 $OUT/sanitizer_$tool.log
done
