#!/bin/bash
# tools/ncu_capture.sh TAG PHASE SKIP [ENVS]  — one `ncu --set full` capture of the step kernel in a bench run (GPU box).
#   PHASE staggered|staggered-match: SKIP = launches of the step kernel to skip (settle 150 + warm-up) before the capture
#   PHASE lockstep: the capture is game turn SKIP+1 of the first episode
# Writes gpurun_out/TAG.ncu-rep.  Run the same bench command without ncu first (B200_PROFILING.md).
set -u
tag=$1; phase=$2; skip=$3; envs=${4:-262144}
mkdir -p gpurun_out
if [ "$phase" = lockstep ]; then warmup=$skip; else warmup=$(( skip - 150 )); fi
cmd="python bench.py --phase $phase --envs-per-gpu $envs --steps 3 --warmup $warmup --e2e-steps 1 --no-cpu-baseline"
$cmd > gpurun_out/${tag}_plain.json 2> gpurun_out/${tag}_plain.err || { echo "bench failed without ncu"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:evg_step_tpm_kernel -s $skip -c 1 -f -o gpurun_out/$tag $cmd > gpurun_out/${tag}_ncu.log 2>&1
tail -3 gpurun_out/${tag}_ncu.log
ls -la gpurun_out/$tag.ncu-rep
