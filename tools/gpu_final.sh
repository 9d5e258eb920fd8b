#!/bin/bash
# The round's final evidence on one GPU (run under gpurun; outputs in gpurun_out/, copied into profiles/ afterwards).
O=gpurun_out; mkdir -p $O
python bench.py > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err
python bench.py --steps 20 --warmup 5 > $O/r2_bench_n1_steps20.json 2>/dev/null
python bench.py --impl reference --steps 20 --warmup 5 > $O/r2_bench_reference_arm.json 2>/dev/null
python bench.py --phase lockstep --no-cpu-baseline > $O/r2_bench_n1_lockstep.json 2>/dev/null
python bench.py --phase staggered-match --no-cpu-baseline > $O/r2_bench_n1_staggered_match.json 2>/dev/null
python bench.py --agents fused --no-cpu-baseline > $O/r2_bench_n1_agents_fused.json 2>/dev/null
python bench.py --e2e-format i16 --no-cpu-baseline > $O/r2_bench_n1_e2e_i16.json 2>/dev/null
python tools/per_turn_time.py > $O/r2_per_turn.json 2>/dev/null
{ python tools/scripted_rollout.py 4096 900 random random; python tools/scripted_rollout.py 65536 450; python tools/scripted_rollout.py 1048576 300; python tools/scripted_rollout.py 65536 450 random random; python tools/scripted_rollout.py 1048576 300 random random; } > $O/r2_scripted_rollout.jsonl 2>/dev/null
{ for a in "--policy dqn --dtype fp32 --graph" "--policy dqn --dtype bf16 --graph" "--policy dqn --fused --graph" "--policy dqn --fused" "--policy ppo --dtype bf16 --graph" "--policy rppo --dtype bf16 --graph"; do python tools/policy_rollout.py $a --turns 150; done; } > $O/r2_policy_rollout.jsonl 2>/dev/null
python tools/mlp_time.py 32768 > $O/r2_mlp_time.jsonl 2>/dev/null
# ncu: launch list of a short default-workload run, then one full capture of the step kernel at the bench's batch size
cmd="python bench.py --steps 40 --warmup 5 --e2e-steps 2 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r2_final_launches_raw.csv $cmd > /dev/null 2> $O/ncu_launches.err
bash tools/ncu_capture.sh r2_final_1m staggered 160 1048576
bash tools/ncu_capture.sh r2_final_256k staggered 160 262144
ls -la $O | tail -30
NCU_SKIP=60 bash tools/ncu_mlp.sh > /dev/null 2>&1
