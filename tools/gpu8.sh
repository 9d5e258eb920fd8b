#!/bin/bash
# Everything that needs the 8-GPU box, in one call (gpurun --gpus 8): box topology, the weak-scaling bench at N=8,
# the strong-scaling record (1,048,576 matches in total at 1/2/4/8 ranks), BASELINE configs[3] at 8 x 32,768.
OUT=gpurun_out
mkdir -p $OUT
bash tools/gpu_probe.sh $OUT > /dev/null; mv $OUT/box_probe.txt $OUT/box_probe_8gpu.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 > $OUT/r2_bench_n8.json 2> $OUT/r2_bench_n8.err
for n in 8 4 2; do
  $TR --nproc-per-node $n --master-port $((29520+n)) bench.py --gpus $n --total-envs 1048576 --e2e-steps 5 > $OUT/r2_strong_n$n.json 2>> $OUT/r2_strong.err
done
python bench.py --total-envs 1048576 --no-cpu-baseline --e2e-steps 5 > $OUT/r2_strong_n1.json 2>> $OUT/r2_strong.err
for pol in "dqn fp32" "dqn bf16" "ppo bf16" "rppo bf16"; do set -- $pol
  $TR --nproc-per-node 8 --master-port 29531 tools/policy_rollout.py --policy $1 --dtype $2 --graph >> $OUT/r2_policy_rollout_n8.jsonl 2>> $OUT/r2_policy.err
done
$TR --nproc-per-node 8 --master-port 29532 tools/policy_rollout.py --policy dqn --dtype bf16 >> $OUT/r2_policy_rollout_n8.jsonl 2>> $OUT/r2_policy.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_bench_n8.json") + glob.glob("gpurun_out/r2_strong_n*.json")):
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f, "N", d["n_gpus"], "value %.4g" % d["value"], "frac %.3f" % r["frac"], "e2e %.4g" % d["e2e"]["value"], "f32 %.4g" % d["e2e_f32"]["value"],
              "probe %.1f GB/s" % d["e2e"]["link"]["d2h_probe_gbs_per_gpu"], d["e2e"]["pinned_placement"])
    except Exception as e:
        print(f, "failed", e)
print(open("gpurun_out/r2_policy_rollout_n8.jsonl").read())
PY
