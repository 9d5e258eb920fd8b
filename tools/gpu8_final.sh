#!/bin/bash
# Final 8-GPU records of the round (gpurun --gpus 8): weak-scaling bench, strong-scaling point, BASELINE configs[3] with the fused forward.
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 > $O/r2_bench_n8.json 2> $O/r2_bench_n8.err
$TR --nproc-per-node 8 --master-port 29612 bench.py --gpus 8 --total-envs 1048576 --e2e-steps 5 > $O/r2_strong_n8.json 2>> $O/r2_strong.err
rm -f $O/r2_policy_rollout_n8_fused.jsonl
$TR --nproc-per-node 8 --master-port 29613 tools/policy_rollout.py --policy dqn --fused --graph >> $O/r2_policy_rollout_n8_fused.jsonl 2>> $O/r2_policy.err
$TR --nproc-per-node 8 --master-port 29614 tools/policy_rollout.py --policy dqn --dtype bf16 --graph >> $O/r2_policy_rollout_n8_fused.jsonl 2>> $O/r2_policy.err
python - <<'PY'
import json
for f in ["gpurun_out/r2_bench_n8.json", "gpurun_out/r2_strong_n8.json"]:
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f, "N", d["n_gpus"], "value %.4g" % d["value"], "frac %.3f" % r["frac"], "e2e %.4g" % d["e2e"]["value"], "f32 %.4g" % d["e2e_f32"]["value"],
              "probe %.1f GB/s" % d["e2e"]["link"]["d2h_probe_gbs_per_gpu"])
    except Exception as e:
        print(f, "failed", e)
print(open("gpurun_out/r2_policy_rollout_n8_fused.jsonl").read())
PY
