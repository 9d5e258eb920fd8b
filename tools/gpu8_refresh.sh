#!/bin/bash
# Short 8-GPU refresh (gpurun --gpus 8): weak-scaling bench line + BASELINE configs[3] with the fused forward.
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 > $O/r2_bench_n8.json 2> $O/r2_bench_n8.err
rm -f $O/r2_policy_rollout_n8_fused.jsonl
timeout 200 $TR --nproc-per-node 8 --master-port 29613 tools/policy_rollout.py --policy dqn --fused --graph >> $O/r2_policy_rollout_n8_fused.jsonl 2>> $O/r2_policy.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_n8.json").read().strip().splitlines()[-1]); r = d["roofline"]
    print("N", d["n_gpus"], "value %.4g" % d["value"], "frac %.3f" % r["frac"], "e2e %.4g" % d["e2e"]["value"], "rollout %.4g" % d["scripted_rollout"]["value"])
except Exception as e:
    print("bench failed", e)
print(open("gpurun_out/r2_policy_rollout_n8_fused.jsonl").read()[:600])
PY
