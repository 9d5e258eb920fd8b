"""Mid-size batches: step-kernel time of the two thread-per-match instantiations (one-warp CTAs without shared-memory
tables = LITE, and 128-thread CTAs) over batch sizes; random self-play, lock-step episodes, CUDA events around the step."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import evgsim

sizes = [int(x) for x in sys.argv[1:]] or [16384, 32768, 65536, 98304, 131072, 196608, 262144, 524288]
for n in sizes:
    out = {"matches": n}
    for name, small_max in (("lite32", 1 << 30), ("cta128", 0)):
        os.environ["EVG_STEP_KERNEL"] = "tpm"
        os.environ["EVG_TPM_SMALL_MAX"] = str(small_max)
        env = evgsim.BatchedEvergladesEnv(n, seed=0, auto_reset=evgsim._capi.AUTORESET_TERMINAL)
        env.reset()
        for _ in range(150):
            env.step(env.random_actions())
        torch.cuda.synchronize()
        ev = []
        for _ in range(300):
            a = env.random_actions()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            env.step(a)
            e1.record()
            ev.append((e0, e1))
        torch.cuda.synchronize()
        us = sum(a.elapsed_time(b) for a, b in ev) / len(ev) * 1e3
        out[name + "_us"] = round(us, 2)
        out[name + "_env_turns_per_s"] = n / us * 1e6
        env.close()
    print(json.dumps(out), flush=True)
