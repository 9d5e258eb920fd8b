"""End-to-end step (host actions in, host wire rows out) over EVG_HOST_CHUNKS values: ms per step of 1 Mi matches."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import evgsim

n = 1 << 20
fmt = sys.argv[1] if len(sys.argv) > 1 else "wire"
env = evgsim.BatchedEvergladesEnv(n, seed=0, auto_reset=evgsim._capi.AUTORESET_TERMINAL)
env.reset()
for _ in range(40):
    env.step(env.random_actions())
acts = [env.random_actions().cpu().pin_memory() for _ in range(4)]
env.host_buffers(fmt)
for chunks in [int(x) for x in sys.argv[2:]] or [1, 2, 4, 8, 16, 32]:
    os.environ["EVG_HOST_CHUNKS"] = str(chunks)
    for _ in range(3):
        env.step_host(acts[0], obs_format=fmt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sink = 0.0
    for k in range(20):
        o, r, d, _ = env.step_host(acts[k % 4], sync=True, obs_format=fmt)
        sink += float(o[0, 0].sum())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(json.dumps({"format": fmt, "chunks": chunks, "ms_per_step": round(ms, 4), "env_turns_per_s": n / ms * 1e3}), flush=True)
