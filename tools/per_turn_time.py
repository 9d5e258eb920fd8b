"""Step-kernel time per turn of one lock-step episode (all matches start together): where the average comes from."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import evgsim

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
env = evgsim.BatchedEvergladesEnv(n, seed=0, auto_reset=evgsim._capi.AUTORESET_TERMINAL)
env.reset()
for _ in range(150):  # warm-up episode
    env.step(env.random_actions())
torch.cuda.synchronize()
env.reset()
ev = []
for t in range(150):
    a = env.random_actions()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    env.step(a)
    e1.record()
    ev.append((e0, e1))
torch.cuda.synchronize()
ms = [a.elapsed_time(b) for a, b in ev]
print(json.dumps({"matches": n, "mean_ms": sum(ms) / len(ms), "per_turn_ms": [round(x, 4) for x in ms],
                  "fought_unit_slots_per_env_turn": env.episode_stats()["fought_unit_slots"] / (150.0 * n)}))
