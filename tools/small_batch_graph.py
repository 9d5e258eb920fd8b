"""BASELINE.json configs[1]: 4096 lock-step matches on one GPU — launch-bound, so the turn (agent kernel + step
kernel) is captured in a CUDA graph of 50 turns and replayed.  Prints env-turns/s with and without the graph."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import evgsim

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = evgsim.BatchedEvergladesEnv(n, seed=0, auto_reset=evgsim._capi.AUTORESET_TERMINAL)
env.reset()
def turn():
    env.step(env.random_actions())
for _ in range(150):
    turn()
torch.cuda.synchronize()
def timed(fn, reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / 1e3
t_plain = timed(turn, 900)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3):
        turn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for _ in range(50):
            turn()
torch.cuda.synchronize()
t_graph = timed(g.replay, 18)
print(json.dumps({"matches": n, "env_turns_per_s_plain": n * 900 / t_plain, "env_turns_per_s_graph": n * 900 / t_graph,
                  "us_per_turn_plain": t_plain / 900 * 1e6, "us_per_turn_graph": t_graph / 900 * 1e6,
                  "episodes": env.episode_stats()["episodes"]}))
