#!/usr/bin/env python
"""BASELINE.json configs[2]: base_rushV1 (player 0) vs SwarmAgent (player 1), 65,536 lock-step matches with in-place
auto-reset on one GPU.  Both agents run on the device (evg_agents), reading the resident records; prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import evgsim

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
turns = int(sys.argv[2]) if len(sys.argv) > 2 else 450
env = evgsim.BatchedEvergladesEnv(n, seed=0, auto_reset=evgsim._capi.AUTORESET_TERMINAL)
env.reset()


def turn():
    env.step(env.agent_actions(evgsim._capi.AGENT_BASE_RUSH, evgsim._capi.AGENT_SWARM))


for _ in range(150):
    turn()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(turns):
    turn()
b.record()
torch.cuda.synchronize()
sec = a.elapsed_time(b) / 1e3
st = env.episode_stats()
print(json.dumps({"workload": "base_rushV1 vs SwarmAgent, auto-reset (BASELINE.json configs[2])", "matches": n, "turns": turns,
                  "env_turns_per_s": n * turns / sec, "us_per_turn": sec * 1e6 / turns, "episodes": st["episodes"], "wins": st["wins"],
                  "ties": st["ties"], "mean_episode_turns": st["total_turns"] / max(st["episodes"], 1), "status_count": st["status_count"]}))
