#!/usr/bin/env python
"""BASELINE.json configs[1] and configs[2] on one GPU: fully scripted self-play with in-place auto-reset, every agent
on the device, whole rollouts in ONE launch (BatchedEvergladesEnv.rollout -> evg_rollout: the warp-per-match multi-turn
kernel for small batches, the thread-per-match kernel keeping each batch in shared memory for all the turns otherwise).

    python tools/scripted_rollout.py [matches] [turns] [agent0 agent1]     agents: random | base_rush | swarm
defaults: 65536 matches, 450 turns, base_rush vs swarm (configs[2]); `4096 900 random random` is configs[1].
Prints one JSON line per mode: one launch for all the turns, one launch per 50 turns, one launch per turn (step_agents)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import evgsim

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
turns = int(sys.argv[2]) if len(sys.argv) > 2 else 450
names = {"random": evgsim._capi.AGENT_RANDOM, "base_rush": evgsim._capi.AGENT_BASE_RUSH, "swarm": evgsim._capi.AGENT_SWARM}
a0 = names[sys.argv[3]] if len(sys.argv) > 4 else names["base_rush"]
a1 = names[sys.argv[4]] if len(sys.argv) > 4 else names["swarm"]
label = "%s vs %s" % tuple(k for v in (a0, a1) for k, vv in names.items() if vv == v)
for mode, per_launch in (("one launch", turns), ("one launch per 50 turns", 50), ("one launch per turn (step_agents)", 0)):
    env = evgsim.BatchedEvergladesEnv(n, seed=0, auto_reset=evgsim._capi.AUTORESET_TERMINAL)
    env.reset()
    env.rollout(150, a0, a1)  # warm-up episode
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    if per_launch:
        for _ in range(turns // per_launch):
            env.rollout(per_launch, a0, a1)
        if turns % per_launch:
            env.rollout(turns % per_launch, a0, a1)
    else:
        for _ in range(turns):
            env.step_agents(a0, a1)
    b.record()
    torch.cuda.synchronize()
    sec = a.elapsed_time(b) / 1e3
    st = env.episode_stats()
    print(json.dumps({"workload": "%s, auto-reset" % label, "matches": n, "turns": turns, "mode": mode,
                      "step_kernel_kind": env._lib.evg_step_kernel_kind(env._h), "env_turns_per_s": n * turns / sec, "us_per_turn": sec * 1e6 / turns,
                      "episodes": st["episodes"], "wins": st["wins"], "ties": st["ties"],
                      "mean_episode_turns": st["total_turns"] / max(st["episodes"], 1), "status_count": st["status_count"]}), flush=True)
    env.close()
