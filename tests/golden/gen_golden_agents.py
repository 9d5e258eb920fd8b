"""Generate tests/golden/agents_v1.npz: games played by the reference's OWN scripted agents
(agents/State_Machine/base_rush_v1.py, swarm_agent.py) on the unmodified reference env (build container only).

    python tests/golden/gen_golden_agents.py

SwarmAgent's only randomness is ``np.random.shuffle`` of its module-level ATTACK_LIST (swarm_agent.py:86-87);
it is patched for the duration of a game with the tape version (oracle/tape.py: swarm_shuffle), the same way
combat draws are.  base_rushV1 is deterministic.  Not covered: two SwarmAgents in one process share ONE list in
the reference (module global); the batched agents keep one list per (match, player).
"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh, tape  # noqa: E402

SEED = 777
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "agents_v1.npz")
PAIRS = [("base_rush", "swarm"), ("swarm", "base_rush"), ("base_rush", "base_rush"), ("base_rush", "swarm"),
         ("swarm", "base_rush"), ("base_rush", "swarm")]


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(rh.REFERENCE_ROOT, rel))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def play(env_id, kinds, turns=150):
    br = _load("ref_base_rush", "agents/State_Machine/base_rush_v1.py")
    sw = _load("ref_swarm", "agents/State_Machine/swarm_agent.py")  # fresh module: fresh ATTACK_LIST
    agents = [br.base_rushV1(7, p) if k == "base_rush" else sw.SwarmAgent(7, p) for p, k in enumerate(kinds)]
    ctx = {"t": 0, "p": 0}
    orig = np.random.shuffle
    np.random.shuffle = lambda lst: tape.swarm_shuffle(lst, SEED, env_id, ctx["t"] + 1, ctx["p"])
    try:
        def policy(t, obs):
            a = np.zeros((2, 7, 2))
            for p in range(2):
                ctx["t"], ctx["p"] = t, p
                a[p] = agents[p].get_action(obs[p])
            return a
        return rh.run_reference_game(SEED, env_id, policy, n_turns=turns)
    finally:
        np.random.shuffle = orig


def main():
    out = {"seed": np.int64(SEED), "kinds": np.array(["%s,%s" % p for p in PAIRS])}
    for i, kinds in enumerate(PAIRS):
        g = play(i, kinds)
        out["g%d_actions" % i] = g["actions"].astype(np.int8)
        out["g%d_obs" % i] = g["obs"].astype(np.int16)
        out["g%d_reward" % i] = g["reward"]
        out["g%d_done" % i] = g["done"]
        print(i, kinds, "turns", len(g["done"]), "draws", g["n_draws"], "reward", g["reward"][-1])
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
