"""The oracle's scripted agents (evo_agent_base_rush / evo_agent_swarm) reproduce, row for row, what the
reference's own Python agents played in the golden games of tests/golden/agents_v1.npz."""
import os

import numpy as np

from oracle import evg_oracle as eo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "agents_v1.npz")


def load_agent_games():
    z = np.load(GOLD)
    kinds = [str(k).split(",") for k in z["kinds"]]
    games = [{k: z["g%d_%s" % (i, k)] for k in ("actions", "obs", "reward", "done")} for i in range(len(kinds))]
    return int(z["seed"]), kinds, games


def test_oracle_agents_match_reference_agents(cfg):
    seed, kinds, games = load_agent_games()
    for i, (kk, g) in enumerate(zip(kinds, games)):
        o = eo.OracleBatch(cfg, 1, seed=seed, first=i)
        o.reset()
        ag = eo.ScriptedAgents(1)
        for t in range(len(g["done"])):
            rows = np.stack([ag.rows(kk[p], cfg, o.states, seed, i, p)[0] for p in range(2)])
            assert np.array_equal(rows, g["actions"][t]), (i, kk, t)
            obs, rew, done = o.step(rows[None])
            assert np.array_equal(obs[0], g["obs"][t + 1].astype(np.float64)), (i, t)
            assert np.array_equal(rew[0], g["reward"][t]) and done[0] == g["done"][t]


def test_base_rush_first_turn_is_blown_and_counters_cycle(cfg):
    o = eo.OracleBatch(cfg, 1)
    o.reset()
    ag = eo.ScriptedAgents(1)
    r0 = ag.rows("base_rush", cfg, o.states, 0, 0, 0)[0]
    assert not r0.any()                                            # base_rush_v1.py:73-76
    r1 = ag.rows("base_rush", cfg, o.states, 0, 0, 0)[0]
    assert r1.tolist() == [[g, 2] for g in range(1, 8)]           # group_num from 1, node_num 2 (:55-56)
    r2 = ag.rows("base_rush", cfg, o.states, 0, 0, 0)[0]
    assert r2.tolist() == [[8, 2], [9, 2], [10, 2], [11, 2], [0, 3], [1, 3], [2, 3]]
