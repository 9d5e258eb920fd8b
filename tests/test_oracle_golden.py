"""The C oracle (oracle/evg_oracle.c) must reproduce the reference's golden trajectories
bit-for-bit: observations, rewards, done flags, every integer state field, node-list order and
fp64 unit health, at every turn of every fixture game."""
import ctypes as C

import numpy as np
import pytest

from conftest import state_fields, flat_health
from oracle import evg_oracle as eo


def _replay(cfg, seed, env_id, game):
    o = eo.OracleEnv(cfg, seed, env_id)
    assert np.array_equal(o.observe(), game["obs"][0].astype(np.float64))
    T = len(game["done"])
    for t in range(T):
        obs, rew, done, scores, status = o.step(game["actions"][t])
        st = o.state[0]
        where = "turn %d" % (t + 1)
        assert np.array_equal(obs, game["obs"][t + 1].astype(np.float64)), where
        assert np.array_equal(rew, game["reward"][t]), where
        assert done == game["done"][t], where
        assert (status != 0) == bool(done)
        assert np.array_equal(state_fields(st), game["grp"][t + 1]), where
        n = cfg.n_nodes
        assert np.array_equal(st["control_state"][1:n + 1], game["node"][t + 1][:, 0]), where
        assert np.array_equal(st["controlled_by"][1:n + 1], game["node"][t + 1][:, 1]), where
        assert np.array_equal(flat_health(st, cfg), game["health"][t + 1]), where
        assert np.array_equal(eo.list_rank(st), game["rank"][t + 1]), where
        assert st["turn"] == t + 1
    return T


def test_all_golden_games(cfg, golden):
    turns = 0
    for i, game in enumerate(golden.games):
        turns += _replay(cfg, golden.seed, i, game)
    assert turns > 3000


def test_golden_fixture_covers_the_rules(golden):
    done_early = sum(1 for g in golden.games if len(g["done"]) < 150)
    wiped = sum(1 for g in golden.games if (g["grp"][-1][:, :, 5].sum(axis=1) == 12).any())
    ties = sum(1 for g in golden.games if g["done"][-1] and g["reward"][-1][0] == g["reward"][-1][1])
    flips = sum(1 for g in golden.games if ((g["node"][1:, :, 1] == -1) & (g["node"][:-1, :, 1] != -1)).any())
    assert done_early >= 5 and wiped >= 3 and ties >= 1 and flips >= 5


def test_initial_observation_known_answer(cfg):
    """initial obs of both players."""
    o = eo.OracleEnv(cfg).observe()
    p0 = [0, 0, 0, 500, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0] + [0, 0, 0, 0] * 3 + [0, 1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0,
                                                                                      0, 0, -500, 100]
    types = [1, 2, 0] * 4
    for g in range(12):
        p0 += [1, types[g], 100, 0, 12 if g == 11 else 8]
    assert o[0].tolist() == [float(x) for x in p0]
    p1 = list(p0)
    p1[3], p1[43] = -500, 500
    assert o[1].tolist() == [float(x) for x in p1]


def test_noop_turn_scores(cfg):
    o = eo.OracleEnv(cfg)
    obs, rew, done, scores, status = o.step(np.zeros((2, 7, 2), dtype=np.int32))
    assert scores.tolist() == [1100, 1100] and status == 0 and done == 0
    assert rew.tolist() == [1100 / 3700, 1100 / 3700]


@pytest.mark.parametrize("n", list(range(1, 17)))
def test_pairwise_sum_matches_numpy(n):
    rng = np.random.default_rng(n)
    for _ in range(2000):
        a = rng.random(n) * 100
        a[rng.random(n) < 0.3] = 0
        got = eo.lib().evo_np_pairwise_sum(a.ctypes.data_as(C.c_void_p), n)
        assert got == float(np.sum(a))


def test_fp64_hits_to_kill(cfg):
    """a tank (armour 3, no node defence) dies after 31 single hits in fp64."""
    h, hits = 100.0, 0
    while h > 0:
        h = h - (10.0 * 1) / (3.0 + 0.0)
        hits += 1
    assert hits == 31
