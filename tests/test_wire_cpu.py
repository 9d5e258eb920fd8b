"""The packed observation row (EVG_OBS_WIRE): evgsim.wire.pack / expand on the reference's own golden observations.
Host logic only (numpy); the rows the GPU writes are checked in tests/test_gpu_wire.py."""
import numpy as np

import evgsim
from evgsim import wire


def test_row_size_is_one_cache_line_on_demomap(cfg):
    assert wire.row_bytes(cfg.n_nodes) == 128
    assert wire.row_bytes(32) == 224 and wire.row_bytes(2) == 96


def test_pack_expand_round_trip_on_golden_games(golden, cfg):
    """Every observation the unmodified reference produced survives float32 -> wire -> float32 bit for bit."""
    total = 0
    for g in golden.games:
        obs = g["obs"][1:].astype(np.float32)          # [T, 2, 105], post-step observations
        T = obs.shape[0]
        rew = g["reward"].astype(np.float32)
        done = g["done"].astype(np.uint8)
        status = np.where(done, 1, 0).astype(np.uint8)
        rows = wire.pack(obs, rew, done, status, cfg)
        assert rows.shape == (T, 128) and rows.dtype == np.uint8
        o2, r2, d2, s2 = wire.expand(rows, cfg)
        assert np.array_equal(o2, obs)
        assert np.array_equal(r2.view(np.uint32), rew.view(np.uint32))
        assert np.array_equal(d2, done) and np.array_equal(s2, status)
        total += T
    assert total > 4000


def test_wire_row_fields_of_the_initial_state(golden, cfg):
    obs0 = golden.games[0]["obs"][:1].astype(np.float32)
    row = wire.pack(obs0, np.zeros((1, 2), np.float32), [0], [0], cfg)[0]
    assert row[0] == 0 and row[1] == 0                                   # turn 0
    assert row[4:6].view("<i2")[0] == 500 and row[6] == 100 and row[7] == 0          # node 1: player 0's base, its 100 units
    assert row[4 + 40:4 + 42].view("<i2")[0] == -500 and row[4 + 42] == 0 and row[4 + 43] == 100   # node 11
    g0 = 4 + 44
    assert tuple(row[g0:g0 + 3]) == (1, 100, 8)                          # group 0 of player 0: node 1, avg 100, 8 units
    assert tuple(row[g0 + 3 * 23:g0 + 3 * 24]) == (11, 100, 12)          # group 11 of player 1: real node 11, 12 units
    assert not row[g0 + 80:].any()
