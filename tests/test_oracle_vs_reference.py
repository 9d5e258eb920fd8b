"""Live check of the oracle against the unmodified reference (only where /root/reference exists;
the committed golden fixtures carry the same guarantee to machines where it does not)."""
import numpy as np
import pytest

from conftest import state_fields, flat_health
from oracle import ref_harness as rh, evg_oracle as eo

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="reference checkout not present")


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_random_game_matches_live_reference(cfg, seed):
    rng = np.random.default_rng(seed)
    acts = np.zeros((150, 2, 7, 2), dtype=np.int64)
    for t in range(150):
        for p in range(2):
            acts[t, p, :, 0] = rng.permutation(12)[:7]
            acts[t, p, :, 1] = rng.permutation(np.arange(1, 12))[:7]
    ref = rh.run_reference_game(seed, seed * 7, acts)
    o = eo.OracleEnv(cfg, seed, seed * 7)
    for t in range(len(ref["done"])):
        obs, rew, done, _, _ = o.step(acts[t])
        assert np.array_equal(obs, ref["obs"][t + 1])
        assert np.array_equal(rew, ref["reward"][t]) and done == ref["done"][t]
        assert np.array_equal(state_fields(o.state[0]), ref["grp"][t + 1])
        assert np.array_equal(flat_health(o.state[0], cfg), ref["health"][t + 1])
        assert np.array_equal(eo.list_rank(o.state[0]), ref["rank"][t + 1])


def test_float_actions_truncate_like_astype_int(cfg):
    """server.py:232 `action.astype(int)`: (0.9, 2.9) commands group 0 to node 2."""
    acts = np.zeros((1, 2, 7, 2))
    acts[0, 0, 0] = (0.9, 2.9)
    ref = rh.run_reference_game(1, 0, acts)
    assert ref["grp"][1][0, 0].tolist()[:5] == [1, 2, 6, 0, 1]
