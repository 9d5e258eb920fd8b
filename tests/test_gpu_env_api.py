"""The N = 1 drop-in: ``evgsim.EvergladesEnv`` keeps the reference wrapper's signatures, types and error
behaviour (gym_everglades/envs/everglades_env.py:13-116) and replays the golden games through the
dict-in / dict-out API."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def evg():
    import __graft_entry__ as g
    g.build()
    import evgsim
    return evgsim


def test_attributes_match_reference(evg):
    env = evg.EvergladesEnv()
    assert (env.num_turns, env.num_units, env.num_groups, env.num_nodes, env.num_actions_per_turn) == (150, 100, 12, 11, 7)
    assert env.unit_classes == ["controller", "striker", "tank"]
    # spaces as objects (everglades_env.py:25-28,124-143); callers read .shape and pass them on (dqn_training.py:66-67)
    sp = env.observation_space
    assert sp.shape == (105,) and sp.low.shape == (105,) and sp.high.shape == (105,)
    assert sp.low[0] == 1 and sp.high[0] == 151 and sp.low[3] == -100 and sp.high[45] == 11
    acts = env.action_space
    assert len(acts.spaces) == 14 and [s.n for s in acts.spaces] == [12, 12] * 7
    assert evg.MAX_SCORE == 3700


def test_reset_keeps_the_simulator_and_starts_a_new_episode(evg):
    """A training loop calls reset() once per episode: nothing is reallocated while the game files stay the same, and
    the combat tape is keyed on the episode, so identical actions fight differently in successive episodes (the
    reference's global numpy stream carries on across resets too)."""
    env = evg.EvergladesEnv(seed=5)
    kw = dict(players={0: "a", 1: "b"}, config_dir=evg.DEFAULT_CONFIG_DIR, map_file="DemoMap.json", unit_file="UnitDefinitions.json")
    # both armies march to the centre node 5 (its own number for both players): p0 1->2->5, p1 11->8 (its own 2)->5
    script = [np.array([[g, 2] for g in range(7)]), np.array([[g, 5] for g in range(7)])]
    healths = []
    first_obs = None
    for ep in range(3):
        obs = env.reset(**kw)
        if first_obs is None:
            first_obs, sim = obs, env._env
        assert env._env is sim and np.array_equal(obs[0], first_obs[0]) and np.array_equal(obs[1], first_obs[1])
        for t in range(60):
            a = script[0] if t < 12 else script[1]
            env.step({0: a, 1: a})
        st = env._env.get_state()[0]
        assert st["episode"] == ep and st["turn"] == 60
        healths.append(st["health"].copy())
    assert healths[0].min() < 100.0, "the script must lead to combat"
    assert not np.array_equal(healths[0], healths[1]) and not np.array_equal(healths[1], healths[2])
    # other game files: a new simulator
    env.reset(setup_file=None, **kw)
    env.close()


@pytest.mark.parametrize("game_idx", [0, 7, 15, 25, 27, 31])
def test_golden_game_through_dict_api(evg, golden, game_idx):
    g = golden.games[game_idx]
    env = evg.EvergladesEnv(seed=golden.seed)
    obs = env.reset(players={0: "a", 1: "b"}, config_dir=evg.DEFAULT_CONFIG_DIR, map_file="DemoMap.json",
                    unit_file="UnitDefinitions.json", output_dir="/tmp", pnames={0: "a", 1: "b"}, debug=False,
                    env_id=game_idx)
    assert set(obs.keys()) == {0, 1} and obs[0].dtype == np.float64 and obs[0].shape == (105,)
    assert np.array_equal(obs[0], g["obs"][0][0]) and np.array_equal(obs[1], g["obs"][0][1])
    for t in range(len(g["done"])):
        a = g["actions"][t].astype(np.float64) + 0.4  # floats are truncated like astype(int), server.py:232
        obs, reward, done, info = env.step({0: a[0], 1: a[1]})
        assert info == {} and isinstance(done, int)
        assert np.array_equal(obs[0], g["obs"][t + 1][0]) and np.array_equal(obs[1], g["obs"][t + 1][1])
        assert reward[0] == g["reward"][t][0] and reward[1] == g["reward"][t][1]  # float64, bit-exact
        assert done == g["done"][t]
    env.close()


def test_error_behaviour_like_reference(evg):
    env = evg.EvergladesEnv()
    with pytest.raises(AssertionError):
        env.reset(players={0: "a"}, config_dir=evg.DEFAULT_CONFIG_DIR, map_file="DemoMap.json", unit_file="UnitDefinitions.json")
    env.reset(players={0: "a", 1: "b"}, config_dir=evg.DEFAULT_CONFIG_DIR, map_file="DemoMap.json", unit_file="UnitDefinitions.json")
    z = np.zeros((7, 2))
    with pytest.raises(AssertionError):
        env.step({0: np.zeros((7, 3)), 1: z})
    with pytest.raises(IndexError):
        env.step({0: np.array([[12, 2]]), 1: z})       # gid 12, SURVEY Appendix B
    with pytest.raises(IndexError):
        env.step({0: z, 1: np.array([[0, 12]])})       # player-1 nid 12
    # gid -1 commands group 11; p0 nid 12 is ignored; only 7 rows count; missing player is skipped
    obs, _, _, _ = env.step({0: np.array([[-1, 2], [0, 12]] + [[1, 0]] * 6 + [[2, 4]])})
    assert obs[0][45 + 5 * 11 + 3] == 1 and obs[0][45 + 3] == 0 and obs[0][45 + 5 * 2 + 3] == 0
    env.close()
