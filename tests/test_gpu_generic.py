"""Config generality: other maps, unit tables, budgets and GameSetup values must stay
bit-exact with the oracle — including group sizes that hit every branch of numpy's pairwise sum
(n < 8, 8 <= n < 16 with a remainder, n == 16) and the 'DEFEND' fortress bonus DemoMap never enables."""
import json

import numpy as np
import pytest

from test_gpu_parity import adjacent_actions, assert_states_equal

pytestmark = pytest.mark.gpu


def write_configs(tmp_path, budget, turn_limit=60, bonus=300):
    nodes = []
    edges = {1: [(2, 2), (3, 5)], 2: [(1, 2), (3, 3), (4, 1)], 3: [(1, 5), (2, 3), (4, 2), (5, 3)], 4: [(2, 1), (3, 2), (5, 2), (6, 4)],
             5: [(3, 3), (4, 2), (6, 2), (7, 5)], 6: [(4, 4), (5, 2), (7, 2)], 7: [(5, 5), (6, 2)]}
    res = {2: ["DEFEND"], 6: ["DEFEND", "OBSERVE"], 4: ["DEFENSE"], 3: ["OBSERVE"]}
    for i in range(1, 8):
        nodes.append({"ID": i, "Connections": [{"ConnectedID": d, "Distance": w} for d, w in edges[i]],
                      "ControlPoints": 250 if i in (1, 7) else 60 + 10 * i, "Resource": res.get(i, []),
                      "StructureDefense": [0, 1, 1.25, 2, 1.5, 1.75, 0.5, 3][i], "TeamStart": {1: 0, 7: 1}.get(i, -1)})
    (tmp_path / "Map7.json").write_text(json.dumps({"MapName": "seven", "nodes": nodes, "P1NodeMap": [0, 7, 6, 5, 4, 3, 2, 1]}))
    units = [{"Name": "Tank", "Health": 4, "Damage": 2, "Speed": 1, "Control": 1, "Cost": 3},
             {"Name": "Scout", "Health": 0.5, "Damage": 1, "Speed": 3, "Control": 0, "Cost": 1},
             {"Name": "Controller", "Health": 2.5, "Damage": 1, "Speed": 2, "Control": 3, "Cost": 2},
             {"Name": "Striker", "Health": 1, "Damage": 5, "Speed": 2, "Control": 1, "Cost": 2}]
    (tmp_path / "Units4.json").write_text(json.dumps({"units": units}))
    (tmp_path / "Setup.json").write_text(json.dumps({"TurnLimit": turn_limit, "CaptureBonus": bonus, "UnitBudget": budget}))
    return str(tmp_path)


@pytest.mark.parametrize("kernel", ["tpm"])
@pytest.mark.parametrize("budget", [60, 100, 120, 180, 192, 12])
def test_other_map_units_and_budgets(tmp_path, budget, kernel, monkeypatch):
    monkeypatch.setenv("EVG_STEP_KERNEL", kernel)
    import __graft_entry__ as g
    g.build()
    import evgsim
    from oracle import evg_oracle as eo

    d = write_configs(tmp_path, budget)
    cfg = evgsim.load_config(d, "Map7.json", "Units4.json", "Setup.json", auto_reset=1)
    assert cfg.n_nodes == 7 and cfg.turn_limit == 60
    n = 768
    env = evgsim.BatchedEvergladesEnv(n, seed=budget, config=cfg, auto_reset=1, env_id_offset=9)
    ora = eo.OracleBatch(cfg, n, seed=budget, first=9)
    assert env.obs_len == 1 + 4 * 7 + 60
    assert np.array_equal(env.reset().cpu().numpy(), ora.reset().astype(np.float32))
    rng = np.random.default_rng(budget)
    deaths = 0
    for t in range(150):
        acts = adjacent_actions(rng, ora.states, cfg)
        obs, rew, done, info = env.step(acts)
        oobs, orew, odone = ora.step(acts)
        assert np.array_equal(done.cpu().numpy(), odone), t
        assert np.array_equal(obs.cpu().numpy(), oobs.astype(np.float32)), t
        assert np.array_equal(rew.cpu().numpy(), orew.astype(np.float32)), t
        deaths += int((ora.states["groups"]["destroyed"]).sum())
        if t % 25 == 24:
            assert_states_equal(env.get_state(), ora.states, "turn %d" % (t + 1))
    assert deaths > 0  # the scenario really fights
    assert env.episode_stats()["episodes"] >= 2 * n


def write_ring32(tmp_path):
    """The largest supported map: 32 nodes on a ring with chords, bases at 1 and 17, mirror map i -> ((i + 15) % 32) + 1."""
    n = 32
    nodes = []
    for i in range(1, n + 1):
        nb = {(i % n) + 1: 2 + i % 3, ((i - 2) % n) + 1: 2 + (i - 1) % 3, ((i + 7) % n) + 1: 5, ((i - 9) % n) + 1: 5}
        nodes.append({"ID": i, "Connections": [{"ConnectedID": d, "Distance": w} for d, w in sorted(nb.items())],
                      "ControlPoints": 300 if i in (1, 17) else 50 + i, "Resource": ["DEFEND"] if i % 5 == 0 else (["OBSERVE"] if i % 7 == 0 else []),
                      "StructureDefense": 1 + (i % 4) * 0.25, "TeamStart": {1: 0, 17: 1}.get(i, -1)})
    p1 = [0] + [((i + 15) % n) + 1 for i in range(1, n + 1)]
    (tmp_path / "Ring32.json").write_text(json.dumps({"MapName": "ring32", "nodes": nodes, "P1NodeMap": p1}))
    (tmp_path / "Setup.json").write_text(json.dumps({"TurnLimit": 50, "CaptureBonus": 700, "UnitBudget": 100}))
    return str(tmp_path)


@pytest.mark.parametrize("kernel", ["tpm", "warp"])
def test_largest_map_32_nodes(tmp_path, kernel, monkeypatch):
    """32 nodes: 352-byte records (more than one 8-byte word per lane), 189-value observations, the byte-array
    variant of the random agent, run-time-sized kernels."""
    monkeypatch.setenv("EVG_STEP_KERNEL", kernel)
    import __graft_entry__ as g
    g.build()
    import evgsim
    from oracle import evg_oracle as eo

    d = write_ring32(tmp_path)
    cfg = evgsim.load_config(d, "Ring32.json", evgsim.DEFAULT_CONFIG_DIR + "/UnitDefinitions.json", "Setup.json", auto_reset=1)
    n = 300
    env = evgsim.BatchedEvergladesEnv(n, seed=8, config=cfg, auto_reset=1, env_id_offset=40)
    ora = eo.OracleBatch(cfg, n, seed=8, first=40)
    assert env.obs_len == 1 + 4 * 32 + 60 and env.layout.record_bytes == 352
    assert np.array_equal(env.reset().cpu().numpy(), ora.reset().astype(np.float32))
    rng = np.random.default_rng(3)
    for t in range(120):
        if t % 4 == 3:  # the on-device random agent on a map too large for its register-only variant
            acts = env.random_actions().cpu().numpy()
            for i in range(0, n, 37):
                for p in range(2):
                    want = eo.agent_random(cfg, 8, 40 + i, int(ora.states[i]["episode"]), int(ora.states[i]["turn"]) + 1, p)
                    assert np.array_equal(acts[i, p], want.astype(np.int8))
        else:
            acts = adjacent_actions(rng, ora.states, cfg)
        obs, rew, done, info = env.step(acts)
        oobs, orew, odone = ora.step(acts)
        assert np.array_equal(done.cpu().numpy(), odone), t
        assert np.array_equal(obs.cpu().numpy(), oobs.astype(np.float32)), t
        assert np.array_equal(rew.cpu().numpy(), orew.astype(np.float32)), t
    assert_states_equal(env.get_state(), ora.states, "end")
    assert env.episode_stats()["episodes"] >= 2 * n


@pytest.mark.parametrize("kernel", ["tpm", "warp"])
def test_simulators_of_different_configurations_coexist(tmp_path, kernel, monkeypatch):
    """The step kernels are shared by every simulator of a process; one that needs less shared memory, created
    later, must not take the opt-in away from an earlier one that needs more (32-node map vs 7-node map vs DemoMap)."""
    monkeypatch.setenv("EVG_STEP_KERNEL", kernel)
    import __graft_entry__ as g
    g.build()
    import evgsim
    from oracle import evg_oracle as eo

    d32 = write_ring32(tmp_path)
    big_cfg = evgsim.load_config(d32, "Ring32.json", evgsim.DEFAULT_CONFIG_DIR + "/UnitDefinitions.json", "Setup.json")
    n = 200
    big = evgsim.BatchedEvergladesEnv(n, seed=4, config=big_cfg)
    big_ora = eo.OracleBatch(big_cfg, n, seed=4, first=0)
    big.reset()
    big_ora.reset()
    small_cfg = evgsim.load_config()  # DemoMap: the smallest rows
    small = evgsim.BatchedEvergladesEnv(n, seed=4, config=small_cfg)
    small_ora = eo.OracleBatch(small_cfg, n, seed=4, first=0)
    small.reset()
    small_ora.reset()
    rng = np.random.default_rng(8)
    for t in range(40):
        for env, ora, cfg in ((big, big_ora, big_cfg), (small, small_ora, small_cfg)):
            acts = adjacent_actions(rng, ora.states, cfg)
            obs, rew, done, info = env.step(acts)
            oobs, orew, odone = ora.step(acts)
            assert np.array_equal(obs.cpu().numpy(), oobs.astype(np.float32)), t
            assert np.array_equal(done.cpu().numpy(), odone), t


@pytest.mark.parametrize("sizes", [
    [12, 9, 10, 11, 12, 9, 10, 11, 4, 4, 4, 1],     # 8 groups of 9..12 slots: every one fights on two lanes (segments of 8 slots)
    [8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 9],           # budget 97: the remainder group has ONE slot in its second segment
    [10, 10, 10, 10, 10, 10, 10, 10, 10, 3, 3, 3],  # 9 groups above 8 slots: more than a round can split, run-time-sized kernel
    [12, 1, 2, 3, 4, 5, 6, 7, 8, 12, 12, 12],       # every size below 8 next to four full ones
])
@pytest.mark.parametrize("lite", [0, 1])
def test_demomap_loadouts_with_groups_split_over_two_lanes(sizes, lite, monkeypatch):
    """DemoMap with other loadouts: the compile-time-map kernel updates a group of more than 8 unit slots on two lanes
    (slots [0, 8) and [8, size)), the first of which rebuilds numpy's pairwise sum and the alive mask from both — for
    second segments of 1..4 slots, with up to 8 such groups per player side by side with small ones; fp64 health, alive
    masks, averages and everything else must stay equal to the oracle's through fights and auto-resets."""
    monkeypatch.setenv("EVG_STEP_KERNEL", "tpm")
    monkeypatch.setenv("EVG_TPM_SMALL_MAX", str(1 << 30) if lite else "0")
    import __graft_entry__ as g
    g.build()
    import evgsim
    from oracle import evg_oracle as eo

    cfg = evgsim.load_config(auto_reset=1, turn_limit=70)
    for p in (0, 1):
        for gi, sz in enumerate(sizes):
            cfg.group_size[p][gi] = sz
    n = 640
    env = evgsim.BatchedEvergladesEnv(n, seed=sum(sizes), config=cfg, auto_reset=1, env_id_offset=3)
    ora = eo.OracleBatch(cfg, n, seed=sum(sizes), first=3)
    eo.fought_slots(clear=True)
    assert np.array_equal(env.reset().cpu().numpy(), ora.reset().astype(np.float32))
    rng = np.random.default_rng(len(sizes) + sizes[1])
    deaths = 0
    for t in range(150):
        acts = adjacent_actions(rng, ora.states, cfg)
        obs, rew, done, info = env.step(acts)
        oobs, orew, odone = ora.step(acts)
        assert np.array_equal(done.cpu().numpy(), odone), t
        assert np.array_equal(obs.cpu().numpy(), oobs.astype(np.float32)), t
        assert np.array_equal(rew.cpu().numpy(), orew.astype(np.float32)), t
        deaths += int((ora.states["groups"]["destroyed"]).sum())
        if t % 30 == 29:
            assert_states_equal(env.get_state(), ora.states, "turn %d" % (t + 1))
    assert deaths > 0
    st = env.episode_stats()
    assert st["episodes"] >= 2 * n and st["fought_unit_slots"] == eo.fought_slots()


@pytest.mark.parametrize("budget", [100, 192])
def test_multi_turn_rollout_on_another_map(tmp_path, budget, monkeypatch):
    """evg_rollout's K-turns-per-launch path in the run-time-sized instantiations (7-node map, byte and 16-bit damage
    histograms): same state, outputs and statistics as evg_step_agents turn by turn, with in-place resets (turn limit 60)."""
    monkeypatch.setenv("EVG_STEP_KERNEL", "tpm")
    import __graft_entry__ as g
    g.build()
    import evgsim

    A = evgsim._capi
    d = write_configs(tmp_path, budget)
    cfg = evgsim.load_config(d, "Map7.json", "Units4.json", "Setup.json", auto_reset=1)
    n = 1500
    for a0, a1 in ((A.AGENT_RANDOM, A.AGENT_SWARM), (A.AGENT_BASE_RUSH, A.AGENT_RANDOM)):
        r = evgsim.BatchedEvergladesEnv(n, seed=budget, config=cfg, auto_reset=1, env_id_offset=4)
        p = evgsim.BatchedEvergladesEnv(n, seed=budget, config=cfg, auto_reset=1, env_id_offset=4)
        r.reset()
        p.reset()
        r.rollout(47, a0, a1)
        r.rollout(90, a0, a1)
        for _ in range(137):
            p.step_agents(a0, a1)
        assert bool((r.obs == p.obs).all()) and bool((r.reward == p.reward).all()) and bool((r.done == p.done).all())
        assert_states_equal(r.get_state(), p.get_state(), "after the rollout")
        assert r.episode_stats() == p.episode_stats() and r.episode_stats()["episodes"] >= 2 * n
