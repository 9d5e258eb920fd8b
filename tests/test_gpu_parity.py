"""Parity of the CUDA path (through the C ABI) with the reference's golden trajectories and with the
CPU oracle on the same seeded inputs.  Bar: BIT-EXACT for observations, rewards (float32 of the
reference's float64), done flags, every integer state field, node-list order and fp64 unit health.
"""
import numpy as np
import pytest

from conftest import state_fields, flat_health

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def evg():
    import __graft_entry__ as g
    g.build()
    import evgsim
    return evgsim


@pytest.fixture(scope="module")
def eo():
    from oracle import evg_oracle
    return evg_oracle


def rank_of(states):
    from oracle import evg_oracle
    return np.stack([evg_oracle.list_rank(s) for s in states])


def assert_states_equal(gpu, ora, where=""):
    for name in ("turn", "episode", "control_state", "controlled_by", "health"):
        assert np.array_equal(gpu[name], ora[name]), (where, name)
    for name in ("location", "travel_destination", "distance_remaining", "ready", "moving", "destroyed", "count",
                 "arrival", "avg_health"):
        assert np.array_equal(gpu["groups"][name], ora["groups"][name]), (where, name)


# --------------------------------------------------------------------------------------------- golden
def test_golden_trajectories_bit_exact(evg, golden, cfg):
    """All 37 reference games replayed as ONE lock-step batch (match id i = fixture game i)."""
    n = len(golden)
    env = evg.BatchedEvergladesEnv(n, seed=golden.seed)
    obs = env.reset().cpu().numpy()
    for i, g in enumerate(golden.games):
        assert np.array_equal(obs[i], g["obs"][0].astype(np.float32)), i
    T = max(len(g["done"]) for g in golden.games)
    checked = 0
    for t in range(T):
        acts = np.zeros((n, 2, 7, 2), dtype=np.int8)
        for i, g in enumerate(golden.games):
            if t < len(g["done"]):
                acts[i] = g["actions"][t][:, :7, :]
        obs, rew, done, info = env.step(acts)
        obs, rew, done = obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy()
        st = env.get_state()
        for i, g in enumerate(golden.games):
            if t >= len(g["done"]):
                continue
            where = "game %d (%s) turn %d" % (i, golden.names[i], t + 1)
            assert np.array_equal(obs[i], g["obs"][t + 1].astype(np.float32)), where
            assert np.array_equal(rew[i], g["reward"][t].astype(np.float32)), where
            assert done[i] == g["done"][t], where
            assert np.array_equal(state_fields(st[i]), g["grp"][t + 1]), where
            assert np.array_equal(st[i]["control_state"][1:12], g["node"][t + 1][:, 0]), where
            assert np.array_equal(st[i]["controlled_by"][1:12], g["node"][t + 1][:, 1]), where
            assert np.array_equal(flat_health(st[i], cfg), g["health"][t + 1]), where
            assert np.array_equal(rank_of(st[i:i + 1])[0], g["rank"][t + 1]), where
            checked += 1
    assert checked == sum(len(g["done"]) for g in golden.games)


# --------------------------------------------------------------------------------------------- oracle, at scale
def adjacent_actions(rng, states, cfg, p_move=0.9):
    """7 random groups per player, each sent to a random neighbour of where it stands (own numbering)."""
    n = len(states)
    adj = [[b for b in range(1, cfg.n_nodes + 1) if cfg.edge_distance[a][b]] or [0] for a in range(cfg.n_nodes + 1)]
    maxd = max(len(a) for a in adj)
    table = np.zeros((cfg.n_nodes + 1, maxd), dtype=np.int64)
    deg = np.zeros(cfg.n_nodes + 1, dtype=np.int64)
    for a, nb in enumerate(adj):
        table[a, :len(nb)] = nb
        deg[a] = len(nb)
    p1map = np.array(list(cfg.p1_node_map)[:cfg.n_nodes + 1])
    acts = np.zeros((n, 2, 7, 2), dtype=np.int8)
    loc = states["groups"]["location"].astype(np.int64)  # [n,2,12]
    for p in range(2):
        gids = np.argsort(rng.random((n, 12)), axis=1)[:, :7]
        l = np.take_along_axis(loc[:, p], gids, axis=1)
        pick = (rng.random((n, 7)) * deg[l]).astype(np.int64)
        real = table[l, pick]
        node = p1map[real] if p else real
        node = np.where(rng.random((n, 7)) < p_move, node, 0)
        acts[:, p, :, 0] = gids
        acts[:, p, :, 1] = node
    return acts


def uniform_actions(rng, n, n_nodes=11):
    acts = np.zeros((n, 2, 7, 2), dtype=np.int8)
    for p in range(2):
        acts[:, p, :, 0] = np.argsort(rng.random((n, 12)), axis=1)[:, :7]
        acts[:, p, :, 1] = np.argsort(rng.random((n, n_nodes)), axis=1)[:, :7] + 1
    return acts


def run_against_oracle(evg, eo, cfg, n, turns, seed, first, make_actions, auto_reset=0, state_every=10):
    cfg.auto_reset = auto_reset
    env = evg.BatchedEvergladesEnv(n, seed=seed, config=cfg, auto_reset=auto_reset, env_id_offset=first)
    ora = eo.OracleBatch(cfg, n, seed=seed, first=first)
    assert np.array_equal(env.reset().cpu().numpy(), ora.reset().astype(np.float32))
    ndone = 0
    eo.fought_slots(clear=True)
    for t in range(turns):
        acts = make_actions(ora.states)
        obs, rew, done, info = env.step(acts)
        oobs, orew, odone = ora.step(acts)
        where = "turn %d" % (t + 1)
        assert np.array_equal(done.cpu().numpy(), odone), where
        assert np.array_equal(obs.cpu().numpy(), oobs.astype(np.float32)), where
        assert np.array_equal(rew.cpu().numpy(), orew.astype(np.float32)), where
        assert np.array_equal(info["scores"].cpu().numpy(), ora.scores), where
        assert np.array_equal(info["status"].cpu().numpy(), ora.status), where
        ndone += int(odone.sum())
        if (t + 1) % state_every == 0 or t == turns - 1:
            assert_states_equal(env.get_state(), ora.states, where)
    # the device's count of unit slots that fought (the health term of bench.py's algorithmic bytes)
    assert env.episode_stats()["fought_unit_slots"] == eo.fought_slots(), "fought_unit_slots"
    return env, ora, ndone


def test_oracle_parity_4096_adjacent_random(evg, eo, cfg):
    """BASELINE config 2 size: 4096 lock-step matches, a full 150-turn episode, heavy traffic/combat."""
    rng = np.random.default_rng(1)
    run_against_oracle(evg, eo, cfg, 4096, 150, seed=11, first=0, make_actions=lambda s: adjacent_actions(rng, s, cfg))


def test_oracle_parity_16384_default_kernel_choice(evg, eo, cfg, monkeypatch):
    """A mid-size batch takes the thread-per-match kernel by default (evg_create's choice): every one of the 16,384
    matches is followed on the oracle through a full episode with in-place auto-reset."""
    monkeypatch.delenv("EVG_STEP_KERNEL", raising=False)
    monkeypatch.delenv("EVG_TPM_SMALL_MAX", raising=False)
    rng = np.random.default_rng(5)
    try:
        env, ora, ndone = run_against_oracle(evg, eo, cfg, 16384, 155, seed=41, first=1 << 21,
                                             make_actions=lambda s: adjacent_actions(rng, s, cfg), auto_reset=1, state_every=31)
        assert env._lib.evg_step_kernel_kind(env._h) == 1
        assert ndone >= 16384
    finally:
        cfg.auto_reset = 0


def test_oracle_parity_uniform_random_with_offset(evg, eo, cfg):
    """random_actions-style rows (mostly invalid moves), match ids starting at a large offset."""
    rng = np.random.default_rng(2)
    run_against_oracle(evg, eo, cfg, 1024, 150, seed=0xDEADBEEFCAFE, first=3_000_000_000,
                       make_actions=lambda s: uniform_actions(rng, len(s)))


@pytest.mark.parametrize("mode", [1, 2])
def test_auto_reset_modes_match_oracle(evg, eo, cfg, mode):
    """In-place auto-reset (terminal obs / next obs) over several episodes, incl. episode statistics."""
    rng = np.random.default_rng(3 + mode)
    cfg.turn_limit = 40
    try:
        env, ora, ndone = run_against_oracle(evg, eo, cfg, 512, 130, seed=5, first=17,
                                             make_actions=lambda s: adjacent_actions(rng, s, cfg), auto_reset=mode)
    finally:
        cfg.turn_limit = 150
        cfg.auto_reset = 0
    st = env.episode_stats()
    assert st["episodes"] == ndone >= 3 * 512
    assert st["wins"][0] + st["wins"][1] + st["ties"] == ndone
    assert sum(st["status_count"]) == ndone and st["status_count"][0] == 0
    assert st["env_turns"] == 512 * 130
    assert (ora.states["episode"] >= 3).all()


def test_out_of_range_rows_are_noops(evg, eo, cfg):
    """Rows the reference answers with IndexError are ignored (documented divergence) — same as the oracle."""
    rng = np.random.default_rng(9)

    def make(states):
        a = adjacent_actions(rng, states, cfg)
        n = len(states)
        junk = rng.random((n, 2, 7)) < 0.3
        a[..., 0] = np.where(junk, rng.integers(-128, 128, (n, 2, 7)), a[..., 0])
        junk = rng.random((n, 2, 7)) < 0.3
        a[..., 1] = np.where(junk, rng.integers(-128, 128, (n, 2, 7)), a[..., 1])
        dup = rng.random((n, 2)) < 0.5
        a[:, :, 1, 0] = np.where(dup, a[:, :, 0, 0], a[:, :, 1, 0])
        return a.astype(np.int8)

    run_against_oracle(evg, eo, cfg, 512, 100, seed=21, first=0, make_actions=make)


def test_import_export_roundtrip_and_resume(evg, eo, cfg):
    """evg_export_state / evg_import_state: a mid-game oracle snapshot resumes identically on the GPU."""
    rng = np.random.default_rng(4)
    n = 256
    ora = eo.OracleBatch(cfg, n, seed=8, first=100)
    ora.reset()
    for t in range(60):
        ora.step(adjacent_actions(rng, ora.states, cfg))
    env = evg.BatchedEvergladesEnv(n, seed=8, config=cfg, env_id_offset=100)
    env.reset()
    env.set_state(ora.states)
    assert_states_equal(env.get_state(), ora.states, "after import")
    for t in range(60):
        acts = adjacent_actions(rng, ora.states, cfg)
        obs, rew, done, _ = env.step(acts)
        oobs, orew, odone = ora.step(acts)
        assert np.array_equal(obs.cpu().numpy(), oobs.astype(np.float32)), t
        assert np.array_equal(rew.cpu().numpy(), orew.astype(np.float32)), t
    assert_states_equal(env.get_state(), ora.states, "after resume")


def test_masked_reset(evg, eo, cfg):
    rng = np.random.default_rng(6)
    n = 128
    env = evg.BatchedEvergladesEnv(n, seed=2, config=cfg)
    ora = eo.OracleBatch(cfg, n, seed=2)
    env.reset()
    ora.reset()
    for t in range(30):
        acts = adjacent_actions(rng, ora.states, cfg)
        env.step(acts)
        ora.step(acts)
    mask = rng.random(n) < 0.5
    obs = env.reset(mask).cpu().numpy()
    oobs = ora.reset(mask)
    assert np.array_equal(obs[mask], oobs[mask].astype(np.float32))
    assert_states_equal(env.get_state(), ora.states, "after masked reset")
    for t in range(30):
        acts = adjacent_actions(rng, ora.states, cfg)
        obs, _, _, _ = env.step(acts)
        oobs, _, _ = ora.step(acts)
        assert np.array_equal(obs.cpu().numpy(), oobs.astype(np.float32)), t


@pytest.mark.parametrize("n", [300, 301, 1027])  # (2n rows of 14 bytes leave the kernel in 16-byte pieces: with and without a tail)
def test_random_agent_kernel_matches_oracle(evg, eo, cfg, n):
    env = evg.BatchedEvergladesEnv(n, seed=77, config=cfg, env_id_offset=5)
    env.reset()
    for turn in (1, 2):
        a = env.random_actions().cpu().numpy()
        one = env.random_actions(player=1, out=env._actions.clone().zero_()).cpu().numpy()  # one player's rows only
        assert np.array_equal(one[:, 1], a[:, 1]) and not one[:, 0].any()
        for i in range(n):
            for p in range(2):
                want = eo.agent_random(cfg, 77, 5 + i, 0, turn, p)
                assert np.array_equal(a[i, p], want.astype(np.int8)), (i, p)
        assert all(len(set(a[i, p, :, 0])) == 7 and len(set(a[i, p, :, 1])) == 7 for i in range(n) for p in range(2))
        env.step(a)


def test_step_host_equals_step(evg, eo, cfg):
    rng = np.random.default_rng(12)
    n = 200
    env = evg.BatchedEvergladesEnv(n, seed=3, config=cfg)
    ora = eo.OracleBatch(cfg, n, seed=3)
    env.reset()
    ora.reset()
    for t in range(40):
        acts = adjacent_actions(rng, ora.states, cfg)
        obs, rew, done, _ = env.step_host(acts)
        oobs, orew, odone = ora.step(acts)
        assert np.array_equal(obs.numpy(), oobs.astype(np.float32))
        assert np.array_equal(rew.numpy(), orew.astype(np.float32))
        assert np.array_equal(done.numpy(), odone)
    assert env.h2d_bytes_per_step() == n * 28 and env.d2h_bytes_per_step() == n * (840 + 8 + 1)


@pytest.mark.parametrize("kernel", ["tpm", "tpm128", "warp"])
def test_every_step_kernel_matches_oracle(evg, eo, cfg, monkeypatch, kernel):
    """The step kernels (two lanes per match / one thread per match in 32- and in 128-thread CTAs / one warp per match)
    stay selectable (EVG_STEP_KERNEL, EVG_TPM_SMALL_MAX) for A/B profiling; each must match the oracle, with auto-reset
    and a tail batch (n % 128 != 0)."""
    # "tpm" = the one-warp-per-CTA instantiation (forced: by default only batches around 64k matches take it), "tpm128" = 128-thread CTAs
    monkeypatch.setenv("EVG_TPM_SMALL_MAX", "0" if kernel == "tpm128" else str(1 << 30))
    kernel = "tpm" if kernel == "tpm128" else kernel
    monkeypatch.setenv("EVG_STEP_KERNEL", kernel)
    rng = np.random.default_rng(31)
    cfg.turn_limit = 70
    try:
        run_against_oracle(evg, eo, cfg, 1000, 160, seed=19, first=7, make_actions=lambda s: adjacent_actions(rng, s, cfg),
                           auto_reset=2)
    finally:
        cfg.turn_limit = 150
        cfg.auto_reset = 0


@pytest.mark.parametrize("kernel", ["default", "tpm", "tpm128"])
def test_fused_agents_equal_agent_kernel_plus_step(evg, eo, cfg, monkeypatch, kernel):
    """evg_step_agents (rows generated inside the step kernel) == evg_agent_random + evg_step == oracle, for the
    batch-size default (warp kernel: rows through the action buffer) and both thread-per-match CTA sizes (fused)."""
    if kernel != "default":
        monkeypatch.setenv("EVG_STEP_KERNEL", "tpm")
        monkeypatch.setenv("EVG_TPM_SMALL_MAX", "0" if kernel == "tpm128" else str(1 << 30))
    n = 640
    cfg.auto_reset = 1
    cfg.turn_limit = 60
    try:
        env = evg.BatchedEvergladesEnv(n, seed=99, config=cfg, auto_reset=1, env_id_offset=11)
        ora = eo.OracleBatch(cfg, n, seed=99, first=11)
        env.reset()
        ora.reset()
        for t in range(130):
            want = np.zeros((n, 2, 7, 2), dtype=np.int8)
            for i in range(n):
                for p in range(2):
                    want[i, p] = eo.agent_random(cfg, 99, 11 + i, int(ora.states[i]["episode"]), int(ora.states[i]["turn"]) + 1, p)
            if t % 3 == 0:    # both scripted, rows reported
                obs, rew, done, info = env.step_agents(want_actions=True)
                assert np.array_equal(info["actions"].cpu().numpy(), want), t
            elif t % 3 == 1:  # both scripted, nothing written
                obs, rew, done, info = env.step_agents()
            else:             # player 0 external, player 1 scripted
                obs, rew, done, info = env.step_agents(evg._capi.AGENT_EXTERNAL, evg._capi.AGENT_RANDOM, actions=want.copy())
            oobs, orew, odone = ora.step(want)
            assert np.array_equal(obs.cpu().numpy(), oobs.astype(np.float32)), t
            assert np.array_equal(rew.cpu().numpy(), orew.astype(np.float32)), t
            assert np.array_equal(done.cpu().numpy(), odone), t
        assert_states_equal(env.get_state(), ora.states, "end")
    finally:
        cfg.auto_reset = 0
        cfg.turn_limit = 150


def test_scripted_agents_match_reference_agent_games(evg, cfg):
    """Config 3 of BASELINE.json: base_rushV1 vs SwarmAgent.  The device agents + device step reproduce the games
    the reference's own Python agents played on the reference env (tests/golden/agents_v1.npz), row for row."""
    from test_agents_cpu import load_agent_games
    seed, kinds, games = load_agent_games()
    ids = {"base_rush": evg._capi.AGENT_BASE_RUSH, "swarm": evg._capi.AGENT_SWARM}
    for i, (kk, g) in enumerate(zip(kinds, games)):
        env = evg.BatchedEvergladesEnv(1, seed=seed, config=cfg, env_id_offset=i)
        obs = env.reset().cpu().numpy()
        assert np.array_equal(obs[0], g["obs"][0].astype(np.float32))
        for t in range(len(g["done"])):
            a = env.agent_actions(ids[kk[0]], ids[kk[1]])
            assert np.array_equal(a.cpu().numpy()[0], g["actions"][t]), (i, kk, t)
            obs, rew, done, _ = env.step(a)
            assert np.array_equal(obs.cpu().numpy()[0], g["obs"][t + 1].astype(np.float32)), (i, t)
            assert np.array_equal(rew.cpu().numpy()[0], g["reward"][t].astype(np.float32)) and int(done[0]) == g["done"][t]


@pytest.mark.parametrize("kernel", ["default", "tpm", "tpm128"])
def test_scripted_agents_batched_with_autoreset_match_oracle(evg, eo, cfg, monkeypatch, kernel):
    """65,536-style run in small: base_rush vs swarm with in-place auto-reset over several matches; agent state
    persists across matches exactly like the oracle's (and the reference's agent objects).  On the thread-per-match
    kernel the agents' rows are generated inside the step kernel: a whole self-play turn is ONE launch."""
    if kernel != "default":
        monkeypatch.setenv("EVG_STEP_KERNEL", "tpm")
        monkeypatch.setenv("EVG_TPM_SMALL_MAX", "0" if kernel == "tpm128" else str(1 << 30))
    n = 512
    cfg.auto_reset = 1
    try:
        env = evg.BatchedEvergladesEnv(n, seed=41, config=cfg, auto_reset=1, env_id_offset=3)
        ora = eo.OracleBatch(cfg, n, seed=41, first=3)
        ag = eo.ScriptedAgents(n)
        env.reset()
        ora.reset()
        for t in range(230):
            want = np.stack([ag.rows("base_rush", cfg, ora.states, 41, 3, 0), ag.rows("swarm", cfg, ora.states, 41, 3, 1)], axis=1)
            if t % 3 == 1:
                obs, rew, done, info = env.step_agents(evg._capi.AGENT_BASE_RUSH, evg._capi.AGENT_SWARM, want_actions=True)
                assert np.array_equal(info["actions"].cpu().numpy(), want), t
            elif t % 3 == 2:
                before = env.launch_count
                obs, rew, done, info = env.step_agents(evg._capi.AGENT_BASE_RUSH, evg._capi.AGENT_SWARM)
                assert env.launch_count - before == (1 if kernel != "default" else 2)  # fused into the step where possible
            else:
                a = env.agent_actions(evg._capi.AGENT_BASE_RUSH, evg._capi.AGENT_SWARM)
                assert np.array_equal(a.cpu().numpy(), want), t
                obs, rew, done, info = env.step(a)
            oobs, orew, odone = ora.step(want)
            assert np.array_equal(obs.cpu().numpy(), oobs.astype(np.float32)), t
            assert np.array_equal(done.cpu().numpy(), odone), t
        assert env.episode_stats()["episodes"] >= 2 * n
        assert_states_equal(env.get_state(), ora.states, "end")
    finally:
        cfg.auto_reset = 0


def test_import_rejects_out_of_range_fields(evg, cfg):
    """Fields the step kernels use as indices are validated: BatchedEvergladesEnv.set_state raises on the host, the C
    entry point forces them into range, reports EVG_E_ARG, and the simulator keeps stepping on sane state."""
    import ctypes as C
    import torch
    env = evg.BatchedEvergladesEnv(64, seed=2, config=cfg)
    with pytest.raises(RuntimeError):
        env.set_state(np.zeros(3, dtype=evg._capi.env_state_dtype()))  # partial import before any reset
    env.reset()
    st = env.get_state()
    for field, value in (("location", 0), ("location", 12), ("travel_destination", 40)):
        bad = st.copy()
        bad["groups"][field][5, 1, 3] = value
        with pytest.raises(ValueError):
            env.set_state(bad)
    bad = st.copy()
    bad["control_state"][7, 4] = 3000
    with pytest.raises(ValueError):
        env.set_state(bad)
    # straight through the C ABI: imported with the field forced into range, and reported
    bad = st.copy()
    bad["groups"]["location"][9, 0, 0] = 63
    buf = torch.from_numpy(bad.view(np.uint8).reshape(-1).copy()).to(env.device)
    rc = env._lib.evg_import_state(env._h, 0, 64, C.c_void_p(buf.data_ptr()), None)
    assert rc == evg._capi.E_ARG if hasattr(evg._capi, "E_ARG") else rc == -1
    assert b"out of range" in env._lib.evg_last_error()
    assert env.get_state()["groups"]["location"][9, 0, 0] == 11
    for _ in range(5):
        env.step(env.random_actions())
    env.set_state(st)  # a valid snapshot still imports
    assert_states_equal(env.get_state(), st, "re-import")


@pytest.mark.parametrize("n,mode", [(600, 1), (16384 + 5, 1), (16384 + 5, 2), (60000, 1), (60000, 2)])
def test_multi_turn_rollout_equals_plain_turns(evg, cfg, n, mode):
    """BatchedEvergladesEnv.rollout plays K scripted self-play turns in ONE launch — the warp-per-match multi-turn kernel
    (600 matches), the 128-thread thread-per-match kernel (16,389: a partial last warp) and its one-warp-CTA instantiation
    (60,000), whose CTAs keep a batch in shared memory for all K turns.  Final state, outputs and statistics must be those
    of step_agents called turn by turn, across in-place resets of both auto-reset modes and for rollouts of 1, 99 and 70
    turns in a row."""
    A = evg._capi
    cfg.auto_reset = mode
    try:
        for a0, a1 in ((A.AGENT_RANDOM, A.AGENT_RANDOM), (A.AGENT_BASE_RUSH, A.AGENT_SWARM)):
            g = evg.BatchedEvergladesEnv(n, seed=13, config=cfg, auto_reset=mode, env_id_offset=2)
            p = evg.BatchedEvergladesEnv(n, seed=13, config=cfg, auto_reset=mode, env_id_offset=2)
            g.reset()
            p.reset()
            launches = g.launch_count
            for k in (1, 99, 70):
                g.rollout(k, a0, a1)
            assert g.launch_count - launches == 3   # one launch per rollout
            for _ in range(170):
                p.step_agents(a0, a1)
            assert bool((g.obs == p.obs).all()) and bool((g.reward == p.reward).all()) and bool((g.done == p.done).all())
            assert bool((g.status == p.status).all()) and bool((g.scores == p.scores).all())
            assert_states_equal(g.get_state(), p.get_state(), "after the rollout")
            sg, sp = g.episode_stats(), p.episode_stats()
            assert sg == sp and sg["episodes"] >= n and sg["env_turns"] == 170 * n
    finally:
        cfg.auto_reset = 0


def test_loss_quotient_paths_agree(evg, eo, cfg, monkeypatch):
    """The thread-per-match kernel forms (10.*dmg)/divisor as reciprocal + two FMAs when evg_create has proved that
    equal to the IEEE quotient for every reachable dmg (Tables::fast_div); EVG_NO_FAST_DIV forces the generic
    instantiation with the loss table / division.  Both must reproduce the oracle's fp64 health bit for bit."""
    monkeypatch.setenv("EVG_STEP_KERNEL", "tpm")
    monkeypatch.setenv("EVG_TPM_SMALL_MAX", "0")
    rng = np.random.default_rng(77)
    for no_fast in ("", "1"):
        if no_fast:
            monkeypatch.setenv("EVG_NO_FAST_DIV", no_fast)
        run_against_oracle(evg, eo, cfg, 700, 110, seed=23, first=3, make_actions=lambda s: adjacent_actions(rng, s, cfg))


def test_step_kernel_kind_follows_batch_size(evg, cfg, monkeypatch):
    """evg_create picks the warp-per-match kernel for small batches and the thread-per-match kernel for large ones;
    scripted agents work with both (rows passed through the action buffer next to the warp kernel)."""
    monkeypatch.delenv("EVG_STEP_KERNEL", raising=False)
    small = evg.BatchedEvergladesEnv(256, seed=1, config=cfg)
    large = evg.BatchedEvergladesEnv(16384, seed=1, config=cfg)
    assert small._lib.evg_step_kernel_kind(small._h) == 0
    assert large._lib.evg_step_kernel_kind(large._h) == 1
    small.reset()
    large.reset()
    for _ in range(40):
        so, sr, sd, _ = small.step_agents()
        lo, lr, ld, _ = large.step_agents()
    assert bool((so == lo[:256]).all()) and bool((sr == lr[:256]).all()) and bool((sd == ld[:256]).all())


@pytest.mark.parametrize("kind", ["dqn", "ppo"])
def test_policy_in_the_loop_rollout_matches_oracle(evg, eo, cfg, kind):
    """BASELINE config 3 in small: a torch network of the reference's shape reads the observation tensor the step
    kernel wrote (in place), its outputs are decoded on the device (DQNAgent.filter_actions / PPO's unravel) and fed
    back into the step.  The oracle replays the same decoded rows: the whole chain stays on the reference's rules."""
    import importlib.util, os, torch
    spec = importlib.util.spec_from_file_location("policy_rollout", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                                               "tools", "policy_rollout.py"))
    pr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pr)
    n = 384
    cfg.auto_reset = 2
    cfg.turn_limit = 50
    try:
        env = evg.BatchedEvergladesEnv(n, seed=5, config=cfg, auto_reset=2, env_id_offset=40)
        ora = eo.OracleBatch(cfg, n, seed=5, first=40)
        net = pr.build_policy(kind, torch.float32, env.device, seed=3)
        assert np.array_equal(env.reset().cpu().numpy(), ora.reset().astype(np.float32))
        torch.manual_seed(11)
        moved = 0
        for t in range(120):
            assert env.obs.view(-1, env.obs_len).data_ptr() == env.obs.data_ptr()  # consumed in place
            acts = pr.policy_actions(env, net, kind, torch.float32).cpu().numpy().copy()
            obs, rew, done, info = env.step(acts)
            oobs, orew, odone = ora.step(acts)
            assert np.array_equal(obs.cpu().numpy(), oobs.astype(np.float32)), t
            assert np.array_equal(rew.cpu().numpy(), orew.astype(np.float32)), t
            assert np.array_equal(done.cpu().numpy(), odone), t
            moved += int((obs[:, :, 48::5] != 0).sum())
        assert moved > 0  # the policies do get groups under way
        assert_states_equal(env.get_state(), ora.states, "end")
    finally:
        cfg.auto_reset = 0
        cfg.turn_limit = 150


def test_step_host_chunk_pipeline_equals_step(evg, eo, cfg):
    """From 65,536 matches on evg_step_host runs the batch as four sub-range launches on two streams of its own
    (H2D + kernel of one chunk under the D2H of the previous one).  Same results as the one-shot device-side step of a
    twin simulator, a ragged last chunk included, and oracle parity on a sample of the matches."""
    import torch
    n = 70000 + 37
    a = evg.BatchedEvergladesEnv(n, seed=21, config=cfg, env_id_offset=5)
    b = evg.BatchedEvergladesEnv(n, seed=21, config=cfg, env_id_offset=5)
    a.reset()
    b.reset()
    ids = np.unique(np.concatenate([np.arange(8), np.arange(17490, 17530), np.arange(n - 16, n)]))
    oracles = [eo.OracleBatch(cfg, 1, seed=21, first=5 + int(i)) for i in ids]
    for o in oracles:
        o.reset()
    for t in range(45):
        acts = a.random_actions().clone()
        obs, rew, done, _ = a.step_host(acts.cpu())
        bobs, brew, bdone, _ = b.step(acts)
        torch.cuda.synchronize()
        assert bool((obs == bobs.cpu()).all()) and bool((rew == brew.cpu()).all()) and bool((done == bdone.cpu()).all()), t
        assert bool((a.obs == bobs).all())  # the device-side copies agree too
        host_acts = acts.cpu().numpy()
        for i, o in zip(ids, oracles):
            oobs, orew, odone = o.step(host_acts[i:i + 1])
            assert np.array_equal(obs[i].numpy(), oobs[0].astype(np.float32)), (t, i)
            assert int(done[i]) == int(odone[0]), (t, i)
