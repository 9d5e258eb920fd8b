"""BASELINE.json's full size (1,048,576 lock-step matches, random_actions self-play with in-place auto-reset):
size-independent properties over ALL matches, plus exact oracle checks on samples taken from the big batch."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = 1 << 20
SEED = 2026


@pytest.fixture(scope="module")
def big():
    import torch
    import __graft_entry__ as g
    g.build()
    import evgsim
    from oracle import evg_oracle as eo
    cfg = evgsim.load_config(auto_reset=1)
    env = evgsim.BatchedEvergladesEnv(N, seed=SEED, config=cfg, auto_reset=1)
    env.reset()
    return {"torch": torch, "evg": evgsim, "eo": eo, "cfg": cfg, "env": env}


def test_invariants_and_sampled_oracle_parity_over_full_episodes(big):
    torch, evg, eo, cfg, env = big["torch"], big["evg"], big["eo"], big["cfg"], big["env"]
    # matches followed exactly on the CPU oracle: first, last, and a scattered sample of global ids
    rng = np.random.default_rng(1)
    ids = np.unique(np.concatenate([np.arange(64), np.arange(N - 64, N), rng.integers(0, N, 384)]))
    oracles = [eo.OracleBatch(cfg, 1, seed=SEED, first=int(i)) for i in ids]
    for o in oracles:
        o.reset()
    idt = torch.as_tensor(ids, device=env.device)
    prev_alive = None
    dones = 0
    for t in range(170):  # crosses the turn-150 auto-reset of every match
        obs, rew, done, info = env.step_agents(want_actions=(t % 10 == 0))
        # ---- properties over all 1,048,576 matches, evaluated on the device
        p0, p1 = obs[:, 0], obs[:, 1]
        alive0, alive1 = p0[:, 49::5], p1[:, 49::5]                       # alive units per own group
        assert bool((p0[:, 4:45:4].sum(1) == alive1.sum(1)).all())        # units p0 sees on the board == p1's own units
        assert bool((p1[:, 4:45:4].sum(1) == alive0.sum(1)).all())
        assert bool((p0[:, 0] == p1[:, 0]).all())                         # same turn counter for both players
        turn = p0[:, 0]
        assert bool(((turn >= 1) & (turn <= 150)).all())
        assert bool((done == (info["status"] != 0)).all())
        assert bool(((turn == 150) <= (done == 1)).all())                 # time limit ends the match
        cs = p0[:, 3:45:4]
        assert bool((cs[:, 1:10].abs() <= 100).all()) and bool((cs[:, [0, 10]].abs() <= 500).all())
        perm = torch.as_tensor([cfg.p1_node_map[k + 1] - 1 for k in range(11)], device=env.device)
        assert bool((p1[:, 3:45:4] == p0[:, 3:45:4][:, perm]).all())      # p1 sees the board through server.py:89's map, raw sign
        sc = info["scores"].to(torch.float64)
        not_done = done == 0
        assert bool((rew[not_done] == (sc[not_done] / 3700.0).to(torch.float32)).all())  # env.py:58-60: float32(float64 quotient)
        rd = rew[done == 1]
        assert bool(((rd[:, 0] == 0) | (rd[:, 0] == 1)).all()) and bool(((rd[:, 1] == 0) | (rd[:, 1].abs() == 1)).all())
        alive = torch.cat([alive0, alive1], 1)
        if prev_alive is not None:
            same_match = prev_done == 0
            assert bool((alive[same_match] <= prev_alive[same_match]).all())  # units only ever die
        prev_alive, prev_done = alive.clone(), done.clone()
        dones += int(done.sum())
        # ---- exact parity on the sampled matches
        so, sr, sd = obs[idt].cpu().numpy(), rew[idt].cpu().numpy(), done[idt].cpu().numpy()
        for k, o in enumerate(oracles):
            st = o.states[0]
            rows = np.stack([eo.agent_random(cfg, SEED, int(ids[k]), int(st["episode"]), int(st["turn"]) + 1, p) for p in range(2)])
            oo, orr, od = o.step(rows[None].astype(np.int8))
            assert np.array_equal(so[k], oo[0].astype(np.float32)), (t, int(ids[k]))
            assert np.array_equal(sr[k], orr[0].astype(np.float32)) and sd[k] == od[0]
    stats = env.episode_stats()
    assert stats["episodes"] == dones >= N
    assert stats["wins"][0] + stats["wins"][1] + stats["ties"] == dones
    assert stats["env_turns"] == N * 170
    # full resident state of the sampled matches (every field, fp64 health) after 170 turns
    for k, o in enumerate(oracles[::16]):
        g = env.get_state(int(ids[16 * k]), 1)[0]
        for name in ("turn", "episode", "control_state", "controlled_by", "health"):
            assert np.array_equal(g[name], o.states[0][name]), (name, int(ids[16 * k]))


def test_trajectories_do_not_depend_on_batch_or_shard(big):
    """The tape is keyed on the GLOBAL match id: a 4096-match shard at offset 524288 replays the same matches."""
    evg, cfg = big["evg"], big["cfg"]
    full = evg.BatchedEvergladesEnv(1 << 16, seed=5, config=cfg, auto_reset=1, env_id_offset=(1 << 19) - 1000)
    part = evg.BatchedEvergladesEnv(4096, seed=5, config=cfg, auto_reset=1, env_id_offset=1 << 19)
    full.reset()
    part.reset()
    for t in range(160):
        fo, fr, fd, _ = full.step_agents()
        po, pr, pd, _ = part.step_agents()
        assert bool((fo[1000:1000 + 4096] == po).all()) and bool((fr[1000:1000 + 4096] == pr).all())
        assert bool((fd[1000:1000 + 4096] == pd).all())


def test_wire_rows_equal_float32_observations_at_full_size(big):
    """1,048,576 matches stepped twice — float32 vectors and packed wire rows — from the same action stream: at turns
    early, mid-game and across the turn-150 reset EVERY row expands to exactly the float32 observation, reward and done
    flag of its match (evgsim.wire.expand, in slices of 131,072 rows)."""
    torch, evg, cfg = big["torch"], big["evg"], big["cfg"]
    from evgsim import wire
    a = evg.BatchedEvergladesEnv(N, seed=77, config=cfg, auto_reset=1)
    b = evg.BatchedEvergladesEnv(N, seed=77, config=cfg, auto_reset=1)
    a.reset()
    b.reset(obs_format="wire")
    checked = 0
    for t in range(1, 153):
        acts = a.random_actions()
        obs, rew, done, info = a.step(acts)
        rows, _, _, _ = b.step(acts, obs_format="wire")
        if t in (1, 2, 40, 75, 110, 150, 151, 152):
            for lo in range(0, N, 1 << 17):
                o, r, d, s = wire.expand(rows[lo:lo + (1 << 17)].cpu().numpy(), cfg)
                assert np.array_equal(o, obs[lo:lo + (1 << 17)].cpu().numpy()), (t, lo)
                assert np.array_equal(r, rew[lo:lo + (1 << 17)].cpu().numpy()) and np.array_equal(d, done[lo:lo + (1 << 17)].cpu().numpy()), (t, lo)
                assert np.array_equal(s, info["status"][lo:lo + (1 << 17)].cpu().numpy()), (t, lo)
            checked += 1
    assert checked == 8 and a.episode_stats() == b.episode_stats()
