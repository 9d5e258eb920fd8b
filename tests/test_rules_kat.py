"""Known answers of the rules (observed on the live reference) replayed on the CPU
oracle, and — with the same hand-built states imported through evg_import_state — on the CUDA path."""
import numpy as np
import pytest

from oracle import evg_oracle as eo

NOOP = np.zeros((1, 2, 7, 2), dtype=np.int8)


def fresh(cfg, seed=0, first=0):
    o = eo.OracleBatch(cfg, 1, seed=seed, first=first)
    o.reset()
    return o


def teleport(o, player, node, arrival0=1000):
    g = o.states[0]["groups"]
    for k in range(12):
        g[player, k]["location"] = node
        g[player, k]["arrival"] = arrival0 + 16 * k + k


def leave_units(o, player, keep):
    """keep: {gid: [health of the surviving units...]}; every other unit of the player dies."""
    st = o.states[0]
    for g in range(12):
        st["health"][player, g, :] = 0.0
        alive = keep.get(g, [])
        st["health"][player, g, :len(alive)] = alive
        st["groups"][player, g]["count"] = len(alive)
        st["groups"][player, g]["destroyed"] = 0 if alive else 1


SCENARIOS = {}


def scenario(fn):
    SCENARIOS[fn.__name__] = fn
    return fn


@scenario
def p1_capture_from_neutral_is_delayed_one_turn(cfg):
    """A.7 #5: 0 -> negative counts as a sign flip, so player 1's full capture shows controlledBy -1 for a turn."""
    o = fresh(cfg)
    teleport(o, 1, 10)
    checks = [lambda o: (o.states[0]["control_state"][10], o.states[0]["controlled_by"][10], o.scores[0].tolist()) == (-100, -1, [1100, 1300]),
              lambda o: o.states[0]["controlled_by"][10] == 1]
    return o, [NOOP, NOOP], checks


@scenario
def p0_capture_from_neutral_is_immediate(cfg):
    o = fresh(cfg)
    teleport(o, 0, 2)
    checks = [lambda o: (o.states[0]["control_state"][2], o.states[0]["controlled_by"][2]) == (100, 0)]
    return o, [NOOP], checks


@scenario
def annihilation_needs_both_players_at_zero(cfg):
    """A.7 #7/#8: one striker each with a sliver of health, same node: both die -> status 3."""
    o = fresh(cfg)
    teleport(o, 0, 6)
    teleport(o, 1, 6)
    leave_units(o, 0, {1: [1e-9]})
    leave_units(o, 1, {4: [1e-9]})
    checks = [lambda o: (int(o.status[0]), int(o.done[0]), int(o.states[0]["groups"]["destroyed"].sum())) == (3, 1, 24)]
    return o, [NOOP], checks


@scenario
def turn_limit_has_priority_over_annihilation(cfg):
    o = fresh(cfg)
    teleport(o, 0, 6)
    teleport(o, 1, 6)
    leave_units(o, 0, {1: [1e-9]})
    leave_units(o, 1, {4: [1e-9]})
    o.states[0]["turn"] = 149
    checks = [lambda o: int(o.status[0]) == 1]
    return o, [NOOP], checks


@scenario
def a_wiped_out_player_just_loses_slowly(cfg):
    o = fresh(cfg)
    leave_units(o, 1, {})
    checks = [lambda o: (int(o.status[0]), o.scores[0].tolist()) == (0, [1100, 1000])]
    return o, [NOOP], checks


@scenario
def base_capture_ends_the_match_with_bonus(cfg):
    """server.py:299-304,327: an enemy-held base gives +1000 and status 2; reward is 1 / -1 (env.py:41-44)."""
    o = fresh(cfg)
    teleport(o, 0, 11)
    leave_units(o, 1, {})
    o.states[0]["control_state"][11] = 400
    o.states[0]["controlled_by"][11] = -1
    checks = [lambda o: (int(o.status[0]), int(o.states[0]["controlled_by"][11]), o.reward[0].tolist()) == (2, 0, [1.0, -1.0])
              and int(o.scores[0, 0]) == 1000 + 1000 + 1000 + 100]
    return o, [NOOP], checks


@scenario
def tank_survives_thirty_single_hits_in_fp64(cfg):
    """A.6: after 30 hits of damage 1 an armour-3 tank on an undefended node holds ~2.75e-14 health: alive, avg health 0."""
    o = fresh(cfg)
    h = 100.0
    for _ in range(30):
        h = h - (10.0 * 1) / (3.0 + 0.0)
    assert 0 < h < 1e-12
    leave_units(o, 0, {2: [h]})
    checks = [lambda o: (o.obs[0, 0, 45 + 5 * 2 + 2], o.obs[0, 0, 45 + 5 * 2 + 4]) == (0.0, 1.0)]
    return o, [NOOP], checks


@scenario
def a_commanded_group_still_fights_and_a_destroyed_one_keeps_its_orders(cfg):
    """A.2 / A.7 #9: the group ordered away this turn is `ready`, not `moving`, so it fights; it dies ready and stays so."""
    o = fresh(cfg)
    teleport(o, 0, 6)
    teleport(o, 1, 6)
    leave_units(o, 0, {1: [1e-9]})
    leave_units(o, 1, {4: [50.0] * 8, 7: [50.0] * 8})
    acts = NOOP.copy()
    acts[0, 0, 0] = (1, 3)  # p0 group 1: 6 -> 3
    g = lambda o: o.states[0]["groups"][0, 1]
    checks = [lambda o: (int(g(o)["destroyed"]), int(g(o)["ready"]), int(g(o)["travel_destination"]), int(g(o)["location"])) == (1, 1, 3, 6)]
    return o, [acts], checks


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_rule_known_answers_on_the_oracle(cfg, name):
    o, steps, checks = SCENARIOS[name](cfg)
    for a, chk in zip(steps, checks):
        o.step(a)
        assert chk(o), name


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_rule_known_answers_on_the_gpu(cfg, name):
    import __graft_entry__ as g
    g.build()
    import evgsim
    o, steps, checks = SCENARIOS[name](cfg)
    env = evgsim.BatchedEvergladesEnv(1, seed=o.seed, config=cfg, env_id_offset=o.first)
    env.reset()
    env.set_state(o.states)
    for a in steps:
        obs, rew, done, info = env.step(a)
        oobs, orew, odone = o.step(a)
        assert np.array_equal(obs.cpu().numpy(), oobs.astype(np.float32)), name
        assert np.array_equal(rew.cpu().numpy(), orew.astype(np.float32)) and int(done[0]) == int(odone[0])
        assert int(info["status"][0]) == int(o.status[0]) and info["scores"].cpu().numpy().tolist() == o.scores.tolist()
    st = env.get_state()[0]
    for f in ("turn", "control_state", "controlled_by", "health"):
        assert np.array_equal(st[f], o.states[0][f]), (name, f)
    for f in ("location", "travel_destination", "distance_remaining", "ready", "moving", "destroyed", "count", "avg_health"):
        assert np.array_equal(st["groups"][f], o.states[0]["groups"][f]), (name, f)
