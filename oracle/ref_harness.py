"""Drive the UNMODIFIED reference (server.py + everglades_env.py) under the random tape, and time it.

TEST INFRASTRUCTURE ONLY.  The reference is read from ``/root/reference`` where that exists (the
build container) and otherwise from ``oracle/_ref/``, where ``oracle/stage_ref.py`` staged the same
files byte for byte so that they reach the GPU box.  Used by ``tests/golden/gen_golden.py`` to
produce the committed golden trajectories, by ``tests/test_oracle_vs_reference.py`` (skipped when
neither copy is present) and by bench.py's ``cpu_baseline`` / ``--impl reference`` legs
(``time_reference``: the real Python server on the box's host cores).

What is imported from the reference, untouched:
  * ``everglades_server.server.EvergladesGame``   (server.py:11)
  * ``gym_everglades.envs.everglades_env.EvergladesEnv`` (env.py:13) under stub
    ``gym`` modules (gym is not installed here; the stubs only provide the names
    env.py:1-6 imports — no arithmetic lives in gym).
Only shims: ``numpy.int = int`` (server.py:55,79,430,471 use the removed alias)
and the tape patch of ``numpy.random.randint`` for the duration of a game.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

from . import tape

def _pick_root() -> str:
    env = os.environ.get("EVG_REFERENCE_ROOT")
    if env:
        return env
    staged = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
    for root in ("/root/reference", staged):
        if os.path.isfile(os.path.join(root, "everglades-server", "everglades_server", "server.py")):
            return root
    return "/root/reference"


REFERENCE_ROOT = _pick_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "everglades-server", "everglades_server", "server.py"))


def _install_gym_stub() -> None:
    if "gym" in sys.modules:
        return

    class _Space:
        def __init__(self, *a, **k):
            self.args, self.kwargs = a, k

    class Box(_Space):
        def __init__(self, low=None, high=None, **k):
            self.low, self.high = np.asarray(low), np.asarray(high)
            self.shape = self.low.shape

    class Discrete(_Space):
        def __init__(self, n):
            self.n = n

    class Tuple(_Space):
        def __init__(self, spaces):
            self.spaces = tuple(spaces)

    gym = types.ModuleType("gym")
    gym.Env = type("Env", (), {})
    spaces = types.ModuleType("gym.spaces")
    spaces.Box, spaces.Discrete, spaces.Tuple = Box, Discrete, Tuple
    utils = types.ModuleType("gym.utils")
    seeding = types.ModuleType("gym.utils.seeding")
    utils.seeding = seeding
    error = types.ModuleType("gym.error")
    envs = types.ModuleType("gym.envs")
    cc = types.ModuleType("gym.envs.classic_control")
    rendering = types.ModuleType("gym.envs.classic_control.rendering")
    cc.rendering = rendering
    envs.classic_control = cc
    registration = types.ModuleType("gym.envs.registration")
    registration.register = lambda **k: None
    envs.registration = registration
    gym.spaces, gym.utils, gym.error, gym.envs = spaces, utils, error, envs
    for name, mod in [("gym", gym), ("gym.spaces", spaces), ("gym.utils", utils), ("gym.utils.seeding", seeding),
                      ("gym.error", error), ("gym.envs", envs), ("gym.envs.classic_control", cc),
                      ("gym.envs.classic_control.rendering", rendering), ("gym.envs.registration", registration)]:
        sys.modules[name] = mod


def import_reference():
    """Return (server_module, EvergladesEnv class) of the unmodified reference."""
    if not reference_available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    if not hasattr(np, "int"):
        np.int = int  # noqa: the only shim the server needs on numpy >= 1.24
    for sub in ("everglades-server", "gym-everglades"):
        p = os.path.join(REFERENCE_ROOT, sub)
        if p not in sys.path:
            sys.path.insert(0, p)
    _install_gym_stub()
    import everglades_server.server as server
    from gym_everglades.envs.everglades_env import EvergladesEnv
    return server, EvergladesEnv


class TapePatch:
    """Context manager: ``numpy.random.randint`` reads the Philox tape inside ``combat``.

    Calls from ``game_init`` (server.py:205) and ``game_end`` (server.py:338) only set the
    unobservable ``focus`` field and get 0.  The draw position is read from the caller's
    frame (locals of server.py:549-562: ``self, node, pid, gid, j``).
    """

    def __init__(self, seed: int, env_id: int):
        self.seed, self.env_id = int(seed), int(env_id)
        self.n_draws = 0
        self.log = None  # optional list of (turn,node,pid,gid,j,n,uid)

    def _randint(self, n, *a, **k):
        f = sys._getframe(1)
        if f.f_code.co_name != "combat":
            return 0
        loc = f.f_locals
        turn = int(loc["self"].current_turn)
        node = int(loc["node"].ID)
        pid, gid, j = int(loc["pid"]), int(loc["gid"]), int(loc["j"])
        uid = tape.combat_draw(self.seed, self.env_id, turn, node, pid, gid, j, int(n))
        self.n_draws += 1
        if self.log is not None:
            self.log.append((turn, node, pid, gid, j, int(n), uid))
        return uid

    def __enter__(self):
        self._orig = np.random.randint
        np.random.randint = self._randint
        return self

    def __exit__(self, *exc):
        np.random.randint = self._orig
        return False


CONFIG_DIR = os.path.join(REFERENCE_ROOT, "config") + os.sep


def snapshot(game) -> dict:
    """Canonical integer/fp64 state of a live reference game (the fields DESIGN.md lists)."""
    n_nodes = len(game.evgMap.nodes)
    grp = np.zeros((2, 12, 7), dtype=np.int32)   # loc, dest, dist, ready, moving, destroyed, count
    rank = np.full((2, 12), -1, dtype=np.int32)  # position in the node list of its location (-1: unlisted)
    health = np.zeros((2, 100), dtype=np.float64)
    for p in (0, 1):
        off = 0
        for g, group in enumerate(game.players[p].groups):
            u = group.units[0]
            grp[p, g] = (group.location, group.travel_destination, group.distance_remaining, int(group.ready),
                         int(group.moving), int(group.destroyed), u.count)
            health[p, off:off + len(u.unitHealth)] = u.unitHealth
            off += len(u.unitHealth)
    node = np.zeros((n_nodes, 2), dtype=np.int32)  # controlState, controlledBy
    for i, nd in enumerate(game.evgMap.nodes):
        node[i] = (nd.controlState, nd.controlledBy)
        for p in (0, 1):
            for pos, g in enumerate(nd.groups[p]):
                rank[p, g] = pos
    return {"turn": int(game.current_turn), "grp": grp, "rank": rank, "node": node, "health": health}


def run_reference_game(seed: int, env_id: int, actions: np.ndarray, map_file: str = "DemoMap.json",
                       unit_file: str = "UnitDefinitions.json", stop_at_done: bool = True, keep_log: bool = False,
                       n_turns: int | None = None):
    """Play one match on the unmodified reference through ``EvergladesEnv.reset/step``.

    ``actions``: array [T, 2, rows, 2] (group, node) in each player's own numbering, or a
    callable ``policy(t, obs_dict) -> array [2, rows, 2]`` for closed-loop play (needs ``n_turns``);
    the rows actually played are returned under "actions".
    Returns a dict of per-turn arrays (index 0 = after reset).
    """
    _, EvergladesEnv = import_reference()
    import gym_everglades.envs.everglades_env as envmod
    envmod.EvergladesRenderer = lambda game: None  # the viewer is not on the step path (pyglet absent)
    env = EvergladesEnv()
    players = {0: None, 1: None}
    T = n_turns if callable(actions) else actions.shape[0]
    obs_l, rew_l, done_l, snaps, played = [], [], [], [], []
    with TapePatch(seed, env_id) as tp:
        if keep_log:
            tp.log = []
        obs = env.reset(players=players, config_dir=CONFIG_DIR, map_file=CONFIG_DIR + map_file,
                        unit_file=CONFIG_DIR + unit_file, output_dir="/tmp/", pnames={0: "a", 1: "b"}, debug=False)
        obs_l.append(np.stack([obs[0], obs[1]]))
        snaps.append(snapshot(env.game))
        for t in range(T):
            a_t = np.asarray(actions(t, obs) if callable(actions) else actions[t])
            played.append(a_t)
            act = {0: np.array(a_t[0]), 1: np.array(a_t[1])}
            obs, reward, done, _ = env.step(act)
            obs_l.append(np.stack([obs[0], obs[1]]))
            rew_l.append([float(reward[0]), float(reward[1])])
            done_l.append(int(done))
            snaps.append(snapshot(env.game))
            if done and stop_at_done:
                break
        n_draws, log = tp.n_draws, tp.log
    out = {
        "actions": np.stack(played),
        "obs": np.stack(obs_l),                                   # [T+1, 2, 105] float64 (integer valued)
        "reward": np.array(rew_l, dtype=np.float64).reshape(-1, 2),
        "done": np.array(done_l, dtype=np.int8),
        "grp": np.stack([s["grp"] for s in snaps]),
        "rank": np.stack([s["rank"] for s in snaps]),
        "node": np.stack([s["node"] for s in snaps]),
        "health": np.stack([s["health"] for s in snaps]),
        "n_draws": n_draws,
    }
    if keep_log:
        out["log"] = log
    return out


# ------------------------------------------------------------------------------------------------
# Timing the unmodified reference (bench.py cpu_baseline kind "reference"; SURVEY.md 8d "CPU baseline")
# ------------------------------------------------------------------------------------------------
def _time_worker(args):
    """One process: random_actions-vs-random_actions matches through the reference's own EvergladesEnv.reset/step
    (game_turn + both players' board_state/player_state, server.py:211-279,382-501) for about `seconds`; the action
    arrays are pre-drawn, the server draws its combat targets from its own global numpy stream (nothing patched).
    Returns (env_turns, seconds spent inside reset/step)."""
    import io
    import time
    import contextlib

    seconds, seed = args
    _, EvergladesEnv = import_reference()
    import gym_everglades.envs.everglades_env as envmod
    envmod.EvergladesRenderer = lambda game: None  # the viewer is not on the step path (pyglet absent)
    rng = np.random.default_rng(seed)
    pool = np.zeros((256, 2, 7, 2), dtype=np.int64)
    for k in range(256):
        for p in range(2):  # random_actions.py:38-46: 7 distinct groups, 7 distinct nodes
            pool[k, p, :, 0] = rng.permutation(12)[:7]
            pool[k, p, :, 1] = rng.permutation(np.arange(1, 12))[:7]
    np.random.seed(seed)
    env = EvergladesEnv()
    kw = dict(players={0: None, 1: None}, config_dir=CONFIG_DIR, map_file=CONFIG_DIR + "DemoMap.json",
              unit_file=CONFIG_DIR + "UnitDefinitions.json", output_dir="/tmp/", pnames={0: "a", 1: "b"}, debug=False)
    turns, spent, k = 0, 0.0, 0
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        while spent < seconds:
            t0 = time.perf_counter()
            env.reset(**kw)
            done = 0
            while not done:
                a = pool[k & 255]
                k += 1
                _, _, done, _ = env.step({0: a[0], 1: a[1]})
                turns += 1
            spent += time.perf_counter() - t0
    return turns, spent


def time_reference(seconds: float = 10.0, procs: int | None = None):
    """env-turns/s of the unmodified Python reference on `procs` host processes (default: one per CPU), each playing
    whole matches for about `seconds`.  Returns (rate_all, procs, rate_one_process, sample description)."""
    import multiprocessing as mp
    import time

    procs = procs or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        res = pool.map(_time_worker, [(seconds, 1000 + i) for i in range(procs)])
    wall = time.perf_counter() - t0
    turns = sum(r[0] for r in res)
    rate_all = sum(r[0] / r[1] for r in res)       # processes run side by side: aggregate of per-process rates
    rate_one = res[0][0] / res[0][1]
    sample = ("%d processes x ~%.0f s of whole random_actions-vs-random_actions matches through the unmodified "
              "EvergladesEnv.reset/step (%d env-turns in total, %.1f s wall incl. process start)" % (procs, seconds, turns, wall))
    return rate_all, procs, rate_one, sample
