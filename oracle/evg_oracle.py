"""ctypes wrapper of oracle/libevg_oracle.so (the C restatement in evg_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs.  The struct layouts come from the product's public
header mirror (evgsim._capi); the product never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

import evgsim
from evgsim import _capi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    path = os.path.join(_HERE, "libevg_oracle.so")
    src = os.path.join(_HERE, "evg_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "evgsim.h")
    if force or not os.path.isfile(path) or os.path.getmtime(path) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libevg_oracle.so"])
    return path


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        P = C.c_void_p
        L.evo_reset.argtypes = [C.POINTER(_capi.EvgConfig), P, C.c_int32]
        L.evo_reset.restype = None
        L.evo_step.argtypes = [C.POINTER(_capi.EvgConfig), P, C.c_uint64, C.c_uint64, P, C.c_int, P, P, P, P]
        L.evo_step.restype = C.c_int
        L.evo_observe.argtypes = [C.POINTER(_capi.EvgConfig), P, P]
        L.evo_observe.restype = None
        L.evo_obs_len.argtypes = [C.POINTER(_capi.EvgConfig)]
        L.evo_obs_len.restype = C.c_int
        L.evo_np_pairwise_sum.argtypes = [P, C.c_int]
        L.evo_np_pairwise_sum.restype = C.c_double
        L.evo_agent_random.argtypes = [C.POINTER(_capi.EvgConfig), C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_int, P]
        L.evo_agent_random.restype = None
        L.evo_run_random.argtypes = [C.POINTER(_capi.EvgConfig), C.c_uint64, C.c_int64, C.c_int64, C.c_int,
                                     C.POINTER(C.c_double), C.POINTER(C.c_int64)]
        L.evo_run_random.restype = C.c_int64
        L.evo_step_batch.argtypes = [C.POINTER(_capi.EvgConfig), P, C.c_int64, C.c_uint64, C.c_int64, P, P, P, P, P, P]
        L.evo_step_batch.restype = None
        L.evo_agent_base_rush.argtypes = [C.POINTER(_capi.EvgConfig), P, P, C.c_int, P]
        L.evo_agent_base_rush.restype = None
        L.evo_agent_swarm.argtypes = [C.POINTER(_capi.EvgConfig), P, P, C.c_uint64, C.c_uint64, C.c_int, P]
        L.evo_agent_swarm.restype = None
        L.evo_philox.argtypes = [P, P, P]
        L.evo_philox.restype = None
        L.evo_fought_slots.argtypes = [C.c_int]
        L.evo_fought_slots.restype = C.c_int64
        assert L.evo_sizeof_config() == C.sizeof(_capi.EvgConfig)
        assert L.evo_sizeof_state() == C.sizeof(_capi.EvgEnvState)
        _LIB = L
    return _LIB


class OracleEnv:
    """One match on the CPU oracle; state is a 1-element numpy record of EvgEnvState layout."""

    def __init__(self, cfg=None, seed: int = 0, env_id: int = 0):
        self.cfg = cfg if cfg is not None else evgsim.load_config()
        self.seed, self.env_id = int(seed), int(env_id)
        self.state = np.zeros(1, dtype=_capi.env_state_dtype())
        self.obs_len = lib().evo_obs_len(C.byref(self.cfg))
        self.reset()

    def _sp(self):
        return self.state.ctypes.data_as(C.c_void_p)

    def reset(self, episode: int = 0):
        lib().evo_reset(C.byref(self.cfg), self._sp(), int(episode))
        return self.observe()

    def observe(self):
        obs = np.zeros((2, self.obs_len), dtype=np.float64)
        lib().evo_observe(C.byref(self.cfg), self._sp(), obs.ctypes.data_as(C.c_void_p))
        return obs

    def step(self, actions):
        """actions: int array [2, rows, 2]. Returns obs[2,L], reward[2], done, scores[2], status."""
        a = np.ascontiguousarray(np.asarray(actions).astype(np.int32))
        assert a.ndim == 3 and a.shape[0] == 2 and a.shape[2] == 2
        obs = np.zeros((2, self.obs_len), dtype=np.float64)
        reward = np.zeros(2, dtype=np.float64)
        scores = np.zeros(2, dtype=np.int64)
        status = C.c_int(0)
        done = lib().evo_step(C.byref(self.cfg), self._sp(), self.seed, self.env_id, a.ctypes.data_as(C.c_void_p),
                              a.shape[1], obs.ctypes.data_as(C.c_void_p), reward.ctypes.data_as(C.c_void_p),
                              scores.ctypes.data_as(C.c_void_p), C.byref(status))
        return obs, reward, int(done), scores, status.value


class OracleBatch:
    """n matches with global ids first..first+n-1 stepped in lock-step on the CPU oracle."""

    def __init__(self, cfg, n, seed=0, first=0):
        self.cfg, self.n, self.seed, self.first = cfg, int(n), int(seed), int(first)
        self.obs_len = lib().evo_obs_len(C.byref(cfg))
        self.states = np.zeros(self.n, dtype=_capi.env_state_dtype())
        self.obs = np.zeros((self.n, 2, self.obs_len), dtype=np.float64)
        self.reward = np.zeros((self.n, 2), dtype=np.float64)
        self.done = np.zeros(self.n, dtype=np.uint8)
        self.scores = np.zeros((self.n, 2), dtype=np.int32)
        self.status = np.zeros(self.n, dtype=np.uint8)

    def _p(self, a):
        return a.ctypes.data_as(C.c_void_p)

    def reset(self, mask=None):
        st = self.states
        for i in range(self.n):
            if mask is None or mask[i]:
                ep = 0 if mask is None else int(st[i]["episode"]) + 1
                lib().evo_reset(C.byref(self.cfg), C.c_void_p(st.ctypes.data + i * st.itemsize), ep)
                lib().evo_observe(C.byref(self.cfg), C.c_void_p(st.ctypes.data + i * st.itemsize),
                                  C.c_void_p(self.obs.ctypes.data + i * self.obs[0].nbytes))
        return self.obs

    def step(self, actions):
        a = np.ascontiguousarray(np.asarray(actions, dtype=np.int8))
        assert a.shape == (self.n, 2, 7, 2)
        lib().evo_step_batch(C.byref(self.cfg), self._p(self.states), self.n, self.seed, self.first, self._p(a),
                             self._p(self.obs), self._p(self.reward), self._p(self.done), self._p(self.scores),
                             self._p(self.status))
        return self.obs, self.reward, self.done


def agent_random(cfg, seed, env_id, episode, turn, player):
    rows = np.zeros((7, 2), dtype=np.int32)
    lib().evo_agent_random(C.byref(cfg), int(seed), int(env_id), int(episode), int(turn), int(player),
                           rows.ctypes.data_as(C.c_void_p))
    return rows


class ScriptedAgents:
    """Per-match state of the observation-driven scripted agents (one uint32 per (match, player, kind))."""

    def __init__(self, n):
        self.state = np.zeros((n, 2, 2), dtype=np.uint32)  # [..., 0] base_rush, [..., 1] swarm

    def rows(self, kind, cfg, states, seed, first, player):
        """kind: 'base_rush' | 'swarm'. Returns int8 [n, 7, 2] rows for `player`, advancing the agent state."""
        n = len(states)
        out = np.zeros((n, 7, 2), dtype=np.int32)
        for i in range(n):
            sp = C.c_void_p(states.ctypes.data + i * states.itemsize)
            if kind == "base_rush":
                stp = C.c_void_p(self.state.ctypes.data + (i * 4 + player * 2) * 4)
                lib().evo_agent_base_rush(C.byref(cfg), sp, stp, player, C.c_void_p(out.ctypes.data + i * 56))
            else:
                stp = C.c_void_p(self.state.ctypes.data + (i * 4 + player * 2 + 1) * 4)
                lib().evo_agent_swarm(C.byref(cfg), sp, stp, int(seed), int(first + i), player, C.c_void_p(out.ctypes.data + i * 56))
        return out.astype(np.int8)


def run_random(cfg, seed, first, count, n_turns):
    chk, eps = C.c_double(0), C.c_int64(0)
    n = lib().evo_run_random(C.byref(cfg), int(seed), int(first), int(count), int(n_turns), C.byref(chk), C.byref(eps))
    return int(n), chk.value, eps.value


def fought_slots(clear=False):
    """Unit slots of the groups that fought in this thread's evo_step calls so far (checks the device counter)."""
    return int(lib().evo_fought_slots(1 if clear else 0))


def list_rank(state_rec):
    """Position of every group in the (derived) node list of its location, -1 if unlisted.
    Comparable with ref_harness.snapshot()['rank']."""
    rank = np.full((2, 12), -1, dtype=np.int32)
    g = state_rec["groups"]
    for p in range(2):
        for a in range(12):
            if g[p, a]["destroyed"]:
                continue
            r = 0
            for b in range(12):
                if b != a and not g[p, b]["destroyed"] and g[p, b]["location"] == g[p, a]["location"] and \
                        (g[p, b]["arrival"], b) < (g[p, a]["arrival"], a):
                    r += 1
            rank[p, a] = r
    return rank
