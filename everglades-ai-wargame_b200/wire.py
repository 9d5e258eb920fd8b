"""The packed observation row (EVG_OBS_WIRE, include/evgsim.h) <-> the reference's observation vector, in numpy.

A wire row carries one match's step result once — turn, done, status, per node (controlState, units of player 0,
units of player 1), per group (location | moving, avg health, units alive), the two rewards — in
``row_bytes(n_nodes)`` bytes (128 on DemoMap) instead of float32[2][obs_len] (840).  What the reference's vector
(board_state server.py:382-455 + player_state :457-501, concatenated at env.py:158-171) holds besides is static
and comes from the EvgConfig: the nodes' DEFENSE/OBSERVE flags, the groups' unit types and player 1's node
numbering.  ``expand`` rebuilds the float32 vector exactly; ``pack`` is its inverse (host-side checks, tests).
Layout conversion only: no game rule is evaluated here.
"""
from __future__ import annotations

import numpy as np

from . import _capi

NG = _capi.NUM_GROUPS


def row_bytes(n_nodes: int) -> int:
    return (_capi.WIRE_NODE0 + 4 * n_nodes + 3 * 2 * NG + 8 + 15) // 16 * 16


def _static(cfg):
    n = int(cfg.n_nodes)
    p1 = np.array(list(cfg.p1_node_map)[:n + 1], dtype=np.int64)
    flags = np.stack([np.array(list(cfg.node_has_defense)[:n + 1]), np.array(list(cfg.node_has_observe)[:n + 1])], -1).astype(np.float32)
    types = np.array([[cfg.group_type[p][g] for g in range(NG)] for p in range(2)], dtype=np.float32)
    return n, p1, flags, types


def expand(rows, cfg):
    """rows: uint8 [N, row_bytes] -> (obs float32 [N,2,obs_len], reward float32 [N,2], done uint8 [N], status uint8 [N])."""
    n, p1, flags, types = _static(cfg)
    rows = np.ascontiguousarray(rows, dtype=np.uint8)
    N = rows.shape[0]
    assert rows.shape[1] == row_bytes(n), (rows.shape, row_bytes(n))
    L = 1 + 4 * n + 5 * NG
    turn = rows[:, 0:2].copy().view("<u2")[:, 0]
    done, status = rows[:, 2].copy(), rows[:, 3].copy()
    node = rows[:, 4:4 + 4 * n].reshape(N, n, 4)
    cs = node[:, :, 0:2].copy().view("<i2")[:, :, 0]          # [N, n], node id x at column x-1
    units = node[:, :, 2:4]                                    # [N, n, player]
    g0 = 4 + 4 * n
    grp = rows[:, g0:g0 + 3 * 2 * NG].reshape(N, 2, NG, 3)
    reward = rows[:, g0 + 72:g0 + 80].copy().view("<f4").reshape(N, 2)
    obs = np.zeros((N, 2, L), dtype=np.float32)
    obs[:, :, 0] = turn[:, None]
    for p in range(2):
        real = p1[1:n + 1] if p else np.arange(1, n + 1)        # viewer slot k shows real node real[k] (server.py:437-439)
        b = obs[:, p, 1:1 + 4 * n].reshape(N, n, 4)
        b[:, :, 0] = flags[real, 0]
        b[:, :, 1] = flags[real, 1]
        b[:, :, 2] = cs[:, real - 1]                            # raw sign for both viewers
        b[:, :, 3] = units[:, real - 1, 1 - p]                  # the OTHER player's listed units
        g = obs[:, p, 1 + 4 * n:].reshape(N, NG, 5)
        loc = (grp[:, p, :, 0] & 63).astype(np.int64)
        g[:, :, 0] = p1[loc] if p else loc
        g[:, :, 1] = types[p]
        g[:, :, 2] = grp[:, p, :, 1]
        g[:, :, 3] = grp[:, p, :, 0] >> 6
        g[:, :, 4] = grp[:, p, :, 2]
    return obs, reward, done, status


def pack(obs, reward, done, status, cfg):
    """Inverse of expand: the wire rows that carry these observations (uint8 [N, row_bytes])."""
    n, p1, _, _ = _static(cfg)
    obs = np.asarray(obs)
    N = obs.shape[0]
    rows = np.zeros((N, row_bytes(n)), dtype=np.uint8)
    rows[:, 0:2] = obs[:, 0, 0].astype("<u2").reshape(N, 1).view(np.uint8)
    rows[:, 2] = np.asarray(done, dtype=np.uint8)
    rows[:, 3] = np.asarray(status, dtype=np.uint8)
    node = np.zeros((N, n, 4), dtype=np.uint8)
    b0 = obs[:, 0, 1:1 + 4 * n].reshape(N, n, 4)               # player 0's slots are the real nodes
    b1 = obs[:, 1, 1:1 + 4 * n].reshape(N, n, 4)
    node[:, :, 0:2] = b0[:, :, 2].astype("<i2").reshape(N, n, 1).view(np.uint8)
    node[:, :, 3] = b0[:, :, 3].astype(np.uint8)               # player 1's units, as player 0 sees them
    node[:, :, 2] = b1[:, p1[1:n + 1] - 1, 3].astype(np.uint8)  # player 0's units at real node x: player 1's slot p1[x]-1
    rows[:, 4:4 + 4 * n] = node.reshape(N, 4 * n)
    g0 = 4 + 4 * n
    grp = np.zeros((N, 2, NG, 3), dtype=np.uint8)
    for p in range(2):
        g = obs[:, p, 1 + 4 * n:].reshape(N, NG, 5)
        own = g[:, :, 0].astype(np.int64)
        loc = p1[own] if p else own                             # the map is an involution
        grp[:, p, :, 0] = (loc | (g[:, :, 3].astype(np.int64) << 6)).astype(np.uint8)
        grp[:, p, :, 1] = g[:, :, 2].astype(np.uint8)
        grp[:, p, :, 2] = g[:, :, 4].astype(np.uint8)
    rows[:, g0:g0 + 72] = grp.reshape(N, 72)
    rows[:, g0 + 72:g0 + 80] = np.ascontiguousarray(reward, dtype="<f4").reshape(N, 2).view(np.uint8)
    return rows
