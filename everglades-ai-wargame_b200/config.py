"""JSON config -> EvgConfig, with the reference's loading semantics.

Mirrors ``EvergladesGame.board_init`` (server.py:40-100), ``unitTypes_init`` (server.py:103-131)
and the fixed loadout of ``EvergladesEnv._build_groups`` (env.py:145-156).  ``GameSetup.json`` is
declared by the reference but never opened (the server hard-codes 150 turns at server.py:321 and
the 1000 bonus at server.py:304; env.py:18 hard-codes 100 units); here it IS read, with those
same defaults, as BASELINE.json's north_star asks.
"""
from __future__ import annotations

import json
import os

from . import _capi

DEFAULT_CONFIG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config")
UNIT_CLASSES = ["controller", "striker", "tank"]  # env.py:21
DEMO_P1_NODE_MAP = [0, 11, 8, 9, 10, 5, 6, 7, 2, 3, 4, 1]  # server.py:89 (hard-coded for DemoMap)
MAX_SCORE = 3700  # env.py:11


def _resolve(config_dir, name):
    """server.py:24-27: the file is opened as given; config_dir only gates the existence test."""
    if name is None:
        raise ValueError("config file name is None")
    if os.path.exists(name):
        return name
    if config_dir is not None and os.path.exists(os.path.join(config_dir, name)):
        return os.path.join(config_dir, name)
    raise FileNotFoundError("config file %r not found (config_dir=%r)" % (name, config_dir))


def load_config(config_dir=None, map_file="DemoMap.json", unit_file="UnitDefinitions.json",
                setup_file="GameSetup.json", auto_reset=_capi.AUTORESET_OFF, num_units=None,
                turn_limit=None, capture_bonus=None) -> _capi.EvgConfig:
    config_dir = config_dir or DEFAULT_CONFIG_DIR
    cfg = _capi.EvgConfig()
    cfg.abi_version = _capi.ABI_VERSION
    cfg.auto_reset = int(auto_reset)
    cfg.max_score = MAX_SCORE

    # ---- GameSetup.json (optional file; reference defaults)
    setup = {}
    if setup_file is not None:
        try:
            with open(_resolve(config_dir, setup_file)) as fid:
                setup = json.load(fid)
        except FileNotFoundError:
            setup = {}
    cfg.turn_limit = int(turn_limit if turn_limit is not None else setup.get("TurnLimit", 150))
    cfg.capture_bonus = int(capture_bonus if capture_bonus is not None else setup.get("CaptureBonus", 1000))
    budget = int(num_units if num_units is not None else setup.get("UnitBudget", 100))

    # ---- map (server.py:45-84)
    with open(_resolve(config_dir, map_file)) as fid:
        map_dat = json.load(fid)
    nodes = map_dat["nodes"]
    n = len(nodes)
    if not 2 <= n <= _capi.MAX_NODES:
        raise ValueError("map has %d nodes; supported: 2..%d" % (n, _capi.MAX_NODES))
    ids = sorted(int(nd["ID"]) for nd in nodes)
    if ids != list(range(1, n + 1)):
        raise ValueError("node IDs must be exactly 1..%d, got %s" % (n, ids))
    cfg.n_nodes = n
    for i in range(_capi.MAX_NODES + 1):
        cfg.node_team_start[i] = -1
    for nd in nodes:
        i = int(nd["ID"])
        cfg.node_control_points[i] = int(nd["ControlPoints"])
        cfg.node_defense[i] = float(nd["StructureDefense"])
        cfg.node_team_start[i] = int(nd["TeamStart"])
        res = nd["Resource"]  # list membership is exact-string (server.py:442-443,595)
        cfg.node_has_defense[i] = 1 if "DEFENSE" in res else 0
        cfg.node_has_observe[i] = 1 if "OBSERVE" in res else 0
        cfg.node_has_defend[i] = 1 if "DEFEND" in res else 0
        for conn in nd["Connections"]:
            d, dist = int(conn["ConnectedID"]), int(conn["Distance"])
            if not 1 <= d <= n:
                raise ValueError("node %d connects to unknown node %d" % (i, d))
            if not 1 <= dist <= 255:
                raise ValueError("edge %d-%d distance %d outside 1..255" % (i, d, dist))
            if cfg.edge_distance[i][d] == 0:  # first matching connection wins (server.py:246-250)
                cfg.edge_distance[i][d] = dist
    if abs(cfg.node_control_points[1]) > 32767 or any(cfg.node_control_points[i] > 32767 for i in range(n + 1)):
        raise ValueError("ControlPoints above 32767 not supported")
    for p in (0, 1):
        if sum(1 for nd in nodes if int(nd["TeamStart"]) == p) < 1:
            raise ValueError("map has no TeamStart for player %d" % p)
    p1map = map_dat.get("P1NodeMap")
    if p1map is None:
        if n != 11:
            raise ValueError("maps other than DemoMap need a 'P1NodeMap' (the reference hard-codes it, server.py:89)")
        p1map = DEMO_P1_NODE_MAP
    if len(p1map) != n + 1 or p1map[0] != 0 or sorted(p1map[1:]) != list(range(1, n + 1)) \
            or any(p1map[p1map[i]] != i for i in range(n + 1)):
        raise ValueError("P1NodeMap must be an involution of 1..%d with map[0] = 0" % n)
    for i, v in enumerate(p1map):
        cfg.p1_node_map[i] = int(v)

    # ---- unit types (server.py:108-130): type id = position in the file, names lower-cased
    with open(_resolve(config_dir, unit_file)) as fid:
        unit_dat = json.load(fid)
    units = unit_dat["units"]
    if not 1 <= len(units) <= _capi.MAX_UNIT_TYPES:
        raise ValueError("1..%d unit types supported" % _capi.MAX_UNIT_TYPES)
    cfg.n_unit_types = len(units)
    names = {}
    for t, u in enumerate(units):
        names[u["Name"].lower()] = t
        cfg.unit_armor[t] = float(u["Health"])
        cfg.unit_damage[t] = int(u["Damage"])
        cfg.unit_speed[t] = int(u["Speed"])
        cfg.unit_control[t] = int(u["Control"])
        cfg.unit_cost[t] = int(u["Cost"])
        if not (0 <= cfg.unit_damage[t] <= 255 and 0 <= cfg.unit_speed[t] <= 255 and
                0 <= cfg.unit_control[t] <= 63 and 0 <= cfg.unit_cost[t] <= 255 and cfg.unit_armor[t] > 0):
            raise ValueError("unit type %r outside the supported value ranges" % u["Name"])

    # ---- loadout (env.py:145-156): classes cycle, equal split, remainder to the last group
    per = budget // _capi.NUM_GROUPS
    sizes = [per] * (_capi.NUM_GROUPS - 1)
    sizes.append(budget - sum(sizes))
    for g, size in enumerate(sizes):
        cls = UNIT_CLASSES[g % len(UNIT_CLASSES)]
        if cls not in names:  # server.py:164
            raise ValueError("Group type %r not in unit type config file" % cls)
        if not 1 <= size <= _capi.MAX_GROUP_UNITS:
            raise ValueError("group %d would hold %d units; supported 1..%d (UnitBudget=%d)"
                             % (g, size, _capi.MAX_GROUP_UNITS, budget))
        for p in (0, 1):
            cfg.group_type[p][g] = names[cls]
            cfg.group_size[p][g] = size
    return cfg


def unit_type_names(unit_file="UnitDefinitions.json", config_dir=None):
    with open(_resolve(config_dir or DEFAULT_CONFIG_DIR, unit_file)) as fid:
        return [u["Name"].lower() for u in json.load(fid)["units"]]
