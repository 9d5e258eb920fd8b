"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/evgsim.h declares,
validates configs, and FAILS LOUDLY without a GPU (there is no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import evgsim
from evgsim import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    return _capi.load()


def test_header_and_binding_list_the_same_symbols(lib):
    hdr = open(os.path.join(ROOT, "include", "evgsim.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(evg_[a-z_]+)\s*\(", hdr))
    bound = {name for name, _, _ in _capi.SYMBOLS}
    assert declared == bound, (declared ^ bound)
    for name in declared:
        assert hasattr(lib, name)


def test_struct_sizes_match_the_oracle_build(lib):
    from oracle import evg_oracle as eo
    L = eo.lib()
    assert L.evo_sizeof_config() == C.sizeof(_capi.EvgConfig)
    assert L.evo_sizeof_state() == C.sizeof(_capi.EvgEnvState) == _capi.env_state_dtype().itemsize


def test_default_config_equals_json_config(lib):
    d = _capi.EvgConfig()
    assert lib.evg_default_config(C.byref(d)) == 0
    j = evgsim.load_config()
    assert bytes(d) == bytes(j)


def _create(lib, cfg, n=4):
    h = C.c_void_p()
    rc = lib.evg_create(C.byref(cfg), n, 0, 0, 0, C.byref(h))
    return rc, h, (lib.evg_last_error() or b"").decode()


def test_config_validation_reports_reason(lib):
    cfg = evgsim.load_config()
    cfg.n_nodes = 40
    rc, _, msg = _create(lib, cfg)
    assert rc == -2 and "n_nodes" in msg
    cfg = evgsim.load_config()
    cfg.p1_node_map[2] = 3
    rc, _, msg = _create(lib, cfg)
    assert rc == -2 and "involution" in msg
    cfg = evgsim.load_config()
    cfg.group_size[0][3] = 17
    rc, _, msg = _create(lib, cfg)
    assert rc == -2 and "units" in msg
    cfg = evgsim.load_config()
    cfg.node_team_start[11] = -1
    rc, _, msg = _create(lib, cfg)
    assert rc == -2 and "TeamStart" in msg
    cfg = evgsim.load_config()
    rc, _, msg = _create(lib, cfg, n=0)
    assert rc == -1


def test_no_gpu_means_error_not_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    rc, h, msg = _create(lib, evgsim.load_config())
    assert rc == -3 and "no CPU path" in msg and not h.value
    with pytest.raises(RuntimeError):
        evgsim.BatchedEvergladesEnv(4)


def test_null_handle_calls_fail(lib):
    assert lib.evg_step(None, None, None, None, None, None, None, None) == -1
    assert lib.evg_reset(None, None, None, None) == -1
    assert lib.evg_destroy(None) == -1
    assert lib.evg_launch_count(None) == -1
    assert lib.evg_step_kernel_kind(None) == -1


def test_json_loader_semantics(tmp_path):
    import json
    cfg = evgsim.load_config()
    assert cfg.n_nodes == 11 and cfg.turn_limit == 150 and cfg.capture_bonus == 1000
    assert [cfg.group_size[0][g] for g in range(12)] == [8] * 11 + [12]
    assert [cfg.group_type[1][g] for g in range(12)] == [1, 2, 0] * 4  # controller, striker, tank (env.py:21)
    assert cfg.edge_distance[3][6] == 3 and cfg.edge_distance[6][3] == 3 and cfg.edge_distance[1][3] == 0
    assert cfg.node_has_defense[4] == 1 and cfg.node_has_defend[4] == 0  # 'DEFENSE' != 'DEFEND' (server.py:595)
    # GameSetup.json values are honoured (the reference hard-codes them)
    setup = tmp_path / "Setup.json"
    setup.write_text(json.dumps({"TurnLimit": 40, "CaptureBonus": 500, "UnitBudget": 60}))
    cfg2 = evgsim.load_config(setup_file=str(setup))
    assert cfg2.turn_limit == 40 and cfg2.capture_bonus == 500
    assert [cfg2.group_size[0][g] for g in range(12)] == [5] * 12
    with pytest.raises(ValueError):
        evgsim.load_config(num_units=400)
    with pytest.raises(FileNotFoundError):
        evgsim.load_config(map_file="NoSuchMap.json")
