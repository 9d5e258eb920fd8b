import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden", "demomap_v1.npz")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """Without a CUDA device the gpu-marked tests are skipped (the product has no CPU path to run them on)."""
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """tests/golden/demomap_v1.npz: trajectories of the unmodified reference (gen_golden.py)."""

    def __init__(self):
        z = np.load(GOLDEN)
        self.seed = int(z["seed"])
        self.names = [str(x) for x in z["names"]]
        self.games = []
        for i in range(len(self.names)):
            self.games.append({k: z["g%d_%s" % (i, k)] for k in
                               ("actions", "obs", "reward", "done", "grp", "rank", "node", "health")})

    def __len__(self):
        return len(self.games)


@pytest.fixture(scope="session")
def golden():
    return Golden()


@pytest.fixture(scope="session")
def cfg():
    import evgsim
    return evgsim.load_config()


def state_fields(rec):
    """(grp[2,12,7], node[n,2]) integer views of an EvgEnvState record, golden-fixture order."""
    g = rec["groups"]
    grp = np.stack([g["location"], g["travel_destination"], g["distance_remaining"], g["ready"], g["moving"],
                    g["destroyed"], g["count"]], -1).astype(np.int64)
    return grp


def flat_health(rec, cfg):
    return np.concatenate([rec["health"][:, g, :cfg.group_size[0][g]] for g in range(12)], axis=1)
