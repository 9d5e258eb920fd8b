"""evgsim — B200-native batched Everglades turn step (drop-in for EvergladesEnv.reset/step).

Public names mirror the reference's gym wrapper (gym_everglades/envs/everglades_env.py):
``EvergladesEnv`` (one match, dict in / dict out) and ``BatchedEvergladesEnv`` (N matches in
lockstep on one GPU).  The game step itself lives in csrc/ (hand-written sm_100a CUDA behind the
C ABI of include/evgsim.h); there is no CPU implementation in this package.
"""
from . import _capi
from .config import load_config, DEFAULT_CONFIG_DIR

__all__ = ["_capi", "load_config", "DEFAULT_CONFIG_DIR", "EvergladesEnv", "BatchedEvergladesEnv", "register", "wire"]

# gym_everglades/__init__.py:3-6: importing the package registers 'everglades-v0' (only when gym/gymnasium is installed)
try:
    from .spaces import register as _register_env
    _register_env()
except Exception:  # never let an unusual gym install break the import of the simulator
    pass


def __getattr__(name):  # torch is imported only when an env class is first used
    if name in ("EvergladesEnv", "BatchedEvergladesEnv", "MAX_SCORE"):
        from . import env as _env
        return getattr(_env, name)
    if name in ("wire", "hostmem", "spaces", "policy"):
        import importlib
        return importlib.import_module("." + name, __name__)
    if name == "register":
        from .spaces import register as _register
        return _register
    if name in ("shard_range", "gather_episode_stats"):
        from . import dist as _dist
        return getattr(_dist, name)
    raise AttributeError(name)
