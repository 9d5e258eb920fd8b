"""Generate tests/golden/policy_decode_v1.npz from the reference's own decode code (build container only):
``DQNAgent.filter_actions`` (agents/DQN/DQNAgent.py:161-197) called unbound on random Q-vectors, and the
index unravel of ``PPOAgent.get_action`` (agents/PPO/PPOAgent.py:122-127, restated: two integer ops)."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("EVG_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "policy_decode_v1.npz")


def load_dqn_agent():
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from oracle import ref_harness
    ref_harness._install_gym_stub()  # QNetwork.py imports gym at module level; nothing of it is used here
    # site-packages ships an unrelated `agents` package that shadows the reference's namespace package:
    # register the reference's directories explicitly
    for name, sub in (("agents", "agents"), ("agents.DQN", "agents/DQN")):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, sub)]
        sys.modules[name] = m
    # sibling modules pull in gym / matplotlib (absent here); filter_actions uses none of their classes
    for name, attr in (("agents.DQN.NoisyLinear", "NoisyLinear"), ("agents.DQN.PrioritizedMemory", "PrioritizedMemory"),
                       ("agents.DQN.SimpleMemory", "ReplayMemory"), ("agents.DQN.QNetwork", "QNetwork")):
        m = types.ModuleType(name)
        setattr(m, attr, type(attr, (), {}))
        sys.modules[name] = m
    spec = importlib.util.spec_from_file_location("agents.DQN.DQNAgent", os.path.join(REF, "agents/DQN/DQNAgent.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["agents.DQN.DQNAgent"] = mod
    spec.loader.exec_module(mod)
    return mod.DQNAgent


def main():
    DQNAgent = load_dqn_agent()
    fake = types.SimpleNamespace(num_groups=12, num_nodes=11, n_actions=7)
    rng = np.random.default_rng(42)
    qs, acts = [], []
    for i in range(400):
        kind = i % 5
        if kind == 0:
            q = rng.normal(size=132)
        elif kind == 1:
            q = rng.normal(size=132) - 2.0          # mostly negative: few or no slots filled
        elif kind == 2:
            q = np.round(rng.normal(size=132) * 2) / 2  # many exact ties
        elif kind == 3:
            q = rng.random(132) * (rng.random(132) < 0.1)  # sparse positives
        else:
            q = rng.normal(size=132) * 10
        q = q.astype(np.float32)
        a = np.zeros((7, 2))
        out = DQNAgent.filter_actions(fake, a, torch.from_numpy(q.copy()))
        qs.append(q)
        acts.append(np.asarray(out).astype(np.int8))
    idx = rng.integers(0, 132, size=(400, 7))
    np.savez_compressed(OUT, q=np.stack(qs), actions=np.stack(acts), ppo_idx=idx.astype(np.int16),
                        ppo_actions=np.stack([idx // 12, idx % 11], -1).astype(np.int8))
    print("wrote", OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
