"""Policy-in-the-loop glue: the reference agents' action decoding, pinned to the reference's own
code by tests/golden/policy_decode_v1.npz (DQNAgent.filter_actions called unbound on 400 Q-vectors)."""
import os

import numpy as np
import pytest

from oracle import policy_decode as pd

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "policy_decode_v1.npz")


def test_numpy_restatement_matches_reference_filter_actions():
    z = np.load(GOLD)
    for q, a in zip(z["q"], z["actions"]):
        assert np.array_equal(pd.dqn_filter_actions(q), a)
    assert np.array_equal(pd.ppo_unravel(z["ppo_idx"]), z["ppo_actions"])


@pytest.mark.gpu
def test_device_decode_matches_reference():
    import torch
    import __graft_entry__ as g
    g.build()
    import evgsim
    z = np.load(GOLD)
    n = 200
    env = evgsim.BatchedEvergladesEnv(n)
    env.reset()
    q = torch.from_numpy(z["q"][:2 * n].reshape(n, 2, 132))
    rows = env.decode_dqn(q).cpu().numpy()
    assert np.array_equal(rows, z["actions"][:2 * n].reshape(n, 2, 7, 2))
    rows1 = env.decode_dqn(torch.from_numpy(z["q"][2 * n - n:2 * n]), player=1).cpu().numpy()
    assert np.array_equal(rows1[:, 1], z["actions"][n:2 * n]) and np.array_equal(rows1[:, 0], rows[:, 0])
    idx = torch.from_numpy(z["ppo_idx"][:2 * n].astype(np.int64).reshape(n, 2, 7))
    assert np.array_equal(env.decode_indices(idx).cpu().numpy(), z["ppo_actions"][:2 * n].reshape(n, 2, 7, 2))
    # policy in the loop: a DQN-shaped MLP (105 -> 528 -> 132, agents/DQN/QNetwork.py:37,42) on the obs tensor, both players
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(105, 528), torch.nn.ReLU(), torch.nn.Linear(528, 132)).to(env.device)
    obs = env.obs
    for t in range(5):
        with torch.no_grad():
            qv = net(obs.reshape(-1, 105)).reshape(n, 2, 132)
        a = env.decode_dqn(qv)
        want = np.stack([pd.dqn_filter_actions(v) for v in qv.cpu().numpy().reshape(-1, 132)]).reshape(n, 2, 7, 2)
        assert np.array_equal(a.cpu().numpy(), want)
        obs, _, _, _ = env.step(a)
