"""The fused policy forward (evg_policy_mlp: tcgen05 + TMEM) against torch and against the numpy statement of its own
arithmetic.  Tolerances: the kernel multiplies bf16-rounded operands and accumulates in fp32, so against
evgsim.policy.reference_forward (same roundings, other summation order) the bar is 2e-3 of the largest |Q|; against the
plain fp32 torch network it is the bf16 operand error, 3e-2 of the largest |Q|."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def evg():
    import __graft_entry__ as g
    g.build()
    import evgsim
    return evgsim


def make_net(torch, seed, hidden=528, out=132, scale=1.0):
    torch.manual_seed(seed)
    net = torch.nn.Sequential(torch.nn.Linear(105, hidden), torch.nn.ReLU(), torch.nn.Linear(hidden, out))
    with torch.no_grad():
        for p in net.parameters():
            p.mul_(scale)
    return net


@pytest.mark.parametrize("n,hidden,out", [(64, 528, 132), (1000 + 37, 528, 132), (300, 192, 132), (200, 100, 24), (4096, 528, 132)])
def test_fused_forward_matches_torch(evg, cfg, n, hidden, out):
    import torch
    from evgsim import policy
    env = evg.BatchedEvergladesEnv(n, seed=3, config=cfg)
    env.reset()
    for _ in range(30):  # real mid-game observations (turn counter, control states up to +-500, unit counts)
        env.step(env.random_actions())
    net = make_net(torch, n, hidden, out, scale=3.0)
    fused = policy.FusedDQN(env, net)
    q = fused.forward().cpu().numpy().reshape(2 * n, out)
    assert np.array_equal(fused.forward_t().cpu().numpy(), q.T)  # the transposed output holds the same values
    obs = env.obs.view(-1, 105)
    w = [t.detach().float().numpy() for t in (net[0].weight, net[0].bias, net[2].weight, net[2].bias)]
    want = policy.reference_forward(obs.cpu().numpy(), *w)
    scale = np.abs(want).max()
    assert np.isfinite(q).all()
    assert np.abs(q - want).max() <= 2e-3 * scale, (np.abs(q - want).max(), scale)
    with torch.no_grad():
        ref32 = net.to(env.device)(obs).cpu().numpy()
    assert np.abs(q - ref32).max() <= 3e-2 * np.abs(ref32).max()


def test_fused_policy_in_the_loop_equals_decode_of_its_own_q(evg, cfg):
    """FusedDQN() = evg_decode_dqn(evg_policy_mlp(obs)): the rows it plays are exactly DQNAgent.filter_actions of the
    Q-values it computed (oracle/policy_decode.py restates that function and is pinned to the reference's own)."""
    import torch
    from evgsim import policy
    from oracle import policy_decode
    n = 256
    env = evg.BatchedEvergladesEnv(n, seed=9, config=cfg, auto_reset=1)
    env.reset()
    fused = policy.FusedDQN(env, make_net(torch, 1))
    for t in range(40):
        acts = fused()
        q = fused.qt.cpu().numpy().T.reshape(n, 2, 132)
        a = acts.cpu().numpy()
        for i in range(0, n, 17):
            for p in range(2):
                assert np.array_equal(a[i, p], policy_decode.dqn_filter_actions(q[i, p]).astype(np.int8)), (t, i, p)
        env.step(acts)
    assert env.episode_stats()["env_turns"] == 40 * n
