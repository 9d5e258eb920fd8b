"""CPU restatement of the reference agents' action decoding (policy-in-the-loop glue).
TEST INFRASTRUCTURE ONLY.

dqn_filter_actions restates ``DQNAgent.filter_actions`` (agents/DQN/DQNAgent.py:161-197): NOT a plain top-7 —
a greedy insertion over (node, group) in node-major order into 7 slots whose best-Q start at 0 (so only positive
Q-values are ever chosen) and whose group ids start at 0 (so group 0 "occupies" every slot until displaced); a
group already placed in another slot can only improve its own slot.  The emitted node is the 0-based column
index (the reference's off-by-one: node 11 can never be chosen, column 0 is a no-op).
ppo_unravel restates ``PPOAgent.get_action`` (agents/PPO/PPOAgent.py:122-127): units = idx // 12, nodes = idx % 11.
Both are pinned against the reference's own code by tests/golden/policy_decode_v1.npz (gen_golden_policy.py).
"""
import numpy as np


def dqn_filter_actions(q, num_groups=12, num_nodes=11, n_actions=7):
    """q: float array [num_groups * num_nodes] (row-major group x node). Returns int array [n_actions, 2]."""
    q = np.asarray(q, dtype=np.float32).reshape(num_groups, num_nodes)
    units = np.zeros(n_actions)
    nodes = np.zeros(n_actions)
    best = np.zeros(n_actions)
    for n in range(num_nodes):
        for g in range(num_groups):
            for s in range(n_actions):
                if q[g, n] > best[s]:
                    if g in units and units[s] != g:
                        continue
                    best[s], units[s], nodes[s] = q[g, n], g, n
                    break
    return np.stack([units, nodes], axis=1).astype(np.int64)


def ppo_unravel(idx, div=12, mod=11):
    idx = np.asarray(idx).astype(np.int64)
    return np.stack([idx // div, idx % mod], axis=-1)


def shape_reward(mode, player, reward, done, turn_num):
    """utils/reward_shaping.py:17-56 restated: mode 0 normalized_score, 1 basic_reward, 2 penalize_long_games,
    3 reward_short_games."""
    mine, other = reward[player], reward[1 - player]
    if mode == 1:
        return 1.0 if (done and mine > other) else 0.0
    if mode == 2:
        return (100.0 if mine > other else -0.1) if done else -0.001
    if mode == 3:
        return ((150.0 - turn_num) / 150.0 if mine > other else -1.0) if done else 0.0
    return mine
