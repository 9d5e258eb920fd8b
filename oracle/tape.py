"""Random tape shared by the oracle, the reference harness and the CUDA path.

TEST INFRASTRUCTURE ONLY (see oracle/README.md): nothing under
``everglades-ai-wargame_b200/`` may import this module.

The reference server is unseeded: combat draws come from the process-global
``np.random.randint`` (server.py:562).  Parity therefore needs a *tape*: a
pure function that returns the value of every combat draw from its position
in the game.  The position visible at the reference call site (frame locals of
``EvergladesGame.combat``) is (turn, node.ID, pid, gid, j); the tape is

    w   = philox4x32_10(key=(seed_lo, seed_hi),
                        ctr=(env, turn, node | pid<<8 | gid<<16 | (j>>3)<<24,
                             DOMAIN | episode<<8))[(j >> 1) & 3]
    r   = (w >> 16) if (j & 1) else (w & 0xFFFF)        # 8 draws of 16 bits per Philox block
    uid = (r * n) >> 16                      # n = opposing alive units at the node (< 2^16)

Philox4x32-10 is Salmon et al., "Parallel random numbers: as easy as 1, 2, 3"
(SC'11); constants below are the published ones and `philox4x32` is checked
against the Random123 known-answer vectors in tests/test_tape.py.
"""
from __future__ import annotations

M0 = 0xD2511F53
M1 = 0xCD9E8D57
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = 0xFFFFFFFF

DOMAIN_COMBAT = 0          # combat target draws (server.py:562)
DOMAIN_AGENT_RANDOM = 1    # on-device random_actions agent
DOMAIN_AGENT_SWARM = 2     # SwarmAgent's np.random.shuffle (agents/State_Machine/swarm_agent.py:86-87)


def philox4x32(ctr, key, rounds: int = 10):
    """Return the 4 x uint32 Philox block for a 4-word counter and 2-word key."""
    c0, c1, c2, c3 = (int(x) & MASK for x in ctr)
    k0, k1 = (int(x) & MASK for x in key)
    for r in range(rounds):
        if r:
            k0 = (k0 + W0) & MASK
            k1 = (k1 + W1) & MASK
        p0 = M0 * c0
        p1 = M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
    return c0, c1, c2, c3


def combat_word(seed: int, env: int, turn: int, node: int, side: int, gid: int, j: int, episode: int = 0) -> int:
    """The raw 16-bit tape value for the j-th alive attacker of (side, gid) at `node`."""
    ctr = (env & MASK, turn & MASK, (node & 0xFF) | (side & 0xFF) << 8 | (gid & 0xFF) << 16 | ((j >> 3) & 0xFF) << 24,
           DOMAIN_COMBAT | (episode & 0xFFFFFF) << 8)
    key = (seed & MASK, (seed >> 32) & MASK)
    w = philox4x32(ctr, key)[(j >> 1) & 3]
    return (w >> 16) if (j & 1) else (w & 0xFFFF)


def combat_draw(seed: int, env: int, turn: int, node: int, side: int, gid: int, j: int, n: int, episode: int = 0) -> int:
    """Value the patched ``np.random.randint(n)`` returns at server.py:562."""
    return (combat_word(seed, env, turn, node, side, gid, j, episode) * int(n)) >> 16


def block_halves(seed: int, env: int, turn: int, c2: int, domain: int, episode: int = 0):
    """The 8 16-bit tape values of one Philox block: low half then high half of each word."""
    w = philox4x32((env & MASK, turn & MASK, c2 & MASK, domain | (episode & 0xFFFFFF) << 8), (seed & MASK, (seed >> 32) & MASK))
    out = []
    for x in w:
        out += [x & 0xFFFF, x >> 16]
    return out


def swarm_shuffle(lst, seed: int, env: int, turn: int, player: int, episode: int = 0) -> None:
    """In-place stand-in for the legacy ``np.random.shuffle(temp_list)`` of SwarmAgent.get_action
    (numpy's Fisher-Yates: for i = n-1 .. 1: j = uniform{0..i}; swap) with j read from the tape."""
    h = block_halves(seed, env, turn, player, DOMAIN_AGENT_SWARM, episode)
    n = len(lst)
    for k, i in enumerate(range(n - 1, 0, -1)):
        j = (h[k] * (i + 1)) >> 16
        lst[i], lst[j] = lst[j], lst[i]
