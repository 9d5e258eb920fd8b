"""Generate tests/golden/demomap_v1.npz from the UNMODIFIED reference (build container only).

    python tests/golden/gen_golden.py

Every game is played through the reference's own ``EvergladesEnv.reset/step`` (env.py:32-116 ->
server.py) under the Philox tape (oracle/ref_harness.py).  Game i uses tape seed SEED and match id
i, so one lock-step batch of len(games) matches replays the whole fixture.
Scenarios are chosen to exercise what uniform-random play rarely reaches: long marches, big
battles, group destruction, node flips, base capture.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402

SEED = 20261018
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "demomap_v1.npz")

# adjacency in a player's OWN numbering (DemoMap is symmetric under server.py:89's map)
ADJ = {1: [2, 4], 2: [1, 3, 5], 3: [2, 4, 5, 6, 7], 4: [1, 3, 7], 5: [2, 3, 8, 9], 6: [3, 9],
       7: [3, 4, 9, 10], 8: [5, 9, 11], 9: [5, 6, 7, 8, 10], 10: [7, 9, 11], 11: [8, 10]}
NEXT_TOP = {1: 2, 2: 5, 5: 8, 8: 11, 3: 5, 4: 3, 6: 9, 7: 9, 9: 8, 10: 11}
NEXT_BOT = {1: 4, 4: 7, 7: 10, 10: 11, 3: 7, 2: 3, 6: 9, 5: 9, 9: 10, 8: 11}


def loc(obs, g):
    return int(obs[45 + 5 * g])


def uniform_random(rng):
    """random_actions.get_action, agents/State_Machine/random_actions.py:38-46"""
    def pol(t, obs):
        a = np.zeros((2, 7, 2), dtype=np.int64)
        for p in range(2):
            a[p, :, 0] = rng.permutation(12)[:7]
            a[p, :, 1] = rng.permutation(np.arange(1, 12))[:7]
        return a
    return pol


def adjacent_random(rng, p_move=1.0):
    def pol(t, obs):
        a = np.zeros((2, 7, 2), dtype=np.int64)
        for p in range(2):
            gs = rng.permutation(12)[:7]
            for r, g in enumerate(gs):
                nb = ADJ[loc(obs[p], g)]
                a[p, r] = (g, nb[rng.integers(len(nb))] if rng.random() < p_move else 0)
        return a
    return pol


def rush(rng, routes=(NEXT_TOP, NEXT_BOT), jitter=0.0):
    """every group marches on the enemy base; groups alternate between the two flanks"""
    def pol(t, obs):
        a = np.zeros((2, 7, 2), dtype=np.int64)
        for p in range(2):
            gs = [g for g in rng.permutation(12) if obs[p][45 + 5 * g + 3] == 0 and loc(obs[p], g) != 11][:7]
            for r, g in enumerate(gs):
                route = routes[(g + p) % len(routes)]
                nxt = route.get(loc(obs[p], g), 0)
                if jitter and rng.random() < jitter:
                    nb = ADJ[loc(obs[p], g)]
                    nxt = nb[rng.integers(len(nb))]
                a[p, r] = (g, nxt)
        return a
    return pol


TO_CENTER_A = {1: 2, 2: 3, 3: 6, 4: 3, 5: 3, 7: 3, 9: 6, 8: 9, 10: 9, 11: 10}
TO_CENTER_B = {1: 4, 4: 3, 3: 6, 2: 3, 5: 9, 7: 9, 9: 6, 8: 9, 10: 9, 11: 8}


def brawl(rng, target_routes=(TO_CENTER_A, TO_CENTER_B)):
    """everybody converges on the centre node and stays: large multi-group battles, many deaths"""
    def pol(t, obs):
        a = np.zeros((2, 7, 2), dtype=np.int64)
        for p in range(2):
            gs = [g for g in rng.permutation(12) if obs[p][45 + 5 * g + 3] == 0 and loc(obs[p], g) != 6][:7]
            for r, g in enumerate(gs):
                a[p, r] = (g, target_routes[(g + p) % 2].get(loc(obs[p], g), 0))
        return a
    return pol


def mixed(pol0, pol1):
    def pol(t, obs):
        a0, a1 = pol0(t, obs), pol1(t, obs)
        return np.stack([a0[0], a1[1]])
    return pol


def idle_vs(pol1):
    def pol(t, obs):
        a = pol1(t, obs)
        a[0] = 0
        return a
    return pol


def quirks(rng):
    """rows the reference accepts without raising: nid 0, non-adjacent, duplicates, p0 nid > 11,
    more than 7 rows (only the first 7 count, server.py:227)"""
    def pol(t, obs):
        a = np.zeros((2, 9, 2), dtype=np.int64)
        for p in range(2):
            for r in range(9):
                g = int(rng.integers(12))
                nb = ADJ[loc(obs[p], g)]
                k = rng.integers(6)
                n = [0, nb[rng.integers(len(nb))], int(rng.integers(1, 12)), nb[0], 11, 12 if p == 0 else 5][k]
                a[p, r] = (g, n)
            if rng.random() < 0.5:
                a[p, 1, 0] = a[p, 0, 0]  # duplicate group id: the first VALID row wins
        return a
    return pol


def scenarios():
    rng = np.random.default_rng(SEED)
    S = []
    for _ in range(6):
        S.append(("uniform_random", uniform_random(rng)))
    for _ in range(6):
        S.append(("adjacent_random", adjacent_random(rng)))
    for _ in range(3):
        S.append(("adjacent_random_sparse", adjacent_random(rng, 0.4)))
    for _ in range(4):
        S.append(("rush_both", rush(rng)))
    for _ in range(3):
        S.append(("rush_jitter", rush(rng, jitter=0.25)))
    S.append(("rush_top_only", rush(rng, routes=(NEXT_TOP,))))
    S.append(("rush_vs_random", mixed(rush(rng), adjacent_random(rng))))
    S.append(("random_vs_rush", mixed(adjacent_random(rng), rush(rng))))
    S.append(("idle_vs_rush", idle_vs(rush(rng))))
    S.append(("idle_vs_rush_top", idle_vs(rush(rng, routes=(NEXT_TOP,)))))
    for _ in range(4):
        S.append(("quirks", quirks(rng)))
    for _ in range(3):
        S.append(("center_brawl", brawl(rng)))
    S.append(("brawl_vs_rush", mixed(brawl(rng), rush(rng))))
    S.append(("rush_vs_uniform", mixed(rush(rng, jitter=0.1), uniform_random(rng))))
    S.append(("uniform_vs_rush", mixed(uniform_random(rng), rush(rng, jitter=0.1))))
    return S


def main():
    out = {"seed": np.int64(SEED)}
    names = []
    for i, (name, pol) in enumerate(scenarios()):
        g = rh.run_reference_game(SEED, i, pol, n_turns=170 if name.startswith("idle") else 150, stop_at_done=True)
        T = len(g["done"])
        names.append(name)
        assert np.all(g["obs"] == np.round(g["obs"])) and np.abs(g["obs"]).max() < 32767
        assert g["actions"].min() >= 0 and g["actions"].max() <= 12
        out["g%d_actions" % i] = g["actions"].astype(np.int8)
        out["g%d_obs" % i] = g["obs"].astype(np.int16)
        out["g%d_reward" % i] = g["reward"]
        out["g%d_done" % i] = g["done"]
        out["g%d_grp" % i] = g["grp"].astype(np.int16)
        out["g%d_rank" % i] = g["rank"].astype(np.int8)
        out["g%d_node" % i] = g["node"].astype(np.int16)
        out["g%d_health" % i] = g["health"]
        alive = g["grp"][-1, :, :, 6].sum(axis=1)
        destroyed = g["grp"][-1, :, :, 5].sum(axis=1)
        print("game %2d %-24s turns %3d draws %5d done %d reward %s alive %s destroyed %s" %
              (i, name, T, g["n_draws"], g["done"][-1], g["reward"][-1], alive, destroyed))
    out["names"] = np.array(names)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
