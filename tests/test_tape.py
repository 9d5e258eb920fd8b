"""The Philox tape: published known-answer vectors, and C oracle == Python harness."""
import ctypes as C

import numpy as np

from oracle import tape, evg_oracle as eo

# Random123 (Salmon et al., SC'11) kat_vectors for philox4x32-10
KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def _c_philox(ctr, key):
    c = np.array(ctr, dtype=np.uint32)
    k = np.array(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    eo.lib().evo_philox(c.ctypes.data_as(C.c_void_p), k.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p))
    return tuple(int(x) for x in o)


def test_philox_known_answers_python():
    for ctr, key, want in KAT:
        assert tape.philox4x32(ctr, key) == want


def test_philox_known_answers_c():
    for ctr, key, want in KAT:
        assert _c_philox(ctr, key) == want


def test_c_matches_python_random_counters():
    rng = np.random.default_rng(0)
    for _ in range(200):
        ctr = tuple(int(x) for x in rng.integers(0, 2**32, 4))
        key = tuple(int(x) for x in rng.integers(0, 2**32, 2))
        assert _c_philox(ctr, key) == tape.philox4x32(ctr, key)


def test_draw_is_in_range_and_uses_all_lanes():
    seen = set()
    for j in range(12):
        w = tape.combat_word(7, 3, 10, 5, 1, 11, j)
        seen.add(w)
        for n in (1, 2, 13, 100):
            assert 0 <= tape.combat_draw(7, 3, 10, 5, 1, 11, j, n) < n
    assert len(seen) >= 11 and all(0 <= w < 65536 for w in seen)
    assert tape.combat_word(7, 3, 10, 5, 1, 11, 0, episode=1) != tape.combat_word(7, 3, 10, 5, 1, 11, 0)


def test_reward_float32_division_has_no_double_rounding():
    """The kernel computes the non-terminal reward (env.py:58-60, float64 score/3700) as ONE float32 IEEE
    division.  float32(float64(s)/D) == float32(s)/float32(D) for every score s < 2^24: the exact
    quotient can never sit within a float64 half-ulp of a float32 rounding midpoint unless it is on it."""
    s = np.arange(0, 1 << 24, dtype=np.int64)
    for D in (3700, 1, 7, 2999):
        a = (s.astype(np.float64) / float(D)).astype(np.float32)
        b = s.astype(np.float32) / np.float32(D)
        assert np.array_equal(a, b), D
