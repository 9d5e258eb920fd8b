"""Host side of the fused policy forward (evgsim.policy): bf16 rounding, the swizzled weight images and the numpy
statement of what the kernel computes.  The kernel itself is checked in tests/test_gpu_policy.py."""
import numpy as np

from evgsim import policy


def test_bf16_rounding_is_round_to_nearest_even():
    x = np.array([1.0, 1.0 + 2 ** -8, 1.0 + 3 * 2 ** -9, -2.5, 3.140625, 1e-3, 65504.0], dtype=np.float32)
    import torch
    want = torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
    assert np.array_equal(policy.bf16_round(x), want)
    r = np.random.default_rng(0).standard_normal(10000).astype(np.float32) * 37
    assert np.array_equal(policy.bf16_round(r), torch.from_numpy(r).to(torch.bfloat16).to(torch.float32).numpy())


def test_swizzle_is_a_bijection_inside_every_block():
    for rows, cols in ((128, 128), (192, 128), (144, 192), (128, 192)):
        r, k = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
        off = policy.swz_offset(rows, r, k)
        assert off.min() == 0 and off.max() == rows * cols * 2 - 2 and len(np.unique(off)) == rows * cols
        # 8 consecutive columns stay one 16-byte chunk (the kernel stores hidden activations 16 bytes at a time)
        assert ((off[:, ::8] % 16) == 0).all() and (np.diff(off.reshape(rows, cols // 8, 8), axis=2) == 2).all()


def test_pack_mlp_places_every_weight_where_the_header_says():
    rng = np.random.default_rng(1)
    w1 = rng.standard_normal((528, 105)).astype(np.float32)
    w2 = rng.standard_normal((132, 528)).astype(np.float32)
    b1 = rng.standard_normal(528).astype(np.float32)
    b2 = rng.standard_normal(132).astype(np.float32)
    img1, b1p, img2, b2p, hidden, out = policy.pack_mlp(w1, b1, w2, b2)
    assert (hidden, out) == (528, 132) and img1.nbytes == 3 * 49152 and img2.nbytes == 3 * 55296 and b1p.shape == (576,)
    assert np.array_equal(b1p[:528], b1) and not b1p[528:].any()
    i1, i2 = img1.view(np.uint16), img2.view(np.uint16)
    for (h, k) in ((0, 0), (191, 104), (192, 7), (527, 64), (300, 63)):
        c, n = divmod(h, 192)
        assert i1[(c * 49152 + policy.swz_offset(192, n, k)) // 2] == policy.to_bf16_bits(w1[h, k])
    for (o, h) in ((0, 0), (131, 527), (77, 191), (5, 192), (100, 400)):
        c, k = divmod(h, 192)
        assert i2[(c * 55296 + policy.swz_offset(144, o, k)) // 2] == policy.to_bf16_bits(w2[o, h])
    # padding (hidden 528..575, outputs 132..143, inputs 105..127) is zero
    assert i1[(2 * 49152 + policy.swz_offset(192, 150, 3)) // 2] == 0 and i2[policy.swz_offset(144, 140, 9) // 2] == 0
    assert i1[policy.swz_offset(192, 3, 120) // 2] == 0
