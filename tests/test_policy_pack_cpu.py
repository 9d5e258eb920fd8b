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
    img1, img2, hidden, out = policy.pack_mlp(w1, b1, w2, b2)
    assert (hidden, out) == (528, 132) and img1.nbytes == 3 * 49152 and img2.nbytes == 3 * 55296
    i1, i2 = img1.view(np.uint16), img2.view(np.uint16)
    # the biases ride inside: input feature 105 is the constant 1 (weights b1), hidden unit 528 is relu(1) = 1 (weights b2)
    for h in (0, 191, 192, 527):
        c, n = divmod(h, 192)
        assert i1[(c * 49152 + policy.swz_offset(192, n, 105)) // 2] == policy.to_bf16_bits(b1[h])
    assert i1[(2 * 49152 + policy.swz_offset(192, 528 - 384, 105)) // 2] == policy.to_bf16_bits(np.float32(1.0))
    assert not any(i1[(2 * 49152 + policy.swz_offset(192, 528 - 384, k)) // 2] for k in (0, 50, 104, 106))
    for o in (0, 77, 131):
        assert i2[(2 * 55296 + policy.swz_offset(144, o, 528 - 384)) // 2] == policy.to_bf16_bits(b2[o])
    for (h, k) in ((0, 0), (191, 104), (192, 7), (527, 64), (300, 63)):
        c, n = divmod(h, 192)
        assert i1[(c * 49152 + policy.swz_offset(192, n, k)) // 2] == policy.to_bf16_bits(w1[h, k])
    for (o, h) in ((0, 0), (131, 527), (77, 191), (5, 192), (100, 400)):
        c, k = divmod(h, 192)
        assert i2[(c * 55296 + policy.swz_offset(144, o, k)) // 2] == policy.to_bf16_bits(w2[o, h])
    # padding (hidden 529..575, outputs 132..143, inputs 106..127) is zero
    assert i1[(2 * 49152 + policy.swz_offset(192, 150, 3)) // 2] == 0 and i2[policy.swz_offset(144, 140, 9) // 2] == 0
    assert i1[policy.swz_offset(192, 3, 120) // 2] == 0
