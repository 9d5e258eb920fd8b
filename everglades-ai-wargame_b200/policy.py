"""Policy-in-the-loop glue: the reference's DQN network (agents/DQN/QNetwork.py:37,42 — Linear(105, 528) - ReLU -
Linear(528, 132)) evaluated by libevgsim's fused tensor-core kernel (evg_policy_mlp: tcgen05 + TMEM, bf16 operands, fp32
accumulation) on the observation tensor the step kernel writes, and decoded by evg_decode_dqn.

``pack_mlp`` turns the two weight matrices into the images the kernel copies straight into shared memory (bf16, K-major,
128-byte swizzle; layout in include/evgsim.h).  ``FusedDQN`` wraps a torch ``Sequential(Linear, ReLU, Linear)`` (or raw
arrays) for a BatchedEvergladesEnv.  Plumbing only; the network itself is the caller's.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi

IN_PAD, CHUNK, OUT_PAD, ATOM = _capi.MLP_IN_PAD, _capi.MLP_CHUNK, _capi.MLP_OUT_PAD, 64


def to_bf16_bits(x):
    """float32 array -> uint16 bf16 bit patterns, round to nearest even (what __float2bfloat16_rn does)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    return ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)


def bf16_round(x):
    """float32 array rounded to bf16 precision (as float32): the operands the kernel multiplies."""
    return (to_bf16_bits(x).astype(np.uint32) << 16).view(np.float32).reshape(np.shape(x))


def swz_offset(rows, r, k):
    """Byte offset of element (row r, column k) in a K-major [rows x K] bf16 block with the 128-byte swizzle."""
    return (k // ATOM) * rows * 128 + r * 128 + ((((k % ATOM) >> 3) ^ (r & 7)) << 4) + (k & 7) * 2


def _image(mat, rows, cols):
    """mat [<=rows, <=cols] float32 -> the uint8 image of a zero-padded [rows x cols] block."""
    full = np.zeros((rows, cols), dtype=np.float32)
    full[:mat.shape[0], :mat.shape[1]] = mat
    bits = to_bf16_bits(full)
    r, k = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
    img = np.zeros((cols // ATOM) * rows * 128 // 2, dtype=np.uint16)
    img[swz_offset(rows, r, k) // 2] = bits
    return img.view(np.uint8)


def pack_mlp(w1, b1, w2, b2):
    """w1 [hidden, in], b1 [hidden], w2 [out, hidden], b2 [out] (float32, torch.nn.Linear's layout) ->
    (w1_img uint8, w2_img uint8, hidden, out) as evg_policy_mlp expects them.  The biases ride inside the images:
    input feature `in` is the constant 1 (its weights are b1), hidden unit `hidden` is wired to relu(1) = 1 (its outgoing
    weights are b2)."""
    w1, b1, w2, b2 = (np.asarray(a, dtype=np.float32) for a in (w1, b1, w2, b2))
    hidden, in_dim = w1.shape
    out = w2.shape[0]
    assert in_dim < IN_PAD and out <= OUT_PAD and w2.shape[1] == hidden and b1.shape == (hidden,) and b2.shape == (out,)
    n_chunks = -(-(hidden + 1) // CHUNK)
    w1p = np.zeros((n_chunks * CHUNK, in_dim + 1), dtype=np.float32)
    w1p[:hidden, :in_dim] = w1
    w1p[:hidden, in_dim] = b1
    w1p[hidden, in_dim] = 1.0
    w2p = np.zeros((out, n_chunks * CHUNK), dtype=np.float32)
    w2p[:, :hidden] = w2
    w2p[:, hidden] = b2
    img1 = np.concatenate([_image(w1p[c * CHUNK:(c + 1) * CHUNK], CHUNK, IN_PAD) for c in range(n_chunks)])
    img2 = np.concatenate([_image(w2p[:, c * CHUNK:(c + 1) * CHUNK], OUT_PAD, CHUNK) for c in range(n_chunks)])
    return img1, img2, hidden, out


def reference_forward(obs, w1, b1, w2, b2):
    """What the kernel computes, in numpy: bf16-rounded operands (weights, biases, observations, hidden activations),
    float32 accumulation."""
    x = bf16_round(np.asarray(obs, dtype=np.float32))
    h = np.maximum(x @ bf16_round(w1).T + bf16_round(b1), 0.0).astype(np.float32)
    return bf16_round(h) @ bf16_round(w2).T + bf16_round(b2)


class FusedDQN:
    """Q-network forward + decode for a BatchedEvergladesEnv: ``actions = FusedDQN(env, net)()`` reads env.obs in place."""

    def __init__(self, env, net=None, weights=None):
        import torch
        if weights is None:
            lin = [m for m in net if isinstance(m, torch.nn.Linear)]
            assert len(lin) == 2, "expected Sequential(Linear, ReLU, Linear)"
            weights = [t.detach().float().cpu().numpy() for t in (lin[0].weight, lin[0].bias, lin[1].weight, lin[1].bias)]
        img1, img2, self.hidden, self.out_dim = pack_mlp(*weights)
        self.env = env
        dev = env.device
        self._w1 = torch.from_numpy(img1).to(dev)
        self._w2 = torch.from_numpy(img2).to(dev)
        self.q = torch.empty((env.num_envs, 2, self.out_dim), dtype=torch.float32, device=dev)
        self.qt = torch.empty((self.out_dim, env.num_envs * 2), dtype=torch.float32, device=dev)

    def _run(self, obs, out, transposed):
        env = self.env
        obs = env.obs if obs is None else obs
        _capi.check(env._lib.evg_policy_mlp(env._h, C.c_void_p(obs.data_ptr()), env.num_envs * 2, C.c_void_p(self._w1.data_ptr()),
                                            C.c_void_p(self._w2.data_ptr()), self.hidden, self.out_dim, C.c_void_p(out.data_ptr()),
                                            int(transposed), env._stream()))
        return out

    def forward(self, obs=None):
        """Q-values float32 [N, 2, out] for `obs` (default: the environment's own observation tensor)."""
        return self._run(obs, self.q, False)

    def forward_t(self, obs=None):
        """The same Q-values transposed, float32 [out, N * 2]: what the decode reads coalesced."""
        return self._run(obs, self.qt, True)

    def __call__(self):
        """Action rows int8 [N, 2, 7, 2] for both players: forward, then DQNAgent.filter_actions on the device."""
        env = self.env
        qt = self.forward_t()
        assert self.out_dim % _capi.NUM_GROUPS == 0
        _capi.check(env._lib.evg_decode_dqn_layout(env._h, C.c_void_p(qt.data_ptr()), self.out_dim // _capi.NUM_GROUPS, -1, 1,
                                                   C.c_void_p(env._actions.data_ptr()), env._stream()))
        return env._actions
