"""ORACLE (test infrastructure — never imported by the product path).

CPU restatement of the reference's regressor training step (_4_train_model.py:119-127,196-204): the reference's own
building blocks — ``SimpleFC`` (utils/nn_model.py:6-41: Linear -> LeakyReLU -> Dropout per hidden layer, Linear ->
Sigmoid), ``nn.MSELoss`` on ``outputs.squeeze()``, ``torch.optim.Adam(lr, weight_decay)`` — in torch CPU fp32 with
autograd, with ONE substitution: ``nn.Dropout``'s random mask is replaced by an explicit mask drawn from the same
counter-based Philox4x32-10 stream the CUDA kernels use (key = seed, counter = (unit // 4, layer, step lo, step hi),
keep iff (word >> 8) * 2^-24 >= p), because torch's own mask depends on the device and launch geometry.

Pinned by tests/golden/train_ref.npz: the final weights of the UNMODIFIED reference ``train()`` (dropout_prob = 0, CPU,
synthetic labelled ``.pt`` directory; tools/gen_golden.py gen_train) must be reproduced by this restatement driven by the
product's host logic (same pandas shuffle, random_split, init and DataLoader permutations).
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 (Salmon et al., SC'11): uint32 arrays/scalars in, 4 uint32 arrays out."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & _MASK for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return [c.astype(np.uint32) for c in (c0, c1, c2, c3)]


def dropout_scale(seed: int, step: int, layer: int, B: int, H: int, p: float) -> np.ndarray:
    """f32 [B,H]: 1/(1-p) where the unit is kept, 0 where it is dropped (unit index e = b*H + j)."""
    if p <= 0:
        return np.ones((B, H), np.float32)
    e = np.arange(B * H, dtype=np.uint64)
    words = philox4x32_10(e >> np.uint64(2), layer, step & 0xFFFFFFFF, step >> 32, seed & 0xFFFFFFFF, seed >> 32)
    sel = np.choose((e & np.uint64(3)).astype(np.int64), words)
    u = (sel >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    keep = u >= np.float32(p)
    return (keep.astype(np.float32) * (np.float32(1.0) / (np.float32(1.0) - np.float32(p)))).reshape(B, H)


class SimpleFCOracle(nn.Module):
    """utils/nn_model.py:6-41 with the dropout mask supplied from outside."""

    def __init__(self, input_size, hidden_sizes, output_size, slope=0.01):
        super().__init__()
        sizes = [input_size] + list(hidden_sizes) + [output_size]
        self.linears = nn.ModuleList([nn.Linear(sizes[i], sizes[i + 1]) for i in range(len(sizes) - 1)])
        self.slope = slope

    def forward(self, x, scales=None):
        n = len(self.linears)
        for l, lin in enumerate(self.linears):
            x = lin(x)
            if l < n - 1:
                x = torch.nn.functional.leaky_relu(x, self.slope)
                if scales is not None:
                    x = x * scales[l]
            else:
                x = torch.sigmoid(x)
        return x


def train_epoch_oracle(model: SimpleFCOracle, opt: torch.optim.Adam, feats: torch.Tensor, labels: torch.Tensor, order, batch: int,
                       dropout_p: float, seed: int, step0: int):
    """One epoch of _4_train_model.py:196-204 over ``order`` (sample indices); returns (sum of batch losses, steps done).
    ``step0`` = optimiser steps taken before this epoch (the dropout stream position is step0 + 1, + 2, ...)."""
    crit = nn.MSELoss()
    total, step = 0.0, step0
    order = list(order)
    for off in range(0, len(order), batch):
        idx = order[off:off + batch]
        step += 1
        x, y = feats[idx], labels[idx]
        scales = None
        if dropout_p > 0:
            scales = [torch.from_numpy(dropout_scale(seed, step, l, len(idx), lin.out_features, dropout_p))
                      for l, lin in enumerate(model.linears[:-1])]
        opt.zero_grad()
        out = model(x, scales)
        loss = crit(out.squeeze(), y) if len(idx) > 1 else crit(out.reshape(-1), y.reshape(-1))
        loss.backward()
        opt.step()
        total += loss.item()
    return total, step


def cosine_warm_restarts_lr(base_lr: float, eta_min: float, T_0: int, epoch_done: int) -> float:
    """Learning rate in effect AFTER ``epoch_done`` scheduler steps of CosineAnnealingWarmRestarts(T_0, T_mult=1, eta_min)
    (_4_train_model.py:126,206)."""
    import math
    t_cur = epoch_done % T_0
    return eta_min + (base_lr - eta_min) * (1 + math.cos(math.pi * t_cur / T_0)) / 2
