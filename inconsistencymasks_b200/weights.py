"""Layer plan and seeded synthetic weights of the reference U-Net (unet.py:46-67) -- pure NumPy.

This module does NOT load libimk.so: bench.py's reference arm, the oracle tests and the golden
generators import it without touching the CUDA product path (``inconsistencymasks_b200.unet``
re-exports these names for users).
"""
from __future__ import annotations

import numpy as np

__all__ = ["layer_plan", "count_params", "init_weights", "evalnet_plan", "init_evalnet_weights"]


def layer_plan(i_channels, num_outputmasks, alpha, ks=3):
    """Parameterised layers in creation order (unet.py:49-63): ``("conv", k, cin, cout)`` / ``("bn", ch)``."""
    f = [int(k * alpha) for k in (16, 32, 64, 128, 256)]
    plan = [("conv", 1, i_channels, f[0]), ("bn", f[0])]
    cin = f[0]
    for w in f[:4]:
        plan += [("conv", ks, cin, w), ("conv", 1, w, w), ("bn", w)]
        cin = w
    plan += [("conv", ks, cin, f[4]), ("conv", 1, f[4], f[3]), ("bn", f[3])]
    cin = f[3]
    for c1, c2 in ((f[3], f[2]), (f[2], f[1]), (f[1], f[0]), (f[0], f[0])):
        plan += [("conv", 1, cin, c1), ("bn", c1), ("conv", ks, c1, c1), ("conv", 1, c1, c2), ("bn", c2)]
        cin = c2
    plan.append(("conv", 1, cin, num_outputmasks))
    return plan


def count_params(i_channels, num_outputmasks, alpha, ks=3):
    n = 0
    for item in layer_plan(i_channels, num_outputmasks, alpha, ks):
        n += (item[1] ** 2 * item[2] * item[3] + item[3]) if item[0] == "conv" else 4 * item[1]
    return n


def init_weights(i_channels, num_outputmasks, alpha, ks=3, seed=0, trained_like=True):
    """Seeded random weights in get_weights() order.

    Conv kernels follow Keras ``he_normal`` (unet.py:46): truncated normal (2 sigma) with
    stddev ``sqrt(2 / fan_in) / 0.87962566``.  With ``trained_like`` the biases and the
    BatchNormalization statistics are non-trivial (bias N(0, .05), gamma U(.5, 1.5),
    beta N(0, .1), mean N(0, .1), var U(.5, 1.5)) so that every term
    of the inference arithmetic is exercised; otherwise they are Keras' fresh-model values.
    """
    rng = np.random.default_rng(seed)
    out = []
    for item in layer_plan(i_channels, num_outputmasks, alpha, ks):
        if item[0] == "conv":
            _, k, cin, cout = item
            std = np.sqrt(2.0 / (k * k * cin)) / 0.87962566103423978
            w = rng.standard_normal((k, k, cin, cout))
            bad = np.abs(w) > 2.0
            while bad.any():                       # truncated normal by resampling
                w[bad] = rng.standard_normal(int(bad.sum()))
                bad = np.abs(w) > 2.0
            out.append((w * std).astype(np.float32))
            out.append((rng.normal(0, 0.05, cout) if trained_like else np.zeros(cout)).astype(np.float32))
        else:
            ch = item[1]
            if trained_like:
                out += [rng.uniform(0.5, 1.5, ch).astype(np.float32), rng.normal(0, 0.1, ch).astype(np.float32),
                        rng.normal(0, 0.1, ch).astype(np.float32), rng.uniform(0.5, 1.5, ch).astype(np.float32)]
            else:
                out += [np.ones(ch, np.float32), np.zeros(ch, np.float32), np.zeros(ch, np.float32), np.ones(ch, np.float32)]
    return out


def evalnet_plan(a_channels, b_channels, alpha, n_heads=1, ks=3):
    """Parameterised layers of get_evalnet / get_evalnet_miou in the order evalnet.py creates them (evalnet.py:24-73):
    branch A (input block, conv_block), branch B, five trunk conv_blocks, the Dense head(s) as ``("dense", cin, cout)``."""
    f = [int(k * alpha) for k in (16, 32, 64, 128, 256)]
    plan = []
    for c_in in (a_channels, b_channels):
        plan += [("conv", 1, c_in, f[0]), ("bn", f[0]), ("conv", ks, f[0], f[0]), ("conv", 1, f[0], f[0]), ("bn", f[0])]
    cin = 2 * f[0]
    for w in f:
        plan += [("conv", ks, cin, w), ("conv", 1, w, w), ("bn", w)]
        cin = w
    n_out = 1 if n_heads == 1 else b_channels
    plan += [("dense", f[4], n_out)] * n_heads
    return plan


def init_evalnet_weights(a_channels, b_channels, alpha, n_heads=1, ks=3, seed=0):
    """Seeded random EvalNet weights in ``evalnet_plan`` order (he_normal convs, glorot-like Dense, trained-like BN)."""
    rng = np.random.default_rng(seed)
    out = []
    for item in evalnet_plan(a_channels, b_channels, alpha, n_heads, ks):
        if item[0] == "conv":
            _, k, cin, cout = item
            std = np.sqrt(2.0 / (k * k * cin))
            out.append((np.clip(rng.standard_normal((k, k, cin, cout)), -2, 2) * std).astype(np.float32))
            out.append(rng.normal(0, 0.05, cout).astype(np.float32))
        elif item[0] == "bn":
            ch = item[1]
            out += [rng.uniform(0.5, 1.5, ch).astype(np.float32), rng.normal(0, 0.1, ch).astype(np.float32),
                    rng.normal(0, 0.1, ch).astype(np.float32), rng.uniform(0.5, 1.5, ch).astype(np.float32)]
        else:
            _, cin, cout = item
            lim = np.sqrt(6.0 / (cin + cout))
            out.append(rng.uniform(-lim, lim, (cin, cout)).astype(np.float32))
            out.append(rng.normal(0, 0.05, cout).astype(np.float32))
    return out
