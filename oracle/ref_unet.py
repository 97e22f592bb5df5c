"""CPU oracle for the U-Net forward pass (``model.predict``) of the hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/ref_im.py``): never imported by the
product package.

PyTorch **fp32, CPU** restatement of ``/root/reference/unet.py:4-67`` consuming
weights in Keras ``model.get_weights()`` order (conv kernels HWIO + bias, BN
gamma/beta/moving_mean/moving_variance -- 104 arrays).

Parity status: **parity unpinned** for this row.  The layer
arithmetic lives in TensorFlow/Keras, which is not vendored by the reference,
not version-pinned by it, not installed here and not installable (no network);
the reference ships no test, golden vector or saved activation for ``predict``.
What is encoded below is Keras' *documented* inference behaviour:
  * ``Conv2D(padding='same', strides=1, use_bias=True)`` -- zero padding,
    cross-correlation, kernel ``(kh, kw, Cin, Cout)``; activation after bias,
  * ``BatchNormalization(axis=-1, epsilon=1e-3)`` in inference form
    ``gamma * (x - mean) / sqrt(var + eps) + beta``,
  * ``MaxPooling2D((2, 2))`` stride 2 valid; ``UpSampling2D((2, 2))`` nearest,
  * ``Lambda(x / 255)`` on the float-cast uint8 input (unet.py:5),
  * ``sigmoid`` / ``softmax(axis=-1)`` on the fp32 output layer (unet.py:63).
A second, independent plain-C restatement (``oracle/unet_oracle.c``) is checked
against this one in ``tests/test_oracle_unet.py`` to catch layout mistakes.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3  # Keras BatchNormalization default


def widths(alpha):
    """Channel widths ``int(k * alpha)`` -- unet.py:49-61."""
    return {k: int(k * alpha) for k in (16, 32, 64, 128, 256)}


def layer_plan(c, num_out, alpha, ks=3):
    """The chain of parameterised layers in creation order (unet.py:49-63).

    Returns a list of ``("conv", kh, cin, cout)`` / ``("bn", ch)`` tuples; the
    flat weight list has 2 arrays per conv and 4 per bn in exactly this order.
    """
    f = widths(alpha)
    plan = [("conv", 1, c, f[16]), ("bn", f[16])]                      # input_block, unet.py:4-9
    cin = f[16]
    for w in (f[16], f[32], f[64], f[128]):                            # encoder_block x4, unet.py:11-19
        plan += [("conv", ks, cin, w), ("conv", 1, w, w), ("bn", w)]
        cin = w
    plan += [("conv", ks, cin, f[256]), ("conv", 1, f[256], f[128]), ("bn", f[128])]  # bottleneck, unet.py:22-29
    cin = f[128]
    for c1, c2 in ((f[128], f[64]), (f[64], f[32]), (f[32], f[16]), (f[16], f[16])):  # decoder_block x4, unet.py:31-43
        plan += [("conv", 1, cin, c1), ("bn", c1), ("conv", ks, c1, c1), ("conv", 1, c1, c2), ("bn", c2)]
        cin = c2
    plan.append(("conv", 1, cin, num_out))                              # 'out', unet.py:63
    return plan


def count_params(c, num_out, alpha, ks=3):
    n = 0
    for item in layer_plan(c, num_out, alpha, ks):
        if item[0] == "conv":
            _, k, cin, cout = item
            n += k * k * cin * cout + cout
        else:
            n += 4 * item[1]
    return n


class _Cursor:
    def __init__(self, weights):
        self.w = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)) for a in weights]
        self.i = 0

    def take(self, n):
        out = self.w[self.i:self.i + n]
        self.i += n
        return out


def _conv_act(x, cur, relu=True):
    k, b = cur.take(2)                              # HWIO, (Cout,)
    pad = k.shape[0] // 2                           # 'same', stride 1, odd kernels
    y = F.conv2d(x, k.permute(3, 2, 0, 1).contiguous(), b, padding=pad)
    return torch.relu(y) if relu else y


def _bn(x, cur):
    g, b, mu, var = cur.take(4)
    scale = g / torch.sqrt(var + BN_EPS)
    return x * scale.view(1, -1, 1, 1) + (b - mu * scale).view(1, -1, 1, 1)


def _h(x):
    """Round to fp16 and back: the value an fp16 activation / weight buffer holds."""
    return x.to(torch.float16).to(torch.float32)


def _layer(x, cur, bn, storage, first=False):
    """Conv2D + ReLU (+ BatchNormalization) -- unet.py:6-7, 12-16, 24-28, 34-41.

    storage == "fp32": the plain fp32 arithmetic.  storage == "fp16": the same layer with the ROUNDING POINTS of a
    mixed-precision execution that keeps activations and hidden-layer weights in fp16 and accumulates in fp32 (the
    reference runs ``mixed_float16``, 09_ISIC_2018_IM.py:16): inputs are fp16 values already, the weights are rounded to
    fp16 (with the BN scale folded in, ``relu(v) * s + t == max(s * v + t, t)`` for ``s > 0``, ``min`` for ``s < 0``),
    bias / shift are added in fp32 and the layer output is rounded to fp16 ONCE.  The first layer keeps fp32 weights.
    """
    if storage == "fp32":
        y = _conv_act(x, cur)
        return _bn(y, cur) if bn else y
    k, b = cur.take(2)
    pad = k.shape[0] // 2
    if bn:
        g, be, mu, var = cur.take(4)
        scale = g / torch.sqrt(var + BN_EPS)
        shift = be - mu * scale
    else:
        scale, shift = torch.ones_like(b), torch.zeros_like(b)
    w = k * scale.view(1, 1, 1, -1)
    if not first:
        w = _h(w)
    v = F.conv2d(x, w.permute(3, 2, 0, 1).contiguous(), None, padding=pad) + (scale * b + shift).view(1, -1, 1, 1)
    t = shift.view(1, -1, 1, 1)
    pos = (scale > 0).view(1, -1, 1, 1)
    zero = (scale == 0).view(1, -1, 1, 1)
    out = torch.where(pos, torch.maximum(v, t), torch.minimum(v, t))
    out = torch.where(zero, t.expand_as(out), out)
    return _h(out)


@torch.no_grad()
def forward(images, weights, actifuout="sigmoid", return_logits=False, storage="fp32"):
    """``model.predict`` -- unet.py:46-67.

    images: uint8 or float ``[N, H, W, c]`` NHWC; weights: the 104 arrays.
    Returns float32 ``[N, H, W, K]`` probabilities (or pre-activation logits).
    ``storage="fp16"`` models fp16 activation / weight storage with fp32 accumulation (see ``_layer``): the
    mixed-precision twin of the fp32 oracle, used to separate "fp16 storage" from "kernel error" in the GPU tests.
    """
    if storage not in ("fp32", "fp16"):
        raise ValueError(storage)
    x = torch.from_numpy(np.ascontiguousarray(images)).to(torch.float32).permute(0, 3, 1, 2)
    cur = _Cursor(weights)
    x = x / 255.0                                   # unet.py:5
    x = _layer(x, cur, True, storage, first=True)   # unet.py:6-7
    skips = []
    for _ in range(4):                              # unet.py:51-54
        x = _layer(_layer(x, cur, False, storage), cur, True, storage)
        skips.append(x)
        x = F.max_pool2d(x, 2)
    x = _layer(_layer(x, cur, False, storage), cur, True, storage)  # unet.py:56
    for skip in reversed(skips):                    # unet.py:58-61
        u = F.interpolate(x, scale_factor=2, mode="nearest") + skip
        if storage == "fp16":
            u = _h(u)                               # an fp16 add
        x = _layer(u, cur, True, storage)
        x = _layer(_layer(x, cur, False, storage), cur, True, storage)
    logits = _conv_act(x, cur, relu=False)          # unet.py:63 (fp32 output layer)
    assert cur.i == len(cur.w), "weight list length does not match the layer plan"
    if return_logits:
        out = logits
    elif actifuout == "sigmoid":
        out = torch.sigmoid(logits)
    elif actifuout == "softmax":
        out = torch.softmax(logits, dim=1)
    else:
        raise ValueError(f"unsupported output activation {actifuout!r}")
    return out.permute(0, 2, 3, 1).contiguous().numpy()


class OracleModel:
    """Duck-typed stand-in for a loaded Keras model: ``.predict(x) -> float32 NHWC``.
    Accepts an array or a one-element list (functions.py:3157 passes ``[image]``)."""

    def __init__(self, weights, actifuout):
        self.weights = weights
        self.actifuout = actifuout

    def predict(self, x, batch_size=32, verbose=0):
        if isinstance(x, (list, tuple)):
            x = x[0]
        return forward(np.asarray(x), self.weights, self.actifuout)
