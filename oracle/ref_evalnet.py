"""TEST INFRASTRUCTURE ONLY -- PyTorch fp32 CPU restatement of the reference's EvalNet (never imported by the product).

    input_block / conv_block / get_evalnet / get_evalnet_miou    /root/reference/evalnet.py:4-73

Keras inference semantics as in oracle/ref_unet.py (Conv2D same/zero padding, HWIO kernels, ReLU inside the conv, BN eps 1e-3,
MaxPooling2D 2x2 valid, GlobalAvgPool2D, Dense + sigmoid).  PARITY UNPINNED at the TensorFlow boundary for the same
reason as the U-Net oracle: TensorFlow is not installed here and the reference ships no saved activations.
Weights: the order evalnet.py creates its layers (inconsistencymasks_b200.weights.evalnet_plan).
"""
import numpy as np
import torch
import torch.nn.functional as Fn


def forward(a, b, weights, alpha, n_heads=1, normalize_a=True, normalize_b=True, ks=3):
    """a: uint8 [N,H,W,cA]; b: [N,H,W,cB] numeric.  Returns float32 [N,1] or (iou [N,K], detection [N,K])."""
    it = iter(weights)

    def conv(x, relu=True):
        k, bias = torch.from_numpy(next(it)), torch.from_numpy(next(it))
        y = Fn.conv2d(x, k.permute(3, 2, 0, 1).contiguous(), bias, padding=k.shape[0] // 2)
        return torch.relu(y) if relu else y

    def bn(x):
        g, be, mu, var = (torch.from_numpy(next(it)) for _ in range(4))
        return (x - mu[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + 1e-3) * g[None, :, None, None] + be[None, :, None, None]

    def input_block(x, normalize):                      # evalnet.py:4-11
        x = x / 255.0 if normalize else x
        return bn(conv(x))

    def conv_block(x):                                  # evalnet.py:14-21
        return Fn.max_pool2d(bn(conv(conv(x))), 2)

    with torch.no_grad():
        xa = torch.from_numpy(np.asarray(a, np.float32)).permute(0, 3, 1, 2)
        xb = torch.from_numpy(np.asarray(b, np.float32)).permute(0, 3, 1, 2)
        ya = conv_block(input_block(xa, normalize_a))
        yb = conv_block(input_block(xb, normalize_b))
        c = torch.cat([ya, yb], dim=1)
        for _ in range(5):
            c = conv_block(c)
        g = c.mean(dim=(2, 3))
        outs = []
        for _ in range(n_heads):
            w, bias = torch.from_numpy(next(it)), torch.from_numpy(next(it))
            outs.append(torch.sigmoid(g @ w + bias).numpy())
    return outs[0] if n_heads == 1 else tuple(outs)
