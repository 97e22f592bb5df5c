#!/usr/bin/env python
"""Debug helper: fused vs layer-wise engine on one config: python tools/fused_check.py H W c K alpha [n]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inconsistencymasks_b200 import unet
h, w, c, K, alpha = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5])
n = int(sys.argv[6]) if len(sys.argv) > 6 else 3
act = "sigmoid" if K <= 3 else "softmax"
wts = unet.init_weights(c, K, alpha, seed=7)
img = np.random.default_rng(3).integers(0, 256, size=(n, h, w, c), dtype=np.uint8)
m = unet.B200UNet(h, w, c, K, alpha, act, wts)
m.set_engine("tcgen05"); a = m.predict(img)
m.set_engine("fused"); b = m.predict(img)
d = np.abs(a - b)
print(f"kinds={os.environ.get('IMK_BT_KINDS')} {h}x{w} c{c} K{K} a{alpha}: max|d|={d.max():.3e} mean|d|={d.mean():.3e}", flush=True)
