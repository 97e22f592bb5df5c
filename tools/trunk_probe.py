#!/usr/bin/env python
"""One U-Net trunk pass (for ncu captures): python tools/trunk_probe.py [--config hela|suim|city|isic] [--images 64] [--passes 2]"""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inconsistencymasks_b200 import unet  # noqa: E402

CFG = {"hela": (256, 256, 1, 3, 1.0, "sigmoid"), "suim": (256, 256, 3, 9, 2.0, "softmax"),
       "city": (208, 416, 3, 35, 1.0, "softmax"), "isic": (256, 256, 3, 1, 0.5, "sigmoid")}
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="hela")
ap.add_argument("--images", type=int, default=64)
ap.add_argument("--passes", type=int, default=2)
ap.add_argument("--engine", default="tcgen05")
a = ap.parse_args()
h, w, c, k, alpha, act = CFG[a.config]
m = unet.B200UNet(h, w, c, k, alpha, act, unet.init_weights(c, k, alpha, seed=1))
m.set_engine(a.engine)
x = torch.randint(0, 256, (a.images, h, w, c), dtype=torch.uint8, device="cuda")
for _ in range(a.passes):
    y = m.forward_device(x)
torch.cuda.synchronize()
print("ok", float(y.float().mean()))
