#!/usr/bin/env python
"""Tile sweep of the block-fused engine (tuning aid): per-kernel device time of one trunk pass for a list of pinned tiles.

    python tools/tile_sweep.py --config hela --images 64 "0:256:8:128;2:256:8:128" "0:256:16:64;2:256:16:64" ...
"""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inconsistencymasks_b200 import unet, _lib  # noqa: E402

CFG = {"hela": (256, 256, 1, 3, 1.0, "sigmoid"), "suim": (256, 256, 3, 9, 2.0, "softmax"),
       "city": (208, 416, 3, 35, 1.0, "softmax"), "isic": (256, 256, 3, 1, 0.5, "sigmoid")}
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="hela")
ap.add_argument("--images", type=int, default=64)
ap.add_argument("tiles", nargs="*")
a = ap.parse_args()
h, w, c, k, alpha, act = CFG[a.config]
x = torch.randint(0, 256, (a.images, h, w, c), dtype=torch.uint8, device="cuda")
for spec in [""] + a.tiles:
    os.environ["IMK_BT_TILE"] = spec
    m = unet.B200UNet(h, w, c, k, alpha, act, unet.init_weights(c, k, alpha, seed=1))
    m.set_engine("fused")
    for _ in range(2):
        m.forward_device(x)
    torch.cuda.synchronize()
    _lib.profile_begin()
    for _ in range(3):
        m.forward_device(x)
    prof = _lib.profile_end()
    tot = sum(p["total_ms"] for p in prof) / 3
    rows = " ".join(f"{p['name']}[{p['tag']}]={1e3 * p['total_ms'] / p['launches']:.1f}" for p in prof)
    print(f"{spec or 'planner':40s} total {1e3 * tot:8.1f} us | {rows}", flush=True)
    m.close()
