#!/usr/bin/env python
"""Head-stage A/B: device time of the level-0 decoder kernel with the head writing fp32 probabilities (.predict, mode 0)
vs decision bytes (fused ensemble with M = 1, mode 1 / 2):  python tools/head_probe.py [--config hela] [--images 512]"""
import argparse, ctypes as C, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inconsistencymasks_b200 import unet, _lib  # noqa: E402
from inconsistencymasks_b200._lib import lib, check  # noqa: E402
CFG = {"hela": (256, 256, 1, 3, 1.0, "sigmoid", False, 0), "isic": (256, 256, 3, 1, 0.5, "sigmoid", False, 1),
       "suim": (256, 256, 3, 9, 2.0, "softmax", True, 0)}
ap = argparse.ArgumentParser(); ap.add_argument("--config", default="hela"); ap.add_argument("--images", type=int, default=512)
a = ap.parse_args()
H, W, c, K, alpha, act, mc, strict = CFG[a.config]
N = a.images
m = unet.B200UNet(H, W, c, K, alpha, act, unet.init_weights(c, K, alpha, seed=1))
img = torch.randint(0, 256, (N, H, W, c), dtype=torch.uint8, device="cuda")
out = torch.empty_like(img); planes = 1 if mc else K
lab = torch.empty((planes, N, H, W), dtype=torch.uint8, device="cuda"); im = torch.empty((N, H, W), dtype=torch.uint8, device="cuda")
sz = torch.empty(N, dtype=torch.int64, device="cuda"); pred = torch.empty((planes, N), dtype=torch.int64, device="cuda")
hs = (C.c_void_p * 1)(m.handle); s = torch.cuda.current_stream().cuda_stream
def ens():
    if mc: check(lib.imk_ensemble_im_multiclass(hs, 1, img.data_ptr(), N, 0, 1, 1, out.data_ptr(), lab.data_ptr(), im.data_ptr(), sz.data_ptr(), None, s))
    else: check(lib.imk_ensemble_im_binary(hs, 1, img.data_ptr(), N, 0, 0.5, strict, 1, 1, out.data_ptr(), lab.data_ptr(), im.data_ptr(), sz.data_ptr(), pred.data_ptr(), s))
res = {}
for name, fn in (("probs", lambda: m.forward_device(img)), ("decisions", ens)):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    _lib.profile_begin(); fn(); fn(); prof = _lib.profile_end()
    res[name] = {f"{p['name']}:{p['tag']}": round(1e3 * p["total_ms"] / p["launches"], 1) for p in prof if p["name"] in ("block_head", "block_dec", "ensemble_votes", "out_probs", "ensemble_im") and p["tag"] in (20, 23, -1)}
print(json.dumps(dict(config=a.config, images=N, us=res)))
