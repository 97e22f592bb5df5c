#!/usr/bin/env python
"""Fused ensemble path (U-Net trunks + last layer + IM + blanking) on any dataset shape, with the per-kernel profile:

    python tools/ensemble_bench.py --config suim|city|city2|isic|isic5|hela [--images 512]
"""
import argparse, ctypes as C, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inconsistencymasks_b200 import unet, _lib  # noqa: E402
from inconsistencymasks_b200._lib import lib, check  # noqa: E402

CFG = {  # H, W, c, K, alpha, act, M, multiclass, strict
    "hela": (256, 256, 1, 3, 1.0, "sigmoid", 2, False, 0), "isic": (256, 256, 3, 1, 0.5, "sigmoid", 2, False, 1),
    "isic5": (256, 256, 3, 1, 0.5, "sigmoid", 5, False, 1), "suim": (256, 256, 3, 9, 2.0, "softmax", 2, True, 0),
    "city": (208, 416, 3, 35, 1.0, "softmax", 2, True, 0), "city2": (208, 416, 3, 35, 2.0, "softmax", 2, True, 0),
}
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="suim")
ap.add_argument("--images", type=int, default=512)
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
H, W, c, K, alpha, act, M, mc, strict = CFG[a.config]
N = a.images
dev = torch.device("cuda", 0)
models = [unet.B200UNet(H, W, c, K, alpha, act, unet.init_weights(c, K, alpha, seed=7 + j)) for j in range(M)]
handles = (C.c_void_p * M)(*[m.handle for m in models])
img = torch.randint(0, 256, (N, H, W, c), dtype=torch.uint8, device=dev)
out = torch.empty_like(img)
planes = 1 if mc else K
lab = torch.empty((planes, N, H, W), dtype=torch.uint8, device=dev)
im = torch.empty((N, H, W), dtype=torch.uint8, device=dev)
sz = torch.empty(N, dtype=torch.int64, device=dev)
pred = torch.empty((planes, N), dtype=torch.int64, device=dev)
s = torch.cuda.current_stream().cuda_stream


def step():
    if mc:
        check(lib.imk_ensemble_im_multiclass(handles, M, img.data_ptr(), N, 0, 1, 1, out.data_ptr(), lab.data_ptr(), im.data_ptr(), sz.data_ptr(), None, s))
    else:
        check(lib.imk_ensemble_im_binary(handles, M, img.data_ptr(), N, 0, 0.5, strict, 1, 1, out.data_ptr(), lab.data_ptr(), im.data_ptr(), sz.data_ptr(), pred.data_ptr(), s))


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
_lib.profile_begin()
step()
prof = _lib.profile_end()
tot = sum(p["total_ms"] for p in prof)
rows = sorted(prof, key=lambda p: -p["total_ms"])
print(json.dumps(dict(config=a.config, images=N, ms_per_step=ms, images_per_s=N / ms * 1e3,
                      kernels=[dict(kernel=p["name"], layer=p["tag"], launches=p["launches"], total_us=round(1e3 * p["total_ms"], 1),
                                    share=round(p["total_ms"] / tot, 3)) for p in rows[:10]])))
