#!/usr/bin/env python
"""Stand-alone timing of the fused IM kernels on materialised fp32 probabilities
(SURVEY.md 8d byte formula), for ncu captures and quick A/B runs.

    python tools/im_kernel_bench.py [--config hela|isic2|isic5|suim|cityscapes] [--images N] [--iters I]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inconsistencymasks_b200 import _lib  # noqa: E402
from inconsistencymasks_b200._lib import lib, check  # noqa: E402

CONFIGS = {  # H, W, c, K, M, multiclass, strict
    "isic2": (256, 256, 3, 1, 2, False, 1), "isic5": (256, 256, 3, 1, 5, False, 1), "hela": (256, 256, 1, 3, 2, False, 0),
    "suim": (256, 256, 3, 9, 2, True, 0), "cityscapes": (208, 416, 3, 35, 2, True, 0),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="hela")
    ap.add_argument("--images", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=8)
    args = ap.parse_args()
    H, W, c, K, M, mc, strict = CONFIGS[args.config]
    N = args.images
    dev = torch.device("cuda", 0)
    probs = [torch.rand((N, H, W, K), dtype=torch.float32, device=dev) for _ in range(M)]
    ptrs = (C.c_void_p * M)(*[p.data_ptr() for p in probs])
    img = torch.randint(0, 256, (N, H, W, c), dtype=torch.uint8, device=dev)
    out = torch.empty_like(img)
    planes = 1 if mc else K
    lab = torch.empty((planes, N, H, W), dtype=torch.uint8, device=dev)
    im = torch.empty((N, H, W), dtype=torch.uint8, device=dev)
    sz = torch.empty(N, dtype=torch.int64, device=dev)
    pred = torch.empty((planes, N), dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    ms = []
    for it in range(args.iters):
        flush.fill_(it)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if mc:
            check(lib.imk_im_multiclass(ptrs, M, N, H, W, K, img.data_ptr(), c, 1, 1, out.data_ptr(), lab.data_ptr(), im.data_ptr(),
                                        sz.data_ptr(), None, s))
        else:
            check(lib.imk_im_binary(ptrs, M, N, H, W, K, 0.5, strict, img.data_ptr(), c, 1, 1, out.data_ptr(), lab.data_ptr(),
                                    im.data_ptr(), sz.data_ptr(), pred.data_ptr(), s))
        b.record()
        torch.cuda.synchronize()
        if it >= 3:
            ms.append(a.elapsed_time(b))
    bpi = H * W * (4 * K * M + c + c + planes + 1)
    t = float(np.mean(ms))
    peak = 6457.7
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    gbs = N * bpi / (t * 1e-3) / 1e9
    print(json.dumps(dict(config=args.config, images=N, ms=t, bytes_per_image=bpi, gbs=gbs, frac=gbs / peak, peak=peak)))


if __name__ == "__main__":
    main()
