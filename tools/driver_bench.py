#!/usr/bin/env python
"""Files/s of the drop-in per-directory drivers (PNG in -> pseudo-label PNGs out), SURVEY.md 8f-3:

    python tools/driver_bench.py [--config isic2|hela|suim] [--files 2048] [--serial]

Generates `--files` synthetic PNGs of the config's shape in a temporary directory, runs
create_pseudo_labels_im_* on it (thread-pooled decode / encode, pinned double buffers, libimk host pipeline) and prints
one JSON line.  `--serial` times the reference's structure for comparison: one file at a time (decode, one-image call,
encode), which is what round 1's drivers did per batch.
"""
import argparse, json, os, sys, tempfile, time
import cv2
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inconsistencymasks_b200 import functions as F, unet as U  # noqa: E402

CFG = {"isic2": (256, 256, 3, 1, 0.5, "sigmoid", 2, "binary"), "hela": (256, 256, 1, 3, 1.0, "sigmoid", 2, "hela"),
       "suim": (256, 256, 3, 9, 2.0, "softmax", 2, "multiclass")}
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="isic2")
ap.add_argument("--files", type=int, default=2048)
ap.add_argument("--serial", action="store_true")
a = ap.parse_args()
h, w, c, K, alpha, act, M, kind = CFG[a.config]
models = [U.B200UNet(h, w, c, K, alpha, act, U.init_weights(c, K, alpha, seed=11 + j)) for j in range(M)]
fn = {"binary": F.create_pseudo_labels_im_ISIC_2018, "hela": F.create_pseudo_labels_im_hela, "multiclass": F.create_pseudo_labels_im_multiclass}[kind]
with tempfile.TemporaryDirectory() as tmp:
    src, dst = os.path.join(tmp, "in"), os.path.join(tmp, "out")
    os.makedirs(src)
    rng = np.random.default_rng(0)
    for i in range(a.files):
        small = rng.integers(0, 256, size=(h // 8, w // 8, c), dtype=np.uint8)
        img = cv2.resize(small, (w, h), interpolation=cv2.INTER_CUBIC).reshape(h, w, c)
        img = np.clip(img.astype(np.int16) + rng.integers(-8, 9, size=img.shape), 0, 255).astype(np.uint8)
        cv2.imwrite(os.path.join(src, f"im_{i:05d}.png"), img if c == 3 else img[..., 0])
    in_bytes = sum(os.path.getsize(os.path.join(src, f)) for f in os.listdir(src))
    kw = dict(erode_kernel=0, dilate_kernel=0)
    if kind == "binary":
        kw["filter_bad_predictions"] = False
    fn(models, h, w, c, src, os.path.join(tmp, "warm"), **kw)            # warm-up: workspaces, page cache
    if a.serial:
        F._FILES_PER_BATCH, F._IO_THREADS = 1, 1
    t0 = time.perf_counter()
    mean = fn(models, h, w, c, src, dst, **kw)
    dt = time.perf_counter() - t0
    out_files = sum(len(os.listdir(os.path.join(dst, d))) for d in os.listdir(dst))
    print(json.dumps(dict(config=a.config, files=a.files, mode="serial" if a.serial else "pipelined", seconds=dt, files_per_s=a.files / dt,
                          png_in_mb=in_bytes / 1e6, png_out_files=out_files, io_threads=F._IO_THREADS, batch=F._FILES_PER_BATCH,
                          host_cores=os.cpu_count(), mean_im_size=mean)))
