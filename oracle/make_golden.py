#!/usr/bin/env python
"""Generate the golden fixtures under ``tests/golden/`` from the REFERENCE ITSELF.

Runs only in the authoring container (``/root/reference`` does not exist on the
GPU box).  ``/root/reference/functions.py`` cannot be imported (it imports
TensorFlow at line 10), so the functions on the hot path are pulled out of its
AST and executed unmodified under NumPy/cv2 with duck-typed models whose
``.predict`` replays stored probability maps.  No reference source is written
into this repository: only inputs and the outputs the reference computed.

    python oracle/make_golden.py            # rewrites tests/golden/*.npz

Fixtures (all seeded, small):
  im_kat.npz          the worked example of IM_creation.jpg (README.md:16-17)
  im_binary.npz       pred_masks_to_im_binary      functions.py:3104-3120
  im_multiclass.npz   pred_masks_to_im_multiclass  functions.py:3123-3137
  predict.npz         get_im_prediction_{binary,hela,multiclass}  :3140-3238
  dilate_mask.npz     dilate_mask                  functions.py:3075-3100
  drivers.npz         create_pseudo_labels_im_{ISIC_2018,hela,multiclass} :2832-3070
"""
import ast
import contextlib
import io
import os
import sys
import tempfile

import cv2
import numpy as np

REF = "/root/reference/functions.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

WANTED = {
    "pred_masks_to_im_binary", "pred_masks_to_im_multiclass",
    "get_im_prediction_binary", "get_im_prediction_hela", "get_im_prediction_multiclass",
    "dilate_mask", "get_pos_contours", "get_min_dist",
    "create_pseudo_labels_im_ISIC_2018", "create_pseudo_labels_im_hela",
    "create_pseudo_labels_im_multiclass",
}


def load_reference():
    """Exec the wanted top-level defs of functions.py in a namespace that
    provides what they use (np, cv2, os, io, contextlib, tqdm, THRESHOLD)."""
    src = open(REF).read()
    tree = ast.parse(src)
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in WANTED]
    missing = WANTED - {n.name for n in body}
    assert not missing, missing
    ns = {"np": np, "cv2": cv2, "os": os, "io": io, "contextlib": contextlib,
          "tqdm": lambda it, *a, **k: it,
          "THRESHOLD": 0.5}          # config.ini:13 -> functions.py:31
    exec(compile(ast.Module(body=body, type_ignores=[]), REF, "exec"), ns)
    return ns


class ReplayModel:
    """``.predict([uint8 NHWC])`` -> the stored float32 map for that image."""

    def __init__(self):
        self.table = {}

    def add(self, image, prob):
        self.table[np.ascontiguousarray(image).tobytes()] = prob

    def predict(self, x, *a, **k):
        if isinstance(x, (list, tuple)):
            x = x[0]
        return self.table[np.ascontiguousarray(x).tobytes()]


def tricky_probs(rng, shape, softmax):
    """Probabilities with exact-threshold values, ties and a few NaNs."""
    p = rng.random(shape, dtype=np.float32)
    if softmax:
        p = p / p.sum(axis=-1, keepdims=True)
    flat = p.reshape(-1, shape[-1])
    n = flat.shape[0]
    idx = rng.choice(n, size=max(4, n // 16), replace=False)
    q = len(idx) // 4
    flat[idx[:q]] = 0.5                                  # exactly on the threshold / all tied
    flat[idx[q:2 * q], -1] = flat[idx[q:2 * q]].max(axis=-1)  # tie between an early and the last class
    flat[idx[2 * q:3 * q], 0] = np.float32(np.nextafter(np.float32(0.5), np.float32(1)))
    flat[idx[3 * q:3 * q + 2], shape[-1] // 2] = np.nan  # NaN: false for >,>= ; the max for argmax
    return p


def main():
    ns = load_reference()
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20240607)

    # ------------------------------------------------------------------ KAT
    a = np.zeros((12, 12), np.int64)
    b = np.zeros((12, 12), np.int64)
    rows_a = {3: (4, 7), 4: (3, 7), 5: (3, 8), 6: (2, 8), 7: (2, 9), 8: (2, 9), 9: (3, 9), 10: (4, 7)}
    rows_b = {2: (4, 5), 3: (3, 7), 4: (2, 7), 5: (2, 8), 6: (2, 8), 7: (2, 9), 8: (2, 9), 9: (3, 9)}
    for r, (lo, hi) in rows_a.items():
        a[r, lo:hi + 1] = 1
    for r, (lo, hi) in rows_b.items():
        b[r, lo:hi + 1] = 1
    label, im, im_size, pred_size = ns["pred_masks_to_im_binary"]([a[..., None], b[..., None]])
    np.savez_compressed(os.path.join(OUT, "im_kat.npz"), a=a, b=b, label=label, im=im,
                        im_size=im_size, pred_size=pred_size)

    # ------------------------------------------------- pred_masks_to_im_binary
    store = {}
    cases = [(1, (8, 8, 1)), (2, (24, 40, 1)), (3, (17, 33, 1)), (4, (16, 16)), (5, (31, 7, 1))]
    for i, (m, shape) in enumerate(cases):
        masks = [(rng.random(shape) > 0.45).astype(int) for _ in range(m)]
        label, im, im_size, pred_size = ns["pred_masks_to_im_binary"](masks)
        store.update({f"{i}/masks": np.stack(masks), f"{i}/label": label, f"{i}/im": im,
                      f"{i}/im_size": im_size, f"{i}/pred_size": pred_size})
    # values other than 0/1 (the helper is generic over ints): pins the SUM semantics
    masks = [rng.integers(-1, 3, size=(12, 20, 1)) for _ in range(3)]
    label, im, im_size, pred_size = ns["pred_masks_to_im_binary"](masks)
    i = len(cases)
    store.update({f"{i}/masks": np.stack(masks), f"{i}/label": label, f"{i}/im": im,
                  f"{i}/im_size": im_size, f"{i}/pred_size": pred_size})
    store["n"] = np.int64(i + 1)
    np.savez_compressed(os.path.join(OUT, "im_binary.npz"), **store)

    # --------------------------------------------- pred_masks_to_im_multiclass
    store = {}
    cases = [(1, 9, (1, 8, 8)), (2, 9, (1, 24, 40)), (3, 35, (1, 13, 26)), (4, 3, (1, 16, 16)), (2, 256, (1, 9, 9))]
    for i, (m, k, shape) in enumerate(cases):
        base = rng.integers(0, k, size=shape)
        masks = []
        for _ in range(m):
            noise = rng.random(shape) < 0.2
            masks.append(np.where(noise, rng.integers(0, k, size=shape), base).astype(np.int64))
        label, im, im_size = ns["pred_masks_to_im_multiclass"](masks)
        store.update({f"{i}/masks": np.stack(masks), f"{i}/label": label, f"{i}/im": im, f"{i}/im_size": im_size})
    store["n"] = np.int64(len(cases))
    np.savez_compressed(os.path.join(OUT, "im_multiclass.npz"), **store)

    # ------------------------------------------------------ get_im_prediction_*
    store = {}
    h, w = 20, 28
    img = rng.integers(0, 256, size=(1, h, w, 3), dtype=np.uint8)
    for m in (1, 2, 3, 5):
        probs = [tricky_probs(rng, (1, h, w, 1), False) for _ in range(m)]
        models = []
        for p in probs:
            mod = ReplayModel(); mod.add(img, p); models.append(mod)
        for thr in (0.5, 0.3):
            label, im, im_size, pred_size = ns["get_im_prediction_binary"](models, img, thr)
            tag = f"binary/m{m}/t{thr}"
            store.update({f"{tag}/label": label, f"{tag}/im": im, f"{tag}/im_size": im_size, f"{tag}/pred_size": pred_size})
        store[f"binary/m{m}/probs"] = np.stack(probs)
    img1 = rng.integers(0, 256, size=(1, h, w, 1), dtype=np.uint8)
    for m in (1, 2, 4):
        probs = [tricky_probs(rng, (1, h, w, 3), False) for _ in range(m)]
        models = []
        for p in probs:
            mod = ReplayModel(); mod.add(img1, p); models.append(mod)
        alive, dead, pos, cim, im_size = ns["get_im_prediction_hela"](models, img1)
        tag = f"hela/m{m}"
        store.update({f"{tag}/probs": np.stack(probs), f"{tag}/alive": alive, f"{tag}/dead": dead,
                      f"{tag}/pos": pos, f"{tag}/im": cim, f"{tag}/im_size": im_size})
    for m, k in ((1, 9), (2, 9), (3, 35), (2, 35), (5, 2)):
        base = tricky_probs(rng, (1, h, w, k), True)
        probs = []
        for _ in range(m):
            jitter = rng.random((1, h, w, k), dtype=np.float32) * np.float32(0.15)
            probs.append((base + jitter).astype(np.float32))
        models = []
        for p in probs:
            mod = ReplayModel(); mod.add(img, p); models.append(mod)
        for flt in (False, True):
            label, im, im_size, eq = ns["get_im_prediction_multiclass"](models, img, flt)
            tag = f"multi/m{m}k{k}/f{int(flt)}"
            store.update({f"{tag}/label": label, f"{tag}/im": im, f"{tag}/im_size": im_size, f"{tag}/lists_equal": np.bool_(eq)})
        store[f"multi/m{m}k{k}/probs"] = np.stack(probs)
    np.savez_compressed(os.path.join(OUT, "predict.npz"), **store)

    # -------------------------------------------------------------- dilate_mask
    store = {}
    for i, k in enumerate((2, 9, 35)):
        lab = np.where(rng.random((24, 32)) < 0.8, 0, rng.integers(0, k, size=(24, 32))).astype(np.uint8)
        store[f"{i}/label"] = lab
        store[f"{i}/out"] = ns["dilate_mask"](lab)
    store["n"] = np.int64(3)
    np.savez_compressed(os.path.join(OUT, "dilate_mask.npz"), **store)

    # ------------------------------------------------------------------ drivers
    store = {}
    h, w = 32, 48
    names = [f"img_{i:02d}.png" for i in range(5)]

    def smooth_probs(shape, softmax, m):
        """Blobby maps so that erosion / dilation / contours have structure."""
        out = []
        base = cv2.GaussianBlur(rng.random(shape[:2]).astype(np.float32), (0, 0), 3.0)
        base = (base - base.min()) / (base.max() - base.min())
        for _ in range(m):
            chans = []
            for _k in range(shape[2]):
                n = cv2.GaussianBlur(rng.random(shape[:2]).astype(np.float32), (0, 0), 2.0)
                n = (n - n.min()) / (n.max() - n.min())
                chans.append(0.6 * base + 0.4 * n if not softmax else n + 0.5 * base * (_k % 3 == 0))
            p = np.stack(chans, axis=-1).astype(np.float32)
            if softmax:
                p = p / p.sum(axis=-1, keepdims=True)
            out.append(p[None])
        return out

    def run_driver(kind, c, k, m, kwargs_list):
        with tempfile.TemporaryDirectory() as tmp:
            src = os.path.join(tmp, "in")
            os.makedirs(src)
            models = [ReplayModel() for _ in range(m)]
            imgs, probs_all = [], []
            for name in names:
                img = rng.integers(0, 256, size=(h, w, c), dtype=np.uint8)
                cv2.imwrite(os.path.join(src, name), img if c == 3 else img[..., 0])
                disk = cv2.imread(os.path.join(src, name)) if c == 3 else cv2.imread(os.path.join(src, name), 0)
                fed = cv2.cvtColor(disk, cv2.COLOR_BGR2RGB) if c == 3 else disk
                fed = np.array(fed.reshape(-1, h, w, c), dtype=np.uint8)
                probs = smooth_probs((h, w, k), kind == "multiclass", m)
                for mod, p in zip(models, probs):
                    mod.add(fed, p if kind == "multiclass" else p)
                imgs.append(disk)
                probs_all.append(np.concatenate(probs, axis=0))
            store[f"{kind}/images"] = np.stack(imgs)
            store[f"{kind}/probs"] = np.stack(probs_all)        # [n_img, M, H, W, K]
            store[f"{kind}/names"] = np.array(names)
            for j, kw in enumerate(kwargs_list):
                dst = os.path.join(tmp, f"out{j}")
                fn = {"binary": "create_pseudo_labels_im_ISIC_2018", "hela": "create_pseudo_labels_im_hela",
                      "multiclass": "create_pseudo_labels_im_multiclass"}[kind]
                with contextlib.redirect_stdout(io.StringIO()):
                    mean = ns[fn](models, h, w, c, src, dst, **kw)
                store[f"{kind}/run{j}/mean_im_size"] = np.float64(mean)
                store[f"{kind}/run{j}/kwargs"] = np.array(repr(sorted(kw.items())))
                for sub in sorted(os.listdir(dst)):
                    for name in names:
                        path = os.path.join(dst, sub, name)
                        if os.path.exists(path):
                            store[f"{kind}/run{j}/{sub}/{name}"] = cv2.imread(path, cv2.IMREAD_UNCHANGED)
            store[f"{kind}/nruns"] = np.int64(len(kwargs_list))

    morph = [dict(erode_kernel=0, dilate_kernel=0), dict(erode_kernel=3, dilate_kernel=3),
             dict(erode_kernel=5, dilate_kernel=0), dict(erode_kernel=0, dilate_kernel=5),
             dict(erode_kernel=5, dilate_kernel=5)]
    run_driver("binary", 3, 1, 3,
               [dict(kw, filter_bad_predictions=True) for kw in morph]
               + [dict(erode_kernel=0, dilate_kernel=0, filter_bad_predictions=False),
                  dict(erode_kernel=0, dilate_kernel=3, block_input=False, block_output=True, filter_bad_predictions=False),
                  dict(erode_kernel=0, dilate_kernel=3, block_input=True, block_output=False, filter_bad_predictions=False)])
    run_driver("hela", 1, 3, 2,
               morph + [dict(erode_kernel=0, dilate_kernel=3, block_input=False, block_output=True),
                        dict(erode_kernel=0, dilate_kernel=3, block_input=True, block_output=False)])
    run_driver("multiclass", 3, 9, 2,
               morph + [dict(erode_kernel=0, dilate_kernel=0, filter_unequal_class_pred=True),
                        dict(erode_kernel=3, dilate_kernel=0, block_input=False, block_output=True),
                        dict(erode_kernel=0, dilate_kernel=3, block_input=True, block_output=False)])
    np.savez_compressed(os.path.join(OUT, "drivers.npz"), **store)

    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    sys.exit(main())
