#!/usr/bin/env python
"""Golden fixtures for the SURVEY.md 8f rows, generated from the REFERENCE ITSELF (authoring container only).

Same method as make_golden.py: the functions are pulled out of /root/reference/functions.py's AST and executed
unmodified under NumPy / cv2; only inputs and the outputs the reference computed are stored.

    python oracle/make_golden_next.py         # rewrites tests/golden/{metrics,augment,benchmarks}.npz

  metrics.npz      get_IoU_binary, dice_score_numpy_binary, get_IoU_multi_unique, pixel_accuracy   functions.py:1767-1861
  augment.npz      augment_image_and_mask / augment_image_and_masks with seeded random + np.random   functions.py:2725-2828
                   (max_noise = 0: the deterministic part; the noise has distributional parity only)
  benchmarks.npz   benchmark_ISIC2018 / benchmark_hela / benchmark_multiclass on synthetic directories with replayed
                   probability maps                                                                 functions.py:1078-1339
  impp.npz         create_augment_images_and_masks_with_evalnet_ensemble_binary / _multiclass with replayed EvalNet scores:
                   copies written per file                                                           functions.py:5684-6054
"""
import ast
import contextlib
import io
import os
import random
import sys
import tempfile

import cv2
import numpy as np

REF = "/root/reference/functions.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

WANTED = {
    "get_IoU_binary", "dice_score_numpy_binary", "get_IoU_multi_unique", "pixel_accuracy",
    "augment_image_and_mask", "augment_image_and_masks", "add_noise", "add_noise_and_blur",
    "benchmark_ISIC2018", "benchmark_hela", "benchmark_multiclass",
    "get_pos_contours", "get_min_dist", "mod_pos_size", "get_cell_count", "convert_class_to_color_mask",
    "create_augment_images_and_masks_ISIC_2018", "create_augment_images_and_masks_hela", "create_augment_images_and_masks_multiclass",
    "create_augment_images_and_masks_with_evalnet_ensemble_binary", "create_augment_images_and_masks_with_evalnet_ensemble_multiclass",
}


def load_reference():
    tree = ast.parse(open(REF).read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in WANTED]
    missing = WANTED - {n.name for n in body}
    assert not missing, missing
    import shutil
    ns = {"np": np, "cv2": cv2, "os": os, "io": io, "contextlib": contextlib, "random": random, "shutil": shutil,
          "tqdm": lambda it, *a, **k: it}
    exec(compile(ast.Module(body=body, type_ignores=[]), REF, "exec"), ns)
    return ns


class BatchReplay:
    """``.predict(uint8 / float32 [n,H,W,c])`` -> the stored float32 maps of those images, stacked."""

    def __init__(self):
        self.table = {}

    def add(self, image, prob):
        self.table[np.ascontiguousarray(image).astype(np.uint8).tobytes()] = prob

    def predict(self, x, *a, **k):
        x = np.asarray(x)
        return np.stack([self.table[np.ascontiguousarray(x[i]).astype(np.uint8).tobytes()] for i in range(x.shape[0])])


def blobs(rng, h, w, n, r=(3, 7)):
    m = np.zeros((h, w), np.uint8)
    for _ in range(n):
        cv2.circle(m, (int(rng.integers(8, w - 8)), int(rng.integers(8, h - 8))), int(rng.integers(*r)), 255, -1)
    return m


def main():
    ns = load_reference()
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(2024)

    # ---- metrics -------------------------------------------------------------------------------------------
    store = {}
    shapes = [(16, 16), (32, 48), (64, 64), (256, 256)]
    for i, (h, w) in enumerate(shapes):
        gt = (rng.random((h, w)) > 0.6).astype(np.uint8) * 255
        pred = (rng.random((h, w)) > 0.5).astype(np.uint8) * 255
        if i == 1:                                     # grey values around both binarisation rules (non-zero / >= 128)
            gt = rng.choice(np.array([0, 1, 127, 128, 200, 255], np.uint8), size=(h, w))
            pred = rng.choice(np.array([0, 255, 128, 127], np.uint8), size=(h, w))
        if i == 2:
            gt[:] = 0                                  # empty ground truth: union may be 0 -> the +1e-7 / smooth terms
            pred[:, :3] = 255
        store[f"b{i}_gt"], store[f"b{i}_pred"] = gt, pred
        iou, dice = ns["get_IoU_binary"](gt, pred), ns["dice_score_numpy_binary"](gt, pred)
        store[f"b{i}_iou"], store[f"b{i}_dice"] = np.float64(iou), np.asarray(dice)
        store[f"b{i}_iou4"], store[f"b{i}_dice4"] = np.float64(round(iou, 4)), np.asarray(round(dice, 4))
        k = [3, 9, 35, 200][i]
        gtm = rng.integers(0, k, size=(h, w)).astype(np.uint8)
        predm = np.where(rng.random((h, w)) > 0.3, gtm, rng.integers(0, k, size=(h, w))).astype(np.uint8)
        store[f"m{i}_gt"], store[f"m{i}_pred"] = gtm, predm
        miou, pa = ns["get_IoU_multi_unique"](predm, gtm), ns["pixel_accuracy"](predm, gtm)
        store[f"m{i}_iou"], store[f"m{i}_pa"] = np.float64(miou), np.float64(pa)
        store[f"m{i}_iou4"], store[f"m{i}_pa4"] = np.float64(round(miou, 4)), np.float64(round(pa, 4))
    store["n"] = np.int64(len(shapes))
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), **store)

    # ---- augmentation (deterministic part) ----------------------------------------------------------------
    store = {}
    cases = []
    for seed in range(24):
        square = seed % 3 != 2
        h, w = (48, 48) if square else (32, 64)
        c = 3 if seed % 2 == 0 else 1
        image = rng.integers(0, 256, size=(h, w, c) if c == 3 else (h, w), dtype=np.uint8)
        mask = rng.integers(0, 5, size=(h, w), dtype=np.uint8)
        mask2 = (rng.random((h, w)) > 0.5).astype(np.uint8) * 255
        random.seed(1000 + seed)
        np.random.seed(2000 + seed)
        if seed % 4 == 3:
            out, masks = ns["augment_image_and_masks"](image.copy(), [mask.copy(), mask2.copy()], max_noise=0, free_rotation=square)
            store[f"a{seed}_mask_out2"] = masks[1]
            mask_out = masks[0]
        else:
            out, mask_out = ns["augment_image_and_mask"](image.copy(), mask.copy(), max_noise=0, free_rotation=square)
        store[f"a{seed}_image"], store[f"a{seed}_mask"], store[f"a{seed}_mask2"] = image, mask, mask2
        store[f"a{seed}_out"], store[f"a{seed}_mask_out"] = out, mask_out
        cases.append((seed, int(square), int(seed % 4 == 3)))
    store["cases"] = np.array(cases, np.int64)
    # the per-directory drivers: ONE file per directory (os.listdir order must not matter), several augmentations of it
    with tempfile.TemporaryDirectory() as tmp:
        img = rng.integers(0, 256, size=(40, 40, 3), dtype=np.uint8)
        msk = (rng.random((40, 40)) > 0.5).astype(np.uint8) * 255
        for d, m in (("i", img), ("m", msk)):
            os.makedirs(os.path.join(tmp, d)); cv2.imwrite(os.path.join(tmp, d, "one.png"), m)
        random.seed(77); np.random.seed(78)
        ns["create_augment_images_and_masks_ISIC_2018"](os.path.join(tmp, "i"), os.path.join(tmp, "m"), os.path.join(tmp, "o"), 5, True, max_noise=0)
        store["drv_isic_image"], store["drv_isic_mask"] = img, msk
        for n in range(5):
            store[f"drv_isic_out_{n}"] = cv2.imread(os.path.join(tmp, "o", "images", f"one_aug_{n}.png"))
            store[f"drv_isic_mask_out_{n}"] = cv2.imread(os.path.join(tmp, "o", "masks", f"one_aug_{n}.png"))
        store["drv_isic_copy"] = cv2.imread(os.path.join(tmp, "o", "images", "one.png"))
        bf = rng.integers(0, 256, size=(32, 32), dtype=np.uint8)
        hm = [(rng.random((32, 32)) > 0.6).astype(np.uint8) * 255 for _ in range(3)]
        for d, m in zip(("brightfield", "alive", "dead", "mod_position"), [bf] + hm):
            os.makedirs(os.path.join(tmp, "h", d)); cv2.imwrite(os.path.join(tmp, "h", d, "c.png"), m)
        random.seed(79); np.random.seed(80)
        ns["create_augment_images_and_masks_hela"](os.path.join(tmp, "h"), os.path.join(tmp, "ho"), 4, False, max_noise=0)
        store["drv_hela_bf"] = bf
        for j, m in enumerate(hm):
            store[f"drv_hela_m{j}"] = m
        for n in range(4):
            for d in ("brightfield", "alive", "dead", "mod_position"):
                store[f"drv_hela_{d}_{n}"] = cv2.imread(os.path.join(tmp, "ho", d, f"c_aug_{n}.png"))
    np.savez_compressed(os.path.join(OUT, "augment.npz"), **store)

    # ---- benchmark_* drivers -------------------------------------------------------------------------------
    store = {}
    h = w = 64
    with tempfile.TemporaryDirectory() as tmp, contextlib.redirect_stdout(io.StringIO()):
        # ISIC: images/ + masks/
        n = 11
        idir, mdir, pdir = (os.path.join(tmp, "isic", d) for d in ("images", "masks", "pred"))
        os.makedirs(idir); os.makedirs(mdir)
        model = BatchReplay()
        imgs, gts, probs = [], [], []
        for i in range(n):
            img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
            gt = blobs(rng, h, w, 4, (5, 12))
            prob = np.clip(gt[..., None] / 255.0 * 0.6 + rng.random((h, w, 1)) * 0.5, 0, 1).astype(np.float32)
            prob.reshape(-1)[rng.choice(h * w, 8, replace=False)] = 0.5          # exact-threshold values: strict >
            cv2.imwrite(os.path.join(idir, f"im_{i:03d}.png"), img)
            cv2.imwrite(os.path.join(mdir, f"im_{i:03d}.png"), gt)
            model.add(cv2.cvtColor(img, cv2.COLOR_BGR2RGB), prob)
            imgs.append(img); gts.append(gt); probs.append(prob)
        miou, mdice = ns["benchmark_ISIC2018"](model, idir, mdir, pdir, h, w, 3, batch_size=4)
        store["isic_images"], store["isic_gt"], store["isic_probs"] = np.stack(imgs), np.stack(gts), np.stack(probs)
        store["isic_result"] = np.array([miou, mdice], np.float64)
        store["isic_pred"] = np.stack([cv2.imread(os.path.join(pdir, f"im_{i:03d}.png"), 0) for i in range(n)])

        # multiclass: images/ + masks/ with class ids
        k = 6
        idir, mdir, pdir = (os.path.join(tmp, "mc", d) for d in ("images", "masks", "pred"))
        os.makedirs(idir); os.makedirs(mdir)
        model = BatchReplay()
        imgs, gts, probs = [], [], []
        mapping = {(int(40 * j), int(255 - 30 * j), int(17 * j)): j for j in range(k)}
        for i in range(n):
            img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
            gt = rng.integers(0, k if i % 3 else 3, size=(h // 8, w // 8)).astype(np.uint8).repeat(8, 0).repeat(8, 1)
            logits = rng.random((h, w, k)).astype(np.float32)
            logits[np.arange(h)[:, None], np.arange(w)[None, :], gt] += 0.4
            prob = (logits / logits.sum(-1, keepdims=True)).astype(np.float32)
            cv2.imwrite(os.path.join(idir, f"im_{i:03d}.png"), img)
            cv2.imwrite(os.path.join(mdir, f"im_{i:03d}.png"), gt)
            model.add(cv2.cvtColor(img, cv2.COLOR_BGR2RGB), prob)
            imgs.append(img); gts.append(gt); probs.append(prob)
        mpa, miou = ns["benchmark_multiclass"](model, idir, mdir, pdir, h, w, 3, mapping, batch_size=4, print_results=False)
        store["mc_images"], store["mc_gt"], store["mc_probs"] = np.stack(imgs), np.stack(gts), np.stack(probs)
        store["mc_result"] = np.array([mpa, miou], np.float64)
        store["mc_pred"] = np.stack([cv2.imread(os.path.join(pdir, f"im_{i:03d}.png"), 0) for i in range(n)])
        store["mc_color"] = np.stack([cv2.imread(os.path.join(pdir, f"im_{i:03d}_color.png")) for i in range(n)])
        store["mc_mapping"] = np.array([list(c) + [v] for c, v in mapping.items()], np.int64)

        # HeLa: brightfield/ alive/ dead/ mod_position/
        root, pdir = os.path.join(tmp, "hela"), os.path.join(tmp, "hela_pred")
        for d in ("brightfield", "alive", "dead", "mod_position"):
            os.makedirs(os.path.join(root, d))
        model = BatchReplay()
        imgs, gta, gtd, gtp, probs = [], [], [], [], []
        for i in range(7):
            img = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
            alive, dead, pos = blobs(rng, h, w, 3, (5, 9)), blobs(rng, h, w, 2, (4, 8)), blobs(rng, h, w, 5, (2, 4))
            prob = np.stack([np.clip(m / 255.0 * 0.7 + rng.random((h, w)) * 0.4, 0, 1) for m in (alive, dead, pos)], -1).astype(np.float32)
            for d, m in (("brightfield", img), ("alive", alive), ("dead", dead), ("mod_position", pos)):
                cv2.imwrite(os.path.join(root, d, f"c_{i:03d}.png"), m)
            model.add(img.reshape(h, w, 1), prob)
            imgs.append(img); gta.append(alive); gtd.append(dead); gtp.append(pos); probs.append(prob)
        res = ns["benchmark_hela"](model, root, pdir, h, w, 1, batch_size=3)
        store["hela_images"], store["hela_probs"] = np.stack(imgs), np.stack(probs)
        store["hela_gt_alive"], store["hela_gt_dead"], store["hela_gt_pos"] = np.stack(gta), np.stack(gtd), np.stack(gtp)
        store["hela_result"] = np.array(res, np.float64)
        for d in ("alive", "dead", "mod_position"):
            store[f"hela_pred_{d}"] = np.stack([cv2.imread(os.path.join(pdir, d, f"c_{i:03d}.png"), 0) for i in range(7)])
    np.savez_compressed(os.path.join(OUT, "benchmarks.npz"), **store)

    # ---- IM++ drivers: EvalNet scores -> number of augmented copies (functions.py:5684-5759, 5946-6054) ------------------
    # replayed EvalNet outputs; augmentation reduced to the horizontal flip (no blur / noise / brightness change, no free
    # rotation) so that the outputs do not depend on the order in which os.listdir hands the files to the RNG
    class ScoreNet:
        def __init__(self):
            self.table = {}

        def add(self, image, out):
            self.table[np.ascontiguousarray(image).astype(np.uint8).tobytes()] = out

        def predict(self, x, *a, **k):
            a0 = np.asarray(x[0])
            outs = [self.table[np.ascontiguousarray(a0[i]).astype(np.uint8).tobytes()] for i in range(a0.shape[0])]
            if isinstance(outs[0], tuple):
                return [np.stack([o[0] for o in outs]), np.stack([o[1] for o in outs])]
            return np.stack(outs)

    store = {}
    h = w = 32
    rng2 = np.random.default_rng(4242)
    with tempfile.TemporaryDirectory() as tmp, contextlib.redirect_stdout(io.StringIO()):
        n = 9
        os.makedirs(os.path.join(tmp, "b", "images")); os.makedirs(os.path.join(tmp, "b", "masks"))
        nets = [ScoreNet(), ScoreNet()]
        imgs, msks, scores = [], [], []
        for i in range(n):
            img = rng2.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
            msk = (rng2.random((h, w)) > 0.5).astype(np.uint8) * 255
            sc = np.array([[0.05 + 0.11 * i]], np.float32)[0]              # spans below min .. above max
            cv2.imwrite(os.path.join(tmp, "b", "images", f"p{i}.png"), img); cv2.imwrite(os.path.join(tmp, "b", "masks", f"p{i}.png"), msk)
            for k, net in enumerate(nets):
                net.add(cv2.cvtColor(img, cv2.COLOR_BGR2RGB), sc + np.float32(0.02 * k))
            imgs.append(img); msks.append(msk); scores.append(sc)
        ns["create_augment_images_and_masks_with_evalnet_ensemble_binary"](nets, h, w, 3, 0.3, 0.8, os.path.join(tmp, "b"), os.path.join(tmp, "bo"),
                                                                           (1.0, 1.0), (0.0, 0.0), 0, 0, False, True)
        store["b_images"], store["b_masks"], store["b_scores"] = np.stack(imgs), np.stack(msks), np.stack(scores)
        store["b_counts"] = np.array([len([f for f in os.listdir(os.path.join(tmp, "bo", "images")) if f.startswith(f"p{i}___")]) for i in range(n)], np.int64)
        k = 5
        os.makedirs(os.path.join(tmp, "m", "images")); os.makedirs(os.path.join(tmp, "m", "masks"))
        nets = [ScoreNet(), ScoreNet(), ScoreNet()]
        imgs, msks, ious, dets = [], [], [], []
        for i in range(n):
            img = rng2.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
            msk = rng2.integers(0, k, size=(h, w), dtype=np.uint8)
            iou = rng2.random(k).astype(np.float32) * (0.2 + 0.1 * i)
            det = (rng2.random(k) > 0.4).astype(np.float32) * 0.9
            if i == 3:
                det[:] = 0.1                                               # no class passes the detection check -> mIoU 0
            cv2.imwrite(os.path.join(tmp, "m", "images", f"q{i}.png"), img); cv2.imwrite(os.path.join(tmp, "m", "masks", f"q{i}.png"), msk)
            for j, net in enumerate(nets):
                net.add(cv2.cvtColor(img, cv2.COLOR_BGR2RGB), (iou + np.float32(0.01 * j), det))
            imgs.append(img); msks.append(msk); ious.append(iou); dets.append(det)
        ns["create_augment_images_and_masks_with_evalnet_ensemble_multiclass"](nets, h, w, 3, k, 0.2, 0.6, os.path.join(tmp, "m"), os.path.join(tmp, "mo"),
                                                                               (1.0, 1.0), (0.0, 0.0), 0, 0, False, True)
        store["m_images"], store["m_masks"], store["m_ious"], store["m_dets"] = np.stack(imgs), np.stack(msks), np.stack(ious), np.stack(dets)
        store["m_counts"] = np.array([len([f for f in os.listdir(os.path.join(tmp, "mo", "images")) if f.startswith(f"q{i}___")]) for i in range(n)], np.int64)
    np.savez_compressed(os.path.join(OUT, "impp.npz"), **store)
    print("wrote", [f for f in sorted(os.listdir(OUT)) if f in ("metrics.npz", "augment.npz", "benchmarks.npz", "impp.npz")])


if __name__ == "__main__":
    sys.exit(main())
