"""IM++ augmentation drivers of the reference's ``functions.py``: the EvalNet ensemble scores every pseudo-labelled image
and decides how many augmented copies of it enter the next training set.

    create_augment_images_and_masks_with_evalnet_ensemble_binary       functions.py:5684-5759   (ISIC, `12_ISIC_2018_IM++.py`)
    create_augment_images_and_masks_with_evalnet_ensemble_multiclass   functions.py:5946-6054   (SUIM / Cityscapes)

Same names, argument order, defaults, file names (``<stem>___<j>.png``).  The reference predicts one image at a time per
EvalNet and augments on the host; here a directory is decoded by a thread pool, scored in device batches
(``B200EvalNet.predict`` or any object with a Keras-like ``predict([A, B])``) and augmented in device batches
(``augment.augment_batch``).  The score -> copies rule is the reference's arithmetic on the same NumPy types; the
augmentation decisions are drawn per (file, j) in the reference's loop order.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import cv2
import numpy as np

from .augment import augment_batch, draw_params, _geometry_only

__all__ = ["create_augment_images_and_masks_with_evalnet_ensemble_binary",
           "create_augment_images_and_masks_with_evalnet_ensemble_multiclass", "num_augs_from_score"]

_FILES_PER_BATCH = 512
_IO_THREADS = max(4, min(32, os.cpu_count() or 4))


def num_augs_from_score(score, min_threshold, max_threshold):
    """functions.py:5744-5753 / 6034-6044: 5 copies above max_threshold, 1 below min_threshold, linear steps in between."""
    threshold_step = (max_threshold - min_threshold) / 5
    if score > max_threshold:
        num_augs = 5
    elif score > min_threshold:
        num_augs = 1 + int((score - min_threshold) / threshold_step)
    else:
        num_augs = 1
    return min(num_augs, 5)


def _predict_batched(model, a, b):
    """Batched ``predict([A, B])``; a duck-typed model that only takes one image at a time is called per image."""
    try:
        return model.predict([a, b])
    except Exception:
        outs = [model.predict([a[i:i + 1], b[i:i + 1]]) for i in range(a.shape[0])]
        if isinstance(outs[0], (list, tuple)):
            return [np.concatenate([o[k] for o in outs], axis=0) for k in range(len(outs[0]))]
        return np.concatenate(outs, axis=0)


def _run(evalnets, h, w, c, main_input_path, main_output_path, rgb, score_batch, aug_kw):
    from . import functions as F
    images_in, masks_in = os.path.join(main_input_path, "images"), os.path.join(main_input_path, "masks")
    images_out, masks_out = os.path.join(main_output_path, "images"), os.path.join(main_output_path, "masks")
    os.makedirs(images_out, exist_ok=True)
    os.makedirs(masks_out, exist_ok=True)
    names = os.listdir(images_in)
    with ThreadPoolExecutor(_IO_THREADS) as io:
        for b0 in range(0, len(names), _FILES_PER_BATCH):
            bn = names[b0:b0 + _FILES_PER_BATCH]
            imgs = list(io.map(lambda nm: cv2.imread(os.path.join(images_in, nm)), bn))
            masks = list(io.map(lambda nm: cv2.imread(os.path.join(masks_in, nm), 0), bn))
            bgr = np.ascontiguousarray(np.stack(imgs))
            fed = np.ascontiguousarray(bgr[..., ::-1]) if (rgb and bgr.shape[-1] == 3) else bgr       # cv2.cvtColor(BGR2RGB)
            prepared = fed.reshape(-1, h, w, c).astype(np.uint8)
            mask_arr = np.ascontiguousarray(np.stack(masks))
            num = score_batch(evalnets, prepared, mask_arr)                  # copies per file
            # decisions per (file, j) in the reference's order, pixels per device batch
            rep = np.repeat(np.arange(len(bn)), num)
            ps = [draw_params(**aug_kw) for _ in rep]
            if len(rep) == 0:
                continue
            out_img, _ = augment_batch(F._dev(np.ascontiguousarray(bgr[rep])), None, ps)
            out_msk, _ = augment_batch(F._dev(np.ascontiguousarray(mask_arr[rep][..., None])), None, [_geometry_only(p) for p in ps])
            out_img, out_msk = out_img.cpu().numpy(), out_msk.cpu().numpy()[..., 0]
            js = np.concatenate([np.arange(k) for k in num]) if len(num) else np.zeros(0, int)

            def save(t):
                i, j = int(rep[t]), int(js[t])
                cv2.imwrite(os.path.join(images_out, f"{bn[i][:-4]}___{j}.png"), out_img[t])
                cv2.imwrite(os.path.join(masks_out, f"{bn[i][:-4]}___{j}.png"), out_msk[t])
            list(io.map(save, range(len(rep))))


def create_augment_images_and_masks_with_evalnet_ensemble_binary(evalnets, h, w, c, min_threshold, max_threshold, main_input_path,
                                                                 main_output_path, brightness_range_alpha=(0.6, 1.4),
                                                                 brightness_range_beta=(-20, 20), max_blur=3, max_noise=20,
                                                                 free_rotation=True, rgb=True):
    """functions.py:5684-5759."""
    def score_batch(models, prepared, mask_arr):
        prepared_mask = mask_arr.reshape(-1, h, w, 1)
        preds = np.stack([np.asarray(_predict_batched(m, prepared, prepared_mask)) for m in models], axis=0)      # [M, n, 1]
        mean_pred_ious = np.mean(preds, axis=0)                                                                      # functions.py:5740
        return [num_augs_from_score(mean_pred_ious[i].reshape(-1)[0], min_threshold, max_threshold) for i in range(prepared.shape[0])]   # np.float32 scalar

    _run(evalnets, h, w, c, main_input_path, main_output_path, rgb, score_batch,
         dict(brightness_range_alpha=brightness_range_alpha, brightness_range_beta=brightness_range_beta, max_blur=max_blur,
              max_noise=max_noise, free_rotation=free_rotation))


def create_augment_images_and_masks_with_evalnet_ensemble_multiclass(evalnets, h, w, c, num_classes, min_threshold, max_threshold,
                                                                     main_input_path, main_output_path, brightness_range_alpha=(0.6, 1.4),
                                                                     brightness_range_beta=(-20, 20), max_blur=3, max_noise=20,
                                                                     free_rotation=False, rgb=True):
    """functions.py:5946-6054."""
    def score_batch(models, prepared, mask_arr):
        mask_array = mask_arr.reshape(-1, h, w, 1)
        one_hot = np.stack([(mask_array == cls).astype(np.int32) for cls in range(num_classes)], axis=-1).squeeze(axis=-2)
        ious, dets = [], []
        for m in models:
            p = _predict_batched(m, prepared, one_hot)
            ious.append(np.asarray(p[0])); dets.append(np.asarray(p[1]))
        mean_ious = np.mean(np.stack(ious, axis=0), axis=0)                  # [n, K]   functions.py:6019-6020
        mean_det = np.mean(np.stack(dets, axis=0), axis=0)
        out = []
        for i in range(prepared.shape[0]):
            valid = [mean_ious[i, k] for k in range(mean_ious.shape[1]) if k > 0 and mean_det[i, k] >= 0.5]
            best_miou = (sum(valid) / len(valid)) if valid else 0.0
            out.append(num_augs_from_score(best_miou, min_threshold, max_threshold))
        return out

    _run(evalnets, h, w, c, main_input_path, main_output_path, rgb, score_batch,
         dict(brightness_range_alpha=brightness_range_alpha, brightness_range_beta=brightness_range_beta, max_blur=max_blur,
              max_noise=max_noise, free_rotation=free_rotation))
