"""Batched ``benchmark_*`` helpers of the reference's ``functions.py``.

    benchmark_ISIC2018      functions.py:1078-1151      mean IoU + mean Dice of a binary model
    benchmark_hela          functions.py:1156-1261      mean IoU (alive / dead / position) + cell-count error
    benchmark_multiclass    functions.py:1265-1339      mean pixel accuracy + mean IoU over the classes of the ground truth
    get_IoU_binary / get_IoU_multi_unique / pixel_accuracy / dice_score_numpy_binary   functions.py:1767-1861

Same names, argument order, defaults, printed lines, written files and returned numbers.  The reference runs these three
times per trained student on the whole pool; here a directory is decoded by a thread pool, predicted in device batches
through the model's CUDA engines (threshold / argmax fused into the forward pass: no probability map reaches the host),
and compared with the ground truth by the confusion-count kernels of ``imk_eval.cu``.  The kernels return exact integer
counts; the quotients below are the reference's own expressions on the same NumPy scalar types, so every per-image
value -- and therefore every mean -- is the reference's, bit for bit, given the same predictions.
Host geometry (cell positions from contours) stays cv2 on the host as in the reference (out-of-scope component #4).
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import cv2
import numpy as np

from ._lib import lib, check

__all__ = ["benchmark_ISIC2018", "benchmark_hela", "benchmark_multiclass", "get_IoU_binary", "get_IoU_multi_unique",
           "pixel_accuracy", "dice_score_numpy_binary", "mod_pos_size", "get_cell_count", "convert_class_to_color_mask",
           "seg_counts_binary", "seg_counts_multiclass"]

_IO_THREADS = max(4, min(32, os.cpu_count() or 4))


def _F():
    from . import functions
    return functions


# ----------------------------------------------------------------------------------------------- device counts
def seg_counts_binary(pred, gt):
    """uint8 [N,H,W] (or [H,W]) masks -> int64 [N,5]: |gt!=0 & pred!=0|, |gt!=0 | pred!=0|, |gt>=128 & pred>=128|,
    |gt>=128|, |pred>=128| (``imk_seg_counts_binary``).  Accepts NumPy arrays or CUDA tensors."""
    F = _F()
    torch = F._torch()
    p = pred if isinstance(pred, torch.Tensor) else F._dev(np.asarray(pred, dtype=np.uint8))
    g = gt if isinstance(gt, torch.Tensor) else F._dev(np.asarray(gt, dtype=np.uint8))
    if p.dim() == 2:
        p, g = p[None], g[None]
    if tuple(p.shape) != tuple(g.shape):
        raise ValueError(f"prediction {tuple(p.shape)} and ground truth {tuple(g.shape)} differ in shape")
    n, hw = p.shape[0], int(p[0].numel())
    out = torch.empty((n, 5), dtype=torch.int64, device="cuda")
    check(lib.imk_seg_counts_binary(p.contiguous().data_ptr(), g.contiguous().data_ptr(), n, hw, out.data_ptr(), F._stream()))
    return out.cpu().numpy()


def seg_counts_multiclass(pred, gt):
    """uint8 class-id maps -> int64 [N,3,256]: per value |gt==v|, |pred==v|, |gt==v & pred==v|."""
    F = _F()
    torch = F._torch()
    p = pred if isinstance(pred, torch.Tensor) else F._dev(np.asarray(pred, dtype=np.uint8))
    g = gt if isinstance(gt, torch.Tensor) else F._dev(np.asarray(gt, dtype=np.uint8))
    if p.dim() == 2:
        p, g = p[None], g[None]
    if tuple(p.shape) != tuple(g.shape):
        raise ValueError(f"prediction {tuple(p.shape)} and ground truth {tuple(g.shape)} differ in shape")
    n, hw = p.shape[0], int(p[0].numel())
    out = torch.empty((n, 3, 256), dtype=torch.int64, device="cuda")
    check(lib.imk_seg_counts_multiclass(p.contiguous().data_ptr(), g.contiguous().data_ptr(), n, hw, out.data_ptr(), F._stream()))
    return out.cpu().numpy()


# ------------------------------------------------------------------- the reference's quotients, from the counts
def _iou_binary(c):
    """functions.py:1781-1785: ``np.logical_and(..).sum() / (np.logical_or(..).sum() + 1e-7)`` (np.int64 / np.float64)."""
    return np.int64(c[0]) / (np.int64(c[1]) + 1e-7)


def _dice_binary(c, smooth=1):
    """functions.py:1852-1859 on float32 masks: the sums are float32 (exact: counts < 2^24), so is the quotient."""
    intersection, union = np.float32(c[2]), np.float32(c[3]) + np.float32(c[4])
    return (2 * intersection + smooth) / (union + smooth)


def _iou_multi_unique(h):
    """functions.py:1802-1815: classes present in the ground truth, union = |gt==i| + |pred==i| - |both|."""
    classes = np.nonzero(h[0])[0]
    iou_list = []
    for i in classes:
        intersection = np.int64(h[2, i])
        union = np.int64(h[0, i] + h[1, i] - h[2, i])
        iou_list.append(intersection / (union + 1e-7))
    return sum(iou_list) / len(classes)


def _pixel_accuracy(h, total):
    """functions.py:1831-1833: ``np.sum(pred == gt) / np.prod(gt.shape)``."""
    return np.int64(h[2].sum()) / np.int64(total)


def get_IoU_binary(gt, pred):
    """functions.py:1767-1787."""
    return _iou_binary(seg_counts_binary(np.asarray(pred), np.asarray(gt))[0])


def dice_score_numpy_binary(gt, pred, smooth=1, threshold=128):
    """functions.py:1837-1861 (the kernel binarises at the reference's default 128)."""
    if threshold != 128:
        gt, pred = (np.asarray(gt) >= threshold).astype(np.uint8) * 255, (np.asarray(pred) >= threshold).astype(np.uint8) * 255
    return _dice_binary(seg_counts_binary(np.asarray(pred), np.asarray(gt))[0], smooth)


def get_IoU_multi_unique(pred, gt):
    """functions.py:1790-1815."""
    return _iou_multi_unique(seg_counts_multiclass(np.asarray(pred), np.asarray(gt))[0])


def pixel_accuracy(pred_mask, gt_mask):
    """functions.py:1819-1834."""
    g = np.asarray(gt_mask)
    return _pixel_accuracy(seg_counts_multiclass(np.asarray(pred_mask), g)[0], g.size)


# ------------------------------------------------------------------------------ host geometry (component #4)
def mod_pos_size(gray_img, max_pos_circle_size=8, min_pos_circle_size=3):
    """functions.py:6256-6294."""
    F = _F()
    positions = F.get_pos_contours(gray_img)
    h, w = gray_img.shape
    out_img = np.zeros((h, w), np.uint8)
    for pos in positions:
        try:
            circle_size = int(F.get_min_dist(pos, positions) // 4)
            circle_size = max(min(circle_size, max_pos_circle_size), min_pos_circle_size)
            cv2.circle(out_img, (pos[0], pos[1]), circle_size, (255), -1)
        except Exception as e:                        # the reference prints and continues
            print(e)
    out_img = cv2.blur(out_img, (2, 2))
    out_img[out_img < 254] = 0
    return out_img


def get_cell_count(positions, img_alive, img_dead, measuring_range=3):
    """functions.py:6299-6370."""
    def gray(img):
        return cv2.cvtColor(img, cv2.COLOR_BGR2GRAY) if (img.ndim == 3 and img.shape[2] > 1) else img
    img_h, img_w = img_dead.shape[:2]
    _, bin_alive = cv2.threshold(gray(img_alive), 10, 255, cv2.THRESH_BINARY)
    _, bin_dead = cv2.threshold(gray(img_dead), 10, 255, cv2.THRESH_BINARY)
    alive_count = dead_count = unclear_count = 0
    r = measuring_range
    for pos in positions:
        x, y = pos[0], pos[1]
        if x - r <= 0:
            x += r
        if x + r > img_w:
            x = img_w - r
        if y - r < 0:
            y += r
        if y + r > img_h:
            y = img_h - r
        a, d = np.sum(bin_alive[y - r:y + r, x - r:x + r]), np.sum(bin_dead[y - r:y + r, x - r:x + r])
        alive_count += int(a > d)
        dead_count += int(d > a)
        unclear_count += int(d == a)
    return alive_count, dead_count, unclear_count


def convert_class_to_color_mask(class_mask, output_path, class_to_color_mapping):
    """functions.py:6127-6149."""
    color_mask = np.zeros(list(class_mask.shape) + [3], dtype=np.uint8)
    for color, class_value in class_to_color_mapping.items():
        color_mask[class_mask == class_value] = color
    cv2.imwrite(output_path, cv2.cvtColor(color_mask, cv2.COLOR_RGB2BGR))


# --------------------------------------------------------------------------------------------- batched drivers
def _read_all(paths, flags):
    with ThreadPoolExecutor(_IO_THREADS) as io:
        return list(io.map(lambda p: cv2.imread(p, *flags), paths))


def _predict_masks(model, images, kind, threshold, strict, swap_rb):
    """One device batch: decisions of ONE model as uint8 planes [planes][n,H,W] kept on the device + a host copy."""
    F = _F()
    torch = F._torch()
    n, h, w, c = images.shape
    planes = 3 if kind == "hela" else 1
    d_img = F._dev(images)
    d_labels = torch.empty((planes, n, h, w), dtype=torch.uint8, device="cuda")
    d_im = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
    d_sz = torch.empty(n, dtype=torch.int64, device="cuda")
    s = F._stream()
    if F._all_b200([model]):
        hs = F._handles([model])
        if kind == "multiclass":
            check(lib.imk_ensemble_im_multiclass(hs, 1, d_img.data_ptr(), n, int(swap_rb), 0, 0, None, d_labels.data_ptr(),
                                                 d_im.data_ptr(), d_sz.data_ptr(), None, s))
        else:
            d_pred = torch.empty((planes, n), dtype=torch.int64, device="cuda")
            check(lib.imk_ensemble_im_binary(hs, 1, d_img.data_ptr(), n, int(swap_rb), float(threshold), int(strict), 0, 0, None,
                                             d_labels.data_ptr(), d_im.data_ptr(), d_sz.data_ptr(), d_pred.data_ptr(), s))
    else:
        # duck-typed model: its own batched .predict, decisions by the stand-alone IM kernels with M = 1
        import ctypes as C
        fed = np.ascontiguousarray(images[..., ::-1]) if swap_rb else images
        probs = F._dev(np.ascontiguousarray(np.asarray(model.predict(fed), dtype=np.float32)))
        k = probs.shape[-1]
        ptrs = (C.c_void_p * 1)(probs.data_ptr())
        if kind == "multiclass":
            check(lib.imk_im_multiclass(ptrs, 1, n, h, w, k, d_img.data_ptr(), c, 0, 0, None, d_labels.data_ptr(), d_im.data_ptr(),
                                        d_sz.data_ptr(), None, s))
        else:
            d_pred = torch.empty((planes, n), dtype=torch.int64, device="cuda")
            check(lib.imk_im_binary(ptrs, 1, n, h, w, k, float(threshold), int(strict), d_img.data_ptr(), c, 0, 0, None,
                                    d_labels.data_ptr(), d_im.data_ptr(), d_sz.data_ptr(), d_pred.data_ptr(), s))
    return d_labels


def benchmark_ISIC2018(model, images_dir, masks_dir, pred_path, h, w, c, batch_size=64, create_images=True, print_results=False):
    """functions.py:1078-1151 -> ``(mIoU, mdice_score)``."""
    F = _F()
    ious, dice_scores = [], []
    os.makedirs(pred_path, exist_ok=True)
    names = os.listdir(images_dir)
    bs = max(int(batch_size), 1) * 8                     # device batches: the reference's batch is a host-loop detail
    with ThreadPoolExecutor(_IO_THREADS) as io:
        for b0 in range(0, len(names), bs):
            bn = names[b0:b0 + bs]
            imgs = _read_all([os.path.join(images_dir, nm) for nm in bn], ())
            gts = _read_all([os.path.join(masks_dir, nm) for nm in bn], (0,))
            batch = np.ascontiguousarray(np.stack([im.reshape(h, w, c) for im in imgs]).astype(np.uint8))
            gt = np.ascontiguousarray(np.stack(gts).astype(np.uint8))
            d_labels = _predict_masks(model, batch, "binary", 0.5, 1, c == 3)      # cvtColor(BGR2RGB), pred > 0.5
            counts = seg_counts_binary(d_labels[0], F._dev(gt))
            if create_images:
                pred = d_labels[0].cpu().numpy()
                list(io.map(lambda i: cv2.imwrite(os.path.join(pred_path, f"{bn[i]}"), pred[i]), range(len(bn))))
            for i, nm in enumerate(bn):
                dice_score = round(_dice_binary(counts[i]), 4)
                dice_scores.append(dice_score)
                iou = round(_iou_binary(counts[i]), 4)
                ious.append(iou)
                if print_results:
                    print(f"{nm} IoU: {iou}    DS: {dice_score}")
    mIoU = round((np.sum(ious) / len(ious)), 3)
    mdice_score = round((np.sum(dice_scores) / len(dice_scores)), 3)
    print(f"------------------------------------------------------------  mIoU: {mIoU}    mdice score: {mdice_score}  ------------------------------------------------------------")
    return mIoU, mdice_score


def benchmark_hela(model, gt_main_dir, pred_dir, h, w, c, threshold=0.5, batch_size=64, save_output=True, benchmark=True, mod_position=True):
    """functions.py:1156-1261 -> ``(mIoU, mIoU_ad, mean_cell_count_error)``."""
    F = _F()
    mIoUs, mIoUs_ad = [], []
    cell_count_delta = 0
    pos_dir = "mod_position" if mod_position else "position"
    for d in ("alive", "dead", pos_dir):
        os.makedirs(os.path.join(pred_dir, d), exist_ok=True)
    image_names = os.listdir(os.path.join(gt_main_dir, "brightfield"))
    bs = max(int(batch_size), 1) * 8
    with ThreadPoolExecutor(_IO_THREADS) as io:
        for b0 in range(0, len(image_names), bs):
            bn = image_names[b0:b0 + bs]
            rd = lambda sub: _read_all([os.path.join(gt_main_dir, sub, nm) for nm in bn], (0,))
            imgs = rd("brightfield")
            batch = np.ascontiguousarray(np.stack([im.reshape(h, w, c) for im in imgs]).astype(np.uint8))
            d_labels = _predict_masks(model, batch, "hela", threshold, 1, False)    # (x > threshold) * 255 per head
            gt_alive, gt_dead, gt_pos = (np.ascontiguousarray(np.stack(rd(s))) for s in ("alive", "dead", "mod_position")) \
                if benchmark else (None, None, None)
            lab = d_labels.cpu().numpy()
            pos_masks = list(io.map(mod_pos_size, lab[2])) if mod_position else list(lab[2])
            if benchmark:
                c_alive = seg_counts_binary(d_labels[0], F._dev(gt_alive))
                c_dead = seg_counts_binary(d_labels[1], F._dev(gt_dead))
                c_pos = seg_counts_binary(np.stack(pos_masks), gt_pos)
            for i, nm in enumerate(bn):
                alive_uint, dead_uint, pos_uint = lab[0, i], lab[1, i], pos_masks[i]
                if benchmark:
                    iou_alive, iou_dead, iou_pos = (round(_iou_binary(cc[i]), 4) for cc in (c_alive, c_dead, c_pos))
                    mIoUs.append((iou_alive + iou_dead + iou_pos) / 3)
                    mIoUs_ad.append((iou_alive + iou_dead) / 2)
                    pred_alive_count, pred_dead_count, _ = get_cell_count(F.get_pos_contours(pos_uint), alive_uint, dead_uint)
                    gt_alive_count, gt_dead_count, _ = get_cell_count(F.get_pos_contours(gt_pos[i]), gt_alive[i], gt_dead[i])
                    cell_count_delta += abs(pred_alive_count - gt_alive_count) + abs(pred_dead_count - gt_dead_count)
                if save_output:
                    cv2.imwrite(os.path.join(pred_dir, "alive", nm), alive_uint)
                    cv2.imwrite(os.path.join(pred_dir, "dead", nm), dead_uint)
                    cv2.imwrite(os.path.join(pred_dir, pos_dir, nm), pos_uint)
    mIoU = round(np.sum(mIoUs) / len(mIoUs), 3)
    mIoU_ad = round(np.sum(mIoUs_ad) / len(mIoUs_ad), 3)
    mean_cell_count_error = round(cell_count_delta / len(mIoUs), 3)
    return mIoU, mIoU_ad, mean_cell_count_error


def benchmark_multiclass(model, image_path, gt_path, pred_path, h, w, c, class_to_color_mapping, batch_size=64, create_images=True, print_results=True):
    """functions.py:1265-1339 -> ``(mPA, mIoU)``."""
    F = _F()
    ious, PAs = [], []
    os.makedirs(pred_path, exist_ok=True)
    names = os.listdir(image_path)
    bs = max(int(batch_size), 1) * 8
    with ThreadPoolExecutor(_IO_THREADS) as io:
        for b0 in range(0, len(names), bs):
            bn = names[b0:b0 + bs]
            imgs = _read_all([os.path.join(image_path, nm) for nm in bn], ())
            gts = _read_all([os.path.join(gt_path, nm) for nm in bn], (0,))
            batch = np.ascontiguousarray(np.stack([im.reshape(h, w, c) for im in imgs]).astype(np.uint8))
            gt = np.ascontiguousarray(np.stack(gts).astype(np.uint8))
            d_labels = _predict_masks(model, batch, "multiclass", 0.0, 0, c == 3)   # np.argmax(pred, axis=-1)
            hist = seg_counts_multiclass(d_labels[0], F._dev(gt))
            if create_images:
                pred = d_labels[0].cpu().numpy()

                def save(i):
                    cv2.imwrite(os.path.join(pred_path, f"{bn[i]}"), pred[i])
                    convert_class_to_color_mask(pred[i], os.path.join(pred_path, f"{bn[i][:-4]}_color.png"), class_to_color_mapping)
                list(io.map(save, range(len(bn))))
            for i, nm in enumerate(bn):
                pa = round(_pixel_accuracy(hist[i], h * w), 4)
                PAs.append(pa)
                iou = round(_iou_multi_unique(hist[i]), 4)
                ious.append(iou)
                if print_results:
                    print(f"{nm} IoU: {iou}    PA: {pa}")
    mPA = round((np.sum(PAs) / len(PAs)), 3)
    mIoU = round((np.sum(ious) / len(ious)), 3)
    print(f"------------------------------------------------------------   mPA: {mPA}      mIoU: {mIoU}  ------------------------------------------------------------")
    return mPA, mIoU
