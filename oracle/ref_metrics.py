"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference's segmentation metrics (never imported by the product).

    get_IoU_binary            /root/reference/functions.py:1767-1787
    get_IoU_multi_unique      /root/reference/functions.py:1790-1815
    pixel_accuracy            /root/reference/functions.py:1819-1834
    dice_score_numpy_binary   /root/reference/functions.py:1837-1861

Pinned by tests/golden/metrics.npz (oracle/make_golden_next.py runs the reference's own functions).
"""
import numpy as np


def get_IoU_binary(gt, pred):
    mask_gt, mask_pred = np.array(gt), np.array(pred)
    intersection = np.logical_and(mask_gt, mask_pred).sum()
    union = np.logical_or(mask_gt, mask_pred).sum()
    return intersection / (union + 1e-7)


def get_IoU_multi_unique(pred, gt):
    unique_classes = np.unique(gt)
    iou_list = []
    for i in unique_classes:
        temp_gt = np.array(gt == i, dtype=np.float32)
        temp_pred = np.array(pred == i, dtype=np.float32)
        intersection = np.logical_and(temp_gt, temp_pred).sum()
        union = np.logical_or(temp_gt, temp_pred).sum()
        iou_list.append(intersection / (union + 1e-7))
    return sum(iou_list) / len(unique_classes)


def pixel_accuracy(pred_mask, gt_mask):
    return np.sum(pred_mask == gt_mask) / np.prod(gt_mask.shape)


def dice_score_numpy_binary(gt, pred, smooth=1, threshold=128):
    gt = (gt >= threshold).astype(np.float32)
    pred = (pred >= threshold).astype(np.float32)
    intersection = np.sum(gt * pred)
    union = np.sum(gt) + np.sum(pred)
    return (2 * intersection + smooth) / (union + smooth)
