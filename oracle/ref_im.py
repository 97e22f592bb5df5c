"""CPU oracle for the Inconsistency-Mask (IM) arithmetic of the pseudo-label hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``inconsistencymasks_b200/`` may import
this module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline``
/ ``--impl reference`` legs of ``bench.py`` do, and only as the checker.

This is a NumPy *restatement* (written from scratch, not copied) of the
reference's post-``predict`` arithmetic.  Every function cites the
``/root/reference`` lines it follows.  Parity status: **pinned** -- the
restatement is checked in ``tests/test_oracle_golden.py`` against fixtures under
``tests/golden/`` that ``oracle/make_golden.py`` produced by executing the
reference's own functions (AST-extracted from ``/root/reference/functions.py``;
the module itself cannot be imported because TensorFlow is absent), plus the
worked example of ``IM_creation.jpg`` (README.md:16-17).

Conventions shared with the CUDA path:
  * probabilities: float32 NHWC, ``[N, H, W, K]`` (``[H, W, K]`` per image),
  * images: uint8 ``[H, W, c]`` as ``cv2.imread`` returns them (BGR on disk),
  * labels / IM: uint8 ``[H, W]``; binary labels and IM are 0/255, multiclass
    labels are class ids with class 0 == "inconsistent / unknown",
  * sizes: ``numpy.int64`` pixel counts taken BEFORE morphology and blanking.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "im_binary", "im_multiclass", "decide_binary", "decide_multiclass",
    "im_prediction_binary", "im_prediction_hela", "im_prediction_multiclass",
    "erode", "dilate", "dilate_label", "blank_binary", "blank_multiclass",
    "blank_hela", "write_decision_binary", "mean_im_size",
]


# --------------------------------------------------------------------------- a5
def im_binary(masks):
    """Ensemble agreement for 0/1 masks -- functions.py:3104-3120.

    ``S = sum_m mask_m``; label where ``S == M``; IM where ``S`` is neither 0
    nor ``M``.  Sizes are counted on the 0/1 maps before the x255 scaling.
    Returns ``(label u8[H,W], im u8[H,W], im_size i64, pred_size i64)`` -- note
    the reference returns im_size *before* pred_size (functions.py:3120).
    """
    stack = np.stack([np.asarray(m) for m in masks], axis=0).astype(np.int64)
    n = stack.shape[0]
    total = stack.sum(axis=0)
    agree_on = total == n
    mixed = (total != 0) & (total != n)
    pred_size = np.int64(agree_on.sum())
    im_size = np.int64(mixed.sum())
    label = (agree_on.astype(np.uint8) * np.uint8(255)).squeeze()
    im = (mixed.astype(np.uint8) * np.uint8(255)).squeeze()
    return label, im, im_size, pred_size


# --------------------------------------------------------------------------- a6
def im_multiclass(masks):
    """Ensemble agreement for class-id masks -- functions.py:3123-3137.

    A pixel agrees when every model equals model 0; the label keeps model 0's
    class there and is 0 elsewhere; the IM is the complement (x255).
    Returns ``(label u8[H,W], im u8[H,W], im_size i64)``.
    """
    stack = np.stack([np.asarray(m) for m in masks], axis=0)
    agree = np.ones(stack.shape[1:], dtype=bool)
    for m in range(1, stack.shape[0]):
        agree &= stack[m] == stack[0]
    label = np.where(agree, stack[0], 0)
    im = np.where(agree, 0, 255)
    im_size = np.int64((~agree).sum())
    return np.squeeze(label).astype(np.uint8), np.squeeze(im).astype(np.uint8), im_size


# ------------------------------------------------------------- per-model decisions
def decide_binary(prob, threshold, strict):
    """Per-model threshold.  ISIC compares with ``>`` (functions.py:3157), HeLa
    with ``>=`` (functions.py:3187-3189).  NaN compares false either way."""
    prob = np.asarray(prob, dtype=np.float32)
    thr = np.float32(threshold)     # NumPy compares a float32 array with a Python float in float32
    return (prob > thr) if strict else (prob >= thr)


def decide_multiclass(prob):
    """``np.argmax(axis=-1)`` -- functions.py:3225.  First index of the maximum
    wins ties and NaN counts as the maximum (NumPy semantics)."""
    return np.argmax(np.asarray(prob), axis=-1)


# --------------------------------------------------------------------------- a2
def im_prediction_binary(probs, threshold=0.5):
    """functions.py:3140-3162 with ``model.predict`` replaced by its output.

    ``probs``: sequence of M float32 arrays ``[H, W, 1]`` (one image).
    Returns ``(label, im, im_size, pred_size)``.
    """
    masks = [decide_binary(p, threshold, strict=True).astype(np.int64) for p in probs]
    return im_binary(masks)


# --------------------------------------------------------------------------- a3
def im_prediction_hela(probs, threshold=0.5):
    """functions.py:3165-3202.  ``probs``: M float32 arrays ``[H, W, 3]`` with
    heads (alive, dead, position).  Each head is thresholded with ``>=`` and
    combined independently; the combined IM is the per-pixel maximum of the three
    IMs and ``im_size`` is the SUM of the three IM sizes (not the union's size).
    Returns ``(alive, dead, pos_raw, combined_im, im_size)``.
    """
    heads = []
    for k in range(3):
        masks = [decide_binary(np.asarray(p)[..., k], threshold, strict=False).astype(np.int64)
                 for p in probs]
        heads.append(im_binary(masks))
    (alive, im_a, sz_a, _), (dead, im_d, sz_d, _), (pos, im_p, sz_p, _) = heads
    combined = np.maximum(np.maximum(im_a, im_d), im_p)
    return alive, dead, pos, combined, np.int64(sz_a + sz_d + sz_p)


# --------------------------------------------------------------------------- a4
def im_prediction_multiclass(probs, filter_unequal_class_pred=False):
    """functions.py:3206-3238.  ``probs``: M float32 arrays ``[1, H, W, K]`` or
    ``[H, W, K]``.  ``lists_equal`` is True unless the filter flag is on and the
    models predict different SETS of classes over the image (functions.py:3231-3234).
    Returns ``(label, im, im_size, lists_equal)``.
    """
    masks = [decide_multiclass(p) for p in probs]
    if filter_unequal_class_pred:
        sets = [set(np.unique(m).tolist()) for m in masks]
        lists_equal = all(s == sets[0] for s in sets)
    else:
        lists_equal = True
    label, im, im_size = im_multiclass(masks)
    return label, im, im_size, lists_equal


# ------------------------------------------------------------------ morphology
def _window_reduce(mask, k, reducer, pad_value):
    """k x k rectangular min/max filter, anchor at the centre ``k // 2`` like
    ``cv2.erode/dilate(mask, ones((k, k)), iterations=1)`` (functions.py:2858-2864).
    cv2's default border is a constant that never wins: +max for erosion (image
    borders do not erode), -max for dilation."""
    mask = np.asarray(mask)
    h, w = mask.shape
    a = k // 2                      # anchor; window covers [-a, k-1-a]
    padded = np.full((h + k - 1, w + k - 1), pad_value, dtype=mask.dtype)
    padded[a:a + h, a:a + w] = mask
    out = None
    for dy in range(k):
        for dx in range(k):
            view = padded[dy:dy + h, dx:dx + w]
            out = view.copy() if out is None else reducer(out, view)
    return out


def erode(mask, k):
    """``cv2.erode`` with a k x k ones kernel; k <= 0 is the identity (the
    reference skips the call, functions.py:2858)."""
    if k <= 0:
        return np.asarray(mask).copy()
    return _window_reduce(mask, k, np.minimum, np.iinfo(np.asarray(mask).dtype).max)


def dilate(mask, k):
    """``cv2.dilate`` with a k x k ones kernel; k <= 0 is the identity."""
    if k <= 0:
        return np.asarray(mask).copy()
    return _window_reduce(mask, k, np.maximum, np.iinfo(np.asarray(mask).dtype).min)


def dilate_label(label, kernel_size=3):
    """``dilate_mask`` -- functions.py:3075-3100.  Every non-zero class is dilated
    on its own with a 3x3 kernel, in ascending id order, later ids overwriting
    earlier ones -- i.e. each pixel takes the LARGEST class id present in its
    3x3 neighbourhood (0 when there is none)."""
    return dilate(np.asarray(label), kernel_size)


# --------------------------------------------------------------------------- a7
def write_decision_binary(pred_size, im_size, filter_bad_predictions=True):
    """functions.py:2876-2882: image and mask are written only when the agreed
    area is non-empty and larger than the IM; the IM file is always written."""
    if filter_bad_predictions:
        return bool(pred_size > im_size and pred_size > 0)
    return True


def blank_binary(image, label, im, erode_kernel=0, dilate_kernel=0,
                 block_input=True, block_output=True):
    """functions.py:2858-2874.  ``image`` uint8 ``[H, W, c]`` (or ``[H, W]``).
    Returns new ``(image, label, im)``; inputs are not modified."""
    image = np.array(image, copy=True)
    label = np.array(label, copy=True)
    im = erode(im, erode_kernel)
    im = dilate(im, dilate_kernel)
    hit = im > 0
    if block_input:
        image[hit] = 0
    if block_output:
        label[hit] = 0
    return image, label, im


# --------------------------------------------------------------------------- a8
def blank_multiclass(image, label, im, erode_kernel=0, dilate_kernel=0,
                     block_input=True, block_output=True):
    """functions.py:3043-3061.  As the binary case, but when ``erode_kernel > 0``
    the label is additionally passed through ``dilate_mask`` (functions.py:3047)."""
    image = np.array(image, copy=True)
    label = np.array(label, copy=True)
    if erode_kernel > 0:
        im = erode(im, erode_kernel)
        label = dilate_label(label)
    else:
        im = np.array(im, copy=True)
    im = dilate(im, dilate_kernel)
    hit = im > 0
    if block_input:
        image[hit] = 0
    if block_output:
        label[hit] = 0
    return image, label, im


# --------------------------------------------------------------------------- a9
def blank_hela(brightfield, alive, dead, pos_drawn, combined_im, erode_kernel=0,
               dilate_kernel=0, block_input=True, block_output=True):
    """functions.py:2942-2974 minus the host-side circle drawing (component #4,
    functions.py:2953-2965), whose result is passed in as ``pos_drawn``
    (uint8 ``[H, W, 3]``).  Returns ``(brightfield, alive, dead, pos, im)``."""
    brightfield = np.array(brightfield, copy=True)
    alive = np.array(alive, copy=True)
    dead = np.array(dead, copy=True)
    pos_drawn = np.array(pos_drawn, copy=True)
    im = combined_im
    if erode_kernel > 0:
        im = erode(im, erode_kernel)
        alive = dilate_label(alive)
        dead = dilate_label(dead)
    im = dilate(im, dilate_kernel)
    hit = im > 0
    if block_input:
        brightfield[hit] = 0
    if block_output:
        alive[hit] = 0
        dead[hit] = 0
        pos_drawn[hit] = 0
    return brightfield, alive, dead, pos_drawn, np.array(im, copy=True)


# -------------------------------------------------------------------------- a10
def mean_im_size(im_sizes):
    """functions.py:2889 (also :2982, :3068): Python ``round(sum / len, 0)`` --
    banker's rounding on a Python float."""
    im_sizes = [int(s) for s in im_sizes]
    return round(sum(im_sizes) / len(im_sizes), 0)
