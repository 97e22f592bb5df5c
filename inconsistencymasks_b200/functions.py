"""Drop-in replacements for the pseudo-label helpers of the reference's ``functions.py``.

Same names, argument order, defaults, return types and on-disk results as

    pred_masks_to_im_binary / _multiclass        functions.py:3104-3137
    get_im_prediction_binary / _hela / _multiclass   functions.py:3140-3238
    dilate_mask                                   functions.py:3075-3100
    create_pseudo_labels_im_ISIC_2018 / _hela / _multiclass   functions.py:2832-3070

but every array operation runs in libimk's CUDA kernels (include/imk.h).  ``models`` may be
``B200UNet`` objects (inconsistencymasks_b200.unet) -- then the ensemble forward and the IM
are one fused device pipeline -- or any objects with a Keras-like ``.predict`` -- then
their float32 outputs are uploaded and the standalone IM kernels are used.  PNG decode /
encode stays ``cv2`` on the host exactly as in the reference (out of scope).
There is no NumPy fallback for the arithmetic.
"""
from __future__ import annotations

import ctypes as C
import os

import cv2
import numpy as np

from . import _lib
from ._lib import lib, check
from .unet import B200UNet

def _config_threshold(default=0.5):
    """functions.py:25-31 reads ``THRESHOLD`` from ``config.ini`` in the working directory (DEFAULT section,
    config.ini:13); the same file is honoured here when present, else the reference's shipped value 0.5."""
    try:
        import configparser
        cp = configparser.ConfigParser()
        if cp.read("config.ini") and cp.has_option("DEFAULT", "THRESHOLD"):
            return cp.getfloat("DEFAULT", "THRESHOLD")
    except Exception:
        pass
    return default


THRESHOLD = _config_threshold()

__all__ = [
    "pred_masks_to_im_binary", "pred_masks_to_im_multiclass", "dilate_mask",
    "get_im_prediction_binary", "get_im_prediction_hela", "get_im_prediction_multiclass",
    "create_pseudo_labels_im_ISIC_2018", "create_pseudo_labels_im_hela", "create_pseudo_labels_im_multiclass",
    "get_pos_contours", "get_min_dist", "THRESHOLD",
]

from .evaluation import (benchmark_ISIC2018, benchmark_hela, benchmark_multiclass, get_IoU_binary, get_IoU_multi_unique,  # noqa: E402,F401
                         pixel_accuracy, dice_score_numpy_binary, mod_pos_size, get_cell_count, convert_class_to_color_mask)
from .augment import (augment_image_and_mask, augment_image_and_masks, create_augment_images_and_masks_ISIC_2018,  # noqa: E402,F401
                      create_augment_images_and_masks_hela, create_augment_images_and_masks_multiclass)

from .impp import (create_augment_images_and_masks_with_evalnet_ensemble_binary,  # noqa: E402,F401
                   create_augment_images_and_masks_with_evalnet_ensemble_multiclass)

__all__ += ["create_augment_images_and_masks_with_evalnet_ensemble_binary", "create_augment_images_and_masks_with_evalnet_ensemble_multiclass",
            "benchmark_ISIC2018", "benchmark_hela", "benchmark_multiclass", "get_IoU_binary", "get_IoU_multi_unique", "pixel_accuracy",
            "dice_score_numpy_binary", "mod_pos_size", "get_cell_count", "convert_class_to_color_mask",
            "augment_image_and_mask", "augment_image_and_masks", "create_augment_images_and_masks_ISIC_2018",
            "create_augment_images_and_masks_hela", "create_augment_images_and_masks_multiclass"]


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise _lib.ImkError("no CUDA device: inconsistencymasks_b200 has no CPU fallback")
    return torch


def _dev(arr):
    torch = _torch()
    return torch.from_numpy(np.ascontiguousarray(arr)).cuda()


def _stream():
    return _torch().cuda.current_stream().cuda_stream


def _all_b200(models):
    if not models or not all(isinstance(m, B200UNet) for m in models):
        return False
    c0 = models[0].config
    keys = ("i_height", "i_width", "i_channels", "num_outputmasks", "actifuout")
    return all(all(m.config[k] == c0[k] for k in keys) for m in models) and \
        len({(int(16 * m.config["alpha"]) + 15) // 16 for m in models}) == 1


def _handles(models):
    return (C.c_void_p * len(models))(*[m.handle for m in models])


# --------------------------------------------------------------------------- a5 / a6
def _masks_to_im(pred_masks, multiclass):
    torch = _torch()
    stack = np.stack([np.asarray(m) for m in pred_masks], axis=0)
    out_shape = np.squeeze(np.empty(stack.shape[1:], np.uint8)).shape
    m = stack.shape[0]
    if stack.dtype.kind == "f" and not np.array_equal(stack, np.trunc(stack)):
        # the reference sums the masks in their own dtype (functions.py:3108); only integer-valued masks (0/1 decisions,
        # class ids) have a defined IM, so fractional values are refused instead of being truncated silently
        raise ValueError("pred_masks_to_im_*: masks must hold integer values")
    flat = np.ascontiguousarray(stack.reshape(m, -1).astype(np.int64, copy=False))
    p = flat.shape[1]
    d_masks = _dev(flat)
    d_label = torch.empty(p, dtype=torch.uint8, device="cuda")
    d_im = torch.empty(p, dtype=torch.uint8, device="cuda")
    d_sizes = torch.empty(2, dtype=torch.int64, device="cuda")
    fn = lib.imk_masks_to_im_multiclass if multiclass else lib.imk_masks_to_im_binary
    check(fn(d_masks.data_ptr(), m, p, d_label.data_ptr(), d_im.data_ptr(), d_sizes.data_ptr(), _stream()))
    sizes = d_sizes.cpu().numpy()
    label = d_label.cpu().numpy().reshape(out_shape)
    im = d_im.cpu().numpy().reshape(out_shape)
    return label, im, np.int64(sizes[0]), np.int64(sizes[1])


def pred_masks_to_im_binary(pred_masks):
    """functions.py:3104-3120 -> ``(label u8, im u8, im_size, pred_size)``."""
    return _masks_to_im(pred_masks, False)


def pred_masks_to_im_multiclass(pred_masks):
    """functions.py:3123-3137 -> ``(label u8, im u8, im_size)``."""
    label, im, im_size, _ = _masks_to_im(pred_masks, True)
    return label, im, im_size


def dilate_mask(mask, kernel_size=3, iterations=1):
    """functions.py:3075-3100: per-class k x k dilation in ascending class id, later ids overwrite -- i.e. a k x k
    max filter per iteration (every pixel ends up with the largest class id its window holds).  Returns the input's
    dtype like the reference (``np.zeros_like(mask)``); class ids must fit uint8 (K <= 35 in config.ini)."""
    torch = _torch()
    mask = np.asarray(mask)
    if mask.ndim != 2:
        raise ValueError(f"dilate_mask expects a 2-D class-id map, got shape {mask.shape}")
    if mask.size and (mask.min() < 0 or mask.max() > 255):
        raise ValueError("dilate_mask: class ids outside 0..255")
    h, w = mask.shape
    src = _dev(mask.astype(np.uint8))
    dst = torch.empty_like(src)
    for _ in range(int(iterations)):
        check(lib.imk_dilate_u8(src.data_ptr(), dst.data_ptr(), 1, h, w, int(kernel_size), _stream()))
        src, dst = dst, src
    return src.cpu().numpy().astype(mask.dtype, copy=False)


# ------------------------------------------------------------------ batched device core
class _Result:
    __slots__ = ("labels", "im", "im_size", "pred_size", "lists_equal", "image")


def _predict_stack(models, images):
    """Duck-typed models: per model, per image ``.predict([img[None]])`` like functions.py:3157."""
    outs = []
    for model in models:
        per = [np.asarray(model.predict([images[i:i + 1]]), dtype=np.float32) for i in range(images.shape[0])]
        outs.append(np.ascontiguousarray(np.concatenate(per, axis=0)))
    return outs


def _run_batch(models, model_input, kind, *, threshold=THRESHOLD, blank_image=None, block_input=False,
               block_output=False, erode_kernel=0, dilate_kernel=0, want_lists_equal=False, swap_rb=False, slot=None):
    """One batch through the device path.

    model_input: uint8 [N,H,W,c] in the channel order the models consume (unless swap_rb).
    blank_image: uint8 [N,H,W,c] to blank (the BGR array of the drivers) or None.
    kind: 'binary' (K=1, strict >), 'hela' (K=3, >=), 'multiclass'.
    slot: optional ``_Slot`` whose pinned buffers receive the results of the fused host pipeline (the drivers).
    """
    torch = _torch()
    n, h, w, c = model_input.shape
    res = _Result()
    planes = 3 if kind == "hela" else 1
    morph = erode_kernel > 0 or dilate_kernel > 0
    fused = _all_b200(models)
    strict = 1 if kind == "binary" else 0
    # blanking inside the IM kernel is only valid when the IM is final (no morphology)
    k_bi = int(bool(block_input) and not morph)
    k_bo = int(bool(block_output) and not morph)

    if blank_image is not None and blank_image is not model_input:
        raise ValueError("the image to blank must be the array the models are fed (channel order is handled by swap_rb)")
    want_img = blank_image is not None

    if fused and not morph:
        # host pipeline: uploads / downloads overlapped with compute inside libimk
        src = np.ascontiguousarray(model_input)
        if slot is not None:
            labels = slot.labels[:planes * n * h * w].reshape(planes, n, h, w)
            im, im_size = slot.im[:n], slot.im_size[:n]
            img_out = slot.img_out[:n] if want_img else None
        else:
            labels = np.empty((planes, n, h, w), np.uint8)
            im = np.empty((n, h, w), np.uint8)
            im_size = np.empty(n, np.int64)
            img_out = np.empty_like(src) if want_img else None
        hs = _handles(models)
        if kind == "multiclass":
            leq = (slot.leq[:n] if slot is not None else np.empty(n, np.uint8)) if want_lists_equal else None
            check(lib.imk_pseudo_label_multiclass_host(hs, len(models), src.ctypes.data, n, int(swap_rb), k_bi, k_bo,
                                                       img_out.ctypes.data if want_img else None,
                                                       labels.ctypes.data, im.ctypes.data, im_size.ctypes.data,
                                                       leq.ctypes.data if leq is not None else None, 0))
            res.pred_size, res.lists_equal = None, leq
        else:
            pred = slot.pred[:planes * n].reshape(planes, n) if slot is not None else np.empty((planes, n), np.int64)
            check(lib.imk_pseudo_label_binary_host(hs, len(models), src.ctypes.data, n, int(swap_rb), float(threshold), strict, k_bi, k_bo,
                                                   img_out.ctypes.data if want_img else None,
                                                   labels.ctypes.data, im.ctypes.data, im_size.ctypes.data,
                                                   pred.ctypes.data, 0))
            res.pred_size, res.lists_equal = pred, None
        res.labels, res.im, res.im_size, res.image = labels, im, im_size, img_out
        return res

    # device-buffer path (morphology and / or duck-typed models)
    s = _stream()
    d_labels = torch.empty((planes, n, h, w), dtype=torch.uint8, device="cuda")
    d_im = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
    d_im_size = torch.empty(n, dtype=torch.int64, device="cuda")
    d_pred = torch.empty((planes, n), dtype=torch.int64, device="cuda")
    d_leq = torch.empty(n, dtype=torch.uint8, device="cuda") if (kind == "multiclass" and want_lists_equal) else None
    d_img = _dev(model_input)
    d_img_out = torch.empty_like(d_img) if want_img else None
    out_ptr = d_img_out.data_ptr() if want_img else None
    if fused:
        if kind == "multiclass":
            check(lib.imk_ensemble_im_multiclass(_handles(models), len(models), d_img.data_ptr(), n, int(swap_rb), k_bi, k_bo, out_ptr,
                                                 d_labels.data_ptr(), d_im.data_ptr(), d_im_size.data_ptr(),
                                                 d_leq.data_ptr() if d_leq is not None else None, s))
        else:
            check(lib.imk_ensemble_im_binary(_handles(models), len(models), d_img.data_ptr(), n, int(swap_rb), float(threshold), strict,
                                             k_bi, k_bo, out_ptr, d_labels.data_ptr(), d_im.data_ptr(),
                                             d_im_size.data_ptr(), d_pred.data_ptr(), s))
    else:
        fed = np.ascontiguousarray(model_input[..., ::-1]) if swap_rb else model_input
        probs = _predict_stack(models, fed)
        k = probs[0].shape[-1]
        d_probs = [_dev(p) for p in probs]
        ptrs = (C.c_void_p * len(d_probs))(*[p.data_ptr() for p in d_probs])
        if kind == "multiclass":
            check(lib.imk_im_multiclass(ptrs, len(d_probs), n, h, w, k, d_img.data_ptr(), c, k_bi, k_bo, out_ptr,
                                        d_labels.data_ptr(), d_im.data_ptr(), d_im_size.data_ptr(),
                                        d_leq.data_ptr() if d_leq is not None else None, s))
        else:
            if k != planes:
                raise ValueError(f"models output {k} maps, {kind} IM expects {planes}")
            check(lib.imk_im_binary(ptrs, len(d_probs), n, h, w, k, float(threshold), strict, d_img.data_ptr(), c,
                                    k_bi, k_bo, out_ptr, d_labels.data_ptr(), d_im.data_ptr(), d_im_size.data_ptr(),
                                    d_pred.data_ptr(), s))
    if morph:
        # functions.py:2858-2864 / 2942-2951 / 3043-3051: erode(im), dilate_mask(labels) when EK > 0, dilate(im);
        # the IM kernels above ran with blanking off and copied the image through unchanged
        tmp = torch.empty_like(d_im)
        if erode_kernel > 0:
            check(lib.imk_erode_u8(d_im.data_ptr(), tmp.data_ptr(), n, h, w, int(erode_kernel), s))
            d_im, tmp = tmp, d_im
            if kind in ("multiclass", "hela"):
                nl = 1 if kind == "multiclass" else 2          # alive, dead; the raw position map is not dilated
                lab_tmp = d_labels.clone()
                check(lib.imk_dilate_u8(d_labels.data_ptr(), lab_tmp.data_ptr(), nl * n, h, w, 3, s))
                d_labels = lab_tmp
        if dilate_kernel > 0:
            check(lib.imk_dilate_u8(d_im.data_ptr(), tmp.data_ptr(), n, h, w, int(dilate_kernel), s))
            d_im, tmp = tmp, d_im
        nl = {"binary": 1, "hela": 2, "multiclass": 1}[kind]
        check(lib.imk_blank(d_im.data_ptr(), n, h, w,
                            d_img_out.data_ptr() if (want_img and block_input) else None, c,
                            d_labels.data_ptr() if block_output else None, nl, s))
    res.labels = d_labels.cpu().numpy()
    res.im = d_im.cpu().numpy()
    res.im_size = d_im_size.cpu().numpy()
    res.pred_size = d_pred.cpu().numpy() if kind != "multiclass" else None
    res.lists_equal = d_leq.cpu().numpy() if d_leq is not None else None
    res.image = d_img_out.cpu().numpy() if want_img else None
    return res


# --------------------------------------------------------------------------- a2 / a3 / a4
def _as_batch(prepared_image):
    x = np.asarray(prepared_image)
    if x.ndim != 4:
        raise ValueError(f"prepared_image must be [1,H,W,c], got {x.shape}")
    return np.ascontiguousarray(x.astype(np.uint8, copy=False))


def get_im_prediction_binary(models, prepared_image, threshold=0.5):
    """functions.py:3140-3162 -> ``(label u8[H,W], im u8[H,W], im_size, pred_size)``."""
    r = _run_batch(models, _as_batch(prepared_image), "binary", threshold=threshold)
    return r.labels[0, 0], r.im[0], np.int64(r.im_size[0]), np.int64(r.pred_size[0, 0])


def get_im_prediction_hela(models, prepared_image, threshold=0.5):
    """functions.py:3165-3202 -> ``(alive, dead, pos_raw, combined_im, im_size)``."""
    r = _run_batch(models, _as_batch(prepared_image), "hela", threshold=threshold)
    return r.labels[0, 0], r.labels[1, 0], r.labels[2, 0], r.im[0], np.int64(r.im_size[0])


def get_im_prediction_multiclass(models, prepared_image, filter_unequal_class_pred=False):
    """functions.py:3206-3238 -> ``(label u8[H,W], im u8[H,W], im_size, lists_equal)``."""
    r = _run_batch(models, _as_batch(prepared_image), "multiclass", want_lists_equal=bool(filter_unequal_class_pred))
    eq = bool(r.lists_equal[0]) if filter_unequal_class_pred else True
    return r.labels[0, 0], r.im[0], np.int64(r.im_size[0]), eq


# ------------------------------------------------------- component #4 (host geometry, cv2)
def get_pos_contours(img, erode_kernel=3):
    """functions.py:6181-6218: centres (+1, +1) of the connected blobs of a position map."""
    if img.ndim not in (2, 3):
        raise AssertionError("Invalid image dimensions.")
    gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY) if (img.ndim == 3 and img.shape[2] > 1) else img
    if erode_kernel > 0:
        gray = cv2.erode(cv2.convertScaleAbs(gray), np.ones((erode_kernel, erode_kernel), "uint8"), iterations=1)
    _, binary = cv2.threshold(gray, 10, 255, 0)
    contours, _ = cv2.findContours(binary.astype("uint8"), cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
    centres = []
    for contour in contours:
        mom = cv2.moments(contour)
        if mom["m00"] != 0:
            centres.append((int(mom["m10"] / mom["m00"]) + 1, int(mom["m01"] / mom["m00"]) + 1))
    return centres


def get_min_dist(xy, positions):
    """functions.py:6221-6250: smallest non-zero Euclidean distance from ``xy`` to ``positions``."""
    d = np.linalg.norm(np.array(positions) - np.array(xy), axis=1)
    d = d[d > 0]
    if d.size == 0:
        print("No other points or positions list is empty.")
        return 0
    return np.min(d)


def _draw_positions(pos_raw, h, w, max_pos_circle_size, min_pos_circle_size):
    """functions.py:2953-2965."""
    positions = get_pos_contours(pos_raw)
    canvas = np.zeros((h, w, 3), np.uint8)
    for pos in positions:
        min_dist = get_min_dist(pos, positions) if len(positions) > 1 else 99
        radius = max(min(int(min_dist // 4), max_pos_circle_size), min_pos_circle_size)
        cv2.circle(canvas, (pos[0], pos[1]), radius, (255, 255, 255), -1)
    return canvas


# --------------------------------------------------------------------------- a7 / a8 / a9 / a10
# The per-directory drivers.  The reference decodes, predicts and encodes one file at a time (functions.py:2844-2887);
# here a directory runs as a 3-slot software pipeline: a thread pool decodes the PNGs of batch i+1 into pinned host
# buffers and encodes / writes the results of batch i-1 (cv2 releases the GIL) while libimk's host pipeline streams
# batch i through the GPU.  File contents, file names and the returned statistic are those of the reference.
_FILES_PER_BATCH = 512
_IO_THREADS = max(4, min(32, os.cpu_count() or 4))


def _pinned(shape, dtype):
    """Page-locked host array (asynchronous copies): a NumPy view of a pinned torch tensor."""
    torch = _torch()
    tdt = {np.uint8: torch.uint8, np.int64: torch.int64}[dtype]
    return torch.empty(shape, dtype=tdt, pin_memory=True).numpy()


class _Slot:
    """Host buffers of one batch in flight (flat, so that a short last batch still gets dense [planes][n] planes)."""

    def __init__(self, cap, h, w, c, planes):
        self.cap, self.h, self.w, self.c, self.planes = cap, h, w, c, planes
        self.img = _pinned((cap, h, w, c), np.uint8)
        self.img_out = _pinned((cap, h, w, c), np.uint8)
        self.labels = _pinned((planes * cap * h * w,), np.uint8)
        self.im = _pinned((cap, h, w), np.uint8)
        self.im_size = _pinned((cap,), np.int64)
        self.pred = _pinned((planes * cap,), np.int64)
        self.leq = _pinned((cap,), np.uint8)


def _shard_names(names, shard):
    """(names of this rank, rank, world).  shard: None = automatic (an initialised torch.distributed group with more
    than one rank shards the directory), False = never, (rank, world) = explicit."""
    if shard is False:
        return names, 0, 1
    if shard is None or shard is True:
        rank, world = 0, 1
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                rank, world = dist.get_rank(), dist.get_world_size()
        except ImportError:
            pass
    else:
        rank, world = int(shard[0]), int(shard[1])
    if world <= 1:
        return names, 0, 1
    from .pool import shard_indices
    ordered = sorted(names)                      # os.listdir order is arbitrary: every rank must see the same list
    return [ordered[i] for i in shard_indices(len(ordered), rank, world)], rank, world


def _mean_im_size(im_sizes, world=1):
    """functions.py:2889: ``round(sum(im_sizes.values()) / len(im_sizes), 0)`` (banker's rounding); with a sharded
    directory the integer sums are all-reduced first (one int64[3] exchange, no collective on the data path)."""
    total, count = sum(int(v) for v in im_sizes.values()), len(im_sizes)
    if world > 1:
        from .pool import allreduce_stats
        total, _, count = allreduce_stats(total, 0, count)
    return round(total / count, 0)


def _run_directory(models, kind, h, w, c, images_path, read_flags, run_kwargs, want_image, decide, store, shard=None):
    """Pipelined per-directory loop shared by the three drivers.

    decide(result, i) -> per-file context (or None to skip the gated outputs); store(slot_view, ctx, i, name) writes the
    files of image i on a pool thread.  Returns ``(im_sizes dict, world)``."""
    from concurrent.futures import ThreadPoolExecutor
    names, _, world = _shard_names(os.listdir(images_path), shard)
    im_sizes = {}
    if not names:
        return im_sizes, world
    planes = 3 if kind == "hela" else 1
    batches = [names[i:i + _FILES_PER_BATCH] for i in range(0, len(names), _FILES_PER_BATCH)]
    cap = max(len(b) for b in batches)
    slots = [_Slot(cap, h, w, c, planes) for _ in range(min(3, len(batches)))]
    busy = [[] for _ in slots]

    def load_one(slot, i, name):
        image = cv2.imread(os.path.join(images_path, name), *read_flags)
        if image is None:                            # the reference fails on `image.reshape` / cvtColor of None
            raise AttributeError(f"cv2.imread returned None for {os.path.join(images_path, name)!r}")
        slot.img[i] = image.reshape(h, w, c)

    with ThreadPoolExecutor(_IO_THREADS) as io:
        def start_load(bi):
            k = bi % len(slots)
            for f in busy[k]:
                f.result()                           # the slot's previous results are on disk (re-raises I/O errors)
            busy[k] = []
            return [io.submit(load_one, slots[k], i, nm) for i, nm in enumerate(batches[bi])]

        loads = start_load(0)
        for bi, bnames in enumerate(batches):
            for f in loads:
                f.result()
            if bi + 1 < len(batches):
                loads = start_load(bi + 1)
            slot, n = slots[bi % len(slots)], len(bnames)
            batch = slot.img[:n]
            r = _run_batch(models, batch, kind, blank_image=batch if want_image else None, slot=slot, **run_kwargs)
            ctxs = []
            for i, nm in enumerate(bnames):
                im_sizes[nm[:-4]] = int(r.im_size[i])
                ctxs.append(decide(r, i))
            busy[bi % len(slots)] = [io.submit(store, r, ctxs[i], i, nm) for i, nm in enumerate(bnames)]
        for b in busy:
            for f in b:
                f.result()
    return im_sizes, world


def create_pseudo_labels_im_ISIC_2018(models, h, w, c, images_path, main_output_path, rgb=True, erode_kernel=5,
                                      dilate_kernel=5, block_input=True, block_output=True, filter_bad_predictions=True,
                                      shard=None):
    """functions.py:2832-2891.  Returns ``mean_im_size`` (float).  ``shard``: see ``_shard_names``."""
    out_img, out_mask, out_im = (os.path.join(main_output_path, d) for d in ("images", "masks", "im"))
    for d in (out_img, out_mask, out_im):
        os.makedirs(d, exist_ok=True)
    swap = bool(rgb) and c == 3                          # cv2.cvtColor(image, COLOR_BGR2RGB), functions.py:2847-2848

    def decide(r, i):
        im_size, pred_size = int(r.im_size[i]), int(r.pred_size[0, i])
        return (pred_size > im_size and pred_size > 0) if filter_bad_predictions else True

    def store(r, write, i, nm):
        if write:
            cv2.imwrite(os.path.join(out_img, nm), r.image[i] if c == 3 else r.image[i, ..., 0])
            cv2.imwrite(os.path.join(out_mask, nm), r.labels[0, i])
        cv2.imwrite(os.path.join(out_im, nm), r.im[i])

    kw = dict(threshold=THRESHOLD, block_input=block_input, block_output=block_output, erode_kernel=erode_kernel,
              dilate_kernel=dilate_kernel, swap_rb=swap)
    im_sizes, world = _run_directory(models, "binary", h, w, c, images_path, (), kw, True, decide, store, shard)
    return _mean_im_size(im_sizes, world)


def create_pseudo_labels_im_hela(models, h, w, c, images_path, main_output_path, erode_kernel=5, dilate_kernel=5,
                                 block_input=True, block_output=True, max_pos_circle_size=8, min_pos_circle_size=3,
                                 shard=None):
    """functions.py:2895-2984.  Returns ``mean_im_size`` (float)."""
    outs = {d: os.path.join(main_output_path, d) for d in ("brightfield", "alive", "dead", "mod_position", "im")}
    for d in outs.values():
        os.makedirs(d, exist_ok=True)

    def store(r, _ctx, i, nm):
        pos = _draw_positions(r.labels[2, i], h, w, max_pos_circle_size, min_pos_circle_size)   # host geometry, on the pool
        if block_output:
            pos[r.im[i] > 0] = 0                         # functions.py:2974, on the host-drawn circles
        cv2.imwrite(os.path.join(outs["brightfield"], nm), r.image[i, ..., 0])
        cv2.imwrite(os.path.join(outs["alive"], nm), r.labels[0, i])
        cv2.imwrite(os.path.join(outs["dead"], nm), r.labels[1, i])
        cv2.imwrite(os.path.join(outs["mod_position"], nm), pos)
        cv2.imwrite(os.path.join(outs["im"], nm), r.im[i])

    kw = dict(threshold=THRESHOLD, block_input=block_input, block_output=block_output, erode_kernel=erode_kernel,
              dilate_kernel=dilate_kernel, swap_rb=False)
    im_sizes, world = _run_directory(models, "hela", h, w, c, images_path, (0,), kw, True, lambda r, i: None, store, shard)
    return _mean_im_size(im_sizes, world)


def create_pseudo_labels_im_multiclass(models, h, w, c, images_path, main_output_path, rgb=True, erode_kernel=5,
                                       dilate_kernel=5, block_input=True, block_output=True, filter_unequal_class_pred=False,
                                       shard=None):
    """functions.py:2988-3070.  Returns ``mean_im_size`` (float)."""
    out_img, out_mask, out_im = (os.path.join(main_output_path, d) for d in ("images", "masks", "im"))
    for d in (out_img, out_mask, out_im):
        os.makedirs(d, exist_ok=True)
    swap = bool(rgb) and c == 3

    def decide(r, i):
        return bool(r.lists_equal[i]) if filter_unequal_class_pred else True

    def store(r, write, i, nm):
        if write:
            cv2.imwrite(os.path.join(out_img, nm), r.image[i] if c == 3 else r.image[i, ..., 0])
            cv2.imwrite(os.path.join(out_mask, nm), r.labels[0, i])
        cv2.imwrite(os.path.join(out_im, nm), r.im[i])

    kw = dict(block_input=block_input, block_output=block_output, erode_kernel=erode_kernel, dilate_kernel=dilate_kernel,
              want_lists_equal=bool(filter_unequal_class_pred), swap_rb=swap)
    im_sizes, world = _run_directory(models, "multiclass", h, w, c, images_path, (), kw, True, decide, store, shard)
    return _mean_im_size(im_sizes, world)
