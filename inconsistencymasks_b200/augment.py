"""IM+ / IM++ augmentation step: ``augment_image_and_mask(s)`` of the reference's ``functions.py``.

    augment_image_and_masks   functions.py:2725-2776
    augment_image_and_mask    functions.py:2779-2828
    add_noise / add_noise_and_blur / apply_random_flip_and_rotation primitives   functions.py:1463-1506

Same names, argument order, defaults and return types.  RNG contract: the reference draws its decisions from the
(unseeded) ``random`` and ``np.random`` modules; ``draw_params`` makes exactly the same calls in the same order, so a
caller that seeds both modules gets the reference's flips, rotation, brightness and blur choice.  The per-pixel noise
(``np.random.randint(size=image.shape)`` in the reference) is replaced by ONE ``np.random.randint`` seed for the
device's counter-based generator: uniform integers in [-max_noise, max_noise) -- distributional parity only.
The pixels are computed by ``imk_augment_u8`` (csrc/imk_augment.cu): bit-exact against cv2 for everything but the noise.

``augment_batch`` is the device-resident form the pipeline uses: it takes the blanked images and labels of a pseudo-label
batch as CUDA tensors straight from the IM kernels and returns CUDA tensors -- no PNG round trip in between.
"""
from __future__ import annotations

import ctypes as C
import random

import numpy as np

from ._lib import lib, check, AugParams

__all__ = ["augment_image_and_mask", "augment_image_and_masks", "draw_params", "augment_batch",
           "create_augment_images_and_masks_ISIC_2018", "create_augment_images_and_masks_hela",
           "create_augment_images_and_masks_multiclass"]


def draw_params(brightness_range_alpha=(0.5, 1.5), brightness_range_beta=(-25, 25), max_blur=3, max_noise=25, free_rotation=True):
    """The decisions of functions.py:2795-2826 + 1496-1505, drawn with the reference's calls in the reference's order."""
    p = AugParams()
    if free_rotation:
        p.flip_v = int(random.randint(0, 1) == 1)
    p.flip_h = int(random.randint(0, 1) == 1)
    if free_rotation:
        p.rot = random.randint(0, 3)
    p.alpha = np.random.uniform(brightness_range_alpha[0], brightness_range_alpha[1])
    p.beta = np.random.uniform(brightness_range_beta[0], brightness_range_beta[1])
    p.scale_on = int(random.randint(0, 1) == 1)
    p.blur_k = {0: 0, 1: 3, 2: 5, 3: 7}.get(random.randint(0, max_blur), 0)
    p.noise_max = int(max_noise) if max_noise > 0 else 0
    p.seed = int(np.random.randint(0, 2 ** 31 - 1)) if max_noise > 0 else 0
    return p


def augment_batch(images, masks, params):
    """images: CUDA uint8 [N,H,W,c] (or None); masks: CUDA uint8 [P,N,H,W] (or None); params: list of N AugParams.
    Returns ``(images_out, masks_out)`` as CUDA tensors (rotations by 90 degrees need H == W)."""
    from . import functions as F
    torch = F._torch()
    ref = images if images is not None else masks[0]
    n = ref.shape[0]
    h, w = ref.shape[1], ref.shape[2]
    c = images.shape[3] if images is not None else 1
    planes = masks.shape[0] if masks is not None else 0
    arr = (AugParams * n)(*params)
    img_out = torch.empty_like(images) if images is not None else None
    masks_out = torch.empty_like(masks) if masks is not None else None
    scratch = torch.empty_like(images) if images is not None and any(p.blur_k or p.noise_max for p in params) else None
    check(lib.imk_augment_u8(images.data_ptr() if images is not None else None, masks.data_ptr() if masks is not None else None,
                             n, h, w, c, planes, arr, img_out.data_ptr() if img_out is not None else None,
                             masks_out.data_ptr() if masks_out is not None else None,
                             scratch.data_ptr() if scratch is not None else None, F._stream()))
    return img_out, masks_out


def _one(image, masks, kw):
    from . import functions as F
    image = np.ascontiguousarray(np.asarray(image, dtype=np.uint8))
    squeeze = image.ndim == 2
    img4 = image.reshape(1, image.shape[0], image.shape[1], 1 if squeeze else image.shape[2])
    p = draw_params(**kw)
    d_masks = F._dev(np.stack([np.asarray(m, dtype=np.uint8) for m in masks])[:, None]) if masks else None
    out, mo = augment_batch(F._dev(img4), d_masks, [p])
    out = out.cpu().numpy()[0]
    return (out[..., 0] if squeeze else out), ([m for m in mo.cpu().numpy()[:, 0]] if masks else [])


def augment_image_and_masks(image, masks, brightness_range_alpha=(0.5, 1.5), brightness_range_beta=(-25, 25), max_blur=3, max_noise=25,
                            free_rotation=True):
    """functions.py:2725-2776 -> ``(image, [masks])``."""
    return _one(image, list(masks), dict(brightness_range_alpha=brightness_range_alpha, brightness_range_beta=brightness_range_beta,
                                         max_blur=max_blur, max_noise=max_noise, free_rotation=free_rotation))


def augment_image_and_mask(image, mask, brightness_range_alpha=(0.5, 1.5), brightness_range_beta=(-25, 25), max_blur=3, max_noise=25,
                           free_rotation=True):
    """functions.py:2779-2828 -> ``(image, mask)``."""
    out, masks = _one(image, [mask], dict(brightness_range_alpha=brightness_range_alpha, brightness_range_beta=brightness_range_beta,
                                          max_blur=max_blur, max_noise=max_noise, free_rotation=free_rotation))
    return out, masks[0]


# ------------------------------------------------------------------------------------ per-directory drivers (IM+ step)
_SAMPLES_PER_BATCH = 2048


def _geometry_only(p):
    q = AugParams()
    q.flip_v, q.flip_h, q.rot = p.flip_v, p.flip_h, p.rot
    return q


def _augment_directory(dirs_in, dirs_out, num_images, copy_org, kw):
    """dirs_in[0] holds the images, the others their masks (same file names); every file is read with ``cv2.imread``'s
    default flags like the reference (3-channel BGR), augmented ``num_images`` times and written as ``<stem>_aug_<n>.png``.
    The decisions are drawn per (file, n) in the reference's loop order; the pixels of a whole batch of samples are
    computed on the device (image: all operations; masks: the geometry only)."""
    import os
    import shutil
    from concurrent.futures import ThreadPoolExecutor
    import cv2
    from . import functions as F
    for d in dirs_out:
        os.makedirs(d, exist_ok=True)
    names = os.listdir(dirs_in[0])
    if copy_org:
        for nm in names:
            for di, do in zip(dirs_in, dirs_out):
                shutil.copy(os.path.join(di, nm), os.path.join(do, nm))
    if not names or num_images <= 0:
        return
    threads = max(4, min(32, os.cpu_count() or 4))
    per_batch = max(1, _SAMPLES_PER_BATCH // num_images)
    with ThreadPoolExecutor(threads) as io:
        for b0 in range(0, len(names), per_batch):
            bn = names[b0:b0 + per_batch]
            stacks = [list(io.map(lambda nm, d=d: cv2.imread(os.path.join(d, nm)), bn)) for d in dirs_in]
            params = [draw_params(**kw) for _ in bn for _ in range(num_images)]        # file-major, n inner: functions.py:2604-2609
            # group by shape (a batch on the device shares H x W)
            shapes = {}
            for i, im in enumerate(stacks[0]):
                shapes.setdefault(im.shape, []).append(i)
            for shape, idx in shapes.items():
                rep = np.repeat(np.asarray(idx), num_images)
                ps = [params[i * num_images + n] for i in idx for n in range(num_images)]
                outs = []
                for k, st in enumerate(stacks):
                    batch = F._dev(np.ascontiguousarray(np.stack([st[i] for i in rep])))
                    out, _ = augment_batch(batch, None, ps if k == 0 else [_geometry_only(p) for p in ps])
                    outs.append(out.cpu().numpy())
                jobs = []
                for j, i in enumerate(rep):
                    n = j % num_images
                    for k, do in enumerate(dirs_out):
                        jobs.append((os.path.join(do, f"{bn[i][:-4]}_aug_{n}.png"), outs[k][j]))
                list(io.map(lambda job: cv2.imwrite(job[0], job[1]), jobs))


def create_augment_images_and_masks_ISIC_2018(images_path, masks_path, main_output_path, num_images=9, copy_org=True,
                                              brightness_range_alpha=(0.5, 1.5), brightness_range_beta=(-25, 25), max_blur=3, max_noise=25,
                                              free_rotation=True):
    """functions.py:2567-2609."""
    import os
    _augment_directory([images_path, masks_path], [os.path.join(main_output_path, "images"), os.path.join(main_output_path, "masks")],
                       num_images, copy_org, dict(brightness_range_alpha=brightness_range_alpha, brightness_range_beta=brightness_range_beta,
                                                  max_blur=max_blur, max_noise=max_noise, free_rotation=free_rotation))


def create_augment_images_and_masks_multiclass(images_path, masks_path, main_output_path, num_images=9, copy_org=True, free_rotation=False,
                                               brightness_range_alpha=(0.5, 1.5), brightness_range_beta=(-25, 25), max_blur=3, max_noise=25):
    """functions.py:2678-2722."""
    import os
    _augment_directory([images_path, masks_path], [os.path.join(main_output_path, "images"), os.path.join(main_output_path, "masks")],
                       num_images, copy_org, dict(brightness_range_alpha=brightness_range_alpha, brightness_range_beta=brightness_range_beta,
                                                  max_blur=max_blur, max_noise=max_noise, free_rotation=free_rotation))


def create_augment_images_and_masks_hela(main_input_path, main_output_path, num_images=9, copy_org=True, free_rotation=True,
                                         brightness_range_alpha=(0.7, 1.3), brightness_range_beta=(-15, 15), max_blur=3, max_noise=25):
    """functions.py:2613-2675."""
    import os
    subs = ("brightfield", "alive", "dead", "mod_position")
    _augment_directory([os.path.join(main_input_path, s) for s in subs], [os.path.join(main_output_path, s) for s in subs],
                       num_images, copy_org, dict(brightness_range_alpha=brightness_range_alpha, brightness_range_beta=brightness_range_beta,
                                                  max_blur=max_blur, max_noise=max_noise, free_rotation=free_rotation))
