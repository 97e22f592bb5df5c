"""TEST INFRASTRUCTURE ONLY -- restatement of the reference's augmentation step with cv2 (never imported by the product).

    augment_image_and_masks / augment_image_and_mask   /root/reference/functions.py:2725-2828
    add_noise_and_blur                                  /root/reference/functions.py:1480-1506

``draw`` makes the reference's random draws in the reference's order and returns them as a dict; ``apply`` runs the
deterministic operations with the same cv2 calls.  Pinned by tests/golden/augment.npz (seeded runs of the reference).
"""
import random

import cv2
import numpy as np


def draw(brightness_range_alpha=(0.5, 1.5), brightness_range_beta=(-25, 25), max_blur=3, free_rotation=True):
    p = dict(flip_v=0, flip_h=0, rot=0)
    if free_rotation:
        p["flip_v"] = int(random.randint(0, 1) == 1)              # :2795-2798
    p["flip_h"] = int(random.randint(0, 1) == 1)                  # :2800-2802
    if free_rotation:
        p["rot"] = random.randint(0, 3)                           # :2804-2817
    p["alpha"] = np.random.uniform(brightness_range_alpha[0], brightness_range_alpha[1])   # :2819
    p["beta"] = np.random.uniform(brightness_range_beta[0], brightness_range_beta[1])      # :2820
    p["scale_on"] = int(random.randint(0, 1) == 1)                # :2822
    p["blur_k"] = {0: 0, 1: 3, 2: 5, 3: 7}.get(random.randint(0, max_blur), 0)            # :1496-1502
    return p


def apply(image, masks, p):
    if p["flip_v"]:
        image, masks = cv2.flip(image, 0), [cv2.flip(m, 0) for m in masks]
    if p["flip_h"]:
        image, masks = cv2.flip(image, 1), [cv2.flip(m, 1) for m in masks]
    code = {1: cv2.ROTATE_90_CLOCKWISE, 2: cv2.ROTATE_180, 3: cv2.ROTATE_90_COUNTERCLOCKWISE}.get(p["rot"])
    if code is not None:
        image, masks = cv2.rotate(image, code), [cv2.rotate(m, code) for m in masks]
    if p["scale_on"]:
        image = cv2.convertScaleAbs(image, alpha=p["alpha"], beta=p["beta"])
    if p["blur_k"]:
        image = cv2.GaussianBlur(image, (p["blur_k"], p["blur_k"]), 0)
    return image, masks
