"""CPU checks of the oracle restatements behind the SURVEY.md 8f rows against fixtures the reference generated
(oracle/make_golden_next.py), and of the host-side logic of the product that needs no GPU."""
import os
import random

import numpy as np
import pytest

from oracle import ref_metrics, ref_augment

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_metrics_oracle_vs_reference():
    z = np.load(os.path.join(G, "metrics.npz"))
    for i in range(int(z["n"])):
        gt, pred = z[f"b{i}_gt"], z[f"b{i}_pred"]
        iou, dice = ref_metrics.get_IoU_binary(gt, pred), ref_metrics.dice_score_numpy_binary(gt, pred)
        assert iou == z[f"b{i}_iou"] and round(iou, 4) == z[f"b{i}_iou4"]
        assert dice.dtype == z[f"b{i}_dice"].dtype and dice == z[f"b{i}_dice"] and round(dice, 4) == z[f"b{i}_dice4"]
        gtm, predm = z[f"m{i}_gt"], z[f"m{i}_pred"]
        assert ref_metrics.get_IoU_multi_unique(predm, gtm) == z[f"m{i}_iou"]
        assert ref_metrics.pixel_accuracy(predm, gtm) == z[f"m{i}_pa"]


def test_quotients_from_counts_match_the_reference():
    """The product forms the metrics from integer counts (evaluation._iou_binary ...): same bits as the reference."""
    from inconsistencymasks_b200 import evaluation as E
    z = np.load(os.path.join(G, "metrics.npz"))
    for i in range(int(z["n"])):
        gt, pred = z[f"b{i}_gt"], z[f"b{i}_pred"]
        c = np.array([np.sum((gt != 0) & (pred != 0)), np.sum((gt != 0) | (pred != 0)), np.sum((gt >= 128) & (pred >= 128)),
                      np.sum(gt >= 128), np.sum(pred >= 128)], np.int64)
        assert E._iou_binary(c) == z[f"b{i}_iou"] and round(E._iou_binary(c), 4) == z[f"b{i}_iou4"]
        d = E._dice_binary(c)
        assert d.dtype == z[f"b{i}_dice"].dtype and d == z[f"b{i}_dice"] and round(d, 4) == z[f"b{i}_dice4"]
        gtm, predm = z[f"m{i}_gt"], z[f"m{i}_pred"]
        h = np.zeros((3, 256), np.int64)
        for v in range(256):
            h[0, v], h[1, v], h[2, v] = np.sum(gtm == v), np.sum(predm == v), np.sum((gtm == v) & (predm == v))
        assert E._iou_multi_unique(h) == z[f"m{i}_iou"] and round(E._iou_multi_unique(h), 4) == z[f"m{i}_iou4"]
        assert E._pixel_accuracy(h, gtm.size) == z[f"m{i}_pa"] and round(E._pixel_accuracy(h, gtm.size), 4) == z[f"m{i}_pa4"]


def test_augment_oracle_vs_reference_and_draw_order():
    from inconsistencymasks_b200 import augment as A
    z = np.load(os.path.join(G, "augment.npz"))
    for seed, square, multi in (tuple(int(v) for v in row) for row in z["cases"]):
        image, mask, mask2 = z[f"a{seed}_image"], z[f"a{seed}_mask"], z[f"a{seed}_mask2"]
        random.seed(1000 + seed); np.random.seed(2000 + seed)
        p = ref_augment.draw(free_rotation=bool(square))
        out, masks = ref_augment.apply(image, [mask, mask2], p)
        assert np.array_equal(out, z[f"a{seed}_out"]) and np.array_equal(masks[0], z[f"a{seed}_mask_out"])
        if multi:
            assert np.array_equal(masks[1], z[f"a{seed}_mask_out2"])
        # the product draws the same decisions from the same module states
        random.seed(1000 + seed); np.random.seed(2000 + seed)
        q = A.draw_params(max_noise=0, free_rotation=bool(square))
        assert (q.flip_v, q.flip_h, q.rot, q.scale_on, q.blur_k) == (p["flip_v"], p["flip_h"], p["rot"], p["scale_on"], p["blur_k"])
        assert q.alpha == np.float32(p["alpha"]) and q.beta == np.float32(p["beta"]) and q.noise_max == 0
