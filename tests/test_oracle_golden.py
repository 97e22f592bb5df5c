"""Pin the CPU oracle (oracle/ref_im.py) against outputs of the reference itself.

The fixtures were produced by oracle/make_golden.py executing the reference's own
functions (functions.py:2832-3238).  CPU only."""
import ast
import os

import cv2
import numpy as np
import pytest

from oracle import ref_im


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.dtype == b.dtype, (a.dtype, b.dtype)
    assert a.shape == b.shape, (a.shape, b.shape)
    np.testing.assert_array_equal(a, b)


def test_kat_im_creation_figure(golden_dir):
    """README.md:16-17 / IM_creation.jpg: (a),(b) -> (d) IM = cells with sum 1, (e) = cells with sum 2."""
    g = load(golden_dir, "im_kat.npz")
    label, im, im_size, pred_size = ref_im.im_binary([g["a"][..., None], g["b"][..., None]])
    same(label, g["label"]); same(im, g["im"])
    assert int(im_size) == int(g["im_size"]) == 9
    assert int(pred_size) == int(g["pred_size"]) == 45   # 4+5+6+7+8+8+7 green cells in panel (e)
    expect_im = {(2, 4), (2, 5), (3, 3), (4, 2), (5, 2), (10, 4), (10, 5), (10, 6), (10, 7)}
    assert {tuple(int(v) for v in p) for p in np.argwhere(im == 255)} == expect_im


def test_im_binary(golden_dir):
    g = load(golden_dir, "im_binary.npz")
    for i in range(int(g["n"])):
        label, im, im_size, pred_size = ref_im.im_binary(list(g[f"{i}/masks"]))
        same(label, g[f"{i}/label"]); same(im, g[f"{i}/im"])
        assert isinstance(im_size, np.int64) and isinstance(pred_size, np.int64)
        assert im_size == g[f"{i}/im_size"] and pred_size == g[f"{i}/pred_size"]


def test_im_multiclass(golden_dir):
    g = load(golden_dir, "im_multiclass.npz")
    for i in range(int(g["n"])):
        label, im, im_size = ref_im.im_multiclass(list(g[f"{i}/masks"]))
        same(label, g[f"{i}/label"]); same(im, g[f"{i}/im"])
        assert isinstance(im_size, np.int64) and im_size == g[f"{i}/im_size"]


def test_prediction_binary(golden_dir):
    g = load(golden_dir, "predict.npz")
    for m in (1, 2, 3, 5):
        probs = g[f"binary/m{m}/probs"]
        for thr in (0.5, 0.3):
            label, im, im_size, pred_size = ref_im.im_prediction_binary([p[0] for p in probs], thr)
            tag = f"binary/m{m}/t{thr}"
            same(label, g[f"{tag}/label"]); same(im, g[f"{tag}/im"])
            assert im_size == g[f"{tag}/im_size"] and pred_size == g[f"{tag}/pred_size"]


def test_prediction_hela(golden_dir):
    g = load(golden_dir, "predict.npz")
    for m in (1, 2, 4):
        tag = f"hela/m{m}"
        alive, dead, pos, im, im_size = ref_im.im_prediction_hela([p[0] for p in g[f"{tag}/probs"]])
        same(alive, g[f"{tag}/alive"]); same(dead, g[f"{tag}/dead"]); same(pos, g[f"{tag}/pos"])
        same(im, g[f"{tag}/im"])
        assert im_size == g[f"{tag}/im_size"]


def test_prediction_multiclass(golden_dir):
    g = load(golden_dir, "predict.npz")
    for m, k in ((1, 9), (2, 9), (3, 35), (2, 35), (5, 2)):
        probs = g[f"multi/m{m}k{k}/probs"]
        for flt in (False, True):
            tag = f"multi/m{m}k{k}/f{int(flt)}"
            label, im, im_size, eq = ref_im.im_prediction_multiclass(list(probs), flt)
            same(label, g[f"{tag}/label"]); same(im, g[f"{tag}/im"])
            assert im_size == g[f"{tag}/im_size"]
            assert bool(eq) == bool(g[f"{tag}/lists_equal"])


def test_dilate_label(golden_dir):
    g = load(golden_dir, "dilate_mask.npz")
    for i in range(int(g["n"])):
        same(ref_im.dilate_label(g[f"{i}/label"]), g[f"{i}/out"])


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 7])
def test_morphology_matches_cv2(k):
    """The reference calls cv2.erode / cv2.dilate (functions.py:2858-2864); cv2 is
    the same third-party build on both sides, so it pins the NumPy restatement."""
    rng = np.random.default_rng(k)
    for shape in ((1, 1), (3, 5), (32, 48), (13, 26)):
        im = (rng.random(shape) < 0.6).astype(np.uint8) * 255
        ker = np.ones((k, k), np.uint8)
        same(ref_im.erode(im, k), cv2.erode(im, ker, iterations=1))
        same(ref_im.dilate(im, k), cv2.dilate(im, ker, iterations=1))


def _driver_via_oracle(kind, images, probs, kw):
    """The per-image body of the reference drivers restated with the oracle
    (functions.py:2844-2887, 2932-2980, 3020-3066), file I/O removed."""
    out = {}
    sizes = []
    for i in range(images.shape[0]):
        image = images[i]
        per_model = list(probs[i])
        ek, dk = kw.get("erode_kernel", 5), kw.get("dilate_kernel", 5)
        bi, bo = kw.get("block_input", True), kw.get("block_output", True)
        if kind == "binary":
            label, im, im_size, pred_size = ref_im.im_prediction_binary(per_model, 0.5)
            image, label, im = ref_im.blank_binary(image, label, im, ek, dk, bi, bo)
            write = ref_im.write_decision_binary(pred_size, im_size, kw.get("filter_bad_predictions", True))
            files = {"im": im}
            if write:
                files.update(images=image, masks=label)
        elif kind == "multiclass":
            flt = kw.get("filter_unequal_class_pred", False)
            label, im, im_size, eq = ref_im.im_prediction_multiclass([p[None] for p in per_model], flt)
            image, label, im = ref_im.blank_multiclass(image, label, im, ek, dk, bi, bo)
            files = {"im": im}
            if (not flt) or eq:
                files.update(images=image, masks=label)
        else:
            raise AssertionError(kind)
        sizes.append(im_size)
        out[i] = files
    return out, ref_im.mean_im_size(sizes)


@pytest.mark.parametrize("kind", ["binary", "multiclass"])
def test_driver_bodies(golden_dir, kind):
    g = load(golden_dir, "drivers.npz")
    names = [str(n) for n in g[f"{kind}/names"]]
    images, probs = g[f"{kind}/images"], g[f"{kind}/probs"]
    for j in range(int(g[f"{kind}/nruns"])):
        kw = dict(ast.literal_eval(str(g[f"{kind}/run{j}/kwargs"])))
        got, mean = _driver_via_oracle(kind, images, probs, kw)
        assert mean == float(g[f"{kind}/run{j}/mean_im_size"])
        for i, name in enumerate(names):
            for sub in ("images", "masks", "im"):
                key = f"{kind}/run{j}/{sub}/{name}"
                assert (key in g.files) == (sub in got[i]), (key, kw)
                if key in g.files:
                    same(got[i][sub], g[key])


def test_driver_body_hela_without_circles(golden_dir):
    """HeLa: everything except mod_position (host-side circle drawing, component #4)."""
    g = load(golden_dir, "drivers.npz")
    names = [str(n) for n in g["hela/names"]]
    images, probs = g["hela/images"], g["hela/probs"]
    for j in range(int(g["hela/nruns"])):
        kw = dict(ast.literal_eval(str(g[f"hela/run{j}/kwargs"])))
        sizes = []
        for i, name in enumerate(names):
            alive, dead, pos, cim, im_size = ref_im.im_prediction_hela(list(probs[i]))
            sizes.append(im_size)
            drawn = np.zeros(images[i].shape + (3,), np.uint8)
            bf, alive, dead, _, im = ref_im.blank_hela(images[i], alive, dead, drawn, cim,
                                                      kw["erode_kernel"], kw["dilate_kernel"],
                                                      kw.get("block_input", True), kw.get("block_output", True))
            same(bf, g[f"hela/run{j}/brightfield/{name}"])
            same(alive, g[f"hela/run{j}/alive/{name}"])
            same(dead, g[f"hela/run{j}/dead/{name}"])
            same(im, g[f"hela/run{j}/im/{name}"])
        assert ref_im.mean_im_size(sizes) == float(g[f"hela/run{j}/mean_im_size"])


def test_mean_im_size_bankers_rounding():
    """functions.py:2889 uses Python round(x, 0): halves go to even."""
    assert ref_im.mean_im_size([0, 1]) == 0.0        # 0.5 -> 0
    assert ref_im.mean_im_size([1, 2]) == 2.0        # 1.5 -> 2
    assert ref_im.mean_im_size([2, 3]) == 2.0        # 2.5 -> 2
    assert isinstance(ref_im.mean_im_size([np.int64(3)]), float)
