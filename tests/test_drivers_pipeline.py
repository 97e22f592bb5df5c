"""The pipelined / sharded per-directory drivers (row a7-a10 + SURVEY.md 8f-3): host-side logic on CPU, and on the GPU
box a ~2k-file directory whose outputs must be byte-identical between the single-process run, a 2-rank sharded run and
the per-image reference arithmetic."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_names_partition():
    """Every file is owned by exactly one rank, whatever order os.listdir returned."""
    from inconsistencymasks_b200 import build
    build.build()
    from inconsistencymasks_b200 import functions as F
    names = [f"img_{i:04d}.png" for i in range(37)]
    shuffled = list(np.random.default_rng(0).permutation(names))
    for world in (1, 2, 3, 8):
        owned = []
        for rank in range(world):
            mine, r, w = F._shard_names(shuffled, (rank, world))
            assert (r, w) == ((rank, world) if world > 1 else (0, 1))
            owned += mine
        assert sorted(owned) == names
    assert F._shard_names(shuffled, False)[0] == shuffled
    assert F._shard_names(shuffled, None)[0] == shuffled          # no process group: unsharded


def test_mean_im_size_rounding():
    from inconsistencymasks_b200 import functions as F
    # functions.py:2889: Python round(x, 0) is banker's rounding
    assert F._mean_im_size({"a": 1, "b": 2}) == 2.0 and F._mean_im_size({"a": 2, "b": 3}) == 2.0
    assert isinstance(F._mean_im_size({"a": 5}), float)


# ------------------------------------------------------------------------------------------- GPU
def _make_pngs(path, n, h, w, c, seed):
    import cv2
    os.makedirs(path, exist_ok=True)
    rng = np.random.default_rng(seed)
    # smooth-ish content so that PNG encode / decode cost resembles real images
    base = rng.integers(0, 256, size=(n, h // 8, w // 8, c), dtype=np.uint8)
    for i in range(n):
        img = cv2.resize(base[i], (w, h), interpolation=cv2.INTER_LINEAR)
        img = np.clip(img.astype(np.int16).reshape(h, w, c) + rng.integers(-6, 7, size=(h, w, c)), 0, 255).astype(np.uint8)
        cv2.imwrite(os.path.join(path, f"im_{i:05d}.png"), img if c == 3 else img[..., 0])


def _tree(path):
    out = {}
    for sub in sorted(os.listdir(path)):
        for nm in sorted(os.listdir(os.path.join(path, sub))):
            out[f"{sub}/{nm}"] = open(os.path.join(path, sub, nm), "rb").read()
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["binary", "multiclass", "hela"])
def test_directory_driver_matches_per_image_reference(tmp_path, kind):
    import cv2
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from inconsistencymasks_b200 import functions as F, unet as U
    from oracle import ref_im
    h = w = 64
    n = 2100 if kind == "binary" else 700                     # several 512-file batches, a short last one
    c, K, act = {"binary": (3, 1, "sigmoid"), "multiclass": (3, 9, "softmax"), "hela": (1, 3, "sigmoid")}[kind]
    src = str(tmp_path / "in")
    _make_pngs(src, n, h, w, c, seed=3)
    models = [U.B200UNet(h, w, c, K, 0.5, act, U.init_weights(c, K, 0.5, seed=60 + j)) for j in range(2)]
    fn = {"binary": F.create_pseudo_labels_im_ISIC_2018, "multiclass": F.create_pseudo_labels_im_multiclass,
          "hela": F.create_pseudo_labels_im_hela}[kind]
    kw = dict(erode_kernel=0, dilate_kernel=0)
    if kind == "binary":
        kw["filter_bad_predictions"] = True
    dst = str(tmp_path / "out")
    mean = fn(models, h, w, c, src, dst, **kw)
    # per-image reference arithmetic on the probabilities .predict returns, for a sample of the files
    names = sorted(os.listdir(src))
    sizes = {}
    rng = np.random.default_rng(1)
    sample = set(rng.choice(len(names), size=40, replace=False).tolist())
    for i, nm in enumerate(names):
        img = cv2.imread(os.path.join(src, nm), 0 if c == 1 else 1).reshape(h, w, c)
        if i in sample or kind != "binary":
            fed = img[..., ::-1] if c == 3 else img
            probs = [m.predict(np.ascontiguousarray(fed)[None])[0] for m in models]
            if kind == "binary":
                lab, im, sz, pred = ref_im.im_prediction_binary(probs, 0.5)
                img_b, lab_b, _ = ref_im.blank_binary(img, lab, im)
                write = pred > sz and pred > 0
                assert os.path.exists(os.path.join(dst, "images", nm)) == bool(write)
                if write:
                    np.testing.assert_array_equal(cv2.imread(os.path.join(dst, "images", nm)), img_b)
                    np.testing.assert_array_equal(cv2.imread(os.path.join(dst, "masks", nm), 0), lab_b)
                np.testing.assert_array_equal(cv2.imread(os.path.join(dst, "im", nm), 0), im)
            elif kind == "multiclass":
                lab, im, sz, _ = ref_im.im_prediction_multiclass(probs)
                img_b, lab_b, _ = ref_im.blank_multiclass(img, lab, im)
                if i in sample:
                    np.testing.assert_array_equal(cv2.imread(os.path.join(dst, "images", nm)), img_b)
                    np.testing.assert_array_equal(cv2.imread(os.path.join(dst, "masks", nm), 0), lab_b)
                    np.testing.assert_array_equal(cv2.imread(os.path.join(dst, "im", nm), 0), im)
            else:
                alive, dead, pos, im, sz = ref_im.im_prediction_hela(probs)
                if i in sample:
                    bf, alive_b, dead_b, _, _ = ref_im.blank_hela(img[..., 0], alive, dead, np.zeros((h, w, 3), np.uint8), im)
                    np.testing.assert_array_equal(cv2.imread(os.path.join(dst, "brightfield", nm), 0), bf)
                    np.testing.assert_array_equal(cv2.imread(os.path.join(dst, "alive", nm), 0), alive_b)
                    np.testing.assert_array_equal(cv2.imread(os.path.join(dst, "im", nm), 0), im)
            sizes[nm[:-4]] = int(sz)
    if kind != "binary":
        assert mean == round(sum(sizes.values()) / len(sizes), 0)
    assert len(os.listdir(os.path.join(dst, "im"))) == n


_SHARD_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
rank, world = int(sys.argv[1]), int(sys.argv[2])
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT={port!r}, RANK=str(rank), WORLD_SIZE=str(world))
dist.init_process_group("gloo", rank=rank, world_size=world)
from inconsistencymasks_b200 import functions as F, unet as U
h = w = 64
models = [U.B200UNet(h, w, 3, 1, 0.5, "sigmoid", U.init_weights(3, 1, 0.5, seed=60 + j)) for j in range(2)]
mean = F.create_pseudo_labels_im_ISIC_2018(models, h, w, 3, {src!r}, {dst!r}, erode_kernel=0, dilate_kernel=0, filter_bad_predictions=False)
open({dst!r} + f".mean{{rank}}", "w").write(repr(mean))
dist.destroy_process_group()
"""


@pytest.mark.gpu
def test_sharded_directory_equals_single_process(tmp_path):
    """Two ranks (gloo rendezvous, both on cuda:0) shard one directory: the union of their files and the all-reduced
    mean_im_size equal the single-process run, byte for byte."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from inconsistencymasks_b200 import functions as F, unet as U
    h = w = 64
    src = str(tmp_path / "in")
    _make_pngs(src, 601, h, w, 3, seed=5)
    models = [U.B200UNet(h, w, 3, 1, 0.5, "sigmoid", U.init_weights(3, 1, 0.5, seed=60 + j)) for j in range(2)]
    single = str(tmp_path / "single")
    mean1 = F.create_pseudo_labels_im_ISIC_2018(models, h, w, 3, src, single, erode_kernel=0, dilate_kernel=0, filter_bad_predictions=False)
    sharded = str(tmp_path / "sharded")
    code = _SHARD_WORKER.format(root=ROOT, port="29533", src=src, dst=sharded)
    procs = [subprocess.Popen([sys.executable, "-c", code, str(r), "2"]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    assert _tree(single) == _tree(sharded)
    for r in range(2):
        assert float(open(sharded + f".mean{r}").read()) == mean1
