"""GPU parity tests of the SURVEY.md 8f rows: device confusion counts / benchmark_* drivers (8f-2), the augmentation
step (8f-1) and the bit-packed result layout, against fixtures the reference generated (oracle/make_golden_next.py),
against cv2 itself and against the NumPy oracle."""
import contextlib
import io
import os
import random

import cv2
import numpy as np
import pytest

from oracle import ref_metrics, ref_augment

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def F():
    from inconsistencymasks_b200 import functions
    return functions


@pytest.fixture(scope="module")
def E():
    from inconsistencymasks_b200 import evaluation
    return evaluation


@pytest.fixture(scope="module")
def A():
    from inconsistencymasks_b200 import augment
    return augment


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype, (a.shape, b.shape, a.dtype, b.dtype)
    assert np.array_equal(a, b), f"{int((a != b).sum())} of {a.size} elements differ"


# ------------------------------------------------------------------------------------------------ 8f-2 metrics
def test_metric_helpers_vs_reference_golden(E):
    z = np.load(os.path.join(G, "metrics.npz"))
    for i in range(int(z["n"])):
        gt, pred = z[f"b{i}_gt"], z[f"b{i}_pred"]
        iou, dice = E.get_IoU_binary(gt, pred), E.dice_score_numpy_binary(gt, pred)
        assert iou == z[f"b{i}_iou"] and round(iou, 4) == z[f"b{i}_iou4"]
        assert dice.dtype == z[f"b{i}_dice"].dtype and dice == z[f"b{i}_dice"]
        gtm, predm = z[f"m{i}_gt"], z[f"m{i}_pred"]
        assert E.get_IoU_multi_unique(predm, gtm) == z[f"m{i}_iou"]
        assert E.pixel_accuracy(predm, gtm) == z[f"m{i}_pa"]


@pytest.mark.parametrize("n,h,w", [(1, 16, 16), (5, 32, 48), (64, 256, 256), (3, 208, 416)])
def test_confusion_counts_vs_numpy(E, n, h, w):
    rng = np.random.default_rng(n * h)
    gt = rng.choice(np.array([0, 1, 127, 128, 255], np.uint8), size=(n, h, w))
    pred = rng.choice(np.array([0, 255, 200, 3], np.uint8), size=(n, h, w))
    c = E.seg_counts_binary(pred, gt)
    for i in range(n):
        want = [np.sum((gt[i] != 0) & (pred[i] != 0)), np.sum((gt[i] != 0) | (pred[i] != 0)), np.sum((gt[i] >= 128) & (pred[i] >= 128)),
                np.sum(gt[i] >= 128), np.sum(pred[i] >= 128)]
        assert list(c[i]) == [int(v) for v in want]
    gtm = rng.integers(0, 256, size=(n, h, w), dtype=np.uint8)
    predm = np.where(rng.random((n, h, w)) > 0.5, gtm, rng.integers(0, 256, size=(n, h, w))).astype(np.uint8)
    hist = E.seg_counts_multiclass(predm, gtm)
    for i in range(min(n, 4)):
        same(hist[i, 0], np.bincount(gtm[i].ravel(), minlength=256).astype(np.int64))
        same(hist[i, 1], np.bincount(predm[i].ravel(), minlength=256).astype(np.int64))
        same(hist[i, 2], np.bincount(gtm[i][gtm[i] == predm[i]].ravel(), minlength=256).astype(np.int64))
        assert E._iou_multi_unique(hist[i]) == ref_metrics.get_IoU_multi_unique(predm[i], gtm[i])


class BatchReplay:
    def __init__(self):
        self.table = {}

    def add(self, image, prob):
        self.table[np.ascontiguousarray(image).astype(np.uint8).tobytes()] = prob

    def predict(self, x, *a, **k):
        x = np.asarray(x)
        return np.stack([self.table[np.ascontiguousarray(x[i]).astype(np.uint8).tobytes()] for i in range(x.shape[0])])


def test_benchmark_drivers_vs_reference_golden(E, tmp_path):
    """benchmark_ISIC2018 / _multiclass / _hela on the directories the reference was run on, with the same replayed
    probability maps: returned numbers and written prediction files are the reference's."""
    z = np.load(os.path.join(G, "benchmarks.npz"))
    h = w = 64
    with contextlib.redirect_stdout(io.StringIO()):
        # ISIC
        idir, mdir, pdir = (str(tmp_path / "isic" / d) for d in ("images", "masks", "pred"))
        os.makedirs(idir); os.makedirs(mdir)
        model = BatchReplay()
        for i, (img, gt, prob) in enumerate(zip(z["isic_images"], z["isic_gt"], z["isic_probs"])):
            cv2.imwrite(os.path.join(idir, f"im_{i:03d}.png"), img); cv2.imwrite(os.path.join(mdir, f"im_{i:03d}.png"), gt)
            model.add(cv2.cvtColor(img, cv2.COLOR_BGR2RGB), prob)
        res = E.benchmark_ISIC2018(model, idir, mdir, pdir, h, w, 3, batch_size=4)
        assert [float(v) for v in res] == list(z["isic_result"])
        for i in range(len(z["isic_images"])):
            same(cv2.imread(os.path.join(pdir, f"im_{i:03d}.png"), 0), z["isic_pred"][i])
        # multiclass
        idir, mdir, pdir = (str(tmp_path / "mc" / d) for d in ("images", "masks", "pred"))
        os.makedirs(idir); os.makedirs(mdir)
        model = BatchReplay()
        for i, (img, gt, prob) in enumerate(zip(z["mc_images"], z["mc_gt"], z["mc_probs"])):
            cv2.imwrite(os.path.join(idir, f"im_{i:03d}.png"), img); cv2.imwrite(os.path.join(mdir, f"im_{i:03d}.png"), gt)
            model.add(cv2.cvtColor(img, cv2.COLOR_BGR2RGB), prob)
        mapping = {tuple(int(v) for v in row[:3]): int(row[3]) for row in z["mc_mapping"]}
        res = E.benchmark_multiclass(model, idir, mdir, pdir, h, w, 3, mapping, batch_size=4, print_results=False)
        assert [float(v) for v in res] == list(z["mc_result"])
        for i in range(len(z["mc_images"])):
            same(cv2.imread(os.path.join(pdir, f"im_{i:03d}.png"), 0), z["mc_pred"][i])
            same(cv2.imread(os.path.join(pdir, f"im_{i:03d}_color.png")), z["mc_color"][i])
        # HeLa
        root, pdir = str(tmp_path / "hela"), str(tmp_path / "hela_pred")
        for d in ("brightfield", "alive", "dead", "mod_position"):
            os.makedirs(os.path.join(root, d))
        model = BatchReplay()
        for i in range(len(z["hela_images"])):
            for d, m in (("brightfield", z["hela_images"][i]), ("alive", z["hela_gt_alive"][i]), ("dead", z["hela_gt_dead"][i]),
                         ("mod_position", z["hela_gt_pos"][i])):
                cv2.imwrite(os.path.join(root, d, f"c_{i:03d}.png"), m)
            model.add(z["hela_images"][i].reshape(h, w, 1), z["hela_probs"][i])
        res = E.benchmark_hela(model, root, pdir, h, w, 1, batch_size=3)
        assert [float(v) for v in res] == list(z["hela_result"])
        for d in ("alive", "dead", "mod_position"):
            for i in range(len(z["hela_images"])):
                same(cv2.imread(os.path.join(pdir, d, f"c_{i:03d}.png"), 0), z[f"hela_pred_{d}"][i])


def test_benchmark_with_b200_model_vs_oracle(E, tmp_path):
    """The fused path (threshold / argmax inside the forward, counts on the device) gives the numbers the reference's
    loop gives on the same model's own probabilities."""
    from inconsistencymasks_b200 import unet as U
    h = w = 64
    rng = np.random.default_rng(9)
    n = 37
    with contextlib.redirect_stdout(io.StringIO()):
        idir, mdir, pdir = (str(tmp_path / "isic" / d) for d in ("images", "masks", "pred"))
        os.makedirs(idir); os.makedirs(mdir)
        model = U.B200UNet(h, w, 3, 1, 0.5, "sigmoid", U.init_weights(3, 1, 0.5, seed=3))
        imgs = rng.integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)
        gts = (rng.random((n, h, w)) > 0.5).astype(np.uint8) * 255
        for i in range(n):
            cv2.imwrite(os.path.join(idir, f"i{i:03d}.png"), imgs[i]); cv2.imwrite(os.path.join(mdir, f"i{i:03d}.png"), gts[i])
        miou, mdice = E.benchmark_ISIC2018(model, idir, mdir, pdir, h, w, 3, batch_size=2)
        names = os.listdir(idir)
        probs = model.predict(np.ascontiguousarray(np.stack([cv2.cvtColor(cv2.imread(os.path.join(idir, nm)), cv2.COLOR_BGR2RGB) for nm in names])))
        ious, dices = [], []
        for i, nm in enumerate(names):
            pred = ((probs[i] > 0.5) * 255).astype(np.uint8).squeeze()
            gt = cv2.imread(os.path.join(mdir, nm), 0)
            same(cv2.imread(os.path.join(pdir, nm), 0), pred)
            dices.append(round(ref_metrics.dice_score_numpy_binary(gt, pred), 4)); ious.append(round(ref_metrics.get_IoU_binary(gt, pred), 4))
        assert miou == round(np.sum(ious) / len(ious), 3) and mdice == round(np.sum(dices) / len(dices), 3)
        # multiclass
        k = 9
        idir, mdir, pdir = (str(tmp_path / "mc" / d) for d in ("images", "masks", "pred"))
        os.makedirs(idir); os.makedirs(mdir)
        model = U.B200UNet(h, w, 3, k, 1.0, "softmax", U.init_weights(3, k, 1.0, seed=4))
        gtm = rng.integers(0, k, size=(n, h, w), dtype=np.uint8)
        for i in range(n):
            cv2.imwrite(os.path.join(idir, f"i{i:03d}.png"), imgs[i]); cv2.imwrite(os.path.join(mdir, f"i{i:03d}.png"), gtm[i])
        mpa, miou = E.benchmark_multiclass(model, idir, mdir, pdir, h, w, 3, {(10 * j, 20, 30): j for j in range(k)}, batch_size=8, print_results=False)
        names = os.listdir(idir)
        probs = model.predict(np.ascontiguousarray(np.stack([cv2.cvtColor(cv2.imread(os.path.join(idir, nm)), cv2.COLOR_BGR2RGB) for nm in names])))
        pas, ious = [], []
        for i, nm in enumerate(names):
            pred = np.argmax(probs[i], axis=-1)
            gt = cv2.imread(os.path.join(mdir, nm), 0)
            same(cv2.imread(os.path.join(pdir, nm), 0), pred.astype(np.uint8))
            pas.append(round(ref_metrics.pixel_accuracy(pred, gt), 4)); ious.append(round(ref_metrics.get_IoU_multi_unique(pred, gt), 4))
        assert mpa == round(np.sum(pas) / len(pas), 3) and miou == round(np.sum(ious) / len(ious), 3)


# ------------------------------------------------------------------------------------------------ 8f-1 augmentation
def test_augment_vs_reference_golden(A):
    """Seeded runs of the reference's augment_image_and_mask(s) (max_noise = 0): same module seeds -> same pixels."""
    z = np.load(os.path.join(G, "augment.npz"))
    for seed, square, multi in (tuple(int(v) for v in row) for row in z["cases"]):
        image, mask, mask2 = z[f"a{seed}_image"], z[f"a{seed}_mask"], z[f"a{seed}_mask2"]
        random.seed(1000 + seed); np.random.seed(2000 + seed)
        if multi:
            out, masks = A.augment_image_and_masks(image.copy(), [mask.copy(), mask2.copy()], max_noise=0, free_rotation=bool(square))
            same(masks[1], z[f"a{seed}_mask_out2"]); mask_out = masks[0]
        else:
            out, mask_out = A.augment_image_and_mask(image.copy(), mask.copy(), max_noise=0, free_rotation=bool(square))
        same(out, z[f"a{seed}_out"]); same(mask_out, z[f"a{seed}_mask_out"])


@pytest.mark.parametrize("h,w,c", [(64, 64, 3), (256, 256, 3), (48, 48, 1), (208, 416, 3), (33, 70, 3)])
def test_augment_batch_vs_cv2(F, A, h, w, c):
    """Every operation and combination against the cv2 calls of the reference, a batch at a time on the device."""
    from inconsistencymasks_b200._lib import AugParams
    rng = np.random.default_rng(h + c)
    n = 40
    imgs = rng.integers(0, 256, size=(n, h, w, c), dtype=np.uint8)
    masks = rng.integers(0, 35, size=(2, n, h, w), dtype=np.uint8)
    params, dicts = [], []
    for i in range(n):
        d = dict(flip_v=int(rng.integers(0, 2)), flip_h=int(rng.integers(0, 2)), rot=int(rng.integers(0, 4)) if h == w else int(rng.choice([0, 2])),
                 scale_on=int(rng.integers(0, 2)), alpha=float(rng.uniform(0.5, 1.5)), beta=float(rng.uniform(-25, 25)),
                 blur_k=int(rng.choice([0, 3, 5, 7])))
        p = AugParams(); p.flip_v, p.flip_h, p.rot, p.scale_on, p.alpha, p.beta, p.blur_k = (d[k] for k in ("flip_v", "flip_h", "rot", "scale_on", "alpha", "beta", "blur_k"))
        params.append(p); dicts.append(d)
    out, mo = A.augment_batch(F._dev(imgs), F._dev(masks), params)
    out, mo = out.cpu().numpy(), mo.cpu().numpy()
    for i in range(n):
        img = imgs[i] if c == 3 else imgs[i, ..., 0]
        want, wm = ref_augment.apply(img, [masks[0, i], masks[1, i]], dicts[i])
        got = out[i] if c == 3 else out[i, ..., 0]
        same(got.reshape(want.shape), want)
        same(mo[0, i].reshape(wm[0].shape), wm[0]); same(mo[1, i].reshape(wm[1].shape), wm[1])


def test_augment_noise_contract(F, A):
    """Noise: uniform integers in [-max_noise, max_noise) added and clipped; same seed -> same pixels, another seed ->
    other pixels; the mask never sees it."""
    from inconsistencymasks_b200._lib import AugParams
    h = w = 128
    img = np.full((3, h, w, 3), 128, np.uint8)
    ps = []
    for seed in (7, 7, 8):
        p = AugParams(); p.noise_max = 25; p.seed = seed
        ps.append(p)
    out, _ = A.augment_batch(F._dev(img), None, ps)
    out = out.cpu().numpy().astype(np.int32) - 128
    assert np.array_equal(out[0], out[1]) and not np.array_equal(out[0], out[2])
    assert out.min() == -25 and out.max() == 24
    counts = np.bincount((out[0] + 25).ravel(), minlength=50)
    assert counts.min() > 0.8 * out[0].size / 50 and counts.max() < 1.2 * out[0].size / 50      # flat histogram
    assert abs(float(np.corrcoef(out[0, :, :-1].ravel(), out[0, :, 1:].ravel())[0, 1])) < 0.02   # no neighbour correlation
    lo = np.full((1, h, w, 3), 3, np.uint8); hi = np.full((1, h, w, 3), 250, np.uint8)
    assert A.augment_batch(F._dev(lo), None, ps[:1])[0].cpu().numpy().min() == 0
    assert A.augment_batch(F._dev(hi), None, ps[:1])[0].cpu().numpy().max() == 255
    with pytest.raises(Exception):
        q = AugParams(); q.rot = 1
        A.augment_batch(F._dev(np.zeros((1, 32, 64, 3), np.uint8)), None, [q])


def test_augment_consumes_pseudo_label_batch_on_device(F, A):
    """8f-1's point: the blanked image and label of the IM kernels go straight into the augmentation, no host hop."""
    import ctypes as C
    import torch
    from inconsistencymasks_b200 import unet as U
    from inconsistencymasks_b200._lib import lib, check, AugParams
    h = w = 64; n = 6
    rng = np.random.default_rng(1)
    models = [U.B200UNet(h, w, 3, 1, 0.5, "sigmoid", U.init_weights(3, 1, 0.5, seed=20 + j)) for j in range(2)]
    imgs = rng.integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)
    d_img = F._dev(imgs)
    d_out = torch.empty_like(d_img); d_lab = torch.empty((1, n, h, w), dtype=torch.uint8, device="cuda")
    d_im = torch.empty((n, h, w), dtype=torch.uint8, device="cuda"); d_sz = torch.empty(n, dtype=torch.int64, device="cuda")
    d_pred = torch.empty((1, n), dtype=torch.int64, device="cuda")
    check(lib.imk_ensemble_im_binary(F._handles(models), 2, d_img.data_ptr(), n, 1, 0.5, 1, 1, 1, d_out.data_ptr(), d_lab.data_ptr(),
                                     d_im.data_ptr(), d_sz.data_ptr(), d_pred.data_ptr(), F._stream()))
    ps, ds = [], []
    for i in range(n):
        d = dict(flip_v=i & 1, flip_h=(i >> 1) & 1, rot=i % 4, scale_on=1, alpha=1.2, beta=-7.0, blur_k=[0, 3, 5, 7][i % 4])
        p = AugParams(); p.flip_v, p.flip_h, p.rot, p.scale_on, p.alpha, p.beta, p.blur_k = d["flip_v"], d["flip_h"], d["rot"], 1, 1.2, -7.0, d["blur_k"]
        ps.append(p); ds.append(d)
    a_img, a_lab = A.augment_batch(d_out, d_lab, ps)
    blanked, lab = d_out.cpu().numpy(), d_lab.cpu().numpy()
    for i in range(n):
        want, wm = ref_augment.apply(blanked[i], [lab[0, i]], ds[i])
        same(a_img[i].cpu().numpy(), want); same(a_lab[0, i].cpu().numpy(), wm[0])


# ------------------------------------------------------------------------------------------------ packed layout
def test_pack_bits_roundtrip(F):
    import torch
    from inconsistencymasks_b200._lib import lib, check
    rng = np.random.default_rng(0)
    planes = (rng.random((3, 5, 64, 48)) > 0.5).astype(np.uint8) * 255
    d = F._dev(planes)
    bits = torch.empty(planes.size // 8, dtype=torch.uint8, device="cuda")
    check(lib.imk_pack_bits(d.data_ptr(), planes.size, bits.data_ptr(), F._stream()))
    back = np.unpackbits(bits.cpu().numpy(), bitorder="little").reshape(planes.shape) * 255
    same(back.astype(np.uint8), planes)
    same(bits.cpu().numpy(), np.packbits(planes.ravel() > 0, bitorder="little"))


@pytest.mark.parametrize("kind,c,K,act", [("binary", 3, 1, "sigmoid"), ("hela", 1, 3, "sigmoid"), ("multiclass", 3, 9, "softmax")])
def test_packed_host_pipeline_equals_u8_layout(F, kind, c, K, act):
    """imk_pseudo_label_*_host_packed: the bit planes unpack to exactly the uint8 planes of the standard call, statistics
    identical; img_out may be NULL (the host blanks its own copy with the IM)."""
    from inconsistencymasks_b200 import unet as U
    from inconsistencymasks_b200._lib import lib, check
    h, w, n = 64, 48, 21
    rng = np.random.default_rng(K)
    models = [U.B200UNet(h, w, c, K, 1.0, act, U.init_weights(c, K, 1.0, seed=60 + j)) for j in range(2)]
    imgs = rng.integers(0, 256, size=(n, h, w, c), dtype=np.uint8)
    r = F._run_batch(models, imgs, kind, blank_image=imgs, block_input=True, block_output=True)
    planes = 3 if kind == "hela" else 1
    hs = F._handles(models)
    im_bits = np.zeros(n * h * w // 8, np.uint8); sz = np.zeros(n, np.int64)
    if kind == "multiclass":
        lab = np.zeros((1, n, h, w), np.uint8)
        check(lib.imk_pseudo_label_multiclass_host_packed(hs, 2, imgs.ctypes.data, n, 0, 1, 1, None, lab.ctypes.data, im_bits.ctypes.data,
                                                          sz.ctypes.data, None, 8))
        same(lab, r.labels)
    else:
        lab_bits = np.zeros(planes * n * h * w // 8, np.uint8); pred = np.zeros((planes, n), np.int64)
        check(lib.imk_pseudo_label_binary_host_packed(hs, 2, imgs.ctypes.data, n, 0, 0.5, 1 if kind == "binary" else 0, 1, 1, None,
                                                      lab_bits.ctypes.data, im_bits.ctypes.data, sz.ctypes.data, pred.ctypes.data, 8))
        same((np.unpackbits(lab_bits, bitorder="little") * 255).astype(np.uint8).reshape(planes, n, h, w), r.labels)
        same(pred, r.pred_size)
    im = (np.unpackbits(im_bits, bitorder="little") * 255).astype(np.uint8).reshape(n, h, w)
    same(im, r.im); same(sz, r.im_size)
    blanked = imgs.copy(); blanked[im > 0] = 0                  # functions.py:2867 on the host's own copy
    same(blanked, r.image)


# ------------------------------------------------------------------------------------------------ 8f-4 EvalNet
@pytest.mark.parametrize("h,w,ca,cb,alpha,heads", [(64, 64, 3, 1, 1.0, 1), (128, 64, 3, 1, 2.0, 1), (64, 128, 1, 3, 2.0, 2), (64, 64, 3, 9, 1.0, 2),
                                                   (256, 256, 3, 1, 1.0, 1), (208, 416, 3, 35, 2.0, 2)])
def test_evalnet_forward_vs_fp32_oracle(h, w, ca, cb, alpha, heads):
    """get_evalnet / get_evalnet_miou against the torch fp32 restatement (oracle/ref_evalnet.py): sigmoid scores within
    5e-3 (fp16 activations, fp32 accumulation).  The one-hot input goes through the look-up layer."""
    from oracle import ref_evalnet
    from inconsistencymasks_b200 import evalnet as EV
    rng = np.random.default_rng(h + cb)
    n = 5 if h < 200 else 2
    a = rng.integers(0, 256, size=(n, h, w, ca), dtype=np.uint8)
    weights = EV.init_evalnet_weights(ca, cb, alpha, heads, seed=cb)
    if heads == 1:
        b = (rng.random((n, h, w, cb)) > 0.5).astype(np.uint8) * 255
        model = EV.get_evalnet(h, w, ca, cb, alpha, weights=weights)
        got = model.predict([a, b])
        want = ref_evalnet.forward(a, b, weights, alpha, 1, True, True)
        assert got.shape == (n, 1) and got.dtype == np.float32
        assert float(np.abs(got - want).max()) <= 5e-3, float(np.abs(got - want).max())
    else:
        cls = rng.integers(0, cb, size=(n, h, w))
        b = np.stack([(cls == k).astype(np.int32) for k in range(cb)], axis=-1)       # functions.py:6005
        b[0, :4, :4, :] = 0                                                            # pixels outside every class
        model = EV.get_evalnet_miou(h, w, ca, cb, alpha, weights=weights)
        iou, det = model.predict([a, b])
        w_iou, w_det = ref_evalnet.forward(a, b, weights, alpha, 2, True, False)
        assert iou.shape == (n, cb) and det.shape == (n, cb)
        assert float(np.abs(iou - w_iou).max()) <= 5e-3 and float(np.abs(det - w_det).max()) <= 5e-3
    from inconsistencymasks_b200.weights import evalnet_plan
    assert model.count_params() == sum((it[1] ** 2 * it[2] * it[3] + it[3]) if it[0] == "conv" else (4 * it[1] if it[0] == "bn" else it[1] * it[2] + it[2])
                                       for it in evalnet_plan(ca, cb, alpha, heads))


def test_augment_directory_drivers_vs_reference_golden(A, tmp_path):
    """create_augment_images_and_masks_ISIC_2018 / _hela on the directories the reference was run on (one file each, so that
    os.listdir order cannot matter), same module seeds, max_noise = 0: every written PNG is the reference's."""
    z = np.load(os.path.join(G, "augment.npz"))
    i_dir, m_dir, o_dir = tmp_path / "i", tmp_path / "m", tmp_path / "o"
    i_dir.mkdir(); m_dir.mkdir()
    cv2.imwrite(str(i_dir / "one.png"), z["drv_isic_image"]); cv2.imwrite(str(m_dir / "one.png"), z["drv_isic_mask"])
    random.seed(77); np.random.seed(78)
    A.create_augment_images_and_masks_ISIC_2018(str(i_dir), str(m_dir), str(o_dir), 5, True, max_noise=0)
    for n in range(5):
        same(cv2.imread(str(o_dir / "images" / f"one_aug_{n}.png")), z[f"drv_isic_out_{n}"])
        same(cv2.imread(str(o_dir / "masks" / f"one_aug_{n}.png")), z[f"drv_isic_mask_out_{n}"])
    same(cv2.imread(str(o_dir / "images" / "one.png")), z["drv_isic_copy"])
    h_in, h_out = tmp_path / "h", tmp_path / "ho"
    for d, key in (("brightfield", "drv_hela_bf"), ("alive", "drv_hela_m0"), ("dead", "drv_hela_m1"), ("mod_position", "drv_hela_m2")):
        (h_in / d).mkdir(parents=True)
        cv2.imwrite(str(h_in / d / "c.png"), z[key])
    random.seed(79); np.random.seed(80)
    A.create_augment_images_and_masks_hela(str(h_in), str(h_out), 4, False, max_noise=0)
    for n in range(4):
        for d in ("brightfield", "alive", "dead", "mod_position"):
            same(cv2.imread(str(h_out / d / f"c_aug_{n}.png")), z[f"drv_hela_{d}_{n}"])
    assert not (h_out / "brightfield" / "c.png").exists()                      # copy_org=False
    # many files, noise on: right number of outputs, masks only moved (same multiset of values per file)
    many_i, many_m, many_o = tmp_path / "mi", tmp_path / "mm", tmp_path / "mo"
    many_i.mkdir(); many_m.mkdir()
    rng = np.random.default_rng(3)
    for k in range(23):
        cv2.imwrite(str(many_i / f"f{k:02d}.png"), rng.integers(0, 256, size=(64, 64, 3), dtype=np.uint8))
        cv2.imwrite(str(many_m / f"f{k:02d}.png"), rng.integers(0, 9, size=(64, 64), dtype=np.uint8))
    A.create_augment_images_and_masks_multiclass(str(many_i), str(many_m), str(many_o), 3, True)
    assert len(os.listdir(many_o / "images")) == 23 * 4 and len(os.listdir(many_o / "masks")) == 23 * 4
    for k in (0, 11, 22):
        src = cv2.imread(str(many_m / f"f{k:02d}.png"))
        for n in range(3):
            out = cv2.imread(str(many_o / "masks" / f"f{k:02d}_aug_{n}.png"))
            assert np.array_equal(np.bincount(out.ravel(), minlength=9), np.bincount(src.ravel(), minlength=9))


class ScoreNet:
    """Replays stored EvalNet outputs by image content, a batch at a time."""

    def __init__(self):
        self.table = {}

    def add(self, image, out):
        self.table[np.ascontiguousarray(image).astype(np.uint8).tobytes()] = out

    def predict(self, x, *a, **k):
        a0 = np.asarray(x[0])
        outs = [self.table[np.ascontiguousarray(a0[i]).astype(np.uint8).tobytes()] for i in range(a0.shape[0])]
        if isinstance(outs[0], tuple):
            return [np.stack([o[0] for o in outs]), np.stack([o[1] for o in outs])]
        return np.stack(outs)


def test_impp_drivers_vs_reference_golden(F, tmp_path):
    """IM++: EvalNet ensemble score -> number of augmented copies per file, as the reference decided on the same replayed
    scores; each copy is the file itself or its horizontal flip (augmentation reduced to the flip in this run), image and
    mask moved together."""
    z = np.load(os.path.join(G, "impp.npz"))
    h = w = 32
    # binary (ISIC)
    root, out = tmp_path / "b", tmp_path / "bo"
    (root / "images").mkdir(parents=True); (root / "masks").mkdir()
    nets = [ScoreNet(), ScoreNet()]
    n = len(z["b_images"])
    for i in range(n):
        cv2.imwrite(str(root / "images" / f"p{i}.png"), z["b_images"][i]); cv2.imwrite(str(root / "masks" / f"p{i}.png"), z["b_masks"][i])
        for k, net in enumerate(nets):
            net.add(cv2.cvtColor(z["b_images"][i], cv2.COLOR_BGR2RGB), z["b_scores"][i] + np.float32(0.02 * k))
    F.create_augment_images_and_masks_with_evalnet_ensemble_binary(nets, h, w, 3, 0.3, 0.8, str(root), str(out), (1.0, 1.0), (0.0, 0.0), 0, 0, False, True)
    counts = [len([f for f in os.listdir(out / "images") if f.startswith(f"p{i}___")]) for i in range(n)]
    assert counts == list(z["b_counts"])
    for i in range(n):
        for j in range(counts[i]):
            im, mk = cv2.imread(str(out / "images" / f"p{i}___{j}.png")), cv2.imread(str(out / "masks" / f"p{i}___{j}.png"), 0)
            flipped = np.array_equal(im, z["b_images"][i][:, ::-1])
            assert flipped or np.array_equal(im, z["b_images"][i])
            same(mk, np.ascontiguousarray(z["b_masks"][i][:, ::-1]) if flipped else z["b_masks"][i])
    # multiclass (SUIM / Cityscapes)
    k = z["m_ious"].shape[1]
    root, out = tmp_path / "m", tmp_path / "mo"
    (root / "images").mkdir(parents=True); (root / "masks").mkdir()
    nets = [ScoreNet(), ScoreNet(), ScoreNet()]
    for i in range(n):
        cv2.imwrite(str(root / "images" / f"q{i}.png"), z["m_images"][i]); cv2.imwrite(str(root / "masks" / f"q{i}.png"), z["m_masks"][i])
        for j, net in enumerate(nets):
            net.add(cv2.cvtColor(z["m_images"][i], cv2.COLOR_BGR2RGB), (z["m_ious"][i] + np.float32(0.01 * j), z["m_dets"][i]))
    F.create_augment_images_and_masks_with_evalnet_ensemble_multiclass(nets, h, w, 3, k, 0.2, 0.6, str(root), str(out), (1.0, 1.0), (0.0, 0.0), 0, 0, False, True)
    counts = [len([f for f in os.listdir(out / "images") if f.startswith(f"q{i}___")]) for i in range(n)]
    assert counts == list(z["m_counts"])


def test_impp_driver_with_b200_evalnets(F, tmp_path):
    """The same driver with real EvalNets on the device: runs end to end, 1..5 copies per file, masks keep their classes."""
    from inconsistencymasks_b200 import evalnet as EV
    h = w = 64; k = 4
    rng = np.random.default_rng(8)
    root, out = tmp_path / "in", tmp_path / "out"
    (root / "images").mkdir(parents=True); (root / "masks").mkdir()
    for i in range(12):
        cv2.imwrite(str(root / "images" / f"s{i}.png"), rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8))
        cv2.imwrite(str(root / "masks" / f"s{i}.png"), rng.integers(0, k, size=(h, w), dtype=np.uint8))
    nets = [EV.get_evalnet_miou(h, w, 3, k, 1.0, seed=s) for s in (1, 2)]
    F.create_augment_images_and_masks_with_evalnet_ensemble_multiclass(nets, h, w, 3, k, 0.3, 0.7, str(root), str(out))
    for i in range(12):
        c = len([f for f in os.listdir(out / "images") if f.startswith(f"s{i}___")])
        assert 1 <= c <= 5 and c == len([f for f in os.listdir(out / "masks") if f.startswith(f"s{i}___")])
        src = cv2.imread(str(root / "masks" / f"s{i}.png"), 0)
        assert np.array_equal(np.bincount(cv2.imread(str(out / "masks" / f"s{i}___0.png"), 0).ravel(), minlength=k), np.bincount(src.ravel(), minlength=k))


def test_evalnet_save_load_roundtrip(tmp_path):
    from inconsistencymasks_b200 import evalnet as EV
    rng = np.random.default_rng(0)
    m = EV.get_evalnet_miou(64, 64, 3, 5, 1.0, seed=3)
    a = rng.integers(0, 256, size=(2, 64, 64, 3), dtype=np.uint8)
    cls = rng.integers(0, 5, size=(2, 64, 64))
    b = np.stack([(cls == k).astype(np.int32) for k in range(5)], axis=-1)
    want = m.predict([a, b])
    m.save_weights(str(tmp_path / "e.npz"))
    m2 = EV.load_evalnet(str(tmp_path / "e.npz"))
    got = m2.predict([a, b])
    same(got[0], want[0]); same(got[1], want[1])
