"""GPU parity of the U-Net forward (row a1) and of the fused ensemble path.

Oracle: oracle/ref_unet.py, a PyTorch fp32 CPU restatement of the reference's unet.py.
The CUDA path keeps activations in fp16 with fp32 accumulation (the reference runs
mixed_float16, 09_ISIC_2018_IM.py:16), so probabilities are compared within a STATED
tolerance: |p_cuda - p_fp32| <= PROB_ATOL.  Everything downstream of the probabilities
(threshold / argmax / IM / blanking / sizes) is compared bit for bit on identical
probabilities, as BASELINE.json's north_star asks.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
from oracle import ref_im, ref_unet  # noqa: E402

PROB_ATOL = 2e-2          # stated tolerance on probabilities (fp16 activations vs fp32 oracle)
MEAN_ATOL = 2e-3


@pytest.fixture(scope="module")
def U():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from inconsistencymasks_b200 import unet
    return unet


@pytest.fixture(scope="module")
def F():
    from inconsistencymasks_b200 import functions
    return functions


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.dtype == b.dtype and a.shape == b.shape, (a.dtype, b.dtype, a.shape, b.shape)
    np.testing.assert_array_equal(a, b)


CASES = [
    # h, w, c, K, alpha, act           (reference configs at reduced resolution + odd widths)
    (32, 32, 3, 1, 0.5, "sigmoid"),     # ISIC, alpha 0.5: widths 8..128 (fused engine: 8-channel maps as one plane, taps paired)
    (64, 48, 3, 1, 0.25, "sigmoid"),    # widths 4, 8, 16, ..: 8-channel planes on two levels, a 4-channel map inside one plane
    (64, 96, 1, 3, 0.5, "sigmoid"),     # grayscale + 8 channels: the input-block table feeds a single plane
    (16, 16, 3, 1, 0.5, "sigmoid"),     # smallest legal input with 8-channel planes (2x2 / 1x1 maps at the bottom)
    (208, 416, 3, 1, 0.5, "sigmoid"),   # 8-channel planes at the Cityscapes aspect (odd tile counts, pitch > 128: 4-D TMA boxes)
    (64, 48, 1, 3, 1.0, "sigmoid"),     # HeLa
    (32, 64, 3, 9, 2.0, "softmax"),     # SUIM noisy-student size: widths 32..512
    (48, 96, 3, 35, 1.0, "softmax"),    # Cityscapes aspect, K = 35
    (32, 32, 3, 9, 1.25, "softmax"),    # widths 20, 40, 80, 160, 320: not multiples of 16
    (16, 16, 3, 1, 0.75, "sigmoid"),    # 1x1 bottleneck, widths 12..192
    (208, 416, 3, 35, 1.0, "softmax"),  # Cityscapes full size: 13x26 bottleneck (odd tile counts)
    (32, 32, 1, 2, 1.0, "softmax"),     # head stage, generic class counts: K = 2 (8 blocks per TMEM group)
    (32, 48, 3, 5, 0.5, "softmax"),     # K = 5 (2 blocks per group, 8-column loads)
    (48, 32, 3, 12, 1.0, "sigmoid"),    # K = 12 (one block per group, 16-column loads)
    (32, 32, 3, 16, 2.0, "softmax"),    # K = 16 on 32 channels (two K steps in the head stage)
]


@pytest.mark.parametrize("engine", ["direct", "tcgen05", "fused"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}x{c[1]}c{c[2]}K{c[3]}a{c[4]}")
def test_predict_vs_fp32_oracle(U, case, engine):
    h, w, c, K, alpha, act = case
    n = 2 if h * w > 50000 else 3
    rng = np.random.default_rng(h * 1000 + w + K)
    weights = U.init_weights(c, K, alpha, seed=K + int(alpha * 100))
    images = rng.integers(0, 256, size=(n, h, w, c), dtype=np.uint8)
    model = U.B200UNet(h, w, c, K, alpha, act, weights)
    model.set_engine(engine)
    assert model.count_params() == ref_unet.count_params(c, K, alpha)
    got = model.predict([images])
    want = ref_unet.forward(images, weights, act)
    assert got.dtype == np.float32 and got.shape == want.shape
    assert np.isfinite(got).all()
    err = np.abs(got - want)
    if act == "softmax":
        np.testing.assert_allclose(got.sum(-1), 1.0, atol=1e-5)
        flips = float((got.argmax(-1) != want.argmax(-1)).mean())
    else:
        flips = float(((got > 0.5) != (want > 0.5)).mean())
    print(f"\n[{engine}] {case}: max|dp|={err.max():.3e} mean|dp|={err.mean():.3e} decision flips={flips:.2e}")
    assert err.max() <= PROB_ATOL and err.mean() <= MEAN_ATOL
    assert flips <= 2e-2
    # float32 input is accepted too (benchmark_hela feeds float arrays, functions.py:1199)
    got_f = model.predict(images.astype(np.float32))
    same(got_f, got)


def test_engines_agree(U):
    """tcgen05 and direct engines share weights and rounding points; they differ only in the
    order of the fp32 accumulation."""
    h, w, c, K, alpha = 64, 64, 3, 9, 1.0
    weights = U.init_weights(c, K, alpha, seed=5)
    images = np.random.default_rng(5).integers(0, 256, size=(4, h, w, c), dtype=np.uint8)
    model = U.B200UNet(h, w, c, K, alpha, "softmax", weights)
    model.set_engine("direct")
    a = model.predict(images)
    model.set_engine("tcgen05")
    b = model.predict(images)
    d = np.abs(a - b)
    print(f"\nengines: max|dp|={d.max():.3e} mean|dp|={d.mean():.3e}")
    assert d.max() < PROB_ATOL and d.mean() < 5e-4
    model.set_engine("fused")          # block-fused: same rounding points except the hi/lo-split first layer
    c_ = model.predict(images)
    d = np.abs(b - c_)
    print(f"fused vs layer-wise: max|dp|={d.max():.3e} mean|dp|={d.mean():.3e}")
    assert d.max() < PROB_ATOL and d.mean() < 5e-4


def test_swap_rb(U):
    h, w = 32, 32
    weights = U.init_weights(3, 1, 0.5, seed=1)
    img = np.random.default_rng(1).integers(0, 256, size=(2, h, w, 3), dtype=np.uint8)
    model = U.B200UNet(h, w, 3, 1, 0.5, "sigmoid", weights)
    a = model.predict(np.ascontiguousarray(img[..., ::-1]))
    model.set_swap_rb(True)
    b = model.predict(img)
    same(a, b)


def test_batch_chunking_is_invisible(U, F):
    """N larger than the internal trunk chunk: .predict equals per-image calls, and the fused ensemble path gives the
    same bits whatever the chunk size (64 vs the default)."""
    from inconsistencymasks_b200 import _lib
    h, w = 16, 16
    img = np.random.default_rng(2).integers(0, 256, size=(150, h, w, 1), dtype=np.uint8)
    models = [U.B200UNet(h, w, 1, 3, 0.5, "sigmoid", U.init_weights(1, 3, 0.5, seed=2 + j)) for j in range(2)]
    whole = F._run_batch(models, img, "hela", blank_image=img, block_input=True, block_output=True)
    assert _lib.lib.imk_max_chunk() >= 150
    try:
        _lib.check(_lib.lib.imk_set_max_chunk(64))
        assert _lib.lib.imk_max_chunk() == 64
        full = models[0].predict(img)
        for i in (0, 63, 64, 127, 128, 149):
            same(models[0].predict(img[i:i + 1])[0], full[i])
        parts = F._run_batch(models, img, "hela", blank_image=img, block_input=True, block_output=True)
    finally:
        _lib.check(_lib.lib.imk_set_max_chunk(0))
    same(parts.labels, whole.labels); same(parts.im, whole.im); same(parts.image, whole.image); same(parts.im_size, whole.im_size)


@pytest.mark.parametrize("case", [(256, 256, 1, 3, 1.0, "sigmoid"), (256, 256, 3, 9, 1.0, "softmax")],
                         ids=["hela256", "suim256a1"])
def test_steady_state_pipeline_many_tiles_per_cta(U, case):
    """Enough full-size images that every persistent CTA of the block-fused engine runs many tiles (double-buffered
    operands, tile-parity barriers, pooling lag in steady state): probabilities of images spread over the batch must
    match the fp32 oracle, and two passes must give identical bits (no race)."""
    h, w, c, K, alpha, act = case
    n = 40
    rng = np.random.default_rng(h + K)
    weights = U.init_weights(c, K, alpha, seed=900 + K)
    images = rng.integers(0, 256, size=(n, h, w, c), dtype=np.uint8)
    model = U.B200UNet(h, w, c, K, alpha, act, weights)
    got = model.predict(images)
    again = model.predict(images)
    same(got, again)
    assert np.isfinite(got).all()
    for i in (0, 17, n - 1):
        want = ref_unet.forward(images[i:i + 1], weights, act)[0]
        err = float(np.abs(got[i] - want).max())
        assert err < 2e-2, f"image {i}: max |p - oracle| = {err}"
    # a single-image call (one tile per CTA at most) sees the same bits as the image inside the batch
    same(model.predict(images[17:18])[0], got[17])


@pytest.mark.parametrize("thr", [0.5, 0.3, 0.7, 0.123456, 0.9999, 1e-6])
@pytest.mark.parametrize("kind,c,K", [("binary", 3, 1), ("hela", 1, 3)])
def test_fused_sigmoid_threshold_is_exact(U, F, kind, c, K, thr):
    """The fused binary path decides `p >= thr` / `p > thr` as `1 + exp(-z) <= d*` with d* searched on the host; it must
    agree bit for bit with thresholding the .predict probabilities, for any threshold."""
    h, w, n = 32, 32, 4
    images = np.random.default_rng(int(thr * 1e6) + K).integers(0, 256, size=(n, h, w, c), dtype=np.uint8)
    models = [U.B200UNet(h, w, c, K, 1.0, "sigmoid", U.init_weights(c, K, 1.0, seed=700 + j)) for j in range(2)]
    probs = [mdl.predict(images) for mdl in models]
    r = F._run_batch(models, images, kind, threshold=thr, blank_image=images, block_input=True, block_output=True)
    for i in range(n):
        if kind == "binary":
            lab, im, sz, pred = ref_im.im_prediction_binary([p[i] for p in probs], thr)
            img_b, lab_b, _ = ref_im.blank_binary(images[i], lab, im)
            same(r.labels[0, i], lab_b); same(r.im[i], im); same(r.image[i], img_b)
            assert r.im_size[i] == sz and r.pred_size[0, i] == pred
        else:
            alive, dead, pos, im, sz = ref_im.im_prediction_hela([p[i] for p in probs], thr)
            same(r.labels[2, i], pos); same(r.im[i], im)
            assert r.im_size[i] == sz


@pytest.mark.parametrize("K", [9, 35])
def test_fused_multiclass_ties_and_near_ties(U, F, K):
    """The fused multiclass path takes the argmax of the softmax numerators and forms the quotients only when two
    numerators are within 2^-22 of each other: exact ties (zero last layer: every class equal) and near ties (classes
    that differ by a few ulps of the bias) must still give what np.argmax gives on the .predict probabilities."""
    h, w, n, c = 32, 32, 3, 3
    images = np.random.default_rng(K).integers(0, 256, size=(n, h, w, c), dtype=np.uint8)
    models = []
    for j in range(2):
        wts = [np.array(a, copy=True) for a in U.init_weights(c, K, 1.0, seed=500 + j)]
        wts[-2][...] = 0.0                                   # last-layer kernel: logits == bias everywhere
        bias = np.zeros(K, np.float32)
        if j == 1:                                           # model 1: class 5 above class 2 by one ulp, the rest tied below
            bias[:] = -1.0
            bias[2] = 1.0
            bias[5] = np.nextafter(np.float32(1.0), np.float32(2.0))
        wts[-1][...] = bias
        models.append(U.B200UNet(h, w, c, K, 1.0, "softmax", wts))
    probs = [mdl.predict(images) for mdl in models]
    r = F._run_batch(models, images, "multiclass", blank_image=images, block_input=True, block_output=True)
    for i in range(n):
        lab, im, sz, _ = ref_im.im_prediction_multiclass([p[i] for p in probs], False)
        img_b, lab_b, _ = ref_im.blank_multiclass(images[i], lab, im)
        same(r.labels[0, i], lab_b); same(r.im[i], im); same(r.image[i], img_b)
        assert r.im_size[i] == sz


@pytest.mark.parametrize("kind,c,K,alpha,act", [("binary", 3, 1, 0.5, "sigmoid"), ("hela", 1, 3, 1.0, "sigmoid"),
                                                ("multiclass", 3, 9, 1.0, "softmax"), ("multiclass", 3, 35, 1.0, "softmax"),
                                                ("multiclass", 3, 2, 1.0, "softmax"), ("multiclass", 3, 5, 2.0, "softmax"),
                                                ("multiclass", 1, 12, 1.0, "sigmoid"), ("multiclass", 3, 16, 0.5, "softmax")])
@pytest.mark.parametrize("M", [1, 2, 3])
def test_fused_ensemble_equals_predict_then_im(U, F, kind, c, K, alpha, act, M):
    """The fused path (no fp32 map in HBM) must give, bit for bit, what the reference's own
    arithmetic gives on the probabilities .predict returns."""
    h, w, n = 32, 48, 5
    rng = np.random.default_rng(M * 10 + K)
    images = rng.integers(0, 256, size=(n, h, w, c), dtype=np.uint8)
    models = [U.B200UNet(h, w, c, K, alpha, act, U.init_weights(c, K, alpha, seed=100 + j)) for j in range(M)]
    probs = [mdl.predict(images) for mdl in models]
    r = F._run_batch(models, images, kind, blank_image=images, block_input=True, block_output=True,
                     want_lists_equal=(kind == "multiclass"))
    for i in range(n):
        if kind == "binary":
            lab, im, sz, pred = ref_im.im_prediction_binary([p[i] for p in probs], 0.5)
            img_b, lab_b, _ = ref_im.blank_binary(images[i], lab, im)
            same(r.labels[0, i], lab_b); same(r.im[i], im); same(r.image[i], img_b)
            assert r.im_size[i] == sz and r.pred_size[0, i] == pred
        elif kind == "hela":
            alive, dead, pos, im, sz = ref_im.im_prediction_hela([p[i] for p in probs])
            bf, alive_b, dead_b, _, _ = ref_im.blank_hela(images[i, ..., 0], alive, dead, np.zeros((h, w, 3), np.uint8), im)
            same(r.labels[0, i], alive_b); same(r.labels[1, i], dead_b); same(r.labels[2, i], pos)
            same(r.im[i], im); same(r.image[i, ..., 0], bf)
            assert r.im_size[i] == sz
        else:
            lab, im, sz, eq = ref_im.im_prediction_multiclass([p[i] for p in probs], True)
            img_b, lab_b, _ = ref_im.blank_multiclass(images[i], lab, im)
            same(r.labels[0, i], lab_b); same(r.im[i], im); same(r.image[i], img_b)
            assert r.im_size[i] == sz and bool(r.lists_equal[i]) == bool(eq)
    if M == 1:
        assert r.im.sum() == 0 and r.im_size.sum() == 0
    # the reference-named helper on one image goes through the same path
    if kind == "binary":
        lab, im, sz, pred = F.get_im_prediction_binary(models, images[:1], 0.5)
        e = ref_im.im_prediction_binary([p[0] for p in probs], 0.5)
        same(lab, e[0]); same(im, e[1]); assert sz == e[2] and pred == e[3]
    elif kind == "hela":
        got = F.get_im_prediction_hela(models, images[:1])
        e = ref_im.im_prediction_hela([p[0] for p in probs])
        for a, b in zip(got[:4], e[:4]):
            same(a, b)
        assert got[4] == e[4]
    else:
        got = F.get_im_prediction_multiclass(models, images[:1], True)
        e = ref_im.im_prediction_multiclass([p[0] for p in probs], True)
        same(got[0], e[0]); same(got[1], e[1]); assert got[2] == e[2] and got[3] == e[3]


@pytest.mark.parametrize("kind,c,K,alpha,act,h,w,n", [("binary", 3, 1, 0.5, "sigmoid", 32, 48, 5), ("hela", 1, 3, 1.0, "sigmoid", 256, 256, 10),
                                                        ("binary", 3, 1, 2.0, "sigmoid", 64, 64, 4), ("multiclass", 3, 3, 1.0, "softmax", 48, 32, 4),
                                                        ("multiclass", 3, 2, 2.0, "sigmoid", 32, 32, 3)])
def test_head_in_epilogue_variant(U, F, monkeypatch, kind, c, K, alpha, act, h, w, n):
    """IMK_BT_HEAD=1 (opt-in): the level-0 decoder kernel evaluates the output layer in its last epilogue and writes one
    decision byte per pixel (ensemble_votes then only counts votes).  Same contract as the default path: .predict within
    tolerance of the oracle, fused == predict -> reference IM arithmetic bit for bit; several tiles per CTA at 256x256."""
    from inconsistencymasks_b200 import _lib
    monkeypatch.setenv("IMK_BT_HEAD", "1")
    monkeypatch.setenv("IMK_BT_NO_C8", "1")       # the head variant exists for the 16-channel layout only
    rng = np.random.default_rng(K * 7 + h)
    images = rng.integers(0, 256, size=(n, h, w, c), dtype=np.uint8)
    weights = [U.init_weights(c, K, alpha, seed=800 + j) for j in range(2)]
    models = [U.B200UNet(h, w, c, K, alpha, act, wts) for wts in weights]
    probs = [mdl.predict(images) for mdl in models]
    same(models[0].predict(images), probs[0])
    want = ref_unet.forward(images[:2], weights[0], act)
    assert float(np.abs(probs[0][:2] - want).max()) <= 3e-2
    _lib.profile_begin()
    r = F._run_batch(models, images, kind, blank_image=images, block_input=True, block_output=True)
    names = {p["name"] for p in _lib.profile_end()}
    assert "block_head" in names and "ensemble_votes" in names and "ensemble_im" not in names, names
    for i in range(n):
        if kind == "binary":
            lab, im, sz, pred = ref_im.im_prediction_binary([p[i] for p in probs], 0.5)
            img_b, lab_b, _ = ref_im.blank_binary(images[i], lab, im)
            same(r.labels[0, i], lab_b); same(r.im[i], im); same(r.image[i], img_b)
            assert r.im_size[i] == sz and r.pred_size[0, i] == pred
        elif kind == "hela":
            alive, dead, pos, im, sz = ref_im.im_prediction_hela([p[i] for p in probs])
            bf, alive_b, dead_b, _, _ = ref_im.blank_hela(images[i, ..., 0], alive, dead, np.zeros((h, w, 3), np.uint8), im)
            same(r.labels[0, i], alive_b); same(r.labels[1, i], dead_b); same(r.labels[2, i], pos)
            same(r.im[i], im); same(r.image[i, ..., 0], bf)
            assert r.im_size[i] == sz
        else:
            lab, im, sz, _ = ref_im.im_prediction_multiclass([p[i] for p in probs])
            img_b, lab_b, _ = ref_im.blank_multiclass(images[i], lab, im)
            same(r.labels[0, i], lab_b); same(r.im[i], im); same(r.image[i], img_b)
            assert r.im_size[i] == sz


def test_fused_with_morphology_and_device_path(U, F):
    """EK / DK > 0 routes through the device-buffer calls + morphology kernels + imk_blank."""
    h, w, n, c, K = 32, 48, 4, 3, 9
    rng = np.random.default_rng(9)
    images = rng.integers(0, 256, size=(n, h, w, c), dtype=np.uint8)
    models = [U.B200UNet(h, w, c, K, 1.0, "softmax", U.init_weights(c, K, 1.0, seed=200 + j)) for j in range(2)]
    probs = [mdl.predict(images) for mdl in models]
    for ek, dk in ((3, 3), (5, 0), (0, 5)):
        r = F._run_batch(models, images, "multiclass", blank_image=images, block_input=True, block_output=True,
                         erode_kernel=ek, dilate_kernel=dk)
        for i in range(n):
            lab, im, sz, _ = ref_im.im_prediction_multiclass([p[i] for p in probs])
            img_b, lab_b, im_b = ref_im.blank_multiclass(images[i], lab, im, ek, dk)
            same(r.labels[0, i], lab_b); same(r.im[i], im_b); same(r.image[i], img_b)
            assert r.im_size[i] == sz


def test_host_pipeline_multi_chunk_matches_single_calls(U, F):
    """imk_pseudo_label_*_host with N spanning several pipeline chunks (two slots, events)."""
    import ctypes as C
    from inconsistencymasks_b200 import _lib
    h, w, c, K, n = 16, 32, 1, 3, 37
    rng = np.random.default_rng(4)
    images = rng.integers(0, 256, size=(n, h, w, c), dtype=np.uint8)
    models = [U.B200UNet(h, w, c, K, 0.5, "sigmoid", U.init_weights(c, K, 0.5, seed=300 + j)) for j in range(2)]
    ref = F._run_batch(models, images, "hela", blank_image=images, block_input=True, block_output=True)
    labels = np.empty((3, n, h, w), np.uint8); im = np.empty((n, h, w), np.uint8)
    out = np.empty_like(images); sz = np.empty(n, np.int64); pred = np.empty((3, n), np.int64)
    hs = (C.c_void_p * 2)(*[m.handle for m in models])
    _lib.check(_lib.lib.imk_pseudo_label_binary_host(hs, 2, images.ctypes.data, n, 0, 0.5, 0, 1, 1, out.ctypes.data,
                                                     labels.ctypes.data, im.ctypes.data, sz.ctypes.data, pred.ctypes.data, 5))
    same(labels, ref.labels); same(im, ref.im); same(out, ref.image); same(sz, ref.im_size); same(pred, ref.pred_size)


def test_host_pipeline_default_schedule_equals_fixed_chunks(U, F):
    """chunk <= 0: the library's own schedule (quarter / half chunks at both ends, ragged middle) gives the bytes of a
    fixed-chunk call, binary and multiclass."""
    import ctypes as C
    from inconsistencymasks_b200 import _lib
    h, w, c, n = 16, 32, 1, 530
    rng = np.random.default_rng(14)
    images = rng.integers(0, 256, size=(n, h, w, c), dtype=np.uint8)
    _lib.check(_lib.lib.imk_set_max_chunk(128))
    try:
        for K, act in ((3, "sigmoid"), (4, "softmax")):
            models = [U.B200UNet(h, w, c, K, 0.5, act, U.init_weights(c, K, 0.5, seed=700 + j)) for j in range(2)]
            hs = (C.c_void_p * 2)(*[m.handle for m in models])
            got = []
            for chunk in (0, 128):
                P = K if act == "sigmoid" else 1
                labels = np.zeros((P, n, h, w), np.uint8); im = np.zeros((n, h, w), np.uint8)
                out = np.zeros_like(images); sz = np.zeros(n, np.int64); pred = np.zeros((P, n), np.int64); eq = np.zeros(n, np.uint8)
                if act == "sigmoid":
                    _lib.check(_lib.lib.imk_pseudo_label_binary_host(hs, 2, images.ctypes.data, n, 0, 0.5, 0, 1, 1, out.ctypes.data,
                                                                     labels.ctypes.data, im.ctypes.data, sz.ctypes.data, pred.ctypes.data, chunk))
                else:
                    _lib.check(_lib.lib.imk_pseudo_label_multiclass_host(hs, 2, images.ctypes.data, n, 0, 1, 1, out.ctypes.data,
                                                                         labels.ctypes.data, im.ctypes.data, sz.ctypes.data, eq.ctypes.data, chunk))
                got.append((labels, im, out, sz, pred, eq))
            for x, y in zip(*got):
                same(x, y)
            assert got[0][1].any() and got[0][3].sum() > 0          # the masks are not trivially empty
    finally:
        _lib.check(_lib.lib.imk_set_max_chunk(0))


def test_keras_like_surface(U, tmp_path):
    model = U.get_unet(32, 32, 3, 9, 1.0, "relu", "softmax", seed=3)
    assert model.input_shape == (None, 32, 32, 3) and model.output_shape == (None, 32, 32, 9)
    assert model.count_params() == 681_817
    x = np.random.default_rng(0).integers(0, 256, size=(1, 32, 32, 3), dtype=np.uint8)
    p = model.predict([x])
    path = str(tmp_path / "m.npz")
    model.save_weights(path)
    again = U.load_model(path, custom_objects={"dice_loss": None})
    same(again.predict(x), p)
    with pytest.raises(ValueError):
        model.predict(np.zeros((1, 16, 16, 3), np.uint8))
    with pytest.raises(ValueError):
        U.get_unet(32, 32, 3, 9, 1.0, "elu", "softmax")
    from inconsistencymasks_b200 import _lib
    with pytest.raises(_lib.ImkError):
        U.get_unet(30, 32, 3, 9, 1.0, "relu", "softmax")      # not a multiple of 16


def test_workspace_allocation_failure_is_reported(U, monkeypatch):
    """A workspace that cannot be allocated is IMK_ENOMEM with a message, nothing is left half-initialised, and the same
    model still serves a batch that fits (6.2 GB per model at chunk 512 had no failure-path test)."""
    from inconsistencymasks_b200 import _lib
    h = w = 64
    model = U.B200UNet(h, w, 3, 1, 1.0, "sigmoid", U.init_weights(3, 1, 1.0, seed=2))
    x = np.random.default_rng(0).integers(0, 256, size=(64, h, w, 3), dtype=np.uint8)
    monkeypatch.setenv("IMK_WS_LIMIT_MB", "1")
    with pytest.raises(_lib.ImkError) as e:
        model.predict(x)
    assert "error -3" in str(e.value) and "workspace" in str(e.value)
    monkeypatch.setenv("IMK_WS_LIMIT_MB", "4096")
    p = model.predict(x[:4])
    assert p.shape == (4, h, w, 1) and np.isfinite(p).all()
    same(p, model.predict(x)[:4])


def test_predict_is_stateless_after_pseudo_label_calls(U, F, tmp_path):
    """a driver call with rgb=True (channel swap) must not change what a later .predict returns."""
    import cv2
    h = w = 32
    rng = np.random.default_rng(5)
    models = [U.B200UNet(h, w, 3, 1, 0.5, "sigmoid", U.init_weights(3, 1, 0.5, seed=40 + j)) for j in range(2)]
    x = rng.integers(0, 256, size=(3, h, w, 3), dtype=np.uint8)
    before = models[0].predict(x)
    d = tmp_path / "in"; d.mkdir()
    for i in range(3):
        cv2.imwrite(str(d / f"a{i}.png"), x[i])
    F.create_pseudo_labels_im_ISIC_2018(models, h, w, 3, str(d), str(tmp_path / "out"), rgb=True, erode_kernel=0, dilate_kernel=0)
    same(models[0].predict(x), before)
