"""GPU parity of the Inconsistency-Mask kernels (rows a2-a10) through the C ABI / the
reference-named helpers: bit-exact against the golden fixtures the reference itself
produced (tests/golden) and against the CPU oracle on seeded inputs."""
import ast
import ctypes as C
import os
import tempfile

import cv2
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
from oracle import ref_im  # noqa: E402


@pytest.fixture(scope="module")
def F():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from inconsistencymasks_b200 import functions
    return functions


@pytest.fixture(scope="module")
def lib():
    from inconsistencymasks_b200 import _lib
    return _lib


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.dtype == b.dtype, (a.dtype, b.dtype)
    assert a.shape == b.shape, (a.shape, b.shape)
    np.testing.assert_array_equal(a, b)


class ReplayModel:
    """Keras-like stand-in whose .predict replays stored maps (same as oracle/make_golden.py)."""

    def __init__(self):
        self.table = {}

    def add(self, image, prob):
        self.table[np.ascontiguousarray(image).tobytes()] = prob

    def predict(self, x, *a, **k):
        if isinstance(x, (list, tuple)):
            x = x[0]
        return self.table[np.ascontiguousarray(x).tobytes()]


# ----------------------------------------------------------------------- a5 / a6 golden
def test_kat_im_creation_figure(F, golden_dir):
    g = load(golden_dir, "im_kat.npz")
    label, im, im_size, pred_size = F.pred_masks_to_im_binary([g["a"][..., None], g["b"][..., None]])
    same(label, g["label"]); same(im, g["im"])
    assert isinstance(im_size, np.int64) and im_size == 9 and pred_size == 45


def test_pred_masks_to_im_binary_golden(F, golden_dir):
    g = load(golden_dir, "im_binary.npz")
    for i in range(int(g["n"])):
        label, im, im_size, pred_size = F.pred_masks_to_im_binary(list(g[f"{i}/masks"]))
        same(label, g[f"{i}/label"]); same(im, g[f"{i}/im"])
        assert im_size == g[f"{i}/im_size"] and pred_size == g[f"{i}/pred_size"]


def test_pred_masks_to_im_multiclass_golden(F, golden_dir):
    g = load(golden_dir, "im_multiclass.npz")
    for i in range(int(g["n"])):
        label, im, im_size = F.pred_masks_to_im_multiclass(list(g[f"{i}/masks"]))
        same(label, g[f"{i}/label"]); same(im, g[f"{i}/im"])
        assert isinstance(im_size, np.int64) and im_size == g[f"{i}/im_size"]


def test_dilate_mask_golden(F, golden_dir):
    g = load(golden_dir, "dilate_mask.npz")
    for i in range(int(g["n"])):
        same(F.dilate_mask(g[f"{i}/label"]), g[f"{i}/out"])


# ----------------------------------------------------------------------- a2 / a3 / a4 golden
def _replay_models(img, probs):
    models = []
    for p in probs:
        m = ReplayModel(); m.add(img, p); models.append(m)
    return models


def test_get_im_prediction_binary_golden(F, golden_dir):
    g = load(golden_dir, "predict.npz")
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(1, 20, 28, 3), dtype=np.uint8)
    for m in (1, 2, 3, 5):
        probs = g[f"binary/m{m}/probs"]
        models = _replay_models(img, probs)
        for thr in (0.5, 0.3):
            label, im, im_size, pred_size = F.get_im_prediction_binary(models, img, thr)
            tag = f"binary/m{m}/t{thr}"
            same(label, g[f"{tag}/label"]); same(im, g[f"{tag}/im"])
            assert im_size == g[f"{tag}/im_size"] and pred_size == g[f"{tag}/pred_size"]
            assert isinstance(im_size, np.int64)


def test_get_im_prediction_hela_golden(F, golden_dir):
    g = load(golden_dir, "predict.npz")
    img = np.zeros((1, 20, 28, 1), np.uint8)
    for m in (1, 2, 4):
        tag = f"hela/m{m}"
        alive, dead, pos, im, im_size = F.get_im_prediction_hela(_replay_models(img, g[f"{tag}/probs"]), img)
        same(alive, g[f"{tag}/alive"]); same(dead, g[f"{tag}/dead"]); same(pos, g[f"{tag}/pos"])
        same(im, g[f"{tag}/im"])
        assert im_size == g[f"{tag}/im_size"]


def test_get_im_prediction_multiclass_golden(F, golden_dir):
    g = load(golden_dir, "predict.npz")
    img = np.zeros((1, 20, 28, 3), np.uint8)
    for m, k in ((1, 9), (2, 9), (3, 35), (2, 35), (5, 2)):
        probs = g[f"multi/m{m}k{k}/probs"]
        models = _replay_models(img, probs)
        for flt in (False, True):
            tag = f"multi/m{m}k{k}/f{int(flt)}"
            label, im, im_size, eq = F.get_im_prediction_multiclass(models, img, flt)
            same(label, g[f"{tag}/label"]); same(im, g[f"{tag}/im"])
            assert im_size == g[f"{tag}/im_size"]
            assert bool(eq) == bool(g[f"{tag}/lists_equal"])


# ----------------------------------------------------------------------- drivers golden
@pytest.mark.parametrize("kind", ["binary", "hela", "multiclass"])
def test_create_pseudo_labels_golden(F, golden_dir, kind):
    """The whole create_pseudo_labels_im_* helper, PNG files in and out, against the files
    the reference wrote (functions.py:2832-3070), including morphology, the write filter,
    block_input/output switches and (HeLa) the host-drawn position circles."""
    g = load(golden_dir, "drivers.npz")
    names = [str(n) for n in g[f"{kind}/names"]]
    images, probs = g[f"{kind}/images"], g[f"{kind}/probs"]        # probs [n_img, M, H, W, K]
    h, w = images.shape[1:3]
    c = 1 if kind == "hela" else 3
    fn = {"binary": F.create_pseudo_labels_im_ISIC_2018, "hela": F.create_pseudo_labels_im_hela,
          "multiclass": F.create_pseudo_labels_im_multiclass}[kind]
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "in")
        os.makedirs(src)
        models = [ReplayModel() for _ in range(probs.shape[1])]
        for i, name in enumerate(names):
            cv2.imwrite(os.path.join(src, name), images[i])
            fed = cv2.cvtColor(images[i], cv2.COLOR_BGR2RGB) if c == 3 else images[i]
            fed = np.array(fed.reshape(-1, h, w, c), dtype=np.uint8)
            for m, mod in enumerate(models):
                mod.add(fed, probs[i, m][None])
        for j in range(int(g[f"{kind}/nruns"])):
            kw = dict(ast.literal_eval(str(g[f"{kind}/run{j}/kwargs"])))
            dst = os.path.join(tmp, f"out{j}")
            mean = fn(models, h, w, c, src, dst, **kw)
            assert isinstance(mean, float) and mean == float(g[f"{kind}/run{j}/mean_im_size"]), kw
            for sub in sorted(os.listdir(dst)):
                for name in names:
                    key = f"{kind}/run{j}/{sub}/{name}"
                    path = os.path.join(dst, sub, name)
                    assert (key in g.files) == os.path.exists(path), (key, kw)
                    if key in g.files:
                        same(cv2.imread(path, cv2.IMREAD_UNCHANGED), g[key])
            subs = {k.split("/")[2] for k in g.files if k.startswith(f"{kind}/run{j}/") and k.count("/") == 3}
            assert subs == set(os.listdir(dst))


# ----------------------------------------------------------------------- C ABI vs oracle, seeded
def _tricky(rng, shape, softmax):
    p = rng.random(shape, dtype=np.float32)
    if softmax:
        p = p / p.sum(axis=-1, keepdims=True)
    flat = p.reshape(-1, shape[-1])
    idx = rng.choice(flat.shape[0], size=max(4, flat.shape[0] // 16), replace=False)
    q = len(idx) // 4
    flat[idx[:q]] = 0.5
    flat[idx[q:2 * q], -1] = flat[idx[q:2 * q]].max(axis=-1)
    flat[idx[2 * q:3 * q], 0] = np.nextafter(np.float32(0.5), np.float32(1))
    flat[idx[3 * q:3 * q + 2], shape[-1] // 2] = np.nan
    return p


def _call_im_binary(lib, probs, img, K, thr, strict, bi, bo):
    n, h, w = probs[0].shape[:3]
    c = img.shape[-1]
    d_probs = [torch.from_numpy(p).cuda() for p in probs]
    ptrs = (C.c_void_p * len(d_probs))(*[p.data_ptr() for p in d_probs])
    d_img = torch.from_numpy(img).cuda()
    d_out = torch.empty_like(d_img)
    d_lab = torch.empty((K, n, h, w), dtype=torch.uint8, device="cuda")
    d_im = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
    d_sz = torch.full((n,), -1, dtype=torch.int64, device="cuda")
    d_pred = torch.full((K, n), -1, dtype=torch.int64, device="cuda")
    lib.check(lib.lib.imk_im_binary(ptrs, len(probs), n, h, w, K, thr, strict, d_img.data_ptr(), c, bi, bo, d_out.data_ptr(),
                                    d_lab.data_ptr(), d_im.data_ptr(), d_sz.data_ptr(), d_pred.data_ptr(),
                                    torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return d_out.cpu().numpy(), d_lab.cpu().numpy(), d_im.cpu().numpy(), d_sz.cpu().numpy(), d_pred.cpu().numpy()


@pytest.mark.parametrize("shape", [(3, 16, 16), (5, 32, 48), (2, 256, 256), (3, 17, 33), (1, 208, 416), (7, 8, 2)])
@pytest.mark.parametrize("M", [1, 2, 5])
@pytest.mark.parametrize("c", [1, 3])
def test_im_binary_vs_oracle(lib, shape, M, c):
    """K = 1, strict >: vector path (H*W % 16 == 0) and generic path (odd shapes), chunks that
    straddle image boundaries, NaNs and exact-threshold values."""
    n, h, w = shape
    rng = np.random.default_rng(hash((shape, M, c)) % 2**32)
    probs = [_tricky(rng, (n, h, w, 1), False) for _ in range(M)]
    img = rng.integers(0, 256, size=(n, h, w, c), dtype=np.uint8)
    for thr, bi, bo in ((0.5, 1, 1), (0.3, 0, 1), (0.5, 1, 0)):
        out, lab, im, sz, pred = _call_im_binary(lib, probs, img, 1, thr, 1, bi, bo)
        for i in range(n):
            e_lab, e_im, e_sz, e_pred = ref_im.im_prediction_binary([p[i] for p in probs], thr)
            e_img, e_lab_b, _ = ref_im.blank_binary(img[i], e_lab, e_im, 0, 0, bool(bi), bool(bo))
            same(lab[0, i], e_lab_b); same(im[i], e_im); same(out[i], e_img)
            assert sz[i] == e_sz and pred[0, i] == e_pred


@pytest.mark.parametrize("shape", [(2, 16, 16), (3, 32, 48), (2, 256, 256), (3, 17, 33), (9, 8, 2)])
@pytest.mark.parametrize("M", [1, 2, 4])
def test_im_hela_vs_oracle(lib, shape, M):
    n, h, w = shape
    rng = np.random.default_rng(hash((shape, M)) % 2**32)
    probs = [_tricky(rng, (n, h, w, 3), False) for _ in range(M)]
    img = rng.integers(0, 256, size=(n, h, w, 1), dtype=np.uint8)
    for bi, bo in ((1, 1), (0, 0)):
        out, lab, im, sz, pred = _call_im_binary(lib, probs, img, 3, 0.5, 0, bi, bo)
        for i in range(n):
            alive, dead, pos, cim, e_sz = ref_im.im_prediction_hela([p[i] for p in probs])
            bf, alive_b, dead_b, _, _ = ref_im.blank_hela(img[i, ..., 0], alive, dead, np.zeros((h, w, 3), np.uint8), cim,
                                                          0, 0, bool(bi), bool(bo))
            same(lab[0, i], alive_b); same(lab[1, i], dead_b); same(lab[2, i], pos)     # position head stays raw
            same(im[i], cim); same(out[i, ..., 0], bf)
            assert sz[i] == e_sz
            assert pred[0, i] == (alive > 0).sum() and pred[2, i] == (pos > 0).sum()


def _call_im_multiclass(lib, probs, img, bi, bo, want_eq):
    n, h, w, K = probs[0].shape
    c = img.shape[-1]
    d_probs = [torch.from_numpy(p).cuda() for p in probs]
    ptrs = (C.c_void_p * len(d_probs))(*[p.data_ptr() for p in d_probs])
    d_img = torch.from_numpy(img).cuda()
    d_out = torch.empty_like(d_img)
    d_lab = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
    d_im = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
    d_sz = torch.full((n,), -1, dtype=torch.int64, device="cuda")
    d_eq = torch.full((n,), 7, dtype=torch.uint8, device="cuda")
    lib.check(lib.lib.imk_im_multiclass(ptrs, len(probs), n, h, w, K, d_img.data_ptr(), c, bi, bo, d_out.data_ptr(),
                                        d_lab.data_ptr(), d_im.data_ptr(), d_sz.data_ptr(),
                                        d_eq.data_ptr() if want_eq else None, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return d_out.cpu().numpy(), d_lab.cpu().numpy(), d_im.cpu().numpy(), d_sz.cpu().numpy(), d_eq.cpu().numpy()


@pytest.mark.parametrize("shape,K", [((3, 16, 16), 9), ((2, 32, 48), 35), ((2, 256, 256), 9), ((1, 208, 416), 35),
                                     ((3, 17, 33), 9), ((2, 16, 16), 2), ((2, 32, 32), 4), ((2, 16, 48), 64),
                                     ((1, 16, 16), 200), ((5, 8, 2), 3)])
@pytest.mark.parametrize("M", [1, 2, 3])
def test_im_multiclass_vs_oracle(lib, shape, K, M):
    """TMA path (H*W % 16 == 0), its tail tiles, the generic path, ties, NaNs, large K."""
    n, h, w = shape
    rng = np.random.default_rng(hash((shape, K, M)) % 2**32)
    base = _tricky(rng, (n, h, w, K), True)
    probs = [(base + rng.random((n, h, w, K), dtype=np.float32) * np.float32(0.15)).astype(np.float32) for _ in range(M)]
    img = rng.integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)
    want_eq = K <= 64
    out, lab, im, sz, eq = _call_im_multiclass(lib, probs, img, 1, 1, want_eq)
    for i in range(n):
        e_lab, e_im, e_sz, e_eq = ref_im.im_prediction_multiclass([p[i] for p in probs], True)
        e_img, e_lab_b, _ = ref_im.blank_multiclass(img[i], e_lab, e_im)
        same(lab[i], e_lab_b); same(im[i], e_im); same(out[i], e_img)
        assert sz[i] == e_sz
        if want_eq:
            assert bool(eq[i]) == bool(e_eq)
    # block_input off copies the image through
    out2, *_ = _call_im_multiclass(lib, probs, img, 0, 1, False)
    same(out2, img)


@pytest.mark.parametrize("K", [9, 35, 5])
def test_argmax_edge_semantics(lib, K):
    """np.argmax corner cases (functions.py:3225): exact ties -> first index, -0.0 == +0.0, NaN is the maximum and the
    first NaN wins, infinities, negative values (the C ABI accepts any float32 map, not only softmax outputs)."""
    rng = np.random.default_rng(K)
    h, w = 16, 16
    p = rng.standard_normal((1, h, w, K)).astype(np.float32)
    flat = p.reshape(-1, K)
    flat[0] = 0.0                                    # all equal
    flat[1] = 0.0; flat[1, 2] = -0.0; flat[1, 3] = 0.0
    flat[2] = -0.0; flat[2, 4] = 0.0                 # +0 after -0: still a tie, the first wins
    flat[3] = -1.0; flat[3, K - 1] = -0.0            # -0 is the maximum of negatives
    flat[4, 1] = np.nan
    flat[5, 1] = np.inf; flat[5, 3] = np.nan         # NaN beats +inf
    flat[6, 2] = np.nan; flat[6, 4] = np.nan         # the first NaN wins
    flat[7] = -np.inf; flat[7, 3] = -3.0e38
    flat[8] = np.inf                                 # ties at +inf
    flat[9] = 1.0; flat[9, K - 1] = np.nextafter(np.float32(1), np.float32(2))
    flat[10] = np.float32(1e-45)                     # denormals
    flat[10, 2] = np.float32(3e-45)
    flat[11, 0] = np.nan
    flat[12] = np.nan
    flat[13] = -5.0; flat[13, 1] = -4.0; flat[13, 2] = -4.0
    img = rng.integers(0, 256, size=(1, h, w, 3), dtype=np.uint8)
    out, lab, im, sz, _ = _call_im_multiclass(lib, [p], img, 1, 1, False)
    same(lab[0], np.argmax(p[0], axis=-1).astype(np.uint8))
    assert sz[0] == 0 and not im.any()
    # two models that differ exactly where the corner cases sit
    p2 = p.copy(); p2.reshape(-1, K)[:14] = np.roll(flat[:14], 1, axis=-1)
    out, lab, im, sz, _ = _call_im_multiclass(lib, [p, p2], img, 1, 1, False)
    e_lab, e_im, e_sz, _ = ref_im.im_prediction_multiclass([p[0], p2[0]], False)
    same(lab[0], e_lab); same(im[0], e_im)
    assert sz[0] == e_sz


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 7])
def test_morphology_vs_cv2(lib, k):
    rng = np.random.default_rng(k)
    for shape in ((1, 1, 1), (2, 3, 5), (3, 32, 48), (2, 13, 26)):
        x = ((rng.random(shape) < 0.6).astype(np.uint8) * 255)
        d = torch.from_numpy(x).cuda()
        e = torch.empty_like(d); f = torch.empty_like(d)
        s = torch.cuda.current_stream().cuda_stream
        lib.check(lib.lib.imk_erode_u8(d.data_ptr(), e.data_ptr(), shape[0], shape[1], shape[2], k, s))
        lib.check(lib.lib.imk_dilate_u8(d.data_ptr(), f.data_ptr(), shape[0], shape[1], shape[2], k, s))
        ker = np.ones((k, k), np.uint8)
        for i in range(shape[0]):
            same(e.cpu().numpy()[i], cv2.erode(x[i], ker, iterations=1).reshape(shape[1:]))
            same(f.cpu().numpy()[i], cv2.dilate(x[i], ker, iterations=1).reshape(shape[1:]))


def test_empty_and_invalid(lib):
    s = torch.cuda.current_stream().cuda_stream
    d = torch.zeros(16, dtype=torch.uint8, device="cuda")
    sz = torch.zeros(1, dtype=torch.int64, device="cuda")
    p = torch.zeros(16, dtype=torch.float32, device="cuda")
    ptrs = (C.c_void_p * 1)(p.data_ptr())
    # N == 0 is a no-op
    assert lib.lib.imk_im_binary(ptrs, 1, 0, 4, 4, 1, 0.5, 1, None, 3, 0, 0, None, d.data_ptr(), d.data_ptr(), sz.data_ptr(), None, s) == 0
    # K outside {1, 3} and M out of range are rejected
    assert lib.lib.imk_im_binary(ptrs, 1, 1, 4, 4, 2, 0.5, 1, None, 3, 0, 0, None, d.data_ptr(), d.data_ptr(), sz.data_ptr(), None, s) == -1
    assert lib.lib.imk_im_binary(ptrs, 17, 1, 4, 4, 1, 0.5, 1, None, 3, 0, 0, None, d.data_ptr(), d.data_ptr(), sz.data_ptr(), None, s) == -1
    assert lib.lib.imk_im_multiclass(ptrs, 1, 1, 4, 4, 100, None, 3, 0, 0, None, d.data_ptr(), d.data_ptr(), sz.data_ptr(), d.data_ptr(), s) == -1


def test_full_size_properties(lib):
    """BASELINE-size inputs (HeLa 256x256, K=3, M=2, N=64; Cityscapes 208x416, K=35): size-independent
    invariants -- M=1 gives an empty IM, im and label never overlap, sizes add up, blanked pixels are zero."""
    rng = np.random.default_rng(11)
    n, h, w = 64, 256, 256
    probs = [rng.random((n, h, w, 3), dtype=np.float32) for _ in range(2)]
    img = rng.integers(1, 256, size=(n, h, w, 1), dtype=np.uint8)
    out, lab, im, sz, pred = _call_im_binary(lib, probs, img, 3, 0.5, 0, 1, 1)
    assert ((lab[0] > 0) & (im > 0)).sum() == 0 and ((lab[1] > 0) & (im > 0)).sum() == 0
    assert (out[..., 0][im > 0] == 0).all() and (out[..., 0][im == 0] == img[..., 0][im == 0]).all()
    assert int(sz.sum()) >= int((im > 0).sum())                       # sum of three heads >= union
    _, _, im1, sz1, _ = _call_im_binary(lib, probs[:1], img, 3, 0.5, 0, 1, 1)
    assert im1.sum() == 0 and sz1.sum() == 0
    n, h, w, K = 8, 208, 416, 35
    probs = [rng.random((n, h, w, K), dtype=np.float32) for _ in range(2)]
    img = rng.integers(1, 256, size=(n, h, w, 3), dtype=np.uint8)
    out, lab, im, sz, _ = _call_im_multiclass(lib, probs, img, 1, 1, False)
    assert (lab[im > 0] == 0).all()
    for i in range(n):
        assert sz[i] == (im[i] > 0).sum()
        assert sz[i] + (im[i] == 0).sum() == h * w
    assert (out[im > 0] == 0).all() and (out[im == 0] == img[im == 0]).all()
    a0, a1 = probs[0].argmax(-1), probs[1].argmax(-1)
    same(im, np.where(a0 == a1, 0, 255).astype(np.uint8))
