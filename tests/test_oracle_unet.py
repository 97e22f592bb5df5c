"""Cross-check of the two independently written CPU restatements of the hot path: the
PyTorch fp32 one (oracle/ref_unet.py, oracle/ref_im.py) and the plain-C one
(oracle/unet_oracle.c).  CPU only.  The U-Net row has no golden vector from the reference
(TensorFlow is unavailable: parity at that boundary is unpinned); agreement of
the two restatements guards against layout / ordering mistakes in either."""
import ctypes as C

import numpy as np
import pytest

from oracle import build_oracle, ref_im, ref_unet
from inconsistencymasks_b200 import unet


@pytest.fixture(scope="module")
def clib():
    lib = C.CDLL(build_oracle.build())
    lib.oracle_unet_forward.restype = C.c_int
    lib.oracle_unet_forward.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                        C.c_void_p, C.c_int, C.c_void_p]
    return lib


@pytest.mark.parametrize("h,w,c,k,alpha,act", [(16, 16, 3, 1, 0.5, "sigmoid"), (32, 16, 1, 3, 1.0, "sigmoid"),
                                               (16, 32, 3, 9, 1.25, "softmax"), (16, 16, 3, 35, 0.75, "softmax")])
def test_c_and_torch_restatements_agree(clib, h, w, c, k, alpha, act):
    rng = np.random.default_rng(k)
    weights = unet.init_weights(c, k, alpha, seed=7 + k)
    images = rng.integers(0, 256, size=(2, h, w, c), dtype=np.uint8)
    want = ref_unet.forward(images, weights, act)
    ptrs = (C.c_void_p * len(weights))(*[wt.ctypes.data for wt in weights])
    got = np.empty_like(want)
    rc = clib.oracle_unet_forward(images.ctypes.data, 2, h, w, c, k, float(alpha), 3, 0 if act == "sigmoid" else 1,
                                  ptrs, len(weights), got.ctypes.data)
    assert rc == 0
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-5)
    # a wrong weight count is detected
    assert clib.oracle_unet_forward(images.ctypes.data, 2, h, w, c, k, float(alpha), 3, 0, ptrs, len(weights) - 1, got.ctypes.data) == -1


def test_c_im_matches_numpy_oracle(clib):
    rng = np.random.default_rng(0)
    for m in (1, 2, 5):
        masks = rng.integers(0, 2, size=(m, 23 * 17)).astype(np.int64)
        label = np.empty(masks.shape[1], np.uint8); im = np.empty_like(label); sizes = np.zeros(2, np.int64)
        clib.oracle_im_binary(masks.ctypes.data, m, masks.shape[1], label.ctypes.data, im.ctypes.data, sizes.ctypes.data)
        e_label, e_im, e_sz, e_pred = ref_im.im_binary(list(masks))
        np.testing.assert_array_equal(label, e_label); np.testing.assert_array_equal(im, e_im)
        assert sizes[0] == e_sz and sizes[1] == e_pred
        cls = rng.integers(0, 9, size=(m, 23 * 17)).astype(np.int64)
        clib.oracle_im_multiclass(cls.ctypes.data, m, cls.shape[1], label.ctypes.data, im.ctypes.data, sizes.ctypes.data)
        e_label, e_im, e_sz = ref_im.im_multiclass(list(cls))
        np.testing.assert_array_equal(label, e_label); np.testing.assert_array_equal(im, e_im)
        assert sizes[0] == e_sz


def test_param_counts_match_readme():
    """README.md:25: '0.17 - 2.72 million parameters'."""
    assert ref_unet.count_params(3, 1, 0.5) == 171_561
    assert ref_unet.count_params(3, 35, 2.0) == 2_718_723
