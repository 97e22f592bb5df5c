"""CPU-only checks of the C-ABI boundary: libimk.so loads, exports every symbol that
include/imk.h declares, and fails loudly (no fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from inconsistencymasks_b200 import build
    build.build()
    from inconsistencymasks_b200 import _lib
    return _lib


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "imk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(imk_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 20
    raw = C.CDLL(lib.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), f"{s} declared in include/imk.h but not exported by libimk.so"
    # and the ctypes table binds exactly the declared set
    assert sorted(lib.SIGNATURES) == syms


def test_version_and_error_string(lib):
    assert lib.lib.imk_version() == 200
    assert isinstance(lib.lib.imk_last_error(), bytes)


def test_argument_validation_needs_no_gpu(lib):
    # NULL arguments are rejected before any CUDA call
    rc = lib.lib.imk_im_binary(None, 2, 1, 16, 16, 1, 0.5, 1, None, 3, 1, 1, None, None, None, None, None, None)
    assert rc == -1
    assert b"NULL" in lib.lib.imk_last_error()
    with pytest.raises(lib.ImkError):
        lib.check(rc)
    rc = lib.lib.imk_erode_u8(None, None, 1, 4, 4, 3, None)
    assert rc == -1


@pytest.mark.skipif(os.environ.get("IMK_EXPECT_GPU") == "1", reason="GPU box")
def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from inconsistencymasks_b200 import unet, functions
    assert not lib.device_available()
    with pytest.raises(lib.ImkError):
        unet.get_unet(32, 32, 3, 1, 0.5, "relu", "sigmoid")
    with pytest.raises(lib.ImkError):
        functions.pred_masks_to_im_binary([np.zeros((4, 4, 1), int), np.ones((4, 4, 1), int)])


def test_plan_matches_reference_param_counts():
    """README.md:25 '0.17 - 2.72 million'; exact counts from SURVEY.md section 2.1."""
    from inconsistencymasks_b200 import unet
    assert unet.count_params(3, 1, 0.5) == 171_561
    assert unet.count_params(3, 1, 1.0) == 681_681
    assert unet.count_params(1, 3, 1.0) == 681_683
    assert unet.count_params(3, 9, 1.0) == 681_817
    assert unet.count_params(3, 35, 1.0) == 682_259
    assert unet.count_params(3, 9, 2.0) == 2_717_865
    assert unet.count_params(3, 35, 2.0) == 2_718_723
    assert len(unet.init_weights(3, 1, 0.5)) == 104
    from oracle import ref_unet
    assert unet.layer_plan(3, 9, 1.25) == ref_unet.layer_plan(3, 9, 1.25)
