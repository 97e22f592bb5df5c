"""ctypes binding of libimk.so (include/imk.h).  There is no fallback: if the CUDA
library is missing or fails to load, importing this module raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IMK_LIB") or os.path.join(_HERE, "libimk.so")     # IMK_LIB: A/B runs against another build

IMK_ACT_SIGMOID, IMK_ACT_SOFTMAX = 0, 1
IMK_IN_U8, IMK_IN_F32 = 0, 1
IMK_MAX_MODELS = 16


class ImkError(RuntimeError):
    pass


class UNetDesc(C.Structure):
    _fields_ = [("height", C.c_int), ("width", C.c_int), ("in_channels", C.c_int),
                ("num_outputmasks", C.c_int), ("alpha", C.c_float), ("ks", C.c_int),
                ("act_out", C.c_int), ("swap_rb", C.c_int)]


class ProfileEntry(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("tag", C.c_int), ("launches", C.c_int64), ("total_ms", C.c_double)]


class EvalNetDesc(C.Structure):
    _fields_ = [("height", C.c_int), ("width", C.c_int), ("a_channels", C.c_int), ("b_channels", C.c_int), ("alpha", C.c_float),
                ("ks", C.c_int), ("normalize_a", C.c_int), ("normalize_b", C.c_int), ("b_onehot", C.c_int), ("n_heads", C.c_int)]


class AugParams(C.Structure):
    """imk_aug_params (include/imk.h): the host-chosen operations of one augmented image."""
    _fields_ = [("flip_v", C.c_int), ("flip_h", C.c_int), ("rot", C.c_int), ("scale_on", C.c_int),
                ("alpha", C.c_float), ("beta", C.c_float), ("blur_k", C.c_int), ("noise_max", C.c_int), ("seed", C.c_uint64)]


_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> (restype, argtypes); mirrors include/imk.h one to one
SIGNATURES = {
    "imk_version": (_i, []),
    "imk_last_error": (C.c_char_p, []),
    "imk_launch_count": (_i64, []),
    "imk_max_chunk": (_i64, []),
    "imk_set_max_chunk": (_i, [_i64]),
    "imk_device_available": (_i, []),
    "imk_profile_begin": (_i, []),
    "imk_profile_end": (_i, [C.POINTER(ProfileEntry), _i, C.POINTER(_i)]),
    "imk_masks_to_im_binary": (_i, [_vp, _i, _i64, _vp, _vp, _vp, _vp]),
    "imk_masks_to_im_multiclass": (_i, [_vp, _i, _i64, _vp, _vp, _vp, _vp]),
    "imk_im_binary": (_i, [_vp, _i, _i64, _i, _i, _i, _f, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "imk_im_multiclass": (_i, [_vp, _i, _i64, _i, _i, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "imk_erode_u8": (_i, [_vp, _vp, _i64, _i, _i, _i, _vp]),
    "imk_dilate_u8": (_i, [_vp, _vp, _i64, _i, _i, _i, _vp]),
    "imk_blank": (_i, [_vp, _i64, _i, _i, _vp, _i, _vp, _i, _vp]),
    "imk_unet_create": (_i, [C.POINTER(UNetDesc), _vp, _vp, _i, C.POINTER(_vp)]),
    "imk_unet_destroy": (None, [_vp]),
    "imk_unet_param_count": (_i, [_vp, C.POINTER(_i64)]),
    "imk_unet_set_engine": (_i, [_vp, _i]),
    "imk_unet_set_swap_rb": (_i, [_vp, _i]),
    "imk_unet_forward": (_i, [_vp, _vp, _i, _i64, _vp, _vp]),
    "imk_unet_predict_host": (_i, [_vp, _vp, _i, _i64, _vp]),
    "imk_ensemble_im_binary": (_i, [_vp, _i, _vp, _i64, _i, _f, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "imk_ensemble_im_multiclass": (_i, [_vp, _i, _vp, _i64, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "imk_pseudo_label_binary_host": (_i, [_vp, _i, _vp, _i64, _i, _f, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i64]),
    "imk_pseudo_label_multiclass_host": (_i, [_vp, _i, _vp, _i64, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i64]),
    "imk_pseudo_label_binary_host_packed": (_i, [_vp, _i, _vp, _i64, _i, _f, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i64]),
    "imk_pseudo_label_multiclass_host_packed": (_i, [_vp, _i, _vp, _i64, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i64]),
    "imk_seg_counts_binary": (_i, [_vp, _vp, _i64, _i64, _vp, _vp]),
    "imk_seg_counts_multiclass": (_i, [_vp, _vp, _i64, _i64, _vp, _vp]),
    "imk_pack_bits": (_i, [_vp, _i64, _vp, _vp]),
    "imk_evalnet_create": (_i, [C.POINTER(EvalNetDesc), _vp, _vp, _i, C.POINTER(_vp)]),
    "imk_evalnet_destroy": (None, [_vp]),
    "imk_evalnet_param_count": (_i, [_vp, C.POINTER(_i64)]),
    "imk_evalnet_forward": (_i, [_vp, _vp, _vp, _i64, _i, _vp, _vp, _vp]),
    "imk_augment_u8": (_i, [_vp, _vp, _i64, _i, _i, _i, _i, C.POINTER(AugParams), _vp, _vp, _vp, _vp]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m inconsistencymasks_b200.build` "
            "(nvcc, sm_100a).  inconsistencymasks_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so is stale: loud on purpose
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc: int) -> None:
    if rc != 0:
        msg = lib.imk_last_error()
        raise ImkError(f"libimk error {rc}: {msg.decode(errors='replace') if msg else '?'}")


def device_available() -> bool:
    return bool(lib.imk_device_available())


def launch_count() -> int:
    return int(lib.imk_launch_count())


def profile_begin() -> None:
    check(lib.imk_profile_begin())


def profile_end():
    """-> list of dicts {name, tag, launches, total_ms}: device time per kernel since profile_begin()."""
    cap = 256
    buf = (ProfileEntry * cap)()
    n = C.c_int()
    check(lib.imk_profile_end(buf, cap, C.byref(n)))
    return [dict(name=buf[i].name.decode(), tag=int(buf[i].tag), launches=int(buf[i].launches),
                 total_ms=float(buf[i].total_ms)) for i in range(min(cap, n.value))]
