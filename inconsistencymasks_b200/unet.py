"""Host-side mirror of the reference's ``unet.py`` for the inference path.

``get_unet`` keeps the reference signature (unet.py:46) and returns a ``B200UNet`` whose
``.predict`` behaves like ``tf.keras.Model.predict`` on the hot path (functions.py:3157,
3184, 3224): uint8 or float32 NHWC in (array or one-element list), float32 NHWC out.
The forward pass runs entirely in libimk's CUDA kernels (include/imk.h, row a1); there is
no PyTorch / NumPy fallback.

Weights are exchanged in Keras ``model.get_weights()`` order: per
Conv2D ``kernel (kh,kw,Cin,Cout)``, ``bias``; per BatchNormalization ``gamma, beta,
moving_mean, moving_variance`` -- 104 float32 arrays.  ``.npz`` files written by
``save_weights`` (arrays ``w000`` .. ``w103`` plus the constructor arguments) replace the
reference's ``.h5`` files, which need TensorFlow/h5py to read; ``tools/export_keras_weights.py``
is the one-liner to run where TensorFlow exists.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import lib, check
from .weights import count_params, init_weights, layer_plan

__all__ = ["get_unet", "B200UNet", "layer_plan", "init_weights", "load_model", "count_params"]


class B200UNet:
    """A packed U-Net resident on the current CUDA device."""

    def __init__(self, i_height, i_width, i_channels, num_outputmasks, alpha, actifuout, weights, ks=3, swap_rb=False):
        if actifuout not in ("sigmoid", "softmax"):
            raise ValueError(f"actifuout must be 'sigmoid' or 'softmax' (config.ini ACTIFU_OUTPUT), got {actifuout!r}")
        self.config = dict(i_height=int(i_height), i_width=int(i_width), i_channels=int(i_channels),
                           num_outputmasks=int(num_outputmasks), alpha=float(alpha), actifuout=actifuout, ks=int(ks))
        self._weights = [np.ascontiguousarray(w, dtype=np.float32) for w in weights]
        desc = _lib.UNetDesc(int(i_height), int(i_width), int(i_channels), int(num_outputmasks), float(alpha), int(ks),
                             _lib.IMK_ACT_SIGMOID if actifuout == "sigmoid" else _lib.IMK_ACT_SOFTMAX, int(bool(swap_rb)))
        n = len(self._weights)
        ptrs = (C.c_void_p * n)(*[w.ctypes.data for w in self._weights])
        sizes = (C.c_int64 * n)(*[w.size for w in self._weights])
        handle = C.c_void_p()
        check(lib.imk_unet_create(C.byref(desc), ptrs, sizes, n, C.byref(handle)))
        self.handle = handle
        self._swap_rb = bool(swap_rb)

    # -- Keras-like surface -------------------------------------------------------------
    @property
    def input_shape(self):
        c = self.config
        return (None, c["i_height"], c["i_width"], c["i_channels"])

    @property
    def output_shape(self):
        c = self.config
        return (None, c["i_height"], c["i_width"], c["num_outputmasks"])

    def count_params(self):
        n = C.c_int64()
        check(lib.imk_unet_param_count(self.handle, C.byref(n)))
        return int(n.value)

    def get_weights(self):
        return [w.copy() for w in self._weights]

    def save_weights(self, path):
        arrs = {f"w{i:03d}": w for i, w in enumerate(self._weights)}
        np.savez(path, __config__=np.array(repr(sorted(self.config.items()))), **arrs)

    def set_engine(self, engine):
        """'fused' (default: block-fused tcgen05, layer-wise where a block does not fit), 'tcgen05' (layer-wise
        implicit GEMM) or 'direct' (CUDA cores) -- which convolution engine runs the hidden layers."""
        check(lib.imk_unet_set_engine(self.handle, {"direct": 0, "tcgen05": 1, "fused": 2}[engine]))

    def set_swap_rb(self, flag):
        check(lib.imk_unet_set_swap_rb(self.handle, int(bool(flag))))
        self._swap_rb = bool(flag)

    def predict(self, x, batch_size=None, verbose=0, **_):
        """``model.predict`` (functions.py:3157): host array in, float32 NHWC host array out."""
        if isinstance(x, (list, tuple)):
            if len(x) != 1:
                raise ValueError("predict expects one input array (or a one-element list)")
            x = x[0]
        x = np.asarray(x)
        c = self.config
        if x.ndim != 4 or x.shape[1:] != (c["i_height"], c["i_width"], c["i_channels"]):
            raise ValueError(f"predict: expected [N,{c['i_height']},{c['i_width']},{c['i_channels']}], got {x.shape}")
        if x.dtype == np.uint8:
            dt = _lib.IMK_IN_U8
        else:
            x = x.astype(np.float32, copy=False)
            dt = _lib.IMK_IN_F32
        x = np.ascontiguousarray(x)
        out = np.empty(x.shape[:3] + (c["num_outputmasks"],), np.float32)
        check(lib.imk_unet_predict_host(self.handle, x.ctypes.data, dt, x.shape[0], out.ctypes.data))
        return out

    __call__ = predict

    def forward_device(self, images, stream=None):
        """Device path: ``images`` is a CUDA torch tensor uint8 / float32 [N,H,W,c]; returns float32 [N,H,W,K]."""
        import torch
        assert images.is_cuda and images.is_contiguous()
        dt = _lib.IMK_IN_U8 if images.dtype == torch.uint8 else _lib.IMK_IN_F32
        out = torch.empty(images.shape[:3] + (self.config["num_outputmasks"],), dtype=torch.float32, device=images.device)
        s = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        check(lib.imk_unet_forward(self.handle, images.data_ptr(), dt, images.shape[0], out.data_ptr(), s))
        return out

    def close(self):
        if getattr(self, "handle", None):
            lib.imk_unet_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def get_unet(i_height, i_width, i_channels, num_outputmasks, alpha, actifu, actifuout, ks=3, kernel_ini="he_normal",
             dropout_rate_encoder=0, dropout_rate_decoder=0, dropout_rate_bottleneck=0, weights=None, seed=0):
    """unet.py:46-67 with the same positional arguments.  ``actifu`` must be ``'relu'`` (config.ini:25,47,68,90);
    dropout is an inference no-op.  ``weights`` (get_weights() order) default to a seeded ``he_normal`` init."""
    if actifu != "relu":
        raise ValueError("the B200 path implements the reference configuration actifu='relu' only")
    if kernel_ini != "he_normal" and weights is None:
        raise ValueError("only kernel_ini='he_normal' is available for random initialisation")
    if weights is None:
        weights = init_weights(i_channels, num_outputmasks, alpha, ks, seed=seed, trained_like=False)
    return B200UNet(i_height, i_width, i_channels, num_outputmasks, alpha, actifuout, weights, ks=ks)


def load_model(path, custom_objects=None, compile=False):
    """Stand-in for ``tf.keras.models.load_model`` (09_ISIC_2018_IM.py:74-76) on ``.npz`` weight files."""
    import ast
    z = np.load(path, allow_pickle=False)
    cfg = dict(ast.literal_eval(str(z["__config__"])))
    n = len([k for k in z.files if k.startswith("w")])
    weights = [z[f"w{i:03d}"] for i in range(n)]
    return B200UNet(cfg["i_height"], cfg["i_width"], cfg["i_channels"], cfg["num_outputmasks"], cfg["alpha"],
                    cfg["actifuout"], weights, ks=cfg["ks"])
