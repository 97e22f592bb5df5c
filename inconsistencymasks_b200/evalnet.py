"""EvalNet forward of the IM++ scripts: drop-in for the reference's ``evalnet.py`` at inference time.

    get_evalnet        evalnet.py:24-47     ISIC: image + mask -> one sigmoid score
    get_evalnet_miou   evalnet.py:49-73     HeLa / SUIM / Cityscapes: image + one-hot class map -> per-class 'iou' and
                                            'detection' scores

Same positional arguments; the returned object has the Keras calls the scripts use (``predict([A, B])`` as in
functions.py:5735 and 6010-6013, ``count_params``).  The forward pass runs in libimk (csrc/imk_evalnet.cu) on the U-Net's
convolution engines.  ``weights`` are in the order evalnet.py creates its layers (``weights.evalnet_plan``);
``weights_from_keras`` collects them from a trained Keras model where TensorFlow exists.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import lib, check
from .weights import evalnet_plan, init_evalnet_weights

__all__ = ["B200EvalNet", "get_evalnet", "get_evalnet_miou", "weights_from_keras", "load_evalnet", "evalnet_plan", "init_evalnet_weights"]


class B200EvalNet:
    def __init__(self, i_height, i_width, inputA_channels, inputB_channels, alpha, weights, n_heads=1, ksi=3,
                 normalize_A=True, normalize_B=True):
        self.config = dict(i_height=i_height, i_width=i_width, inputA_channels=inputA_channels, inputB_channels=inputB_channels,
                           alpha=float(alpha), n_heads=n_heads, ksi=ksi, normalize_A=bool(normalize_A), normalize_B=bool(normalize_B))
        self._weights = [np.ascontiguousarray(w, dtype=np.float32) for w in weights]
        self._handles = {}
        self._create(onehot=inputB_channels > 4)          # a dense (uint8 channels) B input has at most 4 channels

    def _create(self, onehot):
        if onehot in self._handles:
            return self._handles[onehot]
        c = self.config
        d = _lib.EvalNetDesc(c["i_height"], c["i_width"], c["inputA_channels"], c["inputB_channels"], c["alpha"], c["ksi"],
                             int(c["normalize_A"]), int(c["normalize_B"]), int(onehot), c["n_heads"])
        ptrs = (C.c_void_p * len(self._weights))(*[w.ctypes.data for w in self._weights])
        sizes = (C.c_int64 * len(self._weights))(*[w.size for w in self._weights])
        h = C.c_void_p()
        check(lib.imk_evalnet_create(C.byref(d), ptrs, sizes, len(self._weights), C.byref(h)))
        self._handles[onehot] = h
        return h

    def count_params(self):
        n = C.c_int64()
        check(lib.imk_evalnet_param_count(next(iter(self._handles.values())), C.byref(n)))
        return int(n.value)

    def get_weights(self):
        return [w.copy() for w in self._weights]

    def save_weights(self, path):
        """``.npz`` with the construction arguments + the weights in ``evalnet_plan`` order (read back by ``load_evalnet``)."""
        arrs = {f"w{i:03d}": w for i, w in enumerate(self._weights)}
        np.savez(path, __config__=np.array(repr(sorted(self.config.items()))), **arrs)

    def predict(self, x, batch_size=None, verbose=0, **_):
        """``model.predict([A, B])``: A uint8 [N,H,W,cA]; B [N,H,W,cB] -- integer 0/1 one-hot maps (functions.py:6004-6006,
        fed to the model that does not normalise B) are sent as a class map and expanded by the first layer itself,
        anything else as uint8 channels.  Returns float32 [N,1] (get_evalnet) or [iou [N,K], detection [N,K]]."""
        from . import functions as F
        torch = F._torch()
        a, b = (np.asarray(v) for v in x)
        c = self.config
        if a.shape[1:] != (c["i_height"], c["i_width"], c["inputA_channels"]) or b.shape[:3] != a.shape[:3] or b.shape[3] != c["inputB_channels"]:
            raise ValueError(f"predict: unexpected input shapes {a.shape}, {b.shape}")
        onehot = (not c["normalize_B"]) and b.dtype.kind in "iub" and c["inputB_channels"] > 1 and bool(((b == 0) | (b == 1)).all()) \
            and bool((b.sum(-1) <= 1).all())
        if onehot:
            cls = np.where(b.any(-1), b.argmax(-1), 255).astype(np.uint8)      # an all-zero row stays all-zero (class >= K)
            d_b = F._dev(cls)
        else:
            if b.dtype != np.uint8 and not np.array_equal(b, b.astype(np.uint8)):
                raise ValueError("input B must hold integers 0..255")
            d_b = F._dev(b.astype(np.uint8))
        d_a = F._dev(a.astype(np.uint8))
        n, k = a.shape[0], (1 if c["n_heads"] == 1 else c["inputB_channels"])
        o0 = torch.empty((n, k), dtype=torch.float32, device="cuda")
        o1 = torch.empty((n, k), dtype=torch.float32, device="cuda") if c["n_heads"] == 2 else None
        check(lib.imk_evalnet_forward(self._create(onehot), d_a.data_ptr(), d_b.data_ptr(), n, 0, o0.data_ptr(),
                                      o1.data_ptr() if o1 is not None else None, F._stream()))
        return o0.cpu().numpy() if o1 is None else [o0.cpu().numpy(), o1.cpu().numpy()]

    __call__ = predict

    def close(self):
        for h in getattr(self, "_handles", {}).values():
            lib.imk_evalnet_destroy(h)
        self._handles = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def get_evalnet(i_height, i_width, inputA_channels, inputB_channels, alpha=2, actifu="relu", ksi=3, kernel_ini="he_normal",
                normalize_A=True, normalize_B=True, weights=None, seed=0):
    """evalnet.py:24-47."""
    if actifu != "relu":
        raise ValueError("the B200 path implements actifu='relu' only")
    if weights is None:
        weights = init_evalnet_weights(inputA_channels, inputB_channels, alpha, 1, ksi, seed)
    return B200EvalNet(i_height, i_width, inputA_channels, inputB_channels, alpha, weights, 1, ksi, normalize_A, normalize_B)


def get_evalnet_miou(i_height, i_width, inputA_channels, inputB_channels, alpha=2, actifu="relu", ksi=3, kernel_ini="he_normal",
                     normalize_A=True, normalize_B=False, weights=None, seed=0):
    """evalnet.py:49-73."""
    if actifu != "relu":
        raise ValueError("the B200 path implements actifu='relu' only")
    if weights is None:
        weights = init_evalnet_weights(inputA_channels, inputB_channels, alpha, 2, ksi, seed)
    return B200EvalNet(i_height, i_width, inputA_channels, inputB_channels, alpha, weights, 2, ksi, normalize_A, normalize_B)


def weights_from_keras(model):
    """Collect a Keras EvalNet's weights in ``evalnet_plan`` order (to be run where TensorFlow exists).  A functional model
    with two branches lists its layers by graph depth, interleaving the branches; this walks each branch from its input."""
    def chain(tensor_layer):
        out, layer = [], tensor_layer
        while True:
            nxt = [n.outbound_layer if hasattr(n, "outbound_layer") else n.operation for n in layer._outbound_nodes]
            if len(nxt) != 1 or type(nxt[0]).__name__ == "Concatenate":
                return out, (nxt[0] if nxt else None)
            layer = nxt[0]
            if layer.get_weights():
                out.append(layer)
    ins = [l for l in model.layers if type(l).__name__ == "InputLayer"]
    a, cat = chain(ins[0])
    b, _ = chain(ins[1])
    trunk, layer = [], cat
    while layer is not None:
        if layer.get_weights():
            trunk.append(layer)
        nxt = [n.outbound_layer if hasattr(n, "outbound_layer") else n.operation for n in layer._outbound_nodes]
        if len(nxt) != 1:
            trunk += [l for l in nxt if l.get_weights()]
            break
        layer = nxt[0]
    ws = []
    for l in a + b + trunk:
        ws += [np.asarray(w, np.float32) for w in l.get_weights()]
    return ws


def load_evalnet(path, custom_objects=None, compile=False):
    """Stand-in for ``tf.keras.models.load_model`` on an EvalNet ``.npz`` (``B200EvalNet.save_weights`` or
    ``tools/export_keras_weights.py --evalnet``)."""
    import ast
    z = np.load(path, allow_pickle=False)
    cfg = dict(ast.literal_eval(str(z["__config__"])))
    n = len([k for k in z.files if k.startswith("w")])
    return B200EvalNet(cfg["i_height"], cfg["i_width"], cfg["inputA_channels"], cfg["inputB_channels"], cfg["alpha"],
                       [z[f"w{i:03d}"] for i in range(n)], cfg["n_heads"], cfg["ksi"], cfg["normalize_A"], cfg["normalize_B"])
