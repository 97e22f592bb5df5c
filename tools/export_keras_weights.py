#!/usr/bin/env python
"""Convert a trained Keras ``.h5`` model of the reference (unet.py:46-67) into the ``.npz`` weight
file ``inconsistencymasks_b200.unet.load_model`` reads.  Run this where TensorFlow exists (it is not
available in the B200 image):

    python tools/export_keras_weights.py MODEL.h5 MODEL.npz --height 256 --width 256 --channels 3 \
           --outputs 1 --alpha 1.0 --activation sigmoid

Arrays ``w000`` .. ``w103`` are ``model.get_weights()`` in order: per Conv2D kernel (kh,kw,Cin,Cout)
and bias, per BatchNormalization gamma, beta, moving_mean, moving_variance.
"""
import argparse

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("h5")
    ap.add_argument("npz")
    ap.add_argument("--height", type=int, required=True)
    ap.add_argument("--width", type=int, required=True)
    ap.add_argument("--channels", type=int, required=True)
    ap.add_argument("--outputs", type=int, required=True)
    ap.add_argument("--alpha", type=float, required=True)
    ap.add_argument("--activation", choices=["sigmoid", "softmax"], required=True)
    ap.add_argument("--ks", type=int, default=3)
    args = ap.parse_args()
    import tensorflow as tf   # noqa: only needed here
    model = tf.keras.models.load_model(args.h5, compile=False)
    weights = [np.asarray(w, np.float32) for w in model.get_weights()]
    if len(weights) != 104:
        raise SystemExit(f"expected 104 arrays (24 convs + 14 batch norms), the model has {len(weights)}")
    cfg = dict(i_height=args.height, i_width=args.width, i_channels=args.channels, num_outputmasks=args.outputs,
               alpha=args.alpha, actifuout=args.activation, ks=args.ks)
    np.savez(args.npz, __config__=np.array(repr(sorted(cfg.items()))), **{f"w{i:03d}": w for i, w in enumerate(weights)})
    print(f"wrote {args.npz}: {sum(w.size for w in weights)} parameters")


if __name__ == "__main__":
    main()
