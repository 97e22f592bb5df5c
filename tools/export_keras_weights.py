#!/usr/bin/env python
"""Convert a trained Keras ``.h5`` model of the reference (unet.py:46-67) into the ``.npz`` weight
file ``inconsistencymasks_b200.unet.load_model`` reads.  Run this where TensorFlow exists (it is not
available in the B200 image):

    python tools/export_keras_weights.py MODEL.h5 MODEL.npz --height 256 --width 256 --channels 3 \
           --outputs 1 --alpha 1.0 --activation sigmoid

Arrays ``w000`` .. ``w103`` are ``model.get_weights()`` in order: per Conv2D kernel (kh,kw,Cin,Cout)
and bias, per BatchNormalization gamma, beta, moving_mean, moving_variance.

EvalNets (evalnet.py:24-73) with ``--evalnet {1,2}`` (1 = get_evalnet, 2 = get_evalnet_miou heads): ``--channels`` is input A's,
``--outputs`` input B's channel count; the weights are collected branch by branch in layer-creation order
(``inconsistencymasks_b200.evalnet.weights_from_keras``; ``model.get_weights()`` interleaves the two branches) and read back
by ``inconsistencymasks_b200.evalnet.load_evalnet``:

    python tools/export_keras_weights.py EVALNET.h5 EVALNET.npz --evalnet 2 --height 256 --width 256 --channels 3 \
           --outputs 9 --alpha 2.0 --activation sigmoid
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
    ap.add_argument("--evalnet", type=int, default=0, choices=[0, 1, 2], help="0: U-Net; 1 / 2: EvalNet with that many Dense heads")
    ap.add_argument("--no-normalize-b", action="store_true", help="EvalNet: input B is not divided by 255 (get_evalnet_miou default)")
    args = ap.parse_args()
    import tensorflow as tf   # noqa: only needed here
    model = tf.keras.models.load_model(args.h5, compile=False)
    if args.evalnet:
        import os, sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from inconsistencymasks_b200.weights import evalnet_plan
        from inconsistencymasks_b200.evalnet import weights_from_keras
        weights = weights_from_keras(model)
        plan = evalnet_plan(args.channels, args.outputs, args.alpha, args.evalnet, args.ks)
        want = sum(2 if it[0] != "bn" else 4 for it in plan)
        if len(weights) != want:
            raise SystemExit(f"expected {want} arrays for this EvalNet, collected {len(weights)}")
        cfg = dict(i_height=args.height, i_width=args.width, inputA_channels=args.channels, inputB_channels=args.outputs, alpha=args.alpha,
                   n_heads=args.evalnet, ksi=args.ks, normalize_A=True, normalize_B=not (args.no_normalize_b or args.evalnet == 2))
        np.savez(args.npz, __config__=np.array(repr(sorted(cfg.items()))), **{f"w{i:03d}": w for i, w in enumerate(weights)})
        print(f"wrote {args.npz}: EvalNet, {sum(w.size for w in weights)} parameters")
        return
    weights = [np.asarray(w, np.float32) for w in model.get_weights()]
    if len(weights) != 104:
        raise SystemExit(f"expected 104 arrays (24 convs + 14 batch norms), the model has {len(weights)}")
    cfg = dict(i_height=args.height, i_width=args.width, i_channels=args.channels, num_outputmasks=args.outputs,
               alpha=args.alpha, actifuout=args.activation, ks=args.ks)
    np.savez(args.npz, __config__=np.array(repr(sorted(cfg.items()))), **{f"w{i:03d}": w for i, w in enumerate(weights)})
    print(f"wrote {args.npz}: {sum(w.size for w in weights)} parameters")


if __name__ == "__main__":
    main()
