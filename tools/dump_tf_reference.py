#!/usr/bin/env python
"""Dumps what the REFERENCE's own TensorFlow graph computes on the committed golden inputs, so that the oracle (and through it
the CUDA path) can be pinned against the real thing the day a TensorFlow box is available.

    python tools/dump_tf_reference.py --reference /path/to/lacibeb-GA3C/ga3c [--out tests/golden/tf_reference_b4.npz]

NOT EXECUTED IN THE BUILD CONTAINER: TensorFlow is not installable there, which is why DESIGN.md says "parity
unpinned" for the network arithmetic.  tests/test_oracle.py::test_tf_reference_pin consumes the file when it exists.

No line of the reference is restated here.  The conv NetworkVP is assembled from the reference's OWN methods:
  * the class is the reference's `NetworkVP_discrate.Network` (heads, losses, RMSProp: NetworkVP_discrate.py:39-130), built
    with Config.DENSE_LAYERS = (256,);
  * its single `dense_layer(self.x, 256, 'dense1_1_p')` call (:55) is intercepted and answered with the trunk of
    NetworkDNav.py:81-90 built from the reference's own `conv2d_layer` (NetworkVP.py:212-228) and `dense_layer`
    (ReLU, as NetworkDNav.py:256-269): reshape [B,84,84,4] -> conv11 8x8/16/s4 -> conv12 4x4/32/s2 -> flatten -> dense1 256.
Weights come from oracle_np.init_params(default_rng(12345)) (initialisers are overwritten), inputs from tests/_parity.make_case(4):
the same case as tests/golden/network_b4.npz.

Works with TensorFlow 1.x, or TensorFlow 2.x through tensorflow.compat.v1 (v2 behaviour disabled).
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# variable scope in the intercepted graph -> name used by ga3c_b200 / the oracle (TF creation order)
NAMES = ["conv11/w:0", "conv11/b:0", "conv12/w:0", "conv12/b:0", "dense1/w:0", "dense1/b:0",
         "logits_v/w:0", "logits_v/b:0", "logits_p/w:0", "logits_p/b:0"]


def import_tf():
    import tensorflow as tf
    if int(tf.__version__.split(".")[0]) >= 2:
        import tensorflow.compat.v1 as tf1
        tf1.disable_v2_behavior()
        sys.modules["tensorflow"] = tf1          # the reference does `import tensorflow as tf` and uses the 1.x API
        return tf1
    return tf


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", required=True, help="the reference's ga3c/ directory")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "tf_reference_b4.npz"))
    ap.add_argument("--batch", type=int, default=4)
    args = ap.parse_args()

    tf = import_tf()
    sys.dont_write_bytecode = True
    sys.path.insert(0, args.reference)
    from Config import Config                       # the reference's
    Config.DENSE_LAYERS = (256,)
    Config.TENSORBOARD = False
    Config.LOAD_CHECKPOINT = False
    Config.SAVE_MODELS = False
    Config.USE_GRAD_CLIP = False
    Config.DUAL_RMSPROP = False
    import NetworkVP_discrate as ref                # the reference's

    class ConvNetworkVP(ref.Network):
        def dense_layer(self, input, out_dim, name, func=tf.nn.sigmoid):
            base = super(ConvNetworkVP, self).dense_layer
            if name == "dense1_1_p":                # NetworkVP_discrate.py:55 -> the trunk of NetworkDNav.py:81-90
                x4 = tf.reshape(input, [-1, 84, 84, 4])
                n1 = self.conv2d_layer(x4, 8, 16, "conv11", strides=[1, 4, 4, 1])
                n2 = self.conv2d_layer(n1, 4, 32, "conv12", strides=[1, 2, 2, 1])
                flat = tf.reshape(n2, [-1, 11 * 11 * 32])
                return base(flat, 256, "dense1", func=tf.nn.relu)
            return base(input, out_dim, name, func=func)

    from oracle import oracle_np as onp
    from _parity import make_case
    params, x, y_r, a = make_case(args.batch)
    net = ConvNetworkVP("/cpu:0", "tfdump", 6, onp.STATE_DIM)
    with net.graph.as_default():
        tvars = {v.name: v for v in tf.global_variables()}
        missing = [n for n in NAMES if n not in tvars]
        if missing:
            raise SystemExit(f"variables not found in the reference graph: {missing}; have {sorted(tvars)}")
        for n in NAMES:
            net.sess.run(tvars[n].assign(params[n]))
        ordered = [tvars[n] for n in NAMES]
        grads = tf.gradients(net.cost_all, ordered)
        feed = {net.x: x, net.y_r: y_r, net.action_index: a, net.var_beta: net.beta, net.var_learning_rate: net.learning_rate}
        p, v, c1, c2, cv, call = net.sess.run([net.softmax_p, net.logits_v, net.cost_p_1_agg, net.cost_p_2_agg, net.cost_v,
                                               net.cost_all], feed_dict=feed)
        g = net.sess.run(grads, feed_dict=feed)
        net.train(x, y_r, a, None, None, 0)                               # NetworkVP.py:254-257
        after = {k: net.sess.run(t) for k, t in {v_.name: v_ for v_ in tf.global_variables()}.items()}
    out = {"p": p, "v": v, "losses": np.array([c1, c2, -(c1 + c2), cv, call], dtype=np.float64),
           "lr": np.float64(net.learning_rate), "beta": np.float64(net.beta), "batch": np.int64(args.batch),
           "tf_version": np.array(tf.__version__)}
    for n, gi in zip(NAMES, g):
        out["grad_" + n] = gi
    for k, val in after.items():                    # weights, '<var>/RMSProp:0' (ms), '<var>/RMSProp_1:0' (mom), 'step:0'
        out["after_" + k] = val
    np.savez(args.out, **out)
    print(f"wrote {args.out}: p {p.shape}, v {v.shape}, cost_all {call:.6f}, {len(after)} variables after one train step")


if __name__ == "__main__":
    main()
