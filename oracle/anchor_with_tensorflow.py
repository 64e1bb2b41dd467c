#!/usr/bin/env python
"""CPU ORACLE tooling (test infrastructure, NOT product code): re-anchor the oracle on the REAL reference graph.

The oracle's parity with TensorFlow's op numerics is unpinned because TensorFlow / tf_slim cannot be installed in the
build container.  Wherever they CAN be imported (tensorflow 2.6..2.12 + tf-slim 1.1.0, a checkout of the
reference), this script runs the reference's own ``HYPELCNNModel.create_tensor_graph`` / ``get_loss_func`` on the
seed-1234 synthetic batch with the oracle's initial variables assigned by TF variable name, and stores logits,
reconstruction, loss and every variable gradient in ``tests/golden/tf_anchor_hypelcnn.npz``.
``tests/test_oracle_tf_anchor.py`` compares the oracle against that file when it exists (rtol 1e-4, the north-star
tolerance) and reports "parity unpinned" by skipping when it does not.

    python oracle/anchor_with_tensorflow.py --reference /path/to/hypelcnn [--batch 48] [--patch 7] [--channels 145]
"""
import argparse
import json
import os
import sys

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "tf_anchor_hypelcnn.npz")

ap = argparse.ArgumentParser()
ap.add_argument("--reference", required=True, help="checkout of aligokalppeker/hypelcnn")
ap.add_argument("--batch", type=int, default=48)
ap.add_argument("--patch", type=int, default=7)
ap.add_argument("--channels", type=int, default=145)
ap.add_argument("--classes", type=int, default=15)
args = ap.parse_args()

try:
    import tensorflow as tf
    import tf_slim  # noqa: F401
except ImportError as e:
    sys.exit(f"tensorflow / tf_slim are not importable here ({e}); the oracle stays unpinned")

sys.path.insert(0, ROOT)
sys.path.insert(0, args.reference)
import torch  # noqa: E402
from oracle import hypelcnn_ref as R  # noqa: E402

alg = json.load(open(os.path.join(args.reference, "nnmodel", "modelconfigs", "alg_param_hypelcnn.json")))
alg["batch_size"] = args.batch
alg["drop_out_ratio"] = 0.0          # keep_prob 1: dropout is the identity, no random stream to reproduce
rng = numpy.random.default_rng(1234)
x = rng.random((args.batch, args.patch, args.patch, args.channels), dtype=numpy.float32)
labels = rng.integers(0, args.classes, args.batch).astype(numpy.int64)
one_hot = numpy.eye(args.classes, dtype=numpy.uint8)[labels]
variables = R.init_variables(args.patch, args.channels, args.classes, alg, seed=1234, dtype=torch.float32)

tf.compat.v1.disable_v2_behavior()
from common.common_nn_ops import ModelInputParams  # noqa: E402  (the reference's)
from nnmodel.HYPELCNNModel import HYPELCNNModel  # noqa: E402

with tf.Graph().as_default():
    x_ph = tf.compat.v1.placeholder(tf.float32, [None, args.patch, args.patch, args.channels])
    y_ph = tf.compat.v1.placeholder(tf.uint8, [None, args.classes])
    model = HYPELCNNModel()
    template = tf.compat.v1.make_template("nn_core", model.create_tensor_graph, class_count=args.classes)
    outputs = template(ModelInputParams(x=x_ph, y=y_ph, device_id="/cpu:0", is_training=True), algorithm_params=alg)
    loss = tf.reduce_mean(model.get_loss_func(outputs, y_ph))
    trainable = tf.compat.v1.trainable_variables()
    everything = tf.compat.v1.global_variables()
    gradients = tf.gradients(loss, trainable)
    missing = [v.op.name for v in everything if v.op.name not in variables]
    extra = [name for name in variables if name not in {v.op.name for v in everything}]
    if missing or extra:
        sys.exit(f"variable names differ: only in TF {missing[:5]}, only in the oracle {extra[:5]}")
    with tf.compat.v1.Session() as session:
        session.run(tf.compat.v1.global_variables_initializer())
        for v in everything:
            v.load(variables[v.op.name].numpy().reshape(v.shape.as_list()), session)
        feed = {x_ph: x, y_ph: one_hot}
        logits, recon, loss_value, grads = session.run([outputs.y_conv, outputs.image_output, loss, gradients], feed)

payload = {"x": x, "labels": labels, "logits": logits, "recon": recon, "loss": numpy.float64(loss_value),
           "alg": numpy.frombuffer(json.dumps(alg).encode(), dtype=numpy.uint8),
           "tf_version": numpy.frombuffer(tf.__version__.encode(), dtype=numpy.uint8)}
payload.update({"var/" + name: t.numpy() for name, t in variables.items()})
payload.update({"grad/" + v.op.name: g for v, g in zip(trainable, grads)})
numpy.savez_compressed(OUT, **payload)
print(f"wrote {OUT}: logits {logits.shape}, loss {loss_value:.6f}, {len(trainable)} gradients (tensorflow {tf.__version__})")
