#!/usr/bin/env python
"""Generate golden fixtures by EXECUTING the reference's own Python code.

Runs only in the build container (needs /root/reference, which does not exist on
the GPU box).  The outputs (``*.json`` / ``*.npz`` beside this file) are committed
and are what the tests read.

TensorFlow / tf_slim / tifffile are not installed, so they are replaced by
*recording stubs*: every slim layer call is logged (scope, kernel size, channel
counts, activation, normaliser) and returns a shape-only fake tensor.  All the
reference's pure-Python arithmetic therefore runs for real:

* ``common/common_nn_ops.py:546-564``  scale_in_to_out index arithmetic
* ``nnmodel/HYPELCNNModel.py:34-183``  the layer sequence, FC stage sizing
* ``common/common_nn_ops.py:45-106,169-185``  BasicDataSet pad/normalise + window slice
* ``loader/GRSS2018DataLoader.py:10-44``  mixed-resolution gather (numba)
* ``common/common_nn_ops.py:280-292``  class accuracies from a confusion matrix
* ``utilities/stat_extractor.py:24-62``  kappa

What stays unpinned: the numerics INSIDE the TF ops (conv, BN, Adam ...), which
are restated from the published semantics of the pinned versions.

usage: python tests/golden/make_golden.py
"""
import json
import os
import sys
import types

import numpy

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


# --------------------------------------------------------------------------- #
# recording stubs
# --------------------------------------------------------------------------- #
class _Stub(types.ModuleType):
    """A module whose every attribute is another stub; calling it returns a stub."""

    def __init__(self, name):
        super().__init__(name)
        self.__path__ = []

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        child = _Stub(self.__name__ + "." + item)
        setattr(self, item, child)
        return child

    def __call__(self, *a, **k):
        return _Stub(self.__name__ + "()")

    def __mro_entries__(self, bases):  # allow `class X(stub.Something)`
        return (object,)


class Dim:
    def __init__(self, v):
        self.value = v

    def __mul__(self, o):
        return Dim(self.value * (o.value if isinstance(o, Dim) else o))

    def __int__(self):
        return self.value

    def __index__(self):
        return self.value


def _as_int(v):
    return v.value if isinstance(v, Dim) else int(v)


TRACE = []
_ids = [0]


class FakeTensor:
    def __init__(self, shape, producer):
        self.shape = list(shape)
        self.id = _ids[0]
        _ids[0] += 1
        self.producer = producer

    def get_shape(self):
        return [Dim(s) for s in self.shape]

    def __add__(self, other):
        assert self.shape == other.shape, (self.shape, other.shape)
        out = FakeTensor(self.shape, "add")
        TRACE.append({"op": "add", "a": self.id, "b": other.id, "out": out.id, "shape": self.shape})
        return out

    def __sub__(self, other):
        out = FakeTensor(self.shape, "sub")
        TRACE.append({"op": "sub", "a": self.id, "b": other.id, "out": out.id})
        return out


class _Ctx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_scope_defaults = [{}]


def _arg_scope(ops, **kwargs):
    class C(_Ctx):
        def __enter__(s):
            _scope_defaults.append({**_scope_defaults[-1], **{(o.__name__, k): v for o in ops for k, v in kwargs.items()}})

        def __exit__(s, *a):
            _scope_defaults.pop()
            return False
    return C()


def _resolve(opname, kwargs, key, default=None):
    if key in kwargs:
        return kwargs[key]
    return _scope_defaults[-1].get((opname, key), default)


def _fn_name(f):
    if f is None:
        return None
    return getattr(f, "_gold_name", getattr(f, "__name__", str(f)))


def conv2d(inputs, num_outputs, kernel_size, scope=None, data_format=None, **kw):
    n = _as_int(num_outputs)
    out = FakeTensor(inputs.shape[:3] + [n], "conv2d")
    act = _resolve("conv2d", kw, "activation_fn", "relu")
    norm = _resolve("conv2d", kw, "normalizer_fn")
    TRACE.append({"op": "conv2d", "scope": scope, "in": inputs.id, "out": out.id,
                  "cin": inputs.shape[3], "cout": n, "kernel": [int(k) for k in kernel_size],
                  "activation": _fn_name(act), "normalizer": _fn_name(norm),
                  "norm_params": {k: (v if not isinstance(v, FakeTensor) else "tensor")
                                  for k, v in (_resolve("conv2d", kw, "normalizer_params") or {}).items()}})
    return out


def fully_connected(inputs, num_outputs, scope=None, **kw):
    n = _as_int(num_outputs)
    out = FakeTensor([inputs.shape[0], n], "fully_connected")
    act = kw["activation_fn"] if "activation_fn" in kw else _resolve("fully_connected", {}, "activation_fn", "relu")
    norm = _resolve("fully_connected", kw, "normalizer_fn")
    reg = kw["weights_regularizer"] if "weights_regularizer" in kw else _resolve("fully_connected", {}, "weights_regularizer")
    TRACE.append({"op": "fully_connected", "scope": scope, "in": inputs.id, "out": out.id,
                  "cin": inputs.shape[1], "cout": n, "activation": _fn_name(act),
                  "normalizer": _fn_name(norm), "regularized": reg is not None})
    return out


def dropout(inputs, keep_prob=0.5, is_training=True, **kw):
    out = FakeTensor(inputs.shape, "dropout")
    TRACE.append({"op": "dropout", "in": inputs.id, "out": out.id, "keep_prob": keep_prob,
                  "is_training": is_training})
    return out


def flatten(inputs, **kw):
    n = 1
    for s in inputs.shape[1:]:
        n *= s
    out = FakeTensor([inputs.shape[0], n], "flatten")
    TRACE.append({"op": "flatten", "in": inputs.id, "out": out.id, "size": n})
    return out


def batch_norm(*a, **k):
    raise AssertionError("batch_norm is only passed as normalizer_fn")


batch_norm._gold_name = "batch_norm"


def l2_regularizer(scale):
    return ("l2", scale)


def tf_repeat(input, repeats, axis):
    shape = list(input.shape)
    shape[axis] *= repeats
    out = FakeTensor(shape, "repeat")
    TRACE.append({"op": "repeat", "in": input.id, "out": out.id, "repeats": repeats, "axis": axis})
    return out


def tf_gather(params, indices, axis):
    shape = list(params.shape)
    shape[axis] = len(indices)
    out = FakeTensor(shape, "gather")
    TRACE.append({"op": "gather", "in": params.id, "out": out.id, "indices": [int(i) for i in indices],
                  "axis": axis})
    return out


def tf_concat(axis, values):
    shape = list(values[0].shape)
    shape[axis] = sum(v.shape[axis] for v in values)
    out = FakeTensor(shape, "concat")
    TRACE.append({"op": "concat", "ins": [v.id for v in values], "out": out.id, "axis": axis})
    return out


def leaky_relu(inp, alpha):
    raise AssertionError("called only inside the real graph")


def install_stubs():
    tf = _Stub("tensorflow")
    tf.device = lambda d: _Ctx()
    tf.compat.v1.name_scope = lambda n: _Ctx()
    tf.repeat = tf_repeat
    tf.gather = tf_gather
    tf.concat = tf_concat
    sig = lambda x: x
    sig._gold_name = "sigmoid"
    tf.sigmoid = sig
    tf.estimator.SessionRunHook = object
    sys.modules["tensorflow"] = tf
    for sub in ["tensorflow.python", "tensorflow.python.ops", "tensorflow.python.ops.gen_nn_ops",
                "tensorflow.python.ops.metrics_impl", "tensorflow.python.training",
                "tensorflow.python.training.summary_io", "tensorflow.python.data",
                "tensorflow.python.data.experimental", "tensorflow.python.training.session_run_hook",
                "tensorflow.python.training.basic_session_run_hooks", "tensorflow.python.ops.control_flow_ops",
                "tensorflow.python.ops.random_ops", "tensorflow_gan", "tensorflow_gan.python",
                "tensorflow_gan.python.namedtuples", "tensorflow_gan.python.train", "tensorflow_gan.python.losses",
                "tensorflow_gan.python.losses.tuple_losses", "tifffile", "tf_slim.learning", "tf_slim.metrics",
                "matplotlib", "matplotlib.pyplot", "matplotlib.ticker"]:
        sys.modules[sub] = _Stub(sub)
    sys.modules["tensorflow.python.training.session_run_hook"].SessionRunHook = object
    lr = sys.modules["tensorflow.python.ops.gen_nn_ops"]
    lr.leaky_relu = leaky_relu
    slim = _Stub("tf_slim")
    slim.conv2d = conv2d
    slim.fully_connected = fully_connected
    slim.dropout = dropout
    slim.flatten = flatten
    slim.batch_norm = batch_norm
    slim.arg_scope = _arg_scope
    slim.l2_regularizer = l2_regularizer
    sys.modules["tf_slim"] = slim
    sys.path.insert(0, REF)


# --------------------------------------------------------------------------- #
def trace_hypelcnn(patch, channels, classes, alg, is_training):
    from common.common_nn_ops import ModelInputParams
    from nnmodel.HYPELCNNModel import HYPELCNNModel
    TRACE.clear()
    _ids[0] = 0
    x = FakeTensor([-1, patch, patch, channels], "input")
    out = HYPELCNNModel().create_tensor_graph(
        ModelInputParams(x=x, y=None, device_id="/cpu:0", is_training=is_training), classes, alg)
    # lambdas (lrelu) are recorded by name "<lambda>"
    return {"patch": patch, "channels": channels, "classes": classes, "is_training": is_training,
            "alg": alg, "input_id": x.id, "y_conv": out.y_conv.id,
            "image_output": None if out.image_output is None else out.image_output.id,
            "histogram": [(h.tensor.id, h.name) for h in out.histogram_tensors],
            "trace": list(TRACE)}


def golden_scale_in_to_out():
    from common.common_nn_ops import scale_in_to_out
    pairs = [(145, 120), (145, 480), (120, 240), (240, 480), (60, 120), (120, 360), (480, 480), (480, 240),
             (240, 120), (120, 60), (360, 180), (180, 90), (480, 120), (49, 120), (65, 120), (51, 120),
             (120, 30), (65, 30), (7, 5), (5, 7), (3, 9), (145, 1200), (1200, 600), (100, 300)]
    res = {}
    for cin, cout in pairs:
        TRACE.clear()
        a = FakeTensor([-1, 7, 7, cin], "in")
        b = FakeTensor([-1, 7, 7, cout], "out")
        r = scale_in_to_out(a, b, axis_no=3)
        if r is a:
            res[f"{cin}->{cout}"] = {"mode": "identity", "idx": list(range(cout))}
        else:
            t = TRACE[-1]
            if t["op"] == "repeat":
                rep = t["repeats"]
                res[f"{cin}->{cout}"] = {"mode": "repeat", "repeats": rep, "idx": [j // rep for j in range(cin * rep)]}
            else:
                res[f"{cin}->{cout}"] = {"mode": "gather", "idx": t["indices"]}
    return res


def golden_dataset(rng):
    """BasicDataSet + window slice and GRSS2018 mixed-resolution gather on small random scenes."""
    from common.common_nn_ops import BasicDataSet, calculate_class_accuracies_using_confusion
    out = {}
    # same-resolution (GRSS2013 style): uint16 HSI, float LiDAR
    H, W, C, n = 9, 11, 6, 2
    casi = rng.integers(0, 16384, (H, W, C)).astype(numpy.uint16)
    lidar = (rng.random((H, W, 1)) * 50).astype(numpy.float32)
    ds = BasicDataSet(shadow_creator_dict=None, casi=casi.copy(), lidar=lidar.copy(), neighborhood=n, normalize=True)
    pts = numpy.array([[0, 0], [W - 1, H - 1], [3, 4], [W - 1, 0], [0, H - 1], [5, 2]], dtype=numpy.int32)
    patches = numpy.stack([ds.get_data_point(int(p[0]), int(p[1])) for p in pts])
    out.update(same_casi=casi, same_lidar=lidar, same_n=numpy.int32(n), same_pts=pts,
               same_patches=patches.astype(numpy.float32), same_patches_dtype=str(patches.dtype),
               same_casi_min=numpy.asarray(ds.casi_min), same_casi_max=numpy.asarray(ds.casi_max),
               same_lidar_min=numpy.float64(ds.lidar_min), same_lidar_max=numpy.float64(ds.lidar_max),
               same_shape=numpy.array(ds.get_data_shape()), same_scene=numpy.array(ds.get_scene_shape()),
               same_padded_casi=ds.casi.astype(numpy.float64), same_padded_lidar=ds.lidar.astype(numpy.float64))
    # un-normalised, float input
    casi_f = rng.random((H, W, C)).astype(numpy.float32)
    ds2 = BasicDataSet(shadow_creator_dict=None, casi=casi_f.copy(), lidar=lidar.copy(), neighborhood=1,
                       normalize=False)
    out.update(raw_casi=casi_f, raw_patches=numpy.stack([ds2.get_data_point(int(p[0]), int(p[1])) for p in pts]))
    # hsi-only
    ds3 = BasicDataSet(shadow_creator_dict=None, casi=casi_f.copy(), lidar=None, neighborhood=1, normalize=True)
    out.update(hsi_patches=numpy.stack([ds3.get_data_point(int(p[0]), int(p[1])) for p in pts]).astype(numpy.float32),
               hsi_shape=numpy.array(ds3.get_data_shape()))
    # GRSS2018 mixed resolution: casi at half the LiDAR resolution
    from loader.GRSS2018DataLoader import GRSS2018DataSet
    for n18 in (2, 5, 3):
        Hc, Wc, Cc = 8, 10, 5
        casi18 = rng.random((Hc, Wc, Cc)).astype(numpy.float32)
        lidar18 = rng.random((2 * Hc, 2 * Wc, 1)).astype(numpy.float32)
        ds18 = GRSS2018DataSet(shadow_creator_dict=None, casi=casi18.copy(), lidar=lidar18.copy(),
                               neighborhood=n18, normalize=True)
        sh = ds18.get_scene_shape()
        pts18 = numpy.array([[0, 0], [sh[1] - 1, sh[0] - 1], [3, 4], [7, 9], [sh[1] - 1, 0], [1, sh[0] - 2],
                             [6, 6], [2, 13]], dtype=numpy.int32)
        p18 = numpy.stack([ds18.get_data_point(int(p[0]), int(p[1])) for p in pts18])
        out.update({f"g18_{n18}_casi": casi18, f"g18_{n18}_lidar": lidar18, f"g18_{n18}_pts": pts18,
                    f"g18_{n18}_patches": p18.astype(numpy.float32), f"g18_{n18}_scene": numpy.array(sh),
                    f"g18_{n18}_padded_casi": ds18.casi.astype(numpy.float64),
                    f"g18_{n18}_padded_lidar": ds18.lidar.astype(numpy.float64)})
    # confusion-derived class accuracies + kappa
    from utilities.stat_extractor import calc_kappa
    conf = rng.integers(0, 40, (15, 15)).astype(numpy.int32)
    conf[4, :] = 0  # a class with no ground truth
    conf[:, 9] = 0  # a class never predicted
    recall, precision = calculate_class_accuracies_using_confusion(conf, range(0, 15))
    out.update(conf=conf, conf_recall=recall, conf_precision=precision, conf_kappa=numpy.float64(calc_kappa(conf)))
    return out


def main():
    install_stubs()
    rng = numpy.random.default_rng(1234)
    alg = json.load(open(os.path.join(REF, "nnmodel/modelconfigs/alg_param_hypelcnn.json")))
    traces = {
        "c2_train": trace_hypelcnn(7, 145, 15, alg, True),
        "c2_eval": trace_hypelcnn(7, 145, 15, alg, False),
        "c3_train": trace_hypelcnn(11, 49, 20, alg, True),
        "c5_train": trace_hypelcnn(3, 65, 11, alg, True),
        "tiny_train": trace_hypelcnn(3, 10, 4, {**alg, "filter_count": 32}, True),
        "nonres_train": trace_hypelcnn(
            7, 145, 15, json.load(open(os.path.join(REF, "nnmodel/modelconfigs/alg_param_hypelcnn_nonres.json"))),
            True),
    }
    with open(os.path.join(HERE, "hypelcnn_graph_trace.json"), "w") as f:
        json.dump(traces, f, indent=0, sort_keys=True)
    with open(os.path.join(HERE, "scale_in_to_out.json"), "w") as f:
        json.dump(golden_scale_in_to_out(), f, sort_keys=True)
    numpy.savez_compressed(os.path.join(HERE, "dataset_golden.npz"), **golden_dataset(rng))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
