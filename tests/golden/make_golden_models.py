#!/usr/bin/env python
"""Graph traces of DUALCNNModel and CONCNNModel produced by EXECUTING the reference's model code
(nnmodel/DUALCNNModel.py:11-104, nnmodel/CONCNNModel.py:23-64) against the recording stubs of make_golden.py: every
slim layer call is logged with its scope, kernel, channel counts, activation and the dropout keep_prob; tensor adds,
concats, the band split / spatial crop of DUALCNN and CONCNN's local_response_normalization calls are logged too.
Build container only; ``dualcnn_graph_trace.json`` / ``concnn_graph_trace.json`` are committed and read by
tests/test_model_traces.py, which holds the oracles' structure (and so the engine's, which is tested against them)
to what the reference's own code builds.

usage: python tests/golden/make_golden_models.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402


def _getitem(self, key):
    """NHWC slicing of a fake tensor: records the crop, returns the cropped shape."""
    key = key if isinstance(key, tuple) else (key,)
    shape, crop = list(self.shape), []
    for axis, k in enumerate(key):
        if isinstance(k, slice) and not (k.start is None and k.stop is None):
            n = shape[axis]
            start, stop, _ = k.indices(n) if n >= 0 else (k.start, k.stop, 1)
            shape[axis] = stop - start
            crop.append([axis, start, stop])
    out = G.FakeTensor(shape, "slice")
    G.TRACE.append({"op": "slice", "in": self.id, "out": out.id, "crop": crop, "shape": shape})
    return out


def _split(axis, num_or_size_splits, value):
    outs = []
    for size in num_or_size_splits:
        shape = list(value.shape)
        shape[axis] = int(size)
        outs.append(G.FakeTensor(shape, "split"))
    G.TRACE.append({"op": "split", "in": value.id, "outs": [o.id for o in outs], "axis": axis,
                    "sizes": [int(s) for s in num_or_size_splits]})
    return outs


def _lrn(inp, *args, **kwargs):
    out = G.FakeTensor(inp.shape, "lrn")
    G.TRACE.append({"op": "lrn", "in": inp.id, "out": out.id, "args": list(args), "kwargs": kwargs})
    return out


def trace(model_cls, patch, channels, classes, alg, is_training):
    from common.common_nn_ops import ModelInputParams
    G.TRACE.clear()
    G._ids[0] = 0
    x = G.FakeTensor([-1, patch, patch, channels], "input")
    out = model_cls().create_tensor_graph(ModelInputParams(x=x, y=None, device_id="/cpu:0", is_training=is_training),
                                          classes, alg)
    return {"patch": patch, "channels": channels, "classes": classes, "is_training": is_training, "alg": alg,
            "input_id": x.id, "y_conv": out.y_conv.id, "image_output": out.image_output, "trace": list(G.TRACE)}


def main():
    G.install_stubs()
    tf = sys.modules["tensorflow"]
    tf.split = _split
    tf.nn.local_response_normalization = _lrn
    G.FakeTensor.__getitem__ = _getitem
    from nnmodel.CONCNNModel import CONCNNModel
    from nnmodel.DUALCNNModel import DUALCNNModel
    dual_alg = json.load(open(os.path.join(G.REF, "alg_param_dualcnn.json"))) if os.path.exists(
        os.path.join(G.REF, "alg_param_dualcnn.json")) else None
    if dual_alg is None:
        dual_alg = {"batch_size": 48, "drop_out_ratio": 0.70, "lrelu_alpha": 0.18, "filter_count": 480, "hs_lidar_diff": 1}
    cases = {"dualcnn_graph_trace.json": [trace(DUALCNNModel, 7, 145, 15, dual_alg, True),
                                          trace(DUALCNNModel, 5, 51, 20, {**dual_alg, "filter_count": 64}, False)],
             "concnn_graph_trace.json": [trace(CONCNNModel, 5, 145, 15, {"drop_out_ratio": 0.5, "filter_count": 128}, True),
                                         trace(CONCNNModel, 3, 10, 4, {"drop_out_ratio": 0.5, "filter_count": 8}, False)]}
    for name, traces in cases.items():
        with open(os.path.join(HERE, name), "w") as f:
            json.dump(traces, f, indent=None, separators=(",", ":"), default=str)
        print(name, [len(t["trace"]) for t in traces])


if __name__ == "__main__":
    main()
