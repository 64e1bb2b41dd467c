#!/usr/bin/env python
"""Golden vectors for the host-side sample-list helpers, produced by EXECUTING the reference's own functions
(common/common_nn_ops.py:455-543, imported with the TensorFlow stubs of make_golden.py; scikit-learn and numpy are
real).  Build container only; ``sample_ops_golden.npz`` is committed and read by tests/test_sample_ops.py.

usage: python tests/golden/make_golden_samples.py"""
import os
import sys
import types

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402


def main():
    G.install_stubs()
    sys.path.insert(0, G.REF)
    if not hasattr(numpy, "int"):
        numpy.int = int
    import common.common_nn_ops as ops
    rng = numpy.random.default_rng(2024)
    out = {}
    labels = rng.integers(0, 7, (23, 31)).astype(numpy.uint8)
    labels[rng.random(labels.shape) < 0.6] = 255                      # unlabelled pixels
    rows = ops.read_targets_from_image(labels, range(0, 6))           # class 6 deliberately outside the range
    out["labels"], out["rows"] = labels, rows
    test, remaining = ops.shuffle_test_data_using_ratio(rows, 0.2)
    out["test_02"], out["remaining_02"] = test, remaining
    t0, r0 = ops.shuffle_test_data_using_ratio(rows, 0.0)
    out["test_0"], out["remaining_0"] = t0, r0
    for name, (size, vsize) in {"a": (12, None), "b": (40, 5), "c": (3, 1000)}.items():
        numpy.random.seed(99)
        tr, va = ops.shuffle_training_data_using_size(range(0, 6), rows, size, vsize)
        out[f"size_train_{name}"], out[f"size_val_{name}"] = tr, va
    numpy.random.seed(5)
    tr, va = ops.shuffle_training_data_using_ratio(rows, 0.3)         # unseeded split object -> global numpy RNG
    out["ratio_train"], out["ratio_val"] = tr, va
    sample_set = types.SimpleNamespace(training_targets=remaining[:40], test_targets=test,
                                       validation_targets=numpy.vstack([remaining[40:], [[3, 2, 4], [3, 2, 1]]]))
    out["target_image"] = ops.create_target_image_via_samples(sample_set, [23, 31])
    colors = rng.integers(0, 256, (5, 3)).astype(numpy.uint8)
    out["colors"], out["colored"] = colors, ops.create_colored_image(out["target_image"], colors)
    casi = rng.integers(1, 4000, (23, 31, 9)).astype(numpy.uint16)
    shadow = (rng.random((23, 31)) < 0.3).astype(numpy.uint8)
    out["casi"], out["shadow"] = casi, shadow
    out["shadow_ratio"] = ops.calculate_shadow_ratio(casi, shadow, numpy.logical_not(shadow).astype(int))
    numpy.savez_compressed(os.path.join(HERE, "sample_ops_golden.npz"), **out)
    print("wrote sample_ops_golden.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
