"""The sample-list / shadow-statistics helpers against golden vectors produced by running the reference's own
functions (tests/golden/make_golden_samples.py)."""
import os
import types

import numpy

G = numpy.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sample_ops_golden.npz"))


def same(a, b):
    a, b = numpy.asarray(a), numpy.asarray(b)
    return a.shape == b.shape and numpy.array_equal(a, b)


def test_targets_and_rasters_match_the_reference():
    from hypelcnn_b200.common import common_nn_ops as ops
    rows = ops.read_targets_from_image(G["labels"], range(0, 6))
    assert same(rows, G["rows"]) and rows.dtype == G["rows"].dtype
    sample_set = types.SimpleNamespace(training_targets=G["remaining_02"][:40], test_targets=G["test_02"],
                                       validation_targets=numpy.vstack([G["remaining_02"][40:], [[3, 2, 4], [3, 2, 1]]]))
    image = ops.create_target_image_via_samples(sample_set, [23, 31])
    assert same(image, G["target_image"]) and image.dtype == numpy.uint8 and image[2, 3] == 1     # last listing wins
    colored = ops.create_colored_image(image, G["colors"])
    assert same(colored, G["colored"]) and colored.dtype == numpy.uint8
    empty = types.SimpleNamespace(training_targets=numpy.empty((0, 3)), test_targets=numpy.empty((0, 3)),
                                  validation_targets=numpy.empty((0, 3)))
    assert (ops.create_target_image_via_samples(empty, [4, 5]) == ops.INVALID_TARGET_VALUE).all()


def test_splits_match_the_reference():
    from hypelcnn_b200.common import common_nn_ops as ops
    rows = G["rows"]
    test, remaining = ops.shuffle_test_data_using_ratio(rows, 0.2)                 # random_state=0: reproducible
    assert same(test, G["test_02"]) and same(remaining, G["remaining_02"])
    t0, r0 = ops.shuffle_test_data_using_ratio(rows, 0.0)
    assert same(t0, G["test_0"]) and same(r0, G["remaining_0"]) and t0.shape == (0, 3)
    for name, (size, vsize) in {"a": (12, None), "b": (40, 5), "c": (3, 1000)}.items():
        numpy.random.seed(99)
        tr, va = ops.shuffle_training_data_using_size(range(0, 6), rows, size, vsize)
        assert same(tr, G[f"size_train_{name}"]) and same(va, G[f"size_val_{name}"]), name
    numpy.random.seed(5)
    tr, va = ops.shuffle_training_data_using_ratio(rows, 0.3)
    assert same(tr, G["ratio_train"]) and same(va, G["ratio_val"])
    assert len(tr) + len(va) == len(rows)


def test_shadow_ratio_matches_the_reference():
    from hypelcnn_b200.common import common_nn_ops as ops
    ratio = ops.calculate_shadow_ratio(G["casi"], G["shadow"], numpy.logical_not(G["shadow"]).astype(int))
    assert ratio.dtype == numpy.float32 and ratio.shape == (9,)
    assert numpy.allclose(ratio, G["shadow_ratio"], rtol=2e-7, atol=0)             # masked-array mean vs plain mean: 1 ulp
