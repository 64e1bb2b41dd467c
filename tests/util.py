import json
import os

import numpy
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

ALG = {"batch_size": 48, "drop_out_ratio": 0.70, "filter_count": 480, "learning_rate": 0.0003,
       "learning_rate_decay_factor": 0.96, "learning_rate_decay_step": 350, "lrelu_alpha": 0.18,
       "optimizer": "AdamOptimizer", "bn_decay": 0.95, "l2regularizer_scale": 0.00001,
       "spectral_hierarchy_level": 3, "spatial_hierarchy_level": 3, "degradation_coeff": 3, "use_residual": True}

# parity tolerance stated by BASELINE.json north_star: logits within fp32 rtol 1e-4 (atol for values near 0)
RTOL, ATOL = 1e-4, 1e-5


def synthetic_batch(B, P, C, classes, seed=1234):
    """S-C2 style synthetic patches: U[0,1) fp32, uniform labels."""
    rng = numpy.random.default_rng(seed)
    x = rng.random((B, P, P, C), dtype=numpy.float32)
    y = rng.integers(0, classes, B).astype(numpy.uint8)
    return x, y


def oracle_variables(engine, dtype=torch.float64):
    return {k: torch.tensor(v, dtype=dtype) for k, v in engine.export_variables().items()}


def assert_close(got, ref, rtol=RTOL, atol=ATOL, what=""):
    got = numpy.asarray(got, dtype=numpy.float64)
    ref = numpy.asarray(ref, dtype=numpy.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    err = numpy.abs(got - ref)
    bound = atol + rtol * numpy.abs(ref)
    bad = err > bound
    assert not bad.any(), (f"{what}: {int(bad.sum())}/{bad.size} outside rtol={rtol} atol={atol}; "
                           f"max err {err.max():.3e} at ref={ref.flat[int(err.argmax())]:.6e}")


def assert_close_scaled(got, ref, rel=RTOL, what=""):
    """|got-ref| <= rel * max|ref| — for gradient tensors whose entries span many magnitudes."""
    got = numpy.asarray(got, dtype=numpy.float64)
    ref = numpy.asarray(ref, dtype=numpy.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    scale = max(float(numpy.abs(ref).max()), 1e-30)
    err = float(numpy.abs(got - ref).max())
    assert err <= rel * scale, f"{what}: max err {err:.3e} vs scale {scale:.3e} (rel {err / scale:.3e} > {rel})"


def assert_grad_close(got, ref64, ref32, what="", floor=2e-4, factor=4.0):
    """Gradients pass through ~25 BatchNorm backward steps whose cancellation amplifies fp32
    rounding, most at the tiny batches the tests use.  Criterion: the CUDA result is as close
    to the fp64 oracle as the fp32 CPU oracle is (within `factor`), never worse than `floor`
    relative to the tensor's largest entry when fp32 itself is that accurate."""
    got = numpy.asarray(got, dtype=numpy.float64)
    ref64 = numpy.asarray(ref64, dtype=numpy.float64)
    ref32 = numpy.asarray(ref32, dtype=numpy.float64)
    assert got.shape == ref64.shape, (what, got.shape, ref64.shape)
    scale = max(float(numpy.abs(ref64).max()), 1e-30)
    err = float(numpy.abs(got - ref64).max()) / scale
    base = float(numpy.abs(ref32 - ref64).max()) / scale
    assert err <= max(floor, factor * base), f"{what}: rel err {err:.3e} (fp32 oracle itself: {base:.3e})"
