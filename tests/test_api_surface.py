"""The caller-side surface of common/common_nn_ops.py: every public function / class of the reference's module
(tests/golden/api_surface.json, recorded by introspecting the reference with make_golden_api.py) is either mirrored
here with the same argument list (extra trailing OPTIONAL arguments allowed) or listed below with the reason it is not."""
import inspect
import json
import os

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "api_surface.json")))

NOT_MIRRORED = {
    # per-sample TF augmentation ops: one kernel launch per batch instead (hyp_augment_patches, AugmentingIterator)
    "add_augmentation_graph": "AugmentingIterator", "perform_rotation_augmentation_random": "hyp_augment_patches",
    "perform_reflection_augmentation_random": "hyp_augment_patches",
    "perform_spectral_augmentation_random": "hyp_augment_patches",
    "perform_shadow_augmentation_random": "AugmentingIterator (shadow_struct ops on the batch)",
    # numba per-pixel gather closures: the gather kernel (hyp_gather_patches) behind DataSet.get_data_points
    "get_data_point_func": "hyp_gather_patches", "get_data_point_func_hsi": "hyp_gather_patches",
    # channel-resampling index arithmetic: lives in the layer plan (hyp_engine.cu), pinned by tests/golden/scale_in_to_out.json
    "scale_in_to_out": "hyp_engine.cu residual tables",
    "objective": "optuna hyper-parameter search (out of scope)", "set_all_gpu_config": "TensorFlow memory-growth switch",
}
NOT_MIRRORED_CLASSES = {"TextSummaryAtStartHook": "hypelcnn_b200.classify.summaries.ClassificationSummaryWriter.add_text"}
# constructor of an internal holder that create_metric_tensors builds (TF ops there, device buffers here)
DIFFERENT_CONSTRUCTOR = {"MetricOpsHolder"}


def _params(fn):
    fn = fn.__func__ if isinstance(fn, (staticmethod, classmethod)) else fn
    return list(inspect.signature(fn).parameters.values())


def _compatible(mine, reference_names):
    names = [p.name for p in mine]
    if names[:len(reference_names)] != reference_names:
        return False
    return all(p.default is not inspect.Parameter.empty for p in mine[len(reference_names):])


def test_functions_are_mirrored_or_accounted_for():
    from hypelcnn_b200.common import common_nn_ops as ops
    missing = []
    for name, ref_params in GOLD["functions"].items():
        if hasattr(ops, name):
            assert _compatible(_params(getattr(ops, name)), ref_params), name
        elif name not in NOT_MIRRORED:
            missing.append(name)
    assert not missing, missing
    assert not [n for n in NOT_MIRRORED if hasattr(ops, n)]          # the list above stays honest


def test_classes_are_mirrored_or_accounted_for():
    from hypelcnn_b200.common import common_nn_ops as ops
    for name, entry in GOLD["classes"].items():
        if name in NOT_MIRRORED_CLASSES:
            continue
        cls = getattr(ops, name)
        for method, ref_params in entry["methods"].items():
            if method == "__init__" and name in DIFFERENT_CONSTRUCTOR:
                continue
            assert _compatible(_params(inspect.getattr_static(cls, method)), ref_params), (name, method)
    m = ops.MetricOpsHolder.__init__.__code__.co_names + tuple(vars(ops.MetricOpsHolder))
    for attribute in ("accuracy", "mean_per_class_accuracy", "kappa", "confusion"):   # what the callers read
        assert attribute in m or hasattr(ops.MetricOpsHolder, attribute), attribute
