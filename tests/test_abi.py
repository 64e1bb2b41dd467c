"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/hypelcnn_b200.h declares (no compute calls — there is no GPU here), argument
validation that needs no device, and the no-CPU-fallback rule."""
import ctypes
import os
import re

import pytest

from hypelcnn_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    N.build_native()
    return N.lib()


def declared_functions():
    src = open(os.path.join(ROOT, "include", "hypelcnn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hyp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    names = declared_functions()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/hypelcnn_b200.h but not exported"
    assert sorted(N.EXPORTS) == names, "ctypes prototypes out of sync with the header"
    assert lib.hyp_version() == 1


def test_desc_struct_layout_matches_header():
    assert ctypes.sizeof(N.ModelDesc) == 12 * 4 + 4 * 4


def test_invalid_arguments_are_rejected_without_a_device(lib):
    assert lib.hyp_model_create(None, None) == N.HYP_E_INVALID
    assert b"null" in lib.hyp_last_error()
    d = N.ModelDesc(kind=7, patch=7, channels=145, classes=15, filter_count=480, spectral_levels=3, spatial_levels=3,
                    degradation=3, use_residual=1, precision_mode=0, max_batch=8)
    h = ctypes.c_void_p()
    assert lib.hyp_model_create(ctypes.byref(d), ctypes.byref(h)) == N.HYP_E_INVALID
    d.kind, d.patch = 0, 4
    assert lib.hyp_model_create(ctypes.byref(d), ctypes.byref(h)) == N.HYP_E_INVALID  # even patch
    assert lib.hyp_adam_step(None, None, None, None, 4, 0.1, 0.9, 0.999, 1e-8, 1, 1.0, None) == N.HYP_E_INVALID
    assert lib.hyp_argmax_confusion(None, None, 4, 15, None, None, None) == N.HYP_E_INVALID
    assert lib.hyp_gather_patches(None, 0, 1, 1, 1, None, None, None, 0, 0, None, 0, 0, None, 0, None, 1,
                                  None) == N.HYP_E_INVALID


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    d = N.ModelDesc(kind=0, patch=7, channels=145, classes=15, filter_count=480, spectral_levels=3, spatial_levels=3,
                    degradation=3, use_residual=1, precision_mode=0, max_batch=8, lrelu_alpha=0.18, bn_decay=0.95,
                    bn_eps=1e-3, drop_out_ratio=0.7)
    h = ctypes.c_void_p()
    assert lib.hyp_model_create(ctypes.byref(d), ctypes.byref(h)) == N.HYP_E_CUDA
    assert b"no CPU fallback" in lib.hyp_last_error()
    from hypelcnn_b200.engine import PatchEngine
    with pytest.raises(N.NativeError):
        PatchEngine(7, 145, 15, {"filter_count": 480}, 8)


def test_product_code_never_imports_the_oracle():
    for base, _, files in os.walk(os.path.join(ROOT, "hypelcnn_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, os.path.join(base, f)


def test_feature_discriminator_layout_matches_the_oracle_and_slim_names(lib):
    from hypelcnn_b200.gan.wrappers.cut_wrapper import feature_discriminator_variable_table
    from oracle import gan_ref
    for bands, patches, E in ((64, 6, 2), (144, 6, 2), (48, 4, 5), (32, 8, 1)):
        layout, n_full = gan_ref.feature_discriminator_layout(bands, patches, E)
        assert lib.hyp_gan_feature_discriminator_weight_count(bands, patches, E) == len(layout) * n_full
        table = feature_discriminator_variable_table(bands, patches, E)
        flat = [(wo, shape) for _, _, layers in layout for wo, shape, _ in layers]
        assert [(off, shape) for name, off, shape in table if name.endswith("weights")] == flat
    names = [n for n, _, _ in feature_discriminator_variable_table(64, 6, 2)]
    assert names[0] == "fully_connected/weights" and names[2] == "fully_connected_1/weights"
    assert names[-1] == "fully_connected_27/biases"       # 7 slices (64 bands, slices of 10: the last is 4 wide) x 4 layers
