"""ctypes binding of include/hypelcnn_b200.h.  No fallbacks: if the shared library is not
built or cannot be loaded this raises; compute entry points fail with NativeError when no
CUDA device is present."""
import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "lib", "libhypelcnn_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC"]

HYP_OK, HYP_E_INVALID, HYP_E_CUDA, HYP_E_STATE, HYP_E_UNSUPPORTED = 0, -1, -2, -3, -4
HYP_DT_F32, HYP_DT_U16 = 0, 1
HYP_GATHER_SAME_RES, HYP_GATHER_GRSS2018 = 0, 1
HYP_PRECISION_FP32, HYP_PRECISION_3XTF32, HYP_PRECISION_BF16, HYP_PRECISION_3XF16 = 0, 1, 2, 3
HYP_MODEL_HYPELCNN, HYP_MODEL_DUALCNN, HYP_MODEL_CONCNN = 0, 1, 2


class NativeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"hypelcnn_b200 native error {code}: {msg}")
        self.code = code


class ModelDesc(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("patch", ctypes.c_int32), ("channels", ctypes.c_int32),
                ("classes", ctypes.c_int32), ("filter_count", ctypes.c_int32), ("spectral_levels", ctypes.c_int32),
                ("spatial_levels", ctypes.c_int32), ("degradation", ctypes.c_int32), ("use_residual", ctypes.c_int32),
                ("precision_mode", ctypes.c_int32), ("max_batch", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("lrelu_alpha", ctypes.c_float), ("bn_decay", ctypes.c_float), ("bn_eps", ctypes.c_float),
                ("drop_out_ratio", ctypes.c_float)]


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))] + \
           [os.path.join(os.path.dirname(_HERE), "include", "hypelcnn_b200.h")]


def build_native(force=False, verbose=False):
    """Compile csrc/*.cu for sm_100a (one object per translation unit, in parallel; nvcc cross-compiles without a
    GPU) and link lib/libhypelcnn_b200.so."""
    from concurrent.futures import ThreadPoolExecutor
    srcs = sources()
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    obj_dir = os.path.join(os.path.dirname(LIB_PATH), "obj")
    os.makedirs(obj_dir, exist_ok=True)
    headers = [s for s in srcs if not s.endswith(".cu")]
    units = [s for s in srcs if s.endswith(".cu")]
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"]

    def compile_unit(unit):
        obj = os.path.join(obj_dir, os.path.basename(unit)[:-3] + ".o")
        fresh = os.path.exists(obj) and all(os.path.getmtime(obj) >= os.path.getmtime(s) for s in [unit] + headers)
        if force or not fresh:
            cmd = [nvcc] + compile_flags + ["-c", "-o", obj, unit]
            if verbose:
                print(" ".join(cmd))
            subprocess.run(cmd, check=True)
        return obj

    with ThreadPoolExecutor(max_workers=len(units)) as pool:
        objects = list(pool.map(compile_unit, units))
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB_PATH] + objects
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB_PATH


_lib = None
_lock = threading.Lock()
_P, _I, _L, _F = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
_PROTOS = {
    "hyp_version": (ctypes.c_int, []),
    "hyp_last_error": (ctypes.c_char_p, []),
    "hyp_launch_count": (ctypes.c_int64, [_I]),
    "hyp_profile_enable": (_I, [_I]),
    "hyp_profile_get": (_I, [_I, ctypes.c_char_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_L),
                            ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]),
    "hyp_scene_minmax": (_I, [_P, _I, _I, _I, _I, _P, _P, _P]),
    "hyp_gather_patches": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _I, _I, _P, _I, _I, _P, _L, _P, _I, _P]),
    "hyp_model_create": (_I, [ctypes.POINTER(ModelDesc), ctypes.POINTER(_P)]),
    "hyp_model_destroy": (None, [_P]),
    "hyp_model_sizes": (_I, [_P, ctypes.POINTER(_L), ctypes.POINTER(_L), ctypes.POINTER(_L),
                             ctypes.POINTER(ctypes.c_int32)]),
    "hyp_model_variable": (_I, [_P, _I, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(_L),
                                ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)]),
    "hyp_model_bind": (_I, [_P, _P, _P, _P, _P, ctypes.c_size_t]),
    "hyp_model_forward": (_I, [_P, _P, _L, _I, _I, ctypes.c_uint64, _P, _P, _P]),
    "hyp_model_loss": (_I, [_P, _P, _P, _P, _P, _L, _P, _P]),
    "hyp_model_loss_backward": (_I, [_P, _P, _P, _L, _P, _P]),
    "hyp_model_set_grad_notify": (_I, [_P, _L, _P]),
    "hyp_adam_step": (_I, [_P, _P, _P, _P, _L, _F, _F, _F, _F, _L, _F, _P]),
    "hyp_momentum_step": (_I, [_P, _P, _P, _L, _F, _F, _F, _P]),
    "hyp_augment_patches": (_I, [_P, _P, _L, _I, _I, _I, _I, _F, ctypes.c_uint64, _P, _P, _P]),
    "hyp_gan_generator_forward": (_I, [_P, _I, _P, _I, _L, _I, _I, _P, _I, _I, _I, _P]),
    "hyp_gan_generator_train_forward": (_I, [_P, _L, _I, _P, _P, _P]),
    "hyp_gan_generator_backward": (_I, [_P, _P, _L, _I, _P, _P, _P, _P]),
    "hyp_gan_discriminator_forward": (_I, [_P, _L, _I, _P, _P, _P, _P]),
    "hyp_gan_discriminator_backward": (_I, [_P, _P, _P, _L, _I, _P, _P, _P, _P]),
    "hyp_gan_loss_grad": (_I, [_I, _P, _P, _F, _F, _L, _P, _I, _P, _P]),
    "hyp_gan_l2_regularizer": (_I, [_P, _P, _L, _F, _P, _P]),
    "hyp_gan_cycle_generator_step": (_I, [_P, _P, _L, _I, _P, _P, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P, _P, _P]),
    "hyp_gan_cycle_discriminator_step": (_I, [_P, _P, _L, _I, _P, _P, _P, _P, _F, _P, _P, _P, _P, _I, _I, _P, _I, _I, _P]),
    "hyp_gan_generator_backward_enc": (_I, [_P, _P, _P, _L, _I, _P, _P, _P, _P]),
    "hyp_gan_feature_discriminator_weight_count": (_L, [_I, _I, _I]),
    "hyp_gan_feature_discriminator_forward": (_I, [_P, _L, _I, _I, _I, _P, _P, _P, _P]),
    "hyp_gan_feature_discriminator_backward": (_I, [_P, _P, _P, _P, _P, _L, _I, _I, _I, _P, _P, _P, _P]),
    "hyp_gan_patchnce": (_I, [_P, _P, _P, _P, _L, _I, _I, _F, _F, _I, _P, _P, _P, _P, _P, _P]),
    "hyp_argmax_confusion": (_I, [_P, _P, _L, _I, _P, _P, _P]),
    "hyp_scatter_class_map": (_I, [_P, _P, _L, _I, _I, _P, _P]),
    "hyp_model_debug_tensor": (_I, [_P, ctypes.c_char_p, _I, ctypes.POINTER(_P), ctypes.POINTER(_L)]),
    "hyp_crc32c": (_I, [_P, ctypes.c_uint64, _P]),
    "hyp_tiff_lzw_decode": (_I, [_P, ctypes.c_uint64, _P, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64)]),
    "hyp_debug_schedule": (_I, [_P, _I, _I, _I, _P, _P]),
    "hyp_debug_level_tap_groups": (_I, [_I, _I, _I, _I, _I, _P, _I, ctypes.POINTER(_I)]),
    "hyp_debug_tc_gemm": (_I, [_I, _P, _P, _I, _I, _I, _P, _P, _I, _I, _I, _I, _P]),
    "hyp_model_dropout_mask": (_I, [_P, ctypes.c_char_p, ctypes.c_uint64, _L, _P, _P]),
}
EXPORTS = sorted(_PROTOS)


def lib():
    """The loaded shared library (loads on first use; raises if it has not been built)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise NativeError(HYP_E_STATE, f"{LIB_PATH} is not built — run `python -c 'import __graft_entry__ as g; "
                                               f"g.build()'` (there is no CPU fallback)")
            handle = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in _PROTOS.items():
                fn = getattr(handle, name)
                fn.restype, fn.argtypes = res, args
            _lib = handle
    return _lib


def check(code):
    if code != HYP_OK:
        raise NativeError(code, lib().hyp_last_error().decode("utf-8", "replace"))
    return code
