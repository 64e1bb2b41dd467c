"""hypelcnn_b200 — B200-native HSI+LiDAR patch engine behind the plug-in surface of
aligokalppeker/hypelcnn (NNModel / DataImporter / DataLoader / DataSet).

Layout:
  csrc/      CUDA kernels + the C ABI (include/hypelcnn_b200.h)
  _native.py ctypes binding of the C ABI (fails loudly when the library or a GPU is missing)
  engine.py  PatchEngine: owns the flat torch buffers a hyp_model is bound to
  common/ nnmodel/ importer/ loader/   host-side mirror of the reference's plug-in API
"""
from hypelcnn_b200._native import NativeError, build_native, lib  # noqa: F401

__all__ = ["NativeError", "build_native", "lib"]
