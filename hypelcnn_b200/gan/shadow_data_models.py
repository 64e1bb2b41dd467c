"""Shadow GAN generator on the device (reference: gan/shadow_data_models.py:43-90), inference only.

The reference builds the generator out of seven slim ``convolution1d`` layers with ONE filter each
(``net1`` .. ``net7``; kernel sizes C, C/2, C/4, C/8, C/4, C/2, C; leaky_relu 0.1; dense residuals; tanh on the last).
Here the whole stack is one kernel of libhypelcnn_b200.so (``hyp_gan_generator_forward``); ``GeneratorVariables`` holds
the ≤ 239 scalars under the reference's variable names (``net1/weights`` [K,1,1], ``net1/biases`` [1], ...)."""
import ctypes

import numpy
import torch

from hypelcnn_b200 import _native as N


def generator_kernel_sizes(band_size, create_only_encoder=False):
    k = band_size
    sizes = [k, k // 2, k // 4, k // 8]
    return sizes if create_only_encoder else sizes + [k // 4, k // 2, k]


class GeneratorVariables:
    """The generator's variables as one flat device buffer in kernel order: net1 w[K1], b, net2 w[K2], b, ..."""

    def __init__(self, band_size, create_only_encoder=False, device=None):
        self.band_size, self.create_only_encoder = int(band_size), bool(create_only_encoder)
        self.sizes = generator_kernel_sizes(self.band_size, self.create_only_encoder)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        # zeros like the reference's weights_initializer (shadow_data_models.py:47): an untrained generator is
        # net7 = tanh(0) = 0 / the encoder passes residual sums through
        self.flat = torch.zeros(sum(k + 1 for k in self.sizes), dtype=torch.float32, device=self.device)

    def names(self):
        out, off = [], 0
        for i, k in enumerate(self.sizes):
            out.append((f"net{i + 1}/weights", off, (k, 1, 1)))
            out.append((f"net{i + 1}/biases", off + k, (1,)))
            off += k + 1
        return out

    def load(self, values):
        """values: {"net1/weights": [K,1,1], "net1/biases": [1], ...} (checkpoint names below the Generator scope)."""
        for name, off, shape in self.names():
            v = torch.as_tensor(numpy.asarray(values[name], dtype=numpy.float32)).reshape(-1)
            if v.numel() != int(numpy.prod(shape)):
                raise ValueError(f"{name}: expected shape {shape}")
            self.flat[off:off + v.numel()].copy_(v)

    def export(self):
        return {name: self.flat[off:off + int(numpy.prod(shape))].cpu().numpy().reshape(shape)
                for name, off, shape in self.names()}


def shadowdata_generator_model(netinput, create_only_encoder, is_training, variables):
    """netinput [B,1,1,C] (CUDA) -> [B,1,1,C]; same argument meaning as the reference plus the variable holder
    (the reference takes them from the enclosing variable scope).  is_training has no effect on the arithmetic
    (it only marks variables trainable, shadow_data_models.py:53)."""
    if netinput.dim() != 4 or netinput.shape[1] != 1 or netinput.shape[2] != 1:
        raise ValueError("generator input must be [B,1,1,C]")
    B, C = netinput.shape[0], netinput.shape[3]
    x = netinput.reshape(B, C).contiguous()
    out = torch.empty_like(x)
    _generator_rows(x, out, C, 0, variables, False, True)
    return out.reshape(B, 1, 1, C)


def _generator_rows(x2d, out2d, bands, copy_extra, variables, clip_invalid_values, is_shadow_graph):
    if not x2d.is_cuda or x2d.dtype != torch.float32:
        raise TypeError("generator input must be a CUDA float32 tensor (no CPU path)")
    if variables.band_size != bands:
        raise ValueError(f"generator was built for {variables.band_size} bands, got {bands}")
    N.check(N.lib().hyp_gan_generator_forward(
        ctypes.c_void_p(x2d.data_ptr()), x2d.stride(0), ctypes.c_void_p(out2d.data_ptr()), out2d.stride(0),
        x2d.shape[0], bands, copy_extra, ctypes.c_void_p(variables.flat.data_ptr()),
        int(variables.create_only_encoder), int(bool(clip_invalid_values)), int(bool(is_shadow_graph)),
        ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out2d
