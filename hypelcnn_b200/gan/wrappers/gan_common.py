"""Inference helpers of gan/wrappers/gan_common.py on the device."""
import torch

from hypelcnn_b200.gan.shadow_data_models import _generator_rows

model_generator_name = "Generator"
model_base_name = "Model"


def adj_shadow_ratio(shadow_ratio, is_shadow):
    return 1. / shadow_ratio if is_shadow else shadow_ratio


def create_inference_for_matrix_input(input_tensor, is_shadow_graph, clip_invalid_values, generator_variables,
                                      copy_extra=0):
    """Reference: gan/wrappers/gan_common.py:282-304 — the generator applied to every pixel of [B,H,W,C(+extra)]
    separately (the reference builds H*W sub-graphs sharing weights; here every pixel is one row of ONE launch).
    With clip_invalid_values a pixel keeps its input spectrum unless the generated mean moved the expected way.
    copy_extra trailing channels (the LiDAR band, gan/gan_utilities.py:31-35) pass through unchanged."""
    if input_tensor.dim() != 4:
        raise ValueError("input must be [B,H,W,C]")
    x = input_tensor.contiguous()
    C = x.shape[3] - copy_extra
    rows = x.reshape(-1, x.shape[3])
    out = torch.empty_like(rows)
    _generator_rows(rows, out, C, copy_extra, generator_variables, clip_invalid_values, is_shadow_graph)
    return out.reshape(x.shape)
