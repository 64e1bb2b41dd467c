"""Shadow augmenter structs handed to the classifier's input pipeline (reference: gan/gan_utilities.py:7-43).
Same names; the ops take and return device batches [B,P,P,C+1] (HSI bands + LiDAR) instead of one tf sample."""
import numpy
import torch

from hypelcnn_b200.gan.wrappers.gan_common import create_inference_for_matrix_input


class ShadowOpHolder:
    """What the classifier's input pipeline receives from a shadow augmenter (reference: gan/gan_utilities.py:7-15,
    same four fields): the two batch transforms, a factory for the object that restores the generator weights, and
    the function that runs the restore."""

    __slots__ = ("shadow_op", "deshadow_op", "shadow_op_creater", "shadow_op_initializer")

    def __init__(self, shadow_op, deshadow_op, shadow_op_creater, shadow_op_initializer) -> None:
        self.shadow_op, self.deshadow_op = shadow_op, deshadow_op
        self.shadow_op_creater, self.shadow_op_initializer = shadow_op_creater, shadow_op_initializer


def create_simple_shadow_struct(shadow_ratio):
    """The ratio augmenter (:18-28): shadowing divides every band by its shadow ratio, de-shadowing multiplies; the
    trailing LiDAR channel passes through (ratio 1).  The ratio vector is uploaded once per device."""
    host_ratio = numpy.concatenate([numpy.asarray(shadow_ratio, dtype=numpy.float32), numpy.ones(1, numpy.float32)])
    on_device = {}

    def ratio_for(batch):
        if batch.device not in on_device:
            on_device[batch.device] = torch.as_tensor(host_ratio, device=batch.device)
        return on_device[batch.device]

    return ShadowOpHolder(shadow_op=lambda batch: batch / ratio_for(batch),
                          deshadow_op=lambda batch: batch * ratio_for(batch),
                          shadow_op_creater=lambda: None,
                          shadow_op_initializer=lambda restorer, session: None)


class GeneratorInferenceWrapper:
    """The InferenceWrapper slice create_gan_struct needs (gan/wrappers/wrapper.py:21-38): a forward (x -> y,
    "shadow") and a backward generator with frozen weights, construct_inference_graph over a [B,H,W,C] matrix."""

    def __init__(self, forward_generator, backward_generator):
        self.forward_generator, self.backward_generator = forward_generator, backward_generator

    def construct_inference_graph(self, input_tensor, is_shadow_graph, clip_invalid_values, copy_extra=0):
        gen = self.forward_generator if is_shadow_graph else self.backward_generator
        return create_inference_for_matrix_input(input_tensor, is_shadow_graph, clip_invalid_values, gen, copy_extra)

    def create_generator_restorer(self):
        return self  # restorer.restore(values) loads {"ModelX2Y/.../net1/weights": ...}-style dicts

    def restore(self, forward_values=None, backward_values=None):
        if forward_values is not None:
            self.forward_generator.load(forward_values)
        if backward_values is not None:
            self.backward_generator.load(backward_values)


def read_generator_checkpoint(path):
    """A generator checkpoint written by gan_train_for_shadow.run_session (``model.ckpt-<step>.npz``, variables named
    ``ModelX2Y/Generator/net1/weights`` ... like the reference's TF checkpoints) -> (forward values, backward values)
    keyed below the Generator scope.  A single-generator checkpoint (``Model/Generator/...``) fills the forward slot."""
    import os
    if not os.path.exists(path) and os.path.exists(path + ".npz"):
        path = path + ".npz"
    if not os.path.exists(path):
        raise IOError(f"generator checkpoint {path} not found")
    stored = numpy.load(path)
    scopes = {"ModelX2Y/Generator/": {}, "ModelY2X/Generator/": {}, "Model/Generator/": {}}
    for name in stored.files:
        for scope, values in scopes.items():
            if name.startswith(scope):
                values[name[len(scope):]] = stored[name]
    forward = scopes["ModelX2Y/Generator/"] or scopes["Model/Generator/"]
    return (forward or None), (scopes["ModelY2X/Generator/"] or None)


def create_gan_struct(gan_inference_wrapper, model_base_dir=None, ckpt_relative_path=None):
    """Reference: gan/gan_utilities.py:30-43.  The LiDAR channel is stripped, every pixel of the patch goes through
    the (frozen) generator, the LiDAR channel is appended again — done inside one kernel launch.
    ``shadow_op_initializer(restorer, session)`` restores the generators from ``model_base_dir + ckpt_relative_path``
    like the reference; without a path, ``session`` may carry the (forward, backward) value dicts directly."""

    def _build_shadowed_inference_graph(input_data, is_shadow_graph):
        return gan_inference_wrapper.construct_inference_graph(input_data, is_shadow_graph=is_shadow_graph,
                                                               clip_invalid_values=False, copy_extra=1)

    def _initializer(restorer, session):
        if model_base_dir is not None and ckpt_relative_path is not None:
            restorer.restore(*read_generator_checkpoint(model_base_dir + ckpt_relative_path))
        elif session is not None:
            restorer.restore(*session)

    return ShadowOpHolder(shadow_op=lambda x: _build_shadowed_inference_graph(x, True),
                          deshadow_op=lambda x: _build_shadowed_inference_graph(x, False),
                          shadow_op_creater=gan_inference_wrapper.create_generator_restorer,
                          shadow_op_initializer=_initializer)
