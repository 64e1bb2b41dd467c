"""HYPELCNNModel plug-in backed by the native engine.

Same class name, no-arg constructor and method signatures as the reference
(nnmodel/HYPELCNNModel.py:12,34,101), so ``get_model_from_name("HYPELCNNModel")`` and the
callers in common_nn_ops keep working.  Where the reference returns symbolic tf.Tensors
this returns CUDA torch.Tensors computed eagerly by libhypelcnn_b200.so.  One instance =
one set of weights: repeated create_tensor_graph calls (train / test / validation branches)
share them, like tf.make_template("nn_core", ...) does (common/common_nn_ops.py:333).
"""
import torch

from hypelcnn_b200.common.common_nn_ops import HistogramTensorPair, ModelOutputTensors, labels_to_ids
from hypelcnn_b200.engine import PatchEngine
from hypelcnn_b200.nnmodel.NNModel import NNModel


class HYPELCNNModel(NNModel):
    precision = "3xf16"  # tcgen05 tensor-core engine, fp16 hi/lo operand planes (fp32-accurate); "3xtf32": TF32 planes; "fp32": FFMA engine

    def __init__(self):
        self.engine = None
        self.seed = 1234  # the reference's graph seed (classify/monitored_session_runner.py:13)

    def engine_for(self, x, algorithm_params):
        P, C = x.shape[1], x.shape[3]
        if self.engine is None:
            self._class_count = getattr(self, "_class_count", None)
            if self._class_count is None:
                raise RuntimeError("call create_tensor_graph (class_count) before training")
            self.engine = PatchEngine(P, C, self._class_count, algorithm_params,
                                      max_batch=max(int(algorithm_params.get("batch_size", 1)), x.shape[0]),
                                      device=x.device, precision=self.precision)
            self.engine.init_variables(self.seed)
        elif (self.engine.patch, self.engine.channels) != (P, C):
            raise ValueError("this model instance was built for a different patch shape")
        return self.engine

    def create_tensor_graph(self, model_input_params, class_count, algorithm_params):
        x = model_input_params.x
        self._class_count = class_count
        eng = self.engine_for(x, algorithm_params)
        x = x.contiguous()
        logits, recon = eng.forward(x, bool(model_input_params.is_training), update_moving=False,
                                    seed=eng.global_step)
        hist = [HistogramTensorPair(eng.debug_tensor(n).view(x.shape[0], -1), label) for n, label in
                ((f"conv_enc_{eng.alg['spectral_hierarchy_level'] - 1}", "spectral_expansion"),
                 (f"conv_dec_{eng.alg['spectral_hierarchy_level'] - 1}", "spectral_reduction"),
                 (f"connector_conv_{eng.alg['spatial_hierarchy_level'] - 1}", "spatial"))] \
            if getattr(model_input_params, "want_histograms", False) else []
        return ModelOutputTensors(y_conv=logits, image_output=recon, image_original=x, histogram_tensors=hist)

    def get_loss_func(self, tensor_output, label):
        """Per-sample loss [B] (reference: nnmodel/HYPELCNNModel.py:101-112)."""
        return self.engine.per_sample_loss(tensor_output.y_conv, tensor_output.image_output,
                                           tensor_output.image_original if tensor_output.image_output is not None
                                           else None, labels_to_ids(label))
