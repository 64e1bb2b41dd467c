"""CONCNNModel plug-in backed by the native engine (reference: nnmodel/CONCNNModel.py:23-64).

Same class name, no-arg constructor and method signatures as the reference; ``algorithm_params`` is the reference's
JSON dict (modelconfigs/alg_param_concnn.json: filter_count, drop_out_ratio, ["MomentumOptimizer", 0.9], ...)."""
from hypelcnn_b200.common.common_nn_ops import ModelOutputTensors, labels_to_ids
from hypelcnn_b200.engine import PatchEngine
from hypelcnn_b200.nnmodel.NNModel import NNModel


class CONCNNModel(NNModel):
    precision = "3xtf32"

    def __init__(self):
        self.engine = None
        self.seed = 1234  # the reference's graph seed (classify/monitored_session_runner.py:13)

    def engine_for(self, x, algorithm_params):
        P, C = x.shape[1], x.shape[3]
        if self.engine is None:
            if getattr(self, "_class_count", None) is None:
                raise RuntimeError("call create_tensor_graph (class_count) before training")
            self.engine = PatchEngine(P, C, self._class_count, algorithm_params,
                                      max_batch=max(int(algorithm_params.get("batch_size", 1)), x.shape[0]),
                                      device=x.device, precision=self.precision, model="concnn")
            self.engine.init_variables(self.seed)
        elif (self.engine.patch, self.engine.channels) != (P, C):
            raise ValueError("this model instance was built for a different patch shape")
        return self.engine

    def create_tensor_graph(self, model_input_params, class_count, algorithm_params):
        x = model_input_params.x
        self._class_count = class_count
        eng = self.engine_for(x, algorithm_params)
        logits, _ = eng.forward(x.contiguous(), bool(model_input_params.is_training), update_moving=False,
                                seed=eng.global_step)
        return ModelOutputTensors(y_conv=logits, image_output=None, image_original=None, histogram_tensors=[])

    def get_loss_func(self, tensor_output, label):
        """Per-sample softmax cross-entropy [B] (CONCNNModel.py:66-68)."""
        return self.engine.per_sample_loss(tensor_output.y_conv, None, None, labels_to_ids(label))
