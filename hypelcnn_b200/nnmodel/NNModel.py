"""The model plug-in contract (reference interface: nnmodel/NNModel.py:4-12; same two hooks, same argument meaning).

``create_tensor_graph`` receives the reference's ModelInputParams (x, y, device_id, is_training) and returns the
reference's ModelOutputTensors (y_conv = logits, image_output = reconstruction or None, image_original = x,
histogram_tensors); here x / y are CUDA tensors and the call runs the engine's forward pass eagerly.
``get_loss_func`` returns the per-sample loss the caller averages (cross entropy + reconstruction MSE for HYPELCNN)."""
import abc


class NNModel(abc.ABC):

    @abc.abstractmethod
    def create_tensor_graph(self, model_input_params, class_count, algorithm_params):
        raise NotImplementedError

    @abc.abstractmethod
    def get_loss_func(self, tensor_output, label):
        raise NotImplementedError
