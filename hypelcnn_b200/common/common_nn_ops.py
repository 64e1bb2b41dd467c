"""Host-side mirror of the reference's common/common_nn_ops.py for the hot path: the same
class / function names and argument meaning, with torch CUDA tensors where the reference
has tf.Tensors and eager execution where it builds a TF1 graph.  All arithmetic is done by
libhypelcnn_b200.so (hypelcnn_b200.engine); nothing here computes on the CPU.

reference                                   here
------------------------------------------  ---------------------------------------------
DataSet / BasicDataSet (:23-106)            same names; the scene lives UNPADDED in HBM,
                                            get_data_point(s) = hyp_gather_patches
ModelInputParams/ModelOutputTensors/...     same plain holders (:109-165)
training_nn_iterator / simple_nn_iterator   DeviceBatchIterator (shuffle+repeat / in-order)
optimize_nn (:208-240)                      optimize_nn -> TrainOp (forward, loss, backward, Adam)
create_metric_tensors/calculate_accuracy    MetricOpsHolder + calculate_accuracy (:243-310)
perform_prediction (:313-327)               perform_prediction (argmax + scatter on device)
create_graph (:330-373)                     create_graph
get_*_from_name (:443-452)                  same
"""
from abc import ABC, abstractmethod
from functools import partial

import numpy
import torch

from hypelcnn_b200 import _native as N
from hypelcnn_b200 import engine as E
from hypelcnn_b200.common.common_ops import get_class

INVALID_TARGET_VALUE = 255


class DataSet(ABC):
    @abstractmethod
    def get_data_shape(self):
        pass

    @abstractmethod
    def get_casi_band_count(self):
        pass

    @abstractmethod
    def get_scene_shape(self):
        pass

    @abstractmethod
    def get_unnormalized_casi_dtype(self):
        pass

    @abstractmethod
    def get_data_point(self, point_x, point_y):
        pass


class BasicDataSet(DataSet):
    """Reference: common/common_nn_ops.py:45-106.  casi [H,W,C] (uint16/float32) and lidar
    [H,W,1] float32 are numpy arrays or tensors; they are moved to the GPU once, kept
    unpadded and un-normalised; min/max come from hyp_scene_minmax and normalisation + the
    symmetric padding happen inside the gather kernel (bit-identical results)."""
    gather_mode = N.HYP_GATHER_SAME_RES

    def __init__(self, shadow_creator_dict, casi, lidar, neighborhood, normalize,
                 casi_min=None, casi_max=None, lidar_min=None, lidar_max=None, device=None) -> None:
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.neighborhood = neighborhood
        self.shadow_creator_dict = shadow_creator_dict
        self.casi = self._to_device(casi)
        self.casi_unnormalized_dtype = numpy.dtype(str(self.casi.dtype).replace("torch.", ""))
        self.lidar = None if lidar is None else self._to_device(lidar).reshape(lidar.shape[0], lidar.shape[1])
        self.normalize = normalize
        self.casi_min, self.casi_max, self.lidar_min, self.lidar_max = 0, 1, 0, 1
        self._cmin = self._cmax = self._lmm = None
        if normalize:
            if self.lidar is not None:
                lmin, lmax = E.scene_minmax(self.lidar.view(self.lidar.shape[0], self.lidar.shape[1], 1))
                if lidar_min is not None:
                    lmin = torch.tensor([lidar_min], dtype=torch.float32, device=self.device)
                if lidar_max is not None:
                    lmax = torch.tensor([lidar_max], dtype=torch.float32, device=self.device)
                self._lmm = torch.cat([lmin, lmax]).contiguous()
                self.lidar_min, self.lidar_max = float(lmin.item()), float(lmax.item())
            cmin, cmax = E.scene_minmax(self.casi)          # per-band min and max of the SHIFTED data (:73-77)
            bands = self.casi.shape[2]
            if casi_min is not None:                        # given offsets (scalar or per band), e.g. AVON's 0
                given = torch.as_tensor(numpy.asarray(casi_min), dtype=torch.float32).to(self.device).expand(bands)
                if casi_max is None:                        # the divisor is the maximum of the data shifted by THEM
                    cmax = (cmin + cmax) - given
                cmin = given
            if casi_max is not None:
                cmax = torch.as_tensor(numpy.asarray(casi_max), dtype=torch.float32).to(self.device).expand(bands)
            self._cmin, self._cmax = cmin.contiguous(), cmax.contiguous()
            self.casi_min, self.casi_max = cmin.cpu().numpy(), cmax.cpu().numpy()

    def _to_device(self, a):
        if isinstance(a, torch.Tensor):
            return a.to(self.device).contiguous()
        a = numpy.ascontiguousarray(a)
        if a.dtype not in (numpy.uint16, numpy.float32):
            raise TypeError(f"scene dtype {a.dtype} not supported (uint16, float32)")
        return torch.from_numpy(a).to(self.device)

    def get_data_shape(self):
        dim = self.neighborhood * 2 + 1
        return [dim, dim, self.casi.shape[2] + (1 if self.lidar is not None else 0)]

    def get_casi_band_count(self):
        return self.casi.shape[2]

    def get_scene_shape(self):
        primary = self.lidar if self.lidar is not None else self.casi
        return [primary.shape[0], primary.shape[1]]

    def get_unnormalized_casi_dtype(self):
        return self.casi_unnormalized_dtype

    def get_data_points(self, targets_xy, out=None):
        """Batched form of get_data_point: targets [N,>=2] (x=column, y=row) -> [N,S,S,C] CUDA tensor."""
        if not isinstance(targets_xy, torch.Tensor):
            targets_xy = torch.from_numpy(numpy.ascontiguousarray(numpy.asarray(targets_xy)[:, :2].astype(numpy.int32)))
        xy = targets_xy[:, :2].to(device=self.device, dtype=torch.int32).contiguous()
        return E.gather_patches(self.casi, self.lidar, self.neighborhood, xy, self._cmin, self._cmax, self._lmm,
                                self.gather_mode, out)

    def get_data_point(self, point_x, point_y):
        xy = torch.tensor([[int(point_x), int(point_y)]], dtype=torch.int32, device=self.device)
        return self.get_data_points(xy)[0]


class NNParams:
    def __init__(self, input_iterator, data_with_labels, metrics, predict_tensor):
        self.predict_tensor = predict_tensor
        self.metrics = metrics
        self.data_with_labels = data_with_labels
        self.input_iterator = input_iterator


class ModelInputParams:
    def __init__(self, x, y, device_id, is_training):
        self.is_training = is_training
        self.device_id = device_id
        self.y = y
        self.x = x


class HistogramTensorPair:
    def __init__(self, tensor, name):
        self.name = name
        self.tensor = tensor


class ModelOutputTensors:
    def __init__(self, y_conv, image_output, image_original, histogram_tensors):
        self.image_original = image_original
        self.image_output = image_output
        self.y_conv = y_conv
        self.histogram_tensors = histogram_tensors


class TrainingResult:
    def __init__(self, validation_accuracy, test_accuracy, loss):
        self.loss = loss
        self.test_accuracy = test_accuracy
        self.validation_accuracy = validation_accuracy


class AugmentationInfo:
    def __init__(self, shadow_struct, perform_shadow_augmentation, perform_rotation_augmentation,
                 perform_spectral_augmentation, perform_reflection_augmentation, augmentation_random_threshold):
        self.perform_reflection_augmentation = perform_reflection_augmentation
        self.perform_rotation_augmentation = perform_rotation_augmentation
        self.perform_shadow_augmentation = perform_shadow_augmentation
        self.perform_spectral_augmentation = perform_spectral_augmentation
        self.shadow_struct = shadow_struct
        self.augmentation_random_threshold = augmentation_random_threshold


# ------------------------------------------------------------------------------------------
def labels_to_ids(labels):
    """The reference feeds one-hot uint8 rows (InMemoryImporter.py:24); the engine takes class ids."""
    if labels.dim() == 2:
        labels = labels.argmax(dim=1)
    return labels.to(torch.uint8).contiguous()


class DeviceBatchIterator:
    """training_nn_iterator / simple_nn_iterator (:188-205): batches of a device-resident
    (images, labels) data set.  shuffle=True: reshuffled every epoch and repeated
    (shuffle_and_repeat); shuffle=False: one in-order pass then StopIteration
    (tf.errors.OutOfRangeError in the reference)."""

    def __init__(self, images, labels, batch_size, shuffle, num_epochs=None, seed=1234):
        self.images, self.labels, self.batch_size = images, labels, batch_size
        self.shuffle, self.num_epochs = shuffle, num_epochs
        self.gen = torch.Generator(device="cpu")
        self.gen.manual_seed(seed)
        self.initializer = self.reset
        self.reset()

    def reset(self):
        self.pos, self.epoch = 0, 0
        self.order = self._order()

    def _order(self):
        n = self.images.shape[0]
        if self.shuffle:
            return torch.randperm(n, generator=self.gen).to(self.images.device)
        return None

    def get_next(self):
        n = self.images.shape[0]
        if self.pos >= n:
            self.epoch += 1
            if not self.shuffle or (self.num_epochs is not None and self.epoch >= self.num_epochs):
                raise StopIteration
            self.pos, self.order = 0, self._order()
        lo, hi = self.pos, min(self.pos + self.batch_size, n)
        self.pos = hi
        if self.order is None:
            return self.images[lo:hi], self.labels[lo:hi]
        idx = self.order[lo:hi]
        return self.images.index_select(0, idx), self.labels.index_select(0, idx)


class AugmentingIterator:
    """add_augmentation_graph (:376-394) on the device: every training batch goes through hyp_augment_patches
    (rotation, reflection, spectral offset — one random draw per sample, as the reference's per-sample tf.data maps).
    The shadow map needs a trained GAN generator (gan/gan_utilities.py); it is applied when
    ``augmentation_info.shadow_struct`` carries a callable ``shadow_op`` taking / returning a batch."""

    def __init__(self, inner, augmentation_info, seed=1234):
        self.inner, self.info, self.seed, self.calls = inner, augmentation_info, seed, 0
        self.initializer = inner.initializer

    def get_next(self):
        images, labels = self.inner.get_next()
        info = self.info
        self.calls += 1
        if info.perform_shadow_augmentation and info.shadow_struct is not None:
            gen = torch.Generator(device="cpu")
            gen.manual_seed(self.seed * 7919 + self.calls)
            pick = torch.rand(images.shape[0], generator=gen) < info.augmentation_random_threshold
            if bool(pick.any()):            # decided on the host draw: no device synchronisation in the input pipeline
                images = torch.where(pick.to(images.device, non_blocking=True).view(-1, 1, 1, 1),
                                     info.shadow_struct.shadow_op(images), images)
        if info.perform_rotation_augmentation or info.perform_reflection_augmentation or \
                info.perform_spectral_augmentation:
            spectral = float(info.perform_spectral_augmentation) if info.perform_spectral_augmentation else 0.0
            images = E.augment_patches(images.contiguous(), info.perform_rotation_augmentation,
                                       info.perform_reflection_augmentation, spectral, self.seed + self.calls)
        return images, labels


def training_nn_iterator(data_set, augmentation_info, batch_size, num_epochs, device, prefetch_size):
    images, labels = data_set
    it = DeviceBatchIterator(images, labels, batch_size, True, num_epochs)
    if augmentation_info is not None and (augmentation_info.perform_rotation_augmentation or
                                          augmentation_info.perform_reflection_augmentation or
                                          augmentation_info.perform_spectral_augmentation or
                                          (augmentation_info.perform_shadow_augmentation and
                                           augmentation_info.shadow_struct is not None)):
        return AugmentingIterator(it, augmentation_info)
    return it


def simple_nn_iterator(data_set, batch_size):
    images, labels = data_set
    return DeviceBatchIterator(images, labels, batch_size, False)


class TrainOp:
    """What optimize_nn's `train_step` is in the reference: running it performs one optimizer
    step on the next training batch and returns the batch loss."""

    def __init__(self, model, iterator, algorithm_params, loss_func, allreduce=None):
        self.model, self.iterator, self.alg, self.loss_func, self.allreduce = model, iterator, algorithm_params, loss_func, allreduce
        self.last_loss = None
        self.last_logits = None

    def run(self):
        images, labels = self.iterator.get_next()
        eng = self.model.engine_for(images, self.alg)
        self.last_loss = eng.train_step(images.contiguous(), labels_to_ids(labels), allreduce=self.allreduce)
        return self.last_loss

    __call__ = run

    @property
    def global_step(self):
        return self.model.engine.global_step

    @property
    def learning_rate(self):
        return self.model.engine.learning_rate()


def optimize_nn(deep_nn_template, images, labels, device_id, name_prefix, algorithm_params, loss_func):
    """Reference: common/common_nn_ops.py:208-240.  Eager: performs ONE train step on
    (images, labels) and returns (y_conv, cross_entropy, learning_rate, global_step)."""
    model = deep_nn_template.func.__self__
    eng = model.engine_for(images, algorithm_params)
    lr = eng.learning_rate()
    loss = eng.train_step(images.contiguous(), labels_to_ids(labels))
    y_conv = eng.debug_tensor("fc_final").view(images.shape[0], -1)
    return y_conv, loss[0], lr, eng.global_step


class HostBatchTrainer:
    """The end-to-end call a user of the plug-in API makes per step when batches arrive in
    HOST memory (the reference's feed path: monitored_session_runner.py:182-184 +
    prefetch_to_device): H2D copy of the batch, one optimize_nn step, D2H read of the loss."""

    def __init__(self, engine, allreduce=None):
        self.engine, self.allreduce = engine, allreduce
        self._bufs = [None, None]   # two device batches: one being trained on, one being filled
        self._cur = 0
        self._pending = None        # (host_x data_ptr, buffer index, copy-done event) of a prefetched batch
        self._copy_stream = None

    def _buffers(self, i, host_x, host_y):
        b = self._bufs[i]
        if b is None or b[0].shape != host_x.shape:
            b = (torch.empty(host_x.shape, dtype=torch.float32, device=self.engine.device),
                 torch.empty(host_y.shape, dtype=torch.uint8, device=self.engine.device))
            self._bufs[i] = b
        return b

    def step(self, host_x, host_y, prefetch=None):
        """One train step on the host batch; returns the loss on the host.  prefetch = (next_host_x, next_host_y)
        starts the NEXT step's host->device copy on a copy stream while this step computes — the device-side
        counterpart of the reference's prefetch_to_device (common/common_nn_ops.py:200)."""
        main = torch.cuda.current_stream()
        if self._pending is not None and self._pending[0] == host_x.data_ptr():
            _, self._cur, ev = self._pending          # this batch is already on its way
            main.wait_event(ev)
            x, y = self._bufs[self._cur]
        else:
            x, y = self._buffers(self._cur, host_x, host_y)
            x.copy_(host_x, non_blocking=True)
            y.copy_(host_y, non_blocking=True)
        self._pending = None
        if prefetch is not None:
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=self.engine.device)
            nxt = 1 - self._cur
            nx, ny = self._buffers(nxt, prefetch[0], prefetch[1])
            self._copy_stream.wait_stream(main)       # the buffer's previous step has been consumed
            with torch.cuda.stream(self._copy_stream):
                nx.copy_(prefetch[0], non_blocking=True)
                ny.copy_(prefetch[1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
            self._pending = (prefetch[0].data_ptr(), nxt, ev)
        loss = self.engine.train_step(x, y, allreduce=self.allreduce)
        return loss.cpu()  # synchronises: the step's result is on the host


class SceneBatchTrainer:
    """The same end-to-end step when the scene lives in HBM (what InMemoryImporter amounts to here): the host hands in
    the step's TARGET LIST — int32 [B, 3] = (x, y, class), the rows of the reference's sample sets
    (loader/DataLoader.py:5-47) — instead of B pre-cut patches.  H2D copy of the list (12 B per patch instead of
    P*P*C*4), patch gather from the resident scene (hyp_gather_patches), one optimize_nn step, D2H read of the loss."""

    def __init__(self, engine, casi, lidar, neighborhood, mode=0, allreduce=None):
        self.engine, self.allreduce = engine, allreduce
        self.casi, self.lidar, self.neighborhood, self.mode = casi, lidar, neighborhood, mode
        self.cmin, self.cmax = E.scene_minmax(casi)
        self.lmm = None
        if lidar is not None:
            lmin, lmax = E.scene_minmax(lidar.reshape(lidar.shape[0], lidar.shape[1], 1))
            self.lmm = torch.cat([lmin, lmax]).contiguous()
        self._targets = self._patches = None

    def step(self, host_targets):
        n = host_targets.shape[0]
        if self._targets is None or self._targets.shape[0] != n:
            S = 2 * self.neighborhood + 1
            channels = self.casi.shape[2] + (0 if self.lidar is None else 1)
            self._targets = torch.empty((n, 3), dtype=torch.int32, device=self.engine.device)
            self._patches = torch.empty((n, S, S, channels), dtype=torch.float32, device=self.engine.device)
        self._targets.copy_(host_targets, non_blocking=True)
        xy = self._targets[:, :2].contiguous()
        labels = self._targets[:, 2].to(torch.uint8)
        E.gather_patches(self.casi, self.lidar, self.neighborhood, xy, self.cmin, self.cmax, self.lmm, self.mode,
                         self._patches)
        loss = self.engine.train_step(self._patches, labels, allreduce=self.allreduce)
        return loss.cpu()  # synchronises: the step's result is on the host


class MetricOpsHolder:
    """create_metric_tensors (:243-277): streaming accuracy / mean-per-class accuracy / kappa and
    an int32 confusion accumulator, all derived from the device-side confusion matrix."""

    def __init__(self, num_classes, device):
        self.num_classes = num_classes
        self.confusion = torch.zeros((num_classes, num_classes), dtype=torch.int32, device=device)

    def metric_variables_reset_op(self):
        self.confusion.zero_()

    def combined_metric_update_op(self, logits, labels):
        return E.argmax_confusion(logits.contiguous(), labels_to_ids(labels), self.confusion)

    def _conf(self):
        return self.confusion.cpu().numpy().astype(numpy.float64)

    @property
    def accuracy(self):
        c = self._conf()
        return float(numpy.trace(c) / max(c.sum(), 1.0))

    @property
    def mean_per_class_accuracy(self):
        c = self._conf()
        rows = c.sum(axis=1)
        return float(numpy.where(rows > 0, numpy.diag(c) / numpy.maximum(rows, 1), 0.0).mean())

    @property
    def kappa(self):
        c = self._conf()
        n = c.sum()
        if n == 0:
            return 0.0
        po = numpy.trace(c) / n
        pe = float((c.sum(axis=1) * c.sum(axis=0)).sum()) / (n * n)
        return float((po - pe) / (1 - pe)) if pe != 1 else 0.0


def create_metric_tensors(labels, y_conv, class_range, name_prefix, device=None):
    return MetricOpsHolder(class_range.stop, device if device is not None else torch.device("cuda"))


def calculate_class_accuracies_using_confusion(confusion_matrix, class_range):
    """Per-class recall (diagonal / row sums: ground truths) and precision (diagonal / column sums: predictions) of a
    confusion matrix, 0 where a class never occurs; only the classes of ``class_range`` are returned
    (reference: common/common_nn_ops.py:280-292)."""
    matrix = numpy.asarray(confusion_matrix, dtype=numpy.float64)
    hits = numpy.diagonal(matrix)

    def ratio(totals):
        return numpy.divide(hits, totals, out=numpy.zeros_like(hits), where=totals != 0)

    picked = numpy.arange(class_range.start, class_range.stop, class_range.step)
    return ratio(matrix.sum(axis=1))[picked], ratio(matrix.sum(axis=0))[picked]


def calculate_accuracy(sess, nn_params, class_range):
    """Reference: :295-310.  `sess` is unused (eager); nn_params.predict_tensor is the callable
    images -> logits, nn_params.input_iterator a DeviceBatchIterator."""
    nn_params.metrics.metric_variables_reset_op()
    nn_params.input_iterator.reset()
    while True:
        try:
            images, labels = nn_params.input_iterator.get_next()
        except StopIteration:
            break
        nn_params.metrics.combined_metric_update_op(nn_params.predict_tensor(images), labels)
    confusion_matrix = nn_params.metrics.confusion.cpu().numpy()
    class_recall, class_precisions = calculate_class_accuracies_using_confusion(confusion_matrix, class_range)
    return nn_params.metrics.accuracy, class_recall, class_precisions, nn_params.metrics.kappa, \
        nn_params.metrics.mean_per_class_accuracy


def perform_prediction(sess, nn_params, prediction_result):
    """Reference: :313-327.  prediction_result: uint8 [H,W] CUDA tensor pre-filled with 255;
    nn_params.data_with_labels.targets [N,3] (x, y, class).  argmax + scatter run on device."""
    targets = nn_params.data_with_labels.targets
    if not isinstance(targets, torch.Tensor):
        targets = torch.from_numpy(numpy.ascontiguousarray(numpy.asarray(targets)[:, :2].astype(numpy.int32)))
    xy = targets[:, :2].to(device=prediction_result.device, dtype=torch.int32).contiguous()
    pos = 0
    nn_params.input_iterator.reset()
    while True:
        try:
            images, _ = nn_params.input_iterator.get_next()
        except StopIteration:
            break
        pred = E.argmax_confusion(nn_params.predict_tensor(images).contiguous())
        E.scatter_class_map(pred, xy[pos:pos + pred.shape[0]].contiguous(), prediction_result)
        pos += pred.shape[0]
    return prediction_result


def create_graph(training_data_set, testing_data_set, validation_data_set, class_range,
                 batch_size, prefetch_size, device_id, num_epochs, algorithm_params, model,
                 augmentation_info, create_separate_validation_branch):
    """Reference: :330-373.  Returns the same 6-tuple; cross_entropy / learning_rate are
    callables reading the TrainOp's latest values, train_step is the TrainOp."""
    deep_nn_template = partial(model.create_tensor_graph, class_count=class_range.stop)
    model._class_count = class_range.stop  # what make_template("nn_core", ..., class_count=...) binds (:333)
    training_input_iter = training_nn_iterator(training_data_set, augmentation_info, batch_size, num_epochs,
                                               device_id, prefetch_size)
    train_step = TrainOp(model, training_input_iter, algorithm_params, model.get_loss_func)
    train_nn_params = NNParams(input_iterator=training_input_iter, data_with_labels=None, metrics=None,
                               predict_tensor=None)

    def predict(images):
        return deep_nn_template(ModelInputParams(x=images, y=None, device_id=device_id, is_training=False),
                                algorithm_params=algorithm_params).y_conv

    dev = training_data_set[0].device
    testing_input_iter = simple_nn_iterator(testing_data_set, batch_size)
    test_metrics = create_metric_tensors(None, None, class_range, "testing", dev)
    testing_nn_params = NNParams(input_iterator=testing_input_iter, data_with_labels=None, metrics=test_metrics,
                                 predict_tensor=predict)
    validation_nn_params = NNParams(input_iterator=testing_input_iter, data_with_labels=None, metrics=test_metrics,
                                    predict_tensor=predict)
    if create_separate_validation_branch:  # the reference passes the bound method -> always truthy (App. B)
        validation_input_iter = simple_nn_iterator(validation_data_set, batch_size)
        validation_nn_params = NNParams(input_iterator=validation_input_iter, data_with_labels=None,
                                        metrics=create_metric_tensors(None, None, class_range, "validation", dev),
                                        predict_tensor=predict)
    cross_entropy = lambda: train_step.last_loss  # noqa: E731
    learning_rate = lambda: train_step.learning_rate  # noqa: E731
    return cross_entropy, learning_rate, testing_nn_params, train_nn_params, validation_nn_params, train_step


def get_model_from_name(model_name):
    return get_class("nnmodel." + model_name + "." + model_name)()


def get_importer_from_name(importer_name):
    return get_class("importer." + importer_name + "." + importer_name)()


def get_loader_from_name(loader_name, path):
    return get_class("loader." + loader_name + "." + loader_name)(path)


def shadow_ratio_of_data_set(data_set, padded_shadow_map, neighborhood):
    """calculate_shadow_ratio (:473-483) over the cube the reference holds — symmetric-padded by ``neighborhood`` and
    min-max normalised — computed from the unpadded, un-normalised cube resident on the device: a padded pixel is a
    copy of scene pixel (rows[i], cols[j]), so the two masked means are weighted sums over the scene (weights = how
    often the padding repeats a row / column), and (mean_lit - min) / (mean_shadow - min) is the ratio of the
    normalised means (the per-band max cancels).  One [H*W] x [H*W, C] product in fp64 per mask."""
    casi = data_set.casi
    height, width, bands = casi.shape
    padded_shadow_map = numpy.asarray(padded_shadow_map)
    if padded_shadow_map.shape != (height + 2 * neighborhood, width + 2 * neighborhood):
        raise ValueError(f"shadow map {padded_shadow_map.shape} does not match the padded scene "
                         f"{(height + 2 * neighborhood, width + 2 * neighborhood)}")
    in_shadow = torch.as_tensor(padded_shadow_map != 0, device=casi.device).to(torch.float64)
    rows = torch.as_tensor(numpy.pad(numpy.arange(height), neighborhood, mode="symmetric"), device=casi.device)
    cols = torch.as_tensor(numpy.pad(numpy.arange(width), neighborhood, mode="symmetric"), device=casi.device)
    flat = (rows[:, None] * width + cols[None, :]).reshape(-1)
    cube = casi.reshape(height * width, bands).to(torch.float64)
    means = []
    for mask in (1.0 - in_shadow, in_shadow):
        weights = torch.zeros(height * width, dtype=torch.float64, device=casi.device).index_add_(0, flat, mask.reshape(-1))
        means.append((weights @ cube) / weights.sum())
    cmin = getattr(data_set, "_cmin", None)
    offset = 0.0 if cmin is None else cmin.to(torch.float64)
    return ((means[0] - offset) / (means[1] - offset)).to(torch.float32).cpu().numpy()


def load_shadow_map_common(data_set, neighborhood, shadow_file_name):
    """Reference :567-571.  ``shadow_file_name`` is a TIFF (or ``.npy``) file, or the [H,W] array of 0 / 1 itself; the
    map is symmetric-padded by ``neighborhood`` like the scene the reference keeps."""
    if isinstance(shadow_file_name, str):
        if shadow_file_name.endswith(".npy"):
            shadow_map = numpy.load(shadow_file_name)
        else:
            from hypelcnn_b200.utilities.tiff_io import imread
            shadow_map = imread(shadow_file_name)
    else:
        shadow_map = numpy.asarray(shadow_file_name)
    shadow_map = numpy.pad(shadow_map, neighborhood, mode="symmetric")
    shadow_ratio = None if data_set is None else shadow_ratio_of_data_set(data_set, shadow_map, neighborhood)
    return shadow_map, shadow_ratio


# host-side sample-list / shadow-statistics helpers of the reference's module, implemented in sample_ops.py
from hypelcnn_b200.common.sample_ops import (calculate_shadow_ratio, create_colored_image,  # noqa: E402,F401
                                             create_target_image_via_samples, read_targets_from_image,
                                             shuffle_test_data_using_ratio, shuffle_training_data_using_ratio,
                                             shuffle_training_data_using_size)
