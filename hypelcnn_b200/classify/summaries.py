"""TensorBoard event files with the reference's tags.

The reference logs through TF1 summary ops (classify/monitored_session_runner.py:16-28,
common/common_nn_ops.py:588-600): scalars ``training_cross_entropy``, ``training_learning_rate``,
``test_overall_accuracy``, ``validation_overall_accuracy``, ``validation_average_accuracy``, ``validation_kappa``; text
summaries ``test_confusion`` / ``validation_confusion`` (the confusion matrix through tf.as_string) and the run's
``flags`` / ``algorithm_params`` JSON wrapped in <pre>; optionally one histogram per model variable.  Here the same
records are written with the TensorBoard package's own event writer and protos (no TensorFlow): a TF1 scalar summary is
``Summary.Value{tag, simple_value}``, a TF1 text summary a DT_STRING tensor value with plugin name "text", a TF1
histogram a ``HistogramProto`` over TensorFlow's default bucket limits (±1e-12 · 1.1^k, empty runs collapsed).
Host-side reporting only: values are read from the device by the caller (one .item() / .cpu() per logged quantity).
"""
import sys

import numpy
from tensorboard.compat.proto import event_pb2, summary_pb2, tensor_pb2, tensor_shape_pb2, types_pb2
from tensorboard.summary.writer.event_file_writer import EventFileWriter


def _default_bucket_limits():
    pos, v = [], 1e-12
    while v < 1e20:
        pos.append(v)
        v *= 1.1
    pos.append(sys.float_info.max)
    return numpy.array([-x for x in reversed(pos)] + [0.0] + pos)


_LIMITS = _default_bucket_limits()


def histogram_proto(values):
    """tensorflow/core/lib/histogram: bucket i counts limit[i-1] <= v < limit[i]; runs of empty buckets collapse."""
    v = numpy.asarray(values, dtype=numpy.float64).reshape(-1)
    h = summary_pb2.HistogramProto()
    if v.size == 0:
        return h
    h.min, h.max, h.num, h.sum, h.sum_squares = float(v.min()), float(v.max()), float(v.size), float(v.sum()), float((v * v).sum())
    counts = numpy.bincount(numpy.searchsorted(_LIMITS, v, side="right"), minlength=len(_LIMITS))[:len(_LIMITS)]
    i = 0
    while i < len(_LIMITS):
        end, count = _LIMITS[i], counts[i]
        i += 1
        if count <= 0:
            while i < len(_LIMITS) and counts[i] <= 0:
                end, count = _LIMITS[i], counts[i]
                i += 1
        h.bucket_limit.append(float(end))
        h.bucket.append(float(count))
    return h


def _text_value(tag, strings, shape):
    tensor = tensor_pb2.TensorProto(dtype=types_pb2.DT_STRING,
                                    tensor_shape=tensor_shape_pb2.TensorShapeProto(
                                        dim=[tensor_shape_pb2.TensorShapeProto.Dim(size=int(s)) for s in shape]))
    tensor.string_val.extend(s.encode() if isinstance(s, str) else s for s in strings)
    meta = summary_pb2.SummaryMetadata(plugin_data=summary_pb2.SummaryMetadata.PluginData(plugin_name="text"))
    return summary_pb2.Summary.Value(tag=tag, tensor=tensor, metadata=meta)


def _number(x):
    return float(x.item() if hasattr(x, "item") else x)


class ClassificationSummaryWriter:
    """One event file per log directory, like summary_io.SummaryWriterCache.get(log_dir)."""

    def __init__(self, log_dir):
        self._writer = EventFileWriter(log_dir)

    def _add(self, values, step):
        self._writer.add_event(event_pb2.Event(step=int(step), summary=summary_pb2.Summary(value=values)))

    def add_scalar(self, tag, value, step):
        self._add([summary_pb2.Summary.Value(tag=tag, simple_value=_number(value))], step)

    def add_text(self, name, value, step):
        """TextSummaryAtStartHook (common_nn_ops.py:588-600): the JSON of the flags / algorithm parameters."""
        self._add([_text_value(name, ["<pre>" + value + "</pre>"], ())], step)

    def add_classification_summaries(self, step, cross_entropy, learning_rate, testing_metrics, validation_metrics,
                                     model_variables=None):
        """add_classification_summaries (monitored_session_runner.py:16-28) evaluated at `step`.  *_metrics: objects
        with .confusion [C, C], .accuracy, .mean_per_class_accuracy, .kappa (MetricOpsHolder); model_variables:
        optional {variable name: array} for log_all_model_variables."""
        def confusion(tag, metrics):
            m = numpy.asarray(metrics.confusion.cpu() if hasattr(metrics.confusion, "cpu") else metrics.confusion)
            return _text_value(tag, [str(int(x)) for x in m.reshape(-1)], m.shape)      # tf.as_string of an int tensor

        scalar = lambda tag, x: summary_pb2.Summary.Value(tag=tag, simple_value=_number(x))  # noqa: E731
        values = [scalar("training_cross_entropy", cross_entropy), scalar("training_learning_rate", learning_rate),
                  confusion("test_confusion", testing_metrics), scalar("test_overall_accuracy", testing_metrics.accuracy),
                  confusion("validation_confusion", validation_metrics),
                  scalar("validation_overall_accuracy", validation_metrics.accuracy),
                  scalar("validation_average_accuracy", validation_metrics.mean_per_class_accuracy),
                  scalar("validation_kappa", validation_metrics.kappa)]
        for name, array in (model_variables or {}).items():
            array = array.detach().cpu().numpy() if hasattr(array, "detach") else array
            values.append(summary_pb2.Summary.Value(tag=name, histo=histogram_proto(array)))
        self._add(values, step)

    def flush(self):
        self._writer.flush()

    def close(self):
        self._writer.close()
