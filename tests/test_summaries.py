"""CPU test of the TensorBoard event files (classify/monitored_session_runner.py:16-28 tags): written with
hypelcnn_b200.classify.summaries, read back with TensorBoard's own event loader."""
import glob
from types import SimpleNamespace

import numpy


def _events(log_dir, upgraded=False):
    """upgraded=False: the records as they are on disk; True: after TensorBoard's data-compat pass (what its UI sees)."""
    from tensorboard.backend.event_processing.event_file_loader import EventFileLoader, RawEventFileLoader
    from tensorboard.compat.proto import event_pb2
    out = []
    for path in sorted(glob.glob(log_dir + "/events.out.tfevents.*")):
        if upgraded:
            out += [e for e in EventFileLoader(path).Load() if e.HasField("summary")]
        else:
            out += [e for e in (event_pb2.Event.FromString(r) for r in RawEventFileLoader(path).Load()) if e.HasField("summary")]
    return out


def test_classification_summaries_round_trip(tmp_path):
    from hypelcnn_b200.classify.summaries import ClassificationSummaryWriter
    conf_t = numpy.array([[5, 1], [0, 7]])
    conf_v = numpy.array([[3, 2], [1, 9]])
    test_m = SimpleNamespace(confusion=conf_t, accuracy=12 / 13, mean_per_class_accuracy=0.9, kappa=0.8)
    val_m = SimpleNamespace(confusion=conf_v, accuracy=0.8, mean_per_class_accuracy=0.75, kappa=0.55)
    w = ClassificationSummaryWriter(str(tmp_path))
    w.add_text("flags", '{"batch_size": 20}', 0)
    w.add_classification_summaries(350, 1.25, 3e-4, test_m, val_m, {"nn_core/fc_final/weights": numpy.linspace(-1, 1, 101)})
    w.close()
    events = _events(str(tmp_path))
    assert [e.step for e in events] == [0, 350]
    flags = events[0].summary.value[0]
    assert flags.tag == "flags" and flags.metadata.plugin_data.plugin_name == "text"
    assert flags.tensor.string_val[0] == b'<pre>{"batch_size": 20}</pre>'
    by_tag = {v.tag: v for v in events[1].summary.value}
    assert set(by_tag) == {"training_cross_entropy", "training_learning_rate", "test_confusion", "test_overall_accuracy",
                           "validation_confusion", "validation_overall_accuracy", "validation_average_accuracy",
                           "validation_kappa", "nn_core/fc_final/weights"}
    assert abs(by_tag["training_cross_entropy"].simple_value - 1.25) < 1e-7
    assert abs(by_tag["validation_kappa"].simple_value - 0.55) < 1e-7
    t = by_tag["validation_confusion"].tensor
    assert [d.size for d in t.tensor_shape.dim] == [2, 2] and list(t.string_val) == [b"3", b"2", b"1", b"9"]
    h = by_tag["nn_core/fc_final/weights"].histo
    assert h.num == 101 and h.min == -1.0 and h.max == 1.0 and abs(h.sum) < 1e-9 and sum(h.bucket) == 101
    assert len(h.bucket) == len(h.bucket_limit) and list(h.bucket_limit) == sorted(h.bucket_limit)
    # TensorBoard's own reader classifies them as scalar / text / histogram time series
    seen = {v.tag: v.metadata.plugin_data.plugin_name for e in _events(str(tmp_path), upgraded=True) for v in e.summary.value}
    assert seen["training_cross_entropy"] == "scalars" and seen["validation_confusion"] == "text"
    assert seen["nn_core/fc_final/weights"] == "histograms"


def test_histogram_buckets_follow_tensorflows_default_limits():
    from hypelcnn_b200.classify.summaries import _LIMITS, histogram_proto
    assert len(_LIMITS) == 2 * 775 + 1 and _LIMITS[775] == 0.0 and abs(_LIMITS[776] - 1e-12) < 1e-24   # 774 + DBL_MAX per side
    h = histogram_proto([0.0, 0.0, 1.0, -1.0])
    # zeros land in the bucket whose upper limit is the first positive limit; empty runs collapse into one entry
    k = list(h.bucket_limit).index(_LIMITS[776])
    assert h.bucket[k] == 2 and sorted(b for b in h.bucket if b > 0) == [1, 1, 2]
    assert histogram_proto([]).num == 0
