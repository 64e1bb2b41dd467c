"""hypelcnn_b200/utilities/stat_extractor.py against the reference's own module, which is plain numpy and imports as
it is — run side by side here when the reference checkout is present (build container), otherwise against the
golden values it produced (tests/golden/stat_extractor_golden.json)."""
import contextlib
import importlib.util
import io
import json
import os

import numpy
import pytest

from hypelcnn_b200.utilities import stat_extractor as S

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "stat_extractor_golden.json")
REFERENCE = "/root/reference/utilities/stat_extractor.py"


def _matrices():
    rng = numpy.random.default_rng(11)
    out = []
    for classes, runs in ((15, 4), (3, 1), (11, 6)):
        runs_of_set = []
        for _ in range(runs):
            m = rng.integers(0, 30, (classes, classes))
            m[numpy.arange(classes), numpy.arange(classes)] += rng.integers(50, 400, classes)
            runs_of_set.append(m)
        out.append(runs_of_set)
    return out


def _report(module, runs):
    holder = module.extract_statistics_info(runs)
    text = io.StringIO()
    with contextlib.redirect_stdout(text):
        module.print_statistics_info(holder)
    return {"oa": holder.oa_array.tolist(), "aa": holder.aa_array.tolist(), "kappa": holder.kappa_array.tolist(),
            "samples": holder.sample_count.tolist(), "printed": text.getvalue(),
            "mean_kappa": float(module.calc_mean_quadratic_weighted_kappa(holder.kappa_array)),
            "weighted_kappa": float(module.calc_mean_quadratic_weighted_kappa(
                holder.kappa_array, numpy.arange(1, len(runs) + 1, dtype=float))),
            "hist": [module.histogram(runs[0], 0).tolist(), module.histogram(runs[0], 1).tolist()]}


def _assert_same(got, want):
    assert got["printed"] == want["printed"] and got["samples"] == want["samples"] and got["hist"] == want["hist"]
    for key in ("oa", "aa", "kappa", "mean_kappa", "weighted_kappa"):
        assert numpy.allclose(got[key], want[key], rtol=1e-12, atol=1e-14), key


def test_against_golden_values():
    golden = json.load(open(GOLDEN))
    for runs, want in zip(_matrices(), golden):
        _assert_same(_report(S, runs), want)


@pytest.mark.skipif(not os.path.exists(REFERENCE), reason="reference checkout not present (GPU box)")
def test_side_by_side_with_the_reference_module():
    spec = importlib.util.spec_from_file_location("reference_stat_extractor", REFERENCE)
    reference = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(reference)
    for runs in _matrices():
        _assert_same(_report(S, runs), _report(reference, runs))


def test_directory_of_csv_files(tmp_path, capsys):
    runs = _matrices()[0]
    for i, m in enumerate(runs):
        numpy.savetxt(tmp_path / f"run_{i}.csv", m, fmt="%d", delimiter=",")
    loaded = S.get_conf_list_from_directory(str(tmp_path))
    assert len(loaded) == 4 and sorted(int(m.sum()) for m in loaded) == sorted(int(m.sum()) for m in runs)
    S.print_statistics_info(S.extract_statistics_info(loaded))
    assert "#Metrics statistics:" in capsys.readouterr().out
    assert S.extract_statistics_info([]).oa_array is None


def test_confusion_matrices_round_trip_through_event_files(tmp_path, capsys, monkeypatch):
    """Training loop -> TensorBoard event file -> read_summary_file -> statistics: what classify/summaries.py writes under
    ``validation_confusion`` comes back as the same integer matrices, filtered by step, saved as CSV."""
    from types import SimpleNamespace
    from hypelcnn_b200.classify.summaries import ClassificationSummaryWriter
    from hypelcnn_b200.utilities import read_summary_file as R
    log_dir = tmp_path / "experiment" / "run_1"
    os.makedirs(log_dir)
    runs = _matrices()[2][:3]
    writer = ClassificationSummaryWriter(str(log_dir))
    for step, m in zip((100, 200, 300), runs):
        metrics = SimpleNamespace(confusion=m.astype(numpy.int32), accuracy=0.5, mean_per_class_accuracy=0.4, kappa=0.3)
        writer.add_classification_summaries(step, 1.0, 3e-4, metrics, metrics)
    writer.close()
    out_dir = tmp_path / "csv"
    os.makedirs(out_dir)
    got = R.collect(str(log_dir), output_dir=str(out_dir))
    assert len(got) == 3 and all(numpy.array_equal(a, b) for a, b in zip(got, runs))
    assert sorted(os.listdir(out_dir)) == ["experiment_run_1_s100.csv", "experiment_run_1_s200.csv", "experiment_run_1_s300.csv"]
    assert numpy.array_equal(numpy.loadtxt(out_dir / "experiment_run_1_s200.csv", dtype=int, delimiter=","), runs[1])
    only = R.collect(str(log_dir), filtered_steps=[300], output_dir=str(out_dir))
    assert len(only) == 1 and numpy.array_equal(only[0], runs[2])
    monkeypatch.chdir(tmp_path)
    R.main([str(log_dir), "100", "200"])
    assert "OA:" in capsys.readouterr().out
