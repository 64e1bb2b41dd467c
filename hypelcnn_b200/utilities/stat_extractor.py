"""Run statistics from saved confusion matrices (reference: utilities/stat_extractor.py): overall accuracy, per-class
accuracy and Cohen's kappa per run, mean +- std over the runs of a directory of ``*.csv`` confusion matrices.
Array arithmetic instead of the reference's element loops; pinned against the reference's own module
(tests/test_stat_extractor.py)."""
import glob
import os
import sys
from collections import namedtuple

import numpy

MetricsHolder = namedtuple('MetricsHolder', ['aa_array', 'kappa_array', 'oa_array', 'sample_count'])


def histogram(confusion_matrix, index):
    """Row sums (index 0: how often each class was the truth) or column sums (index 1: how often it was predicted)."""
    confusion_matrix = numpy.asarray(confusion_matrix)
    if index not in (0, 1):
        return numpy.zeros(confusion_matrix.shape[index], dtype=int)
    return confusion_matrix.sum(axis=1 - index).astype(int)


def calc_kappa(conf_mat):
    """Cohen's kappa: 1 - (observed disagreement) / (disagreement expected from the marginals)."""
    conf_mat = numpy.asarray(conf_mat, dtype=float)
    total = float(conf_mat.sum())
    expected = numpy.outer(histogram(conf_mat, 0).astype(float), histogram(conf_mat, 1).astype(float)) / total
    off_diagonal = ~numpy.eye(len(conf_mat), dtype=bool)
    return 1.0 - (conf_mat[off_diagonal].sum() / total) / (expected[off_diagonal].sum() / total)


def calc_mean_quadratic_weighted_kappa(kappas, weights=None):
    """Mean of kappas in Fisher's z space (kappas capped to +-0.999), optionally weighted (weights normalised to mean 1)."""
    kappas = numpy.clip(numpy.array(kappas, dtype=float), -.999, .999)
    weights = numpy.ones(numpy.shape(kappas)) if weights is None else weights / numpy.mean(weights)
    z = numpy.mean(0.5 * numpy.log((1 + kappas) / (1 - kappas)) * weights)
    return (numpy.exp(2 * z) - 1) / (numpy.exp(2 * z) + 1)


def extract_accuracy_metrics(confusion_matrix):
    """-> (overall accuracy, per-class accuracy = diagonal / row sum, kappa, samples per class)."""
    confusion_matrix = numpy.asarray(confusion_matrix)
    class_based_samples = confusion_matrix.sum(axis=1).astype(int)
    overall_accuracy = numpy.trace(confusion_matrix) / numpy.sum(confusion_matrix)
    with numpy.errstate(divide="ignore", invalid="ignore"):
        class_accuracy = numpy.diag(confusion_matrix) / confusion_matrix.sum(axis=1).astype(float)
    return overall_accuracy, class_accuracy, calc_kappa(confusion_matrix), class_based_samples


def extract_statistics_info(confusion_matrix_list):
    """One row per run.  Like the reference, run i is stored at slot i - 1 (the first run ends up last) and the sample
    counts are those of the first run."""
    runs = [extract_accuracy_metrics(m) for m in confusion_matrix_list]
    if not runs:
        return MetricsHolder(aa_array=None, kappa_array=None, oa_array=None, sample_count=None)
    order = list(range(1, len(runs))) + [0]
    return MetricsHolder(aa_array=numpy.array([runs[i][1] for i in order], dtype=float),
                         kappa_array=numpy.array([runs[i][2] for i in order], dtype=float),
                         oa_array=numpy.array([runs[i][0] for i in order], dtype=float), sample_count=runs[0][3])


def get_conf_list_from_directory(directory):
    return [numpy.loadtxt(filename, dtype=int, delimiter=",") for filename in glob.glob(os.path.join(directory, "*.csv"))]


def calculate_mean_std_metrics(oa_array, aa_array, kappa_array):
    per_run_aa = numpy.mean(aa_array, axis=1)
    return (numpy.mean(oa_array), numpy.std(oa_array), numpy.mean(per_run_aa), numpy.std(per_run_aa),
            numpy.mean(kappa_array), numpy.std(kappa_array))


def print_statistics_info(metrics_holder):
    for oa, aa, kappa in zip(metrics_holder.oa_array, metrics_holder.aa_array, metrics_holder.kappa_array):
        print("OA: %.4f AA: %.4f Kappa: %.4f" % (oa, numpy.mean(aa), kappa))
    print("#Metrics statistics:")
    summary = calculate_mean_std_metrics(metrics_holder.oa_array, metrics_holder.aa_array, metrics_holder.kappa_array)
    for label, (mean, std) in zip(("OA:   ", "AA:   ", "Kappa:"), zip(summary[0::2], summary[1::2])):
        print("%s %.4f +- %.4f" % (label, mean, std))
    print("#Class based accuracy")
    for aa_mean, aa_std, count in zip(numpy.mean(metrics_holder.aa_array, axis=0), numpy.std(metrics_holder.aa_array, axis=0),
                                      metrics_holder.sample_count):
        print("%.4f +- %.4f %d" % (aa_mean, aa_std, count))


def main():
    print_statistics_info(extract_statistics_info(get_conf_list_from_directory(sys.argv[1])))


if __name__ == '__main__':
    main()
