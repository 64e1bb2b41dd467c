"""Confusion matrices back out of TensorBoard event files (reference: utilities/read_summary_file.py): every
``validation_confusion`` text summary (optionally only at the given steps) is decoded, saved as
``<grandparent>_<parent>_s<step>.csv`` and fed to stat_extractor.  Event files are read with the TensorBoard package's
loader (the reference uses tf.compat.v1.train.summary_iterator); the files written by classify/summaries.py and by the
reference's TF run decode the same way.
    python -m hypelcnn_b200.utilities.read_summary_file <log dir> [step ...]"""
import glob
import os
import sys
from pathlib import Path

import numpy
from tensorboard.backend.event_processing.event_file_loader import EventFileLoader

from hypelcnn_b200.utilities.stat_extractor import extract_statistics_info, print_statistics_info


def confusion_matrices_of(event_path, filtered_steps=(), tag="validation_confusion"):
    """-> [(step, [C,C] int matrix)] of one event file."""
    found = []
    for event in EventFileLoader(event_path).Load():
        if not event.HasField("summary") or (filtered_steps and event.step not in filtered_steps):
            continue
        for value in event.summary.value:
            if value.tag == tag:
                shape = [d.size for d in value.tensor.tensor_shape.dim]
                matrix = numpy.array([int(s) for s in value.tensor.string_val], dtype=int).reshape(shape)
                found.append((event.step, matrix))
    return found


def collect(log_dir, filtered_steps=(), output_dir="."):
    matrices = []
    for event_path in sorted(glob.glob(os.path.join(log_dir, "event*"))):
        parent_dir = Path(event_path).parent
        try:
            for step, matrix in confusion_matrices_of(event_path, filtered_steps):
                print("Step %i in %s" % (step, event_path))
                record = os.path.join(output_dir, f"{parent_dir.parent.name}_{parent_dir.name}_s{step}.csv")
                print("Saving to file:", record)
                numpy.savetxt(record, matrix, fmt="%d", delimiter=",")
                matrices.append(matrix)
        except Exception as error:      # a truncated file of a crashed run: report and go on, like the reference
            print("Error reading summary file: ", event_path, error)
    return matrices


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    print_statistics_info(extract_statistics_info(collect(argv[0], [int(step) for step in argv[1:]])))


if __name__ == '__main__':
    main()
