"""A shadow map without the validation targets (reference: utilities/remove_test_targets_from_shadow.py): every
validation target that lies under the shadow map is cleared from it, so GAN pairs drawn from the map never contain
pixels the classifier is later validated on; the result is written as ``shadow_map.tif``.
    python -m hypelcnn_b200.utilities.remove_test_targets_from_shadow --loader_name ... --path ..."""
import argparse

import numpy

from hypelcnn_b200.common.cmd_parser import add_parse_cmds_for_loaders, add_parse_cmds_for_loggers
from hypelcnn_b200.common.common_nn_ops import get_loader_from_name
from hypelcnn_b200.utilities.tiff_io import imwrite


def clear_targets(shadow_map, targets):
    """-> (map with the shadowed targets cleared, number of targets that were NOT under the shadow)."""
    shadow_map = numpy.array(shadow_map)
    targets = numpy.asarray(targets).astype(int).reshape(-1, 3)
    under_shadow = shadow_map[targets[:, 1], targets[:, 0]] == 1
    shadow_map[targets[under_shadow, 1], targets[under_shadow, 0]] = 0
    return shadow_map, int((~under_shadow).sum())


def run(flags, output_path="shadow_map.tif"):
    loader = get_loader_from_name(flags.loader_name, flags.path)
    sample_set = loader.load_samples(0.1, 0.1)
    data_set = loader.load_data(0, True)
    shadow_map, _ = loader.load_shadow_map(0, data_set)
    shadow_map, _ = clear_targets(shadow_map, sample_set.validation_targets)
    imwrite(output_path, shadow_map, planarconfig='contig')
    return shadow_map


def main(argv=None):
    parser = argparse.ArgumentParser()
    add_parse_cmds_for_loggers(parser)
    add_parse_cmds_for_loaders(parser)
    flags, _ = parser.parse_known_args(argv)
    run(flags)


if __name__ == '__main__':
    main()
