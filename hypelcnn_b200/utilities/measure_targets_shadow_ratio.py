"""Measured shadowed / lit ratio per band over the paired samples of a scene (reference:
utilities/measure_targets_shadow_ratio.py): pair with the chosen sampler, divide, drop the pairs with a non-finite band,
report mean +- std per band as the band-ratio curve (CSV, PDF with matplotlib).  The pair matrices live where the
sampler put them (HBM for the device data sets); the statistics are a handful of tensor reductions there.
    python -m hypelcnn_b200.utilities.measure_targets_shadow_ratio --loader_name ... --path ... --pairing_method random"""
import argparse

import torch

from hypelcnn_b200.common.cmd_parser import add_flags, add_parse_cmds_for_loaders, add_parse_cmds_for_loggers
from hypelcnn_b200.common.common_nn_ops import get_loader_from_name
from hypelcnn_b200.gan.wrapper_registry import get_sampling_map
from hypelcnn_b200.gan.wrappers.gan_common import plot_overall_info, read_hsi_data


def ratio_statistics(normal_data_as_matrix, shadow_data_as_matrix):
    """[N,1,1,bands] pairs -> (mean, std) of shadow / normal per band over the pairs whose ratios are all finite."""
    normal = torch.as_tensor(normal_data_as_matrix)
    shadow = torch.as_tensor(shadow_data_as_matrix).to(normal.device)
    ratio = (shadow / normal).reshape(normal.shape[0], -1)
    ratio = ratio[torch.isfinite(ratio).all(dim=1)]
    return ratio.mean(dim=0).cpu().numpy(), ratio.std(dim=0, unbiased=False).cpu().numpy()


def run(flags, output_dir="./"):
    loader = get_loader_from_name(flags.loader_name, flags.path)
    data_set = loader.load_data(0, True)
    shadow_map, _ = loader.load_shadow_map(0, data_set)
    normal, shadow = read_hsi_data(loader, data_set, shadow_map, flags.pairing_method, get_sampling_map())
    mean_res, std_res = ratio_statistics(normal, shadow)
    plot_overall_info(loader.get_band_measurements(), mean_res, mean_res - std_res, mean_res + std_res, 0,
                      f"{flags.loader_name.lower()}_{flags.pairing_method.lower()}", output_dir)
    return mean_res, std_res


def main(argv=None):
    parser = argparse.ArgumentParser()
    add_parse_cmds_for_loggers(parser)
    add_parse_cmds_for_loaders(parser)
    add_flags(parser, (("pairing_method", str, "random", "random, target, dummy or neighbour"),))
    flags, _ = parser.parse_known_args(argv)
    run(flags)


if __name__ == '__main__':
    main()
