"""Shadow / de-shadow a whole scene with a trained generator (reference: gan/gan_infer_image_for_shadow.py).

The reference runs ``session.run`` once per pixel (:73-88).  Here the scene is walked in chunks of up to
``CHUNK`` pixels: one gather launch, one generator launch, a masked select and the de-normalisation on the device;
only the finished chunk crosses to the host, where it is cast to the scene's original dtype and written as
``shadow_image_<mode>_<checkpoint>[_all].tif``.  (The reference's extra RGB rendering needs the ``colour`` package's
CIE tables and is not produced.)"""
import argparse
import os

import numpy
import torch

from hypelcnn_b200.common.cmd_parser import (add_flags, add_parse_cmds_for_loaders, add_parse_cmds_for_loggers,
                                             type_ensure_strtobool)
from hypelcnn_b200.common.common_nn_ops import get_loader_from_name
from hypelcnn_b200.gan.gan_infer_for_shadow import restore_generators
from hypelcnn_b200.gan.wrapper_registry import get_infer_wrapper
from hypelcnn_b200.utilities.tiff_io import imwrite

CHUNK = 1 << 18
APP_FLAGS = (("gan_type", str, "cycle_gan", "cycle_gan, gan_x2y, gan_y2x, cut_x2y, cut_y2x, dcl_gan, dcl_cycle_gan"),
             ("make_them_shadow", str, "", "shadow: shadow the lit pixels, deshadow: light the shadowed ones, else none"),
             ("convert_all", type_ensure_strtobool, False, "Convert every pixel instead of the filtered ones."))


def add_parse_cmds_for_app(parser):
    add_flags(parser, APP_FLAGS)


def conversion_plan(make_them_shadow):
    """-> (use the shadow generator?, the shadow-map value of the pixels to convert, normalised mode name) (:41-51)."""
    if make_them_shadow == "shadow":
        return True, 0, "shadow"
    if make_them_shadow == "deshadow":
        return False, 1, "deshadow"
    return True, -1, "none"


def convert_scene(data_set, shadow_map, infer_model, sign_to_filter_in_shadow_map, convert_all, chunk=CHUNK):
    """[H,W,bands] scene in its un-normalised dtype: converted where the shadow map holds the sign (everywhere with
    ``convert_all``), the original spectrum elsewhere — both de-normalised the way the reference does
    (value * casi_max + casi_min, cast)."""
    height, width = data_set.get_scene_shape()
    bands = data_set.get_casi_band_count()
    target_dtype = data_set.get_unnormalized_casi_dtype()
    casi_min, casi_max = numpy.asarray(data_set.casi_min), numpy.asarray(data_set.casi_max)
    image = numpy.zeros([height * width, bands], dtype=target_dtype)
    selected = numpy.ones(height * width, bool) if convert_all else \
        (numpy.asarray(shadow_map)[:height, :width].reshape(-1) == sign_to_filter_in_shadow_map)
    ys, xs = numpy.divmod(numpy.arange(height * width), width)
    scale = offset = None
    for start in range(0, height * width, chunk):
        stop = min(start + chunk, height * width)
        targets = numpy.stack([xs[start:stop], ys[start:stop]], axis=1).astype(numpy.int32)
        spectra = data_set.get_data_points(targets)[:, :, :, 0:bands]
        centre = spectra.shape[1] // 2
        spectra = spectra[:, centre:centre + 1, centre:centre + 1, :].contiguous()       # neighborhood 0 in the reference
        if scale is None:
            scale = torch.as_tensor(casi_max.astype(numpy.float32)).to(spectra.device)
            offset = torch.as_tensor(casi_min.astype(numpy.float32)).to(spectra.device)
        pick = torch.as_tensor(selected[start:stop]).to(spectra.device)
        if bool(pick.any()):
            spectra = torch.where(pick.view(-1, 1, 1, 1), infer_model(spectra), spectra)
        restored = spectra.reshape(stop - start, bands) * scale + offset
        image[start:stop] = restored.cpu().numpy().astype(target_dtype)
    return image.reshape(height, width, bands)


def run(flags):
    loader = get_loader_from_name(flags.loader_name, flags.path)
    data_set = loader.load_data(0, True)
    shadow_map, _ = loader.load_shadow_map(0, data_set)
    shadow, sign, mode = conversion_plan(flags.make_them_shadow)
    wrapper = get_infer_wrapper(flags.gan_type, bands=data_set.get_casi_band_count())
    if mode != "none":
        restore_generators(wrapper, flags.base_log_path)
    _, infer_model = wrapper.make_inference_graph(data_set, shadow, clip_invalid_values=False)
    hsi_image = convert_scene(data_set, shadow_map, infer_model, sign, flags.convert_all)
    region = "_all" if flags.convert_all else ""
    checkpoint_number = flags.base_log_path.rsplit("-", 1)[-1]
    if checkpoint_number.endswith(".npz"):
        checkpoint_number = checkpoint_number[:-4]
    path = os.path.join(flags.output_path, f"shadow_image_{mode}_{checkpoint_number}{region}.tif")
    print(f"Saving output to {path}")
    imwrite(path, hsi_image, planarconfig="contig")
    return path, hsi_image


def main(argv=None):
    parser = argparse.ArgumentParser()
    add_parse_cmds_for_loaders(parser)
    add_parse_cmds_for_loggers(parser)
    add_parse_cmds_for_app(parser)
    flags, _ = parser.parse_known_args(argv)
    run(flags)


if __name__ == "__main__":
    main()
