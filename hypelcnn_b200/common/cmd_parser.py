"""Command-line flag groups of the reference (common/cmd_parser.py) as tables: one row per flag = (name, type, default,
help).  The ``add_parse_cmds_for_*`` functions keep the reference's names so its ``main`` functions read the same;
defaults are pinned against the reference's own parsers in tests/test_classify_apps.py / test_gan_train_app.py."""
import os


def type_ensure_strtobool(val):
    """distutils.util.strtobool (gone from Python 3.12) over str(val): y/yes/t/true/on/1 and n/no/f/false/off/0."""
    text = str(val).lower()
    if text in ("y", "yes", "t", "true", "on", "1"):
        return True
    if text in ("n", "no", "f", "false", "off", "0"):
        return False
    raise ValueError(f"invalid truth value {text!r}")


LOADER_FLAGS = (("path", str, "/data/2013_DFTC/2013_DFTC", "Input data path"),
                ("loader_name", str, "GRSS2013DataLoader", "Data set loader name: GRSS2013DataLoader, GRSS2018DataLoader, "
                                                           "GULFPORTDataLoader, GULFPORTALTDataLoader, AVONDataLoader, Synthetic*"),
                ("neighborhood", int, 0, "Neighborhood for data extraction, 1 means 3x3 patches"),
                ("test_ratio", float, 0.05, "Ratio of training data to use in testing"),
                ("train_ratio", float, 0.10, "Ratio (< 1) or per-class count (>= 1) of the training samples"))
LOGGER_FLAGS = (("base_log_path", str, None, "Base path for logs / checkpoints, default: working directory"),
                ("output_path", str, None, "Path for output images, default: working directory"))
TRAINER_FLAGS = (("batch_size", int, 20, "Batch size"),
                 ("step", int, 50000, "Number of training steps (this or --epoch)"),
                 ("epoch", int, None, "Number of passes over the data (this or --step)"))
MODEL_FLAGS = (("algorithm_param_path", str, None, "Algorithm parameter (json) file"),
               ("model_name", str, "HYPELCNNModel", "CONCNNModel, DUALCNNModel or HYPELCNNModel"))
IMPORTER_FLAGS = (("importer_name", str, "InMemoryImporter", "GeneratorImporter, InMemoryImporter or TFRecordImporter"),)
JSON_LOADER_FLAGS = (("flag_config_file", str, None, "Flags as json"),)


def add_flags(parser, table):
    for name, kind, default, text in table:
        if default is None and name in ("base_log_path", "output_path"):
            default = os.getcwd()
        parser.add_argument("--" + name, nargs="?", type=kind, default=default, help=text)


def add_parse_cmds_for_loaders(parser):
    add_flags(parser, LOADER_FLAGS)


def add_parse_cmds_for_loggers(parser):
    add_flags(parser, LOGGER_FLAGS)


def add_parse_cmds_for_trainers(parser):
    add_flags(parser, TRAINER_FLAGS)


def add_parse_cmds_for_models(parser):
    add_flags(parser, MODEL_FLAGS)


def add_parse_cmds_for_importers(parser):
    add_flags(parser, IMPORTER_FLAGS)


def add_parse_cmds_for_json_loader(parser):
    add_flags(parser, JSON_LOADER_FLAGS)
