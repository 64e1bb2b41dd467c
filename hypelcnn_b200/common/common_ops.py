"""Small host helpers with the names the reference's callers use (common/common_ops.py:4-31)."""
import importlib
import os


def get_class(kls):
    """Resolve a dotted "pkg.Module.Class" name.  The reference's registry strings ("nnmodel.X.X", "importer.X.X",
    "loader.X.X") are looked up inside this package first and then as top-level modules, so third-party plug-ins on
    PYTHONPATH keep working."""
    module_name, _, attr = kls.rpartition(".")
    failure = None
    for prefix in ("hypelcnn_b200.", ""):
        try:
            return getattr(importlib.import_module(prefix + module_name), attr)
        except (ImportError, AttributeError) as e:
            failure = e
    raise ImportError(f"cannot resolve {kls}: {failure}")


def is_integer_num(n):
    """True for ints and for floats with no fractional part (bools count as ints, like in the reference)."""
    return isinstance(n, int) or (isinstance(n, float) and n.is_integer())


def replace_abbrs(txt, abbrs_dict):
    """Apply every word -> abbreviation substitution of the dict, in its order."""
    for word in abbrs_dict:
        txt = txt.replace(word, abbrs_dict[word])
    return txt


def path_leaf(path):
    """Last component of a path written with either separator (trailing separators and a drive prefix ignored)."""
    unified = path.replace(chr(92), "/")
    if len(unified) >= 2 and unified[1] == ":":
        unified = unified[2:]
    tail = unified.rsplit("/", 1)[-1]
    return tail if tail else unified.rstrip("/").rsplit("/", 1)[-1]
