"""Mirror of the reference's common/common_ops.py:4-31 helpers (same names, same behaviour)."""
import importlib
import ntpath


def get_class(kls):
    """Resolve "pkg.Module.Class" (reference: common/common_ops.py:4-10).  The reference's
    registry names ("nnmodel.X.X", "importer.X.X", "loader.X.X") resolve inside this package
    first, then as top-level modules, so plug-ins on PYTHONPATH keep working."""
    parts = kls.split(".")
    module = ".".join(parts[:-1])
    last = None
    for prefix in ("hypelcnn_b200.", ""):
        try:
            m = importlib.import_module(prefix + module)
            return getattr(m, parts[-1])
        except (ImportError, AttributeError) as e:
            last = e
    raise ImportError(f"cannot resolve {kls}: {last}")


def is_integer_num(n):
    if isinstance(n, int):
        return True
    if isinstance(n, float):
        return n.is_integer()
    return False


def replace_abbrs(txt, abbrs_dict):
    for word, abbr in abbrs_dict.items():
        txt = txt.replace(word, abbr)
    return txt


def path_leaf(path):
    head, tail = ntpath.split(path)
    return tail or ntpath.basename(head)
