#!/usr/bin/env python
"""Golden vectors for the small host helpers, produced by EXECUTING the reference's own modules (they import without
TensorFlow): common/common_ops.py (path_leaf, is_integer_num, replace_abbrs) and the DummySampler of
gan/gan_sampling_methods.py:191-201.  Build container only (/root/reference is not on the GPU box); the output
``host_helpers.json`` beside this file is committed and read by tests/test_host_helpers.py.

usage: python tests/golden/make_golden_host.py"""
import json
import os
import sys

import numpy

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

PATHS = ["a/b/c.txt", "a/b/c/", "c.txt", "a\\b\\c.txt", "a\\b\\", "/", "", "a/b\\c", "C:\\x\\y.tif", "a/b//", "C:file",
         "\\\\srv\\share\\f.tif", "x/", "./log/run1/", "alg_param_hypelcnn.json", "/data/2013_DFTC/2013_DFTC"]
NUMBERS = [0, 1, -3, 2.0, 2.5, -0.0, 1e20, float("inf"), "3", None, True]
ABBRS = [("grss2013dataloader_hypelcnnmodel_trn010_alg_param_hypelcnn_7x7", {"model": "mdl", "dataloader": "ldr", "alg_param_": "p"}),
         ("model_of_models", {"model": "mdl", "mdl_of": "x"}), ("nothing to do", {})]


def main():
    sys.path.insert(0, REF)
    if not hasattr(numpy, "int"):
        numpy.int = int                       # the alias the reference itself patches in (gan_sampling_methods.py:8)
    import common.common_ops as C
    import gan.gan_sampling_methods as S

    def integer(n):
        try:
            return bool(C.is_integer_num(n))
        except Exception as e:                # inf.is_integer() etc.
            return type(e).__name__

    class Shape:
        def get_data_shape(self):
            return [1, 1, 6]
    normal, shadow = S.DummySampler(element_count=5, fill_value=0.5, coefficient=2).get_sample_pairs(Shape(), None, None)
    out = {"path_leaf": [[p, C.path_leaf(p)] for p in PATHS],
           "is_integer_num": [[repr(n), integer(n)] for n in NUMBERS],
           "replace_abbrs": [[t, d, C.replace_abbrs(t, d)] for t, d in ABBRS],
           "dummy_sampler": {"element_count": 5, "fill_value": 0.5, "coefficient": 2, "shape": list(normal.shape),
                             "dtype": str(normal.dtype), "normal": float(normal.flat[0]), "shadow": float(shadow.flat[0]),
                             "constant": bool((normal == normal.flat[0]).all() and (shadow == shadow.flat[0]).all())}}
    # the plug-in boundary itself: abstract method names and argument lists of the reference's four ABC modules
    import inspect
    import importer.DataImporter as DI
    import loader.DataLoader as DL
    import nnmodel.NNModel as NM
    import gan.wrappers.wrapper as GW
    interfaces = {}
    for module, names in ((DI, ["DataImporter"]), (DL, ["DataLoader", "SampleSet", "LoadingMode"]), (NM, ["NNModel"]),
                          (GW, ["Wrapper", "InferenceWrapper"])):
        for name in names:
            cls = getattr(module, name)
            entry = {"abstract": sorted(getattr(cls, "__abstractmethods__", [])), "methods": {}}
            for mname, fn in inspect.getmembers(cls, predicate=inspect.isfunction):
                if not mname.startswith("__") or mname == "__init__":
                    entry["methods"][mname] = list(inspect.signature(fn).parameters)
            if name == "LoadingMode":
                entry["members"] = {m.name: m.value for m in cls}
            interfaces[module.__name__ + "." + name] = entry
    out["interfaces"] = interfaces
    with open(os.path.join(HERE, "host_helpers.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote host_helpers.json")


if __name__ == "__main__":
    main()
