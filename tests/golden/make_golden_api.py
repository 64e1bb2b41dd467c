#!/usr/bin/env python
"""The caller-side API surface of the reference's common/common_nn_ops.py, recorded by importing the module with the
TensorFlow stubs of make_golden.py and introspecting it: argument lists of the public functions and of the classes'
methods / constructors.  tests/test_api_surface.py holds this engine's mirror (hypelcnn_b200/common/common_nn_ops.py)
against it.  Build container only; output ``api_surface.json`` is committed.

usage: python tests/golden/make_golden_api.py"""
import inspect
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402


def main():
    G.install_stubs()
    sys.path.insert(0, G.REF)
    import common.common_nn_ops as ops
    surface = {"functions": {}, "classes": {}}
    for name, obj in vars(ops).items():
        if name.startswith("_") or getattr(obj, "__module__", None) != ops.__name__:
            continue
        if inspect.isfunction(obj):
            surface["functions"][name] = list(inspect.signature(obj).parameters)
        elif inspect.isclass(obj):
            methods = {}
            for mname, fn in vars(obj).items():
                fn = fn.__func__ if isinstance(fn, (staticmethod, classmethod)) else fn
                if inspect.isfunction(fn) and (not mname.startswith("_") or mname == "__init__"):
                    methods[mname] = list(inspect.signature(fn).parameters)
            surface["classes"][name] = {"methods": methods, "fields": list(getattr(obj, "_fields", []))}
    with open(os.path.join(HERE, "api_surface.json"), "w") as f:
        json.dump(surface, f, indent=1, sort_keys=True)
    print("functions", len(surface["functions"]), "classes", len(surface["classes"]))


if __name__ == "__main__":
    main()
