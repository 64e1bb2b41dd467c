"""The host helper mirrors against golden vectors produced by running the reference's own code
(tests/golden/make_golden_host.py -> tests/golden/host_helpers.json)."""
import json
import os

import numpy

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "host_helpers.json")))
VALUES = {"0": 0, "1": 1, "-3": -3, "2.0": 2.0, "2.5": 2.5, "-0.0": -0.0, "1e+20": 1e20, "inf": float("inf"), "'3'": "3",
          "None": None, "True": True}


def test_common_ops_match_the_reference():
    from hypelcnn_b200.common import common_ops as C
    for path, leaf in GOLD["path_leaf"]:
        assert C.path_leaf(path) == leaf, path
    for text, abbrs, result in GOLD["replace_abbrs"]:
        assert C.replace_abbrs(text, abbrs) == result
    for rep, expected in GOLD["is_integer_num"]:
        if isinstance(expected, bool):
            assert C.is_integer_num(VALUES[rep]) == expected, rep
    assert C.get_class("nnmodel.HYPELCNNModel.HYPELCNNModel").__name__ == "HYPELCNNModel"     # registry-style names
    assert C.get_class("collections.OrderedDict").__name__ == "OrderedDict"                   # plain modules too


def test_dummy_sampler_matches_the_reference():
    from hypelcnn_b200.gan.gan_sampling_methods import DummySampler
    g = GOLD["dummy_sampler"]

    class Shape:
        def get_data_shape(self):
            return g["shape"][1:]
    normal, shadow = DummySampler(g["element_count"], g["fill_value"], g["coefficient"]).get_sample_pairs(Shape(), None, None)
    assert list(normal.shape) == g["shape"] and str(normal.dtype) == g["dtype"] and g["constant"]
    assert numpy.all(normal == g["normal"]) and numpy.all(shadow == g["shadow"])


def test_plugin_interfaces_have_the_reference_signatures():
    """Every abstract method of the reference's DataImporter / DataLoader / NNModel / Wrapper / InferenceWrapper exists
    here under the same name with the same argument list (the drop-in boundary)."""
    import importlib
    import inspect
    assert len(GOLD["interfaces"]) == 7
    for qualified, entry in GOLD["interfaces"].items():
        module_name, _, cls_name = qualified.rpartition(".")
        cls = getattr(importlib.import_module("hypelcnn_b200." + module_name), cls_name)
        assert sorted(getattr(cls, "__abstractmethods__", [])) == entry["abstract"], qualified
        for mname, params in entry["methods"].items():
            assert list(inspect.signature(getattr(cls, mname)).parameters) == params, (qualified, mname)
        if "members" in entry:
            assert {m.name: m.value for m in cls} == entry["members"]
