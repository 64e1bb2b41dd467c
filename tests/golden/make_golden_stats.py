#!/usr/bin/env python
"""stat_extractor_golden.json: the reference's own utilities/stat_extractor.py (plain numpy, imports as it is) run over
the seeded confusion matrices of tests/test_stat_extractor.py.  Build container only.
usage: python tests/golden/make_golden_stats.py"""
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.test_stat_extractor import GOLDEN, REFERENCE, _matrices, _report  # noqa: E402

spec = importlib.util.spec_from_file_location("reference_stat_extractor", REFERENCE)
reference = importlib.util.module_from_spec(spec)
spec.loader.exec_module(reference)
json.dump([_report(reference, runs) for runs in _matrices()], open(GOLDEN, "w"), indent=1)
print("wrote", GOLDEN)
