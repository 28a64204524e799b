import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """Reference input/output vectors produced by oracle/make_golden.py from /root/reference."""

    def __init__(self):
        with open(os.path.join(GOLDEN_DIR, "reference_vectors.json")) as f:
            self.meta = json.load(f)
        self.arrays = np.load(os.path.join(GOLDEN_DIR, "reference_vectors.npz"))
        self.cases = self.meta["cases"]

    def inputs(self, case):
        x1 = self.arrays[case["id"] + "_x1"]
        x2 = None if case["self"] else self.arrays[case["id"] + "_x2"]
        return x1, x2

    def output(self, case):
        return self.arrays[case["id"] + "_out"]


_golden = None


def golden():
    global _golden
    if _golden is None:
        _golden = Golden()
    return _golden


def golden_cases(fn_prefix=None):
    g = golden()
    return [c for c in g.cases if fn_prefix is None or c["fn"].startswith(fn_prefix)]


def case_id(case):
    return f'{case["id"]}-{case["note"]}'.replace(" ", "_")


def is_integer_valued(*arrs):
    return all(a is None or bool(np.all(a == np.floor(a))) for a in arrs)
