import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# The parity suite asserts bit-level properties (batch-position invariance, run-to-run equality, bit-exact chunk stitching):
# it runs the bit-reproducible GEMM schedule.  The library default (tail split mode 2, last-bit run-to-run differences in the
# tiles of a partial last wave) is covered by tests/test_model_gpu.py::test_default_tail_split_mode_* within a stated tolerance.
os.environ.setdefault("JAT_GEMM_TAIL", "0")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
