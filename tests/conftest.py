import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
# the hub variant declines tiles whose hub columns serve too few nonzeros; the small matrices of tests/test_new_variants_gpu.py must
# run it anyway.  The library reads the threshold once per process, at its first multiply, so it is set before anything is loaded.
os.environ.setdefault("CB_SPMM_HUB_MIN_COVER_PCT", "0")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver on the GPU box)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref/libcbref.so (the compiled, unmodified reference)")


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    gpu = has_gpu()
    from oracle import oracle as O
    ref = O.ref_available()
    for item in items:
        if "gpu" in item.keywords and not gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "ref" in item.keywords and not ref:
            item.add_marker(pytest.mark.skip(reason="oracle/_ref/libcbref.so not built"))


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
