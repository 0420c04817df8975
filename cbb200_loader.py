"""Imports the hyphen-named package directory ``combblas-spmm-test_b200/`` as module ``combblas_spmm_test_b200``."""
import importlib.util
import os
import sys

_NAME = "combblas_spmm_test_b200"
_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "combblas-spmm-test_b200")


def load_package():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_DIR, "__init__.py"),
                                                  submodule_search_locations=[_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod
