"""B200-native sparse x tall-skinny-dense multiply behind the CombBLAS surface.

Layout
  csrc/              hand-written sm_100a CUDA + the C ABI  -> lib/libcombblas_b200.so
  include/CombBLAS/  header-only C++ host layer mirroring the reference's SpParMat / CommGrid /
                     semiring API (the product's host side: the reference is C++)
  capi.py            ctypes binding of include/combblas_b200.h used by tests/ and bench.py

The directory name carries a hyphen, so Python code loads it through ``load_package()`` in the
repository root module ``cbb200_loader`` (or importlib); nothing here depends on the name.
There is no CPU fallback: importing works anywhere, but every compute call needs a B200.
"""
from . import capi  # noqa: F401
from .capi import (Context, Dense, Tile, CBError, F32, F64, I32, I64, U8, PATTERN,  # noqa: F401
                   PLUS_TIMES, MIN_PLUS, MAX_SEL2ND, OR_AND, build_library, library_path)
