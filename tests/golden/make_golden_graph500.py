"""Goldens of the reference's own generator (run in the build container, where /root/reference and oracle/_ref exist):
  * edges of the packed Graph500 stream (RefGen21::generate_kronecker_range, -DDETERMINISTIC seed) at several scales / offsets;
  * the matrix ReleaseTests/GenWriteMatrix.cpp builds (SpParMat(DEL, false), RemoveLoops, Symmetricize) at scale 10, directed and
    symmetrised, as triples with their multiplicity values.
Written to tests/golden/graph500_ref.npz; the device generator (csrc/cb_gen.cu) and the test-only mock are pinned to it."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
EDGE_CASES = [(10, 0, 4096), (16, 60000, 2048), (20, 70000, 512), (24, (16 << 24) - 512, 512)]       # (scale, first edge, count)
GENWRITE = [(10, 16, 1), (10, 16, 0), (9, 8, 1)]                                                   # (scale, edge factor, symmetric)


def main():
    R = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libcbref.so"))
    R.cbref_graph500_edges.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
    R.cbref_genwrite.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int64), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    out = {}
    for scale, first, count in EDGE_CASES:
        s, d = np.empty(count, np.int64), np.empty(count, np.int64)
        R.cbref_graph500_edges(scale, first, count, s.ctypes.data, d.ctypes.data)
        out[f"edges_s{scale}_{first}"] = np.stack([s, d])
    for scale, ef, sym in GENWRITE:
        nnz = ctypes.c_int64()
        R.cbref_genwrite(scale, ef, sym, ctypes.byref(nnz), None, None, None)
        I, J, V = np.empty(nnz.value, np.int64), np.empty(nnz.value, np.int64), np.empty(nnz.value, np.int32)
        R.cbref_genwrite(scale, ef, sym, ctypes.byref(nnz), I.ctypes.data, J.ctypes.data, V.ctypes.data)
        o = np.lexsort((J, I))
        out[f"genwrite_s{scale}_ef{ef}_sym{sym}"] = np.stack([I[o], J[o], V[o].astype(np.int64)])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "graph500_ref.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
