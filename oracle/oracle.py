"""CPU checker for the SpMM hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package never does (it has no CPU fallback).

Two engines behind one interface:
  * ``spmm`` / ``spmm_summa``  -> oracle/liboracle.so, the plain-C restatement (oracle/spmm_oracle.c)
  * ``ref_spmm`` / ``ref_read_mm`` / ``ref_spgemm_i64`` -> oracle/_ref/libcbref.so, the UNMODIFIED
    reference compiled from /root/reference (oracle/Makefile target ``ref``); present in this
    container and shipped prebuilt to the GPU box, never rebuilt there.

numpy restatements that live here (each cites the reference lines it follows):
  * read_mm ............ Matrix Market ingest with symmetric expansion, include/CombBLAS/SpHelper.h:75-91,147-183
                         and duplicate merge by BinOp, include/CombBLAS/SpParMat.cpp:2962-2967
  * to_dcsc ............ column-compressed tile arrays cp/jc/ir/numx, include/CombBLAS/dcsc.h:124-131
  * synthetic inputs ... counter-based generators shared with the CUDA side (csrc/cb_gen.cu); the
                         recipe (initiator, edge factor, scramble, dedup, loop removal, symmetrise)
                         follows ReleaseTests/GenWriteMatrix.cpp:96-131, the bit stream is our own.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import c_char_p, c_double, c_int, c_int64, c_size_t, c_void_p, POINTER

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

F32, F64, I32, I64, U8, PATTERN = 0, 1, 2, 3, 4, 255
PLUS_TIMES, MIN_PLUS, MAX_SEL2ND, OR_AND = 0, 1, 2, 3
SEMIRING_NAMES = {PLUS_TIMES: "plus_times", MIN_PLUS: "min_plus", MAX_SEL2ND: "select_max", OR_AND: "plus_times"}
NP_OF = {F32: np.float32, F64: np.float64, I32: np.int32, I64: np.int64, U8: np.uint8}
CODE_OF = {np.dtype(np.float32): F32, np.dtype(np.float64): F64, np.dtype(np.int32): I32,
           np.dtype(np.int64): I64, np.dtype(np.uint8): U8, np.dtype(np.bool_): U8}
REF_NAME = {F32: "f32", F64: "f64", I32: "i32", I64: "i64", U8: "bool", PATTERN: "bool"}


def semiring_id(semiring: int, dtype: int):
    """SR::id() (Semirings.h:194,215,239)."""
    t = NP_OF[dtype]
    if semiring in (PLUS_TIMES, OR_AND):
        return t(0)
    if semiring == MIN_PLUS:
        return np.finfo(t).max if np.issubdtype(t, np.floating) else np.iinfo(t).max
    return t(-1)


# ----------------------------------------------------------------------------- library loading
_lib = None
_ref = None


def build(ref: bool | None = None) -> None:
    """Compile liboracle.so, and libcbref.so when /root/reference is present (or ref=True)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "liboracle.so"])
    if ref is None:
        ref = os.path.isdir("/root/reference/include/CombBLAS")
    if ref:
        subprocess.check_call(["make", "-s", "-j4", "-C", HERE, "ref"])


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path) or os.path.getmtime(path) < max(os.path.getmtime(os.path.join(HERE, f)) for f in ("spmm_oracle.c", "gen_oracle.c")):
            subprocess.check_call(["make", "-s", "-C", HERE, "liboracle.so"])
        L = ctypes.CDLL(path)
        L.oracle_spmm_colcompressed.restype = c_int
        L.oracle_spmm_colcompressed.argtypes = [c_int, c_int, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                                c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int]
        L.oracle_spmm_summa.restype = c_int
        L.oracle_spmm_summa.argtypes = [c_int, c_int, c_int, c_size_t, c_size_t, c_int, c_int, c_int64, c_int64, c_int64,
                                        c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]
        L.oracle_owner.restype = c_int
        L.oracle_owner.argtypes = [c_int64, c_int64, c_int, c_int, c_int64, c_int64, POINTER(c_int64), POINTER(c_int64)]
        L.oracle_block_range.restype = None
        L.oracle_block_range.argtypes = [c_int64, c_int, c_int, POINTER(c_int64), POINTER(c_int64)]
        L.oracle_spmv_pt_f64.restype = None
        L.oracle_spmv_pt_f64.argtypes = [c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
        L.oracle_gen_matrix.restype = c_int
        L.oracle_gen_matrix.argtypes = [c_int, c_int, ctypes.c_uint64, c_void_p, c_int, c_int, c_int, POINTER(c_int64),
                                        POINTER(c_void_p), POINTER(c_void_p)]
        L.oracle_free.restype = None
        L.oracle_free.argtypes = [c_void_p]
        L.oracle_set_num_threads.restype = None
        L.oracle_set_num_threads.argtypes = [c_int]
        L.oracle_matrix_values.restype = None
        L.oracle_matrix_values.argtypes = [c_void_p, c_void_p, c_int64, c_int64, ctypes.c_uint64, c_int, c_void_p]
        L.oracle_dense_columns.restype = None
        L.oracle_dense_columns.argtypes = [c_int64, c_int64, c_int64, c_int64, ctypes.c_uint64, c_int, c_int, c_void_p]
        _lib = L
    return _lib


def ref_available() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libcbref.so"))


def ref():
    global _ref
    if _ref is None:
        R = ctypes.CDLL(os.path.join(HERE, "_ref", "libcbref.so"))
        R.cbref_spmm.restype = c_int
        R.cbref_spmm.argtypes = [c_char_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int64,
                                 c_void_p, c_void_p, c_int64, POINTER(c_double)]
        R.cbref_read_mm.restype = c_int
        R.cbref_read_mm.argtypes = [c_char_p, POINTER(c_int64), POINTER(c_int64), POINTER(c_int64), c_void_p, c_void_p, c_void_p]
        R.cbref_spgemm_i64.restype = c_int
        R.cbref_spgemm_i64.argtypes = [c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                                       c_void_p, c_void_p, POINTER(c_int64), c_void_p, c_void_p, c_void_p]
        R.cbref_num_threads.restype = c_int
        R.cbref_set_num_threads.argtypes = [c_int]
        _ref = R
    return _ref


def _p(a):
    return None if a is None else a.ctypes.data_as(c_void_p)


# ----------------------------------------------------------------------------- tile formats
def to_dcsc(m, n, I, J, V=None):
    """COO (unique entries) -> DCSC arrays (cp[nzc+1], jc[nzc], ir[nz], numx[nz]); dcsc.h:124-131.
    Entries are ordered column-major with ascending rows inside a column (SpDCCols.cpp:108-184)."""
    I = np.asarray(I, np.int64)
    J = np.asarray(J, np.int64)
    order = np.lexsort((I, J))
    I, J = I[order], J[order]
    V = None if V is None else np.ascontiguousarray(np.asarray(V)[order])
    jc, counts = np.unique(J, return_counts=True)
    cp = np.zeros(len(jc) + 1, np.int64)
    np.cumsum(counts, out=cp[1:])
    return cp, jc.astype(np.int64), np.ascontiguousarray(I), V


def to_csc(m, n, I, J, V=None):
    """COO -> plain CSC (jc[n+1] pointer array, ir, num); csc.h:71-75."""
    I = np.asarray(I, np.int64)
    J = np.asarray(J, np.int64)
    order = np.lexsort((I, J))
    I, J = I[order], J[order]
    V = None if V is None else np.ascontiguousarray(np.asarray(V)[order])
    cp = np.zeros(n + 1, np.int64)
    np.cumsum(np.bincount(J, minlength=n), out=cp[1:])
    return cp, np.ascontiguousarray(I), V


def a_dtype_code(V):
    return PATTERN if V is None else CODE_OF[np.asarray(V).dtype]


# ----------------------------------------------------------------------------- the oracle multiply
def spmm(semiring, m, n, I, J, V, X, accum_into=None):
    """Y = A (x).(+) X with the C restatement. X: (n,k) C-contiguous; returns (m,k) of X's dtype."""
    X = np.ascontiguousarray(X)
    if X.dtype == np.bool_:
        X = X.view(np.uint8)
    xd = CODE_OF[X.dtype]
    if V is not None and np.asarray(V).dtype == np.bool_:
        V = np.asarray(V).view(np.uint8)
    cp, jc, ir, numx = to_dcsc(m, n, I, J, V)
    k = X.shape[1]
    Y = np.empty((m, k), X.dtype) if accum_into is None else accum_into
    rc = lib().oracle_spmm_colcompressed(semiring, a_dtype_code(V), xd, m, len(jc), _p(cp), _p(jc), _p(ir), _p(numx),
                                         k, _p(X), k, _p(Y), k, 0 if accum_into is None else 1)
    if rc:
        raise ValueError(f"oracle: unsupported combination semiring={semiring} A={a_dtype_code(V)} X={xd}")
    return Y


def spmm_summa(semiring, pr, pc, m, n, I, J, V, X):
    """The same product through the emulated pr x pc SUMMA stage loop (stage partials merged in order)."""
    X = np.ascontiguousarray(X)
    if X.dtype == np.bool_:
        X = X.view(np.uint8)
    if V is not None and np.asarray(V).dtype == np.bool_:
        V = np.asarray(V).view(np.uint8)
    I = np.ascontiguousarray(I, np.int64)
    J = np.ascontiguousarray(J, np.int64)
    V = None if V is None else np.ascontiguousarray(V)
    k = X.shape[1]
    Y = np.empty((m, k), X.dtype)
    rc = lib().oracle_spmm_summa(semiring, a_dtype_code(V), CODE_OF[X.dtype], 0 if V is None else V.dtype.itemsize,
                                 X.dtype.itemsize, pr, pc, m, n, len(I), _p(I), _p(J), _p(V), k, _p(X), _p(Y))
    if rc:
        raise ValueError("oracle: unsupported combination")
    return Y


def owner(total_m, total_n, pr, pc, grow, gcol):
    lr, lc = c_int64(), c_int64()
    r = lib().oracle_owner(total_m, total_n, pr, pc, grow, gcol, ctypes.byref(lr), ctypes.byref(lc))
    return r, lr.value, lc.value


def block_range(total, nb, b):
    s, l = c_int64(), c_int64()
    lib().oracle_block_range(total, nb, b, ctypes.byref(s), ctypes.byref(l))
    return s.value, l.value


def spmv_pt_f64(m, n, I, J, V, x):
    cp, jc, ir, numx = to_dcsc(m, n, I, J, np.asarray(V, np.float64))
    x = np.ascontiguousarray(x, np.float64)
    y = np.empty(m, np.float64)
    lib().oracle_spmv_pt_f64(m, len(jc), _p(cp), _p(jc), _p(ir), _p(numx), _p(x), _p(y))
    return y


# ----------------------------------------------------------------------------- the real reference
def ref_key(semiring, a_code, x_code):
    return f"{SEMIRING_NAMES[semiring]}:{REF_NAME[a_code]}:{REF_NAME[x_code]}"


def ref_spmm(semiring, m, n, I, J, V, X, via=0, panel=0, threads=None):
    """Run the unmodified reference (Mult_AnXBn_Synch via=0, k x SpMV via=1). Returns (Y, seconds)."""
    R = ref()
    if threads:
        R.cbref_set_num_threads(int(threads))
    X = np.ascontiguousarray(X)
    Xb = X.view(np.uint8) if X.dtype == np.bool_ else X
    xd = CODE_OF[Xb.dtype]
    I = np.ascontiguousarray(I, np.int64)
    J = np.ascontiguousarray(J, np.int64)
    if V is None:
        ad, Vb = PATTERN, None
    else:
        Vb = np.ascontiguousarray(V)
        if Vb.dtype == np.bool_:
            Vb = Vb.view(np.uint8)
        ad = CODE_OF[Vb.dtype]
    k = Xb.shape[1]
    Y = np.empty((m, k), Xb.dtype)
    sec = c_double(0)
    rc = R.cbref_spmm(ref_key(semiring, ad, xd).encode(), via, m, n, len(I), _p(I), _p(J), _p(Vb), k, _p(Xb), _p(Y),
                      panel, ctypes.byref(sec))
    if rc:
        raise ValueError(f"reference: rc={rc} for key {ref_key(semiring, ad, xd)}")
    return Y, sec.value


def ref_grid_available() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "cbref_grid"))


def _run_grid(args, ranks, threads=1, timeout=900, extra_env=None):
    env = dict(os.environ, CBMPI_NP=str(ranks), OMP_NUM_THREADS=str(threads), CBMPI_TIMEOUT=str(timeout), **(extra_env or {}))
    r = subprocess.run([os.path.join(HERE, "_ref", "cbref_grid")] + args, env=env, capture_output=True, text=True, timeout=timeout + 30)
    if r.returncode:
        raise RuntimeError(f"cbref_grid {args} on {ranks} ranks: rc={r.returncode}\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
    return r.stdout


def ref_grid_torus(ranks):
    """The SpMMError.cpp program run by the unmodified reference on sqrt(ranks) x sqrt(ranks) processes; returns its report line."""
    return [l for l in _run_grid(["torus"], ranks).splitlines() if l.startswith("torus")][0]


def ref_grid_spmm(semiring, ranks, m, n, I, J, V, X, via=0, threads=1, check_distribution=True, timeout=900):
    """The unmodified reference on a sqrt(ranks) x sqrt(ranks) process grid (oracle/_ref/cbref_grid: one process per rank over
    the shared-memory MPI stand-in).  via: 0 Mult_AnXBn_Synch, 1 k x SpMV, 2 _DoubleBuff, 3 _Overlap.
    Returns (Y, seconds, stdout)."""
    import tempfile
    X = np.ascontiguousarray(X)
    Xb = X.view(np.uint8) if X.dtype == np.bool_ else X
    xd = CODE_OF[Xb.dtype]
    if V is None:
        ad, Vb = PATTERN, None
    else:
        Vb = np.ascontiguousarray(V)
        Vb = Vb.view(np.uint8) if Vb.dtype == np.bool_ else Vb
        ad = CODE_OF[Vb.dtype]
    k = Xb.shape[1]
    with tempfile.TemporaryDirectory(prefix="cbref_grid_") as d:
        with open(os.path.join(d, "meta.txt"), "w") as f:
            f.write(f"{m} {n} {len(I)} {k} {0 if Vb is None else 1}\n")
        np.ascontiguousarray(I, np.int64).tofile(os.path.join(d, "I.bin"))
        np.ascontiguousarray(J, np.int64).tofile(os.path.join(d, "J.bin"))
        if Vb is not None:
            Vb.tofile(os.path.join(d, "V.bin"))
        Xb.tofile(os.path.join(d, "X.bin"))
        out = _run_grid(["spmm", ref_key(semiring, ad, xd), str(via), d], ranks, threads, timeout=timeout,
                        extra_env=None if check_distribution else {"CBREF_SKIP_DISTCHECK": "1"})
        Y = np.empty((m, k), Xb.dtype)
        seen = np.zeros((m, k), bool)
        for r in range(ranks):
            raw = np.fromfile(os.path.join(d, f"Y_{r}.bin"), np.uint8)
            r0, c0, rows, cols = (int(v) for v in raw[:32].view(np.int64))
            Y[r0:r0 + rows, c0:c0 + cols] = raw[32:].view(Xb.dtype).reshape(rows, cols)
            seen[r0:r0 + rows, c0:c0 + cols] = True
        assert seen.all(), "cbref_grid: the ranks' blocks do not tile Y"
    sec = float([l for l in out.splitlines() if l.startswith("multiply seconds")][0].split()[2])
    return Y, sec, out


def ref_grid_layout(glen, ranks):
    """FullyDistVec layout of a length-glen vector as the reference computes it on sqrt(ranks)^2 processes:
    dict of int arrays until[rank], len[rank], owner[index], lind[index]."""
    out = {}
    for line in _run_grid(["layout", str(glen)], ranks).splitlines():
        key, *vals = line.split()
        if key in ("until", "len", "owner", "lind"):
            out[key] = np.array([int(v) for v in vals], np.int64)
    return out


def ref_grid_mm(path, ranks, X):
    """BASELINE config C1 on sqrt(ranks)^2 processes: the reference reads the Matrix Market file itself (ParallelReadMM) and
    multiplies by X (n x k float64) with Mult_AnXBn_Synch.  Returns (Y, report line)."""
    import tempfile
    X = np.ascontiguousarray(X, np.float64)
    n, k = X.shape
    with tempfile.TemporaryDirectory(prefix="cbref_grid_") as d:
        X.tofile(os.path.join(d, "X.bin"))
        out = _run_grid(["mm", path, str(k), d], ranks)
        blocks = []
        for r in range(ranks):
            raw = np.fromfile(os.path.join(d, f"Y_{r}.bin"), np.uint8)
            r0, c0, rows, cols = (int(v) for v in raw[:32].view(np.int64))
            blocks.append((r0, c0, raw[32:].view(np.float64).reshape(rows, cols)))
    m = max(r0 + b.shape[0] for r0, _, b in blocks)
    Y = np.full((m, k), np.nan)
    for r0, c0, b in blocks:
        Y[r0:r0 + b.shape[0], c0:c0 + b.shape[1]] = b
    assert not np.isnan(Y).any()
    return Y, [l for l in out.splitlines() if l.startswith("A:")][0]


class RefMatrix:
    """A matrix resident in the unmodified reference (SpParMat over SpDCCols, built once) for repeated Mult_AnXBn_Synch
    calls with fresh panels - bench.py's CPU legs at BASELINE.json's full sizes."""

    def __init__(self, semiring, m, n, I, J, V, x_dtype, colmajor_sorted=False, threads=None):
        R = ref()
        R.cbref_matrix_create.restype = c_void_p
        R.cbref_matrix_create.argtypes = [c_char_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int]
        R.cbref_matrix_mult.restype = c_int
        R.cbref_matrix_mult.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, POINTER(c_double), POINTER(c_int64)]
        R.cbref_matrix_free.restype = None
        R.cbref_matrix_free.argtypes = [c_void_p]
        if threads:
            R.cbref_set_num_threads(int(threads))
        I = np.ascontiguousarray(I, np.int64)
        J = np.ascontiguousarray(J, np.int64)
        if V is None:
            ad, Vb = PATTERN, None
        else:
            Vb = np.ascontiguousarray(V)
            Vb = Vb.view(np.uint8) if Vb.dtype == np.bool_ else Vb
            ad = CODE_OF[Vb.dtype]
        self.m, self.n, self.nnz, self.xdt = m, n, len(I), np.dtype(x_dtype)
        self.h = R.cbref_matrix_create(ref_key(semiring, ad, CODE_OF[self.xdt]).encode(), m, n, len(I), _p(I), _p(J), _p(Vb),
                                       int(colmajor_sorted))
        if not self.h:
            raise ValueError("reference: no such semiring / type combination")

    def mult(self, X, want_y=True):
        """-> (Y or None, seconds of the Mult_AnXBn_Synch call, nnz of the reference's sparse C)"""
        X = np.ascontiguousarray(X, self.xdt)
        Y = np.empty((self.m, X.shape[1]), self.xdt) if want_y else None
        sec, nnzc = c_double(0), c_int64(0)
        rc = ref().cbref_matrix_mult(self.h, X.shape[1], _p(X), _p(Y), ctypes.byref(sec), ctypes.byref(nnzc))
        if rc:
            raise ValueError(f"reference: rc={rc}")
        return Y, sec.value, nnzc.value

    def free(self):
        if self.h:
            ref().cbref_matrix_free(self.h)
            self.h = None


def ref_best_time(semiring, m, n, I, J, V, X, cores, reps=1):
    """Seconds of the fastest configuration of the unmodified reference on `cores` host cores: 1 process x cores OpenMP
    threads (libcbref.so) or a 2x2 process grid x cores/4 threads (cbref_grid) - its MPI+OpenMP design is faster with
    processes.  Returns (seconds, description); used by bench.py's cpu_baseline / --impl reference legs only."""
    best = None
    for _ in range(max(1, reps)):
        _, sec = ref_spmm(semiring, m, n, I, J, V, X, via=0, threads=cores)
        if best is None or sec < best[0]:
            best = (sec, f"1 process x {cores} OpenMP threads")
    if ref_grid_available() and cores >= 4:
        th = max(1, cores // 4)
        try:
            for _ in range(max(1, reps)):
                _, sec, _ = ref_grid_spmm(semiring, 4, m, n, I, J, V, X, via=0, threads=th, check_distribution=False, timeout=240)
                if sec < best[0]:
                    best = (sec, f"2x2 processes x {th} OpenMP threads")
        except Exception as ex:      # the multi-process run is an extra; the 1-process time stands if it cannot run on this box
            best = (best[0], best[1] + f" (2x2-process run unavailable: {type(ex).__name__})")
    return best


def ref_read_mm(path):
    R = ref()
    m, n, nnz = c_int64(), c_int64(), c_int64()
    R.cbref_read_mm(path.encode(), ctypes.byref(m), ctypes.byref(n), ctypes.byref(nnz), None, None, None)
    I = np.empty(nnz.value, np.int64)
    J = np.empty(nnz.value, np.int64)
    V = np.empty(nnz.value, np.float64)
    R.cbref_read_mm(path.encode(), ctypes.byref(m), ctypes.byref(n), ctypes.byref(nnz), _p(I), _p(J), _p(V))
    return m.value, n.value, I, J, V


def ref_spgemm_i64(m, kd, n, AI, AJ, AV, BI, BJ, BV):
    R = ref()
    a = [np.ascontiguousarray(x, np.int64) for x in (AI, AJ, AV, BI, BJ, BV)]
    nnzc = c_int64()
    R.cbref_spgemm_i64(m, kd, n, len(a[0]), _p(a[0]), _p(a[1]), _p(a[2]), len(a[3]), _p(a[3]), _p(a[4]), _p(a[5]),
                       ctypes.byref(nnzc), None, None, None)
    CI = np.empty(nnzc.value, np.int64)
    CJ = np.empty(nnzc.value, np.int64)
    CV = np.empty(nnzc.value, np.int64)
    R.cbref_spgemm_i64(m, kd, n, len(a[0]), _p(a[0]), _p(a[1]), _p(a[2]), len(a[3]), _p(a[3]), _p(a[4]), _p(a[5]),
                       ctypes.byref(nnzc), _p(CI), _p(CJ), _p(CV))
    return CI, CJ, CV


# ----------------------------------------------------------------------------- Matrix Market ingest
def read_mm(path, dup="max"):
    """Matrix Market coordinate file -> (m, n, I, J, V float64), the way ParallelReadMM does it:
    1-based -> 0-based, symmetric files expanded with the transpose of every off-diagonal entry
    (SpHelper.h:75-91), pattern files get value 1 (:168-176), duplicates merged with BinOp
    (SpParMat.cpp:2962-2967; the drivers pass maximum<double>())."""
    with open(path) as f:
        banner = f.readline().lower().split()
        if len(banner) < 5 or banner[0] != "%%matrixmarket" or banner[2] != "coordinate":
            raise ValueError("not a MatrixMarket coordinate file")
        field, sym = banner[3], banner[4]
        line = f.readline()
        while line.startswith("%"):
            line = f.readline()
        m, n, nz = (int(t) for t in line.split()[:3])
        data = np.loadtxt(f, ndmin=2) if nz else np.zeros((0, 3))
    I = data[:, 0].astype(np.int64) - 1
    J = data[:, 1].astype(np.int64) - 1
    V = np.ones(len(I)) if field == "pattern" else data[:, 2].astype(np.float64)
    if sym in ("symmetric", "skew-symmetric", "hermitian"):
        off = I != J
        I, J, V = np.concatenate([I, J[off]]), np.concatenate([J, I[off]]), np.concatenate([V, V[off]])
    I, J, V = dedup(I, J, V, n, dup)
    return m, n, I, J, V


def dedup(I, J, V, n, how="max"):
    key = I * np.int64(n) + J
    order = np.argsort(key, kind="stable")
    key, I, J = key[order], I[order], J[order]
    first = np.ones(len(key), bool)
    first[1:] = key[1:] != key[:-1]
    if V is None:
        return I[first], J[first], None
    V = np.asarray(V)[order]
    starts = np.flatnonzero(first)
    if how == "max":
        Vd = np.maximum.reduceat(V, starts) if len(V) else V
    elif how == "sum":
        Vd = np.add.reduceat(V, starts) if len(V) else V
    else:
        Vd = V[first]
    return I[first], J[first], Vd


# ----------------------------------------------------------------------------- synthetic inputs
# Counter-based generators.  The SAME integer recipes are implemented in CUDA (csrc/cb_gen.cu);
# tests/test_gen_gpu.py checks the two produce identical edge lists and operand values.
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    with np.errstate(over="ignore"):
        x = (np.asarray(x, np.uint64) + np.uint64(0x9E3779B97F4A7C15))
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def _bitrev(v, bits):
    v = v.astype(np.uint64)
    r = np.zeros_like(v)
    for b in range(bits):
        r |= ((v >> np.uint64(b)) & np.uint64(1)) << np.uint64(bits - 1 - b)
    return r


def scramble(v, scale, seed):
    """Bijection on [0, 2^scale): odd multiply, add, bit reversal, odd multiply, add (all mod 2^scale)."""
    mask = np.uint64((1 << scale) - 1)
    s1 = splitmix64(np.uint64(seed) ^ np.uint64(0x5CA1AB1E))
    s2 = splitmix64(s1)
    with np.errstate(over="ignore"):
        v = (v.astype(np.uint64) * (s1 | np.uint64(1)) + (s1 >> np.uint64(32))) & mask
        v = _bitrev(v, scale)
        v = (v * (s2 | np.uint64(1)) + (s2 >> np.uint64(32))) & mask
    return v


def rmat_edges(scale, edgefactor=16, seed=0, initiator=(0.57, 0.19, 0.19, 0.05), do_scramble=True, first=0, count=None):
    """Kronecker/R-MAT edge list: edge e picks one quadrant per level from 16 fresh bits of
    splitmix64(seed*2^40 ^ (e*8 + word)).  Independent of how edges are split over ranks."""
    nedges = edgefactor << scale
    if count is None:
        count = nedges - first
    e = np.arange(first, first + count, dtype=np.uint64)
    a, b, c, _ = initiator
    t1, t2, t3 = (np.uint64(int(round(a * 65536))), np.uint64(int(round((a + b) * 65536))),
                  np.uint64(int(round((a + b + c) * 65536))))
    i = np.zeros(count, np.uint64)
    j = np.zeros(count, np.uint64)
    base = (np.uint64(seed) << np.uint64(40))
    with np.errstate(over="ignore"):
        for lvl in range(scale):
            if lvl % 4 == 0:
                h = splitmix64(base ^ (e * np.uint64(8) + np.uint64(lvl // 4)))
            u = (h >> np.uint64(16 * (lvl % 4))) & np.uint64(0xFFFF)
            ib = (u >= t2).astype(np.uint64)                       # quadrants c,d -> row bit
            jb = (((u >= t1) & (u < t2)) | (u >= t3)).astype(np.uint64)   # quadrants b,d -> col bit
            i = (i << np.uint64(1)) | ib
            j = (j << np.uint64(1)) | jb
    if do_scramble:
        i, j = scramble(i, scale, seed), scramble(j, scale, seed)
    return i.astype(np.int64), j.astype(np.int64)


def rmat_matrix(scale, edgefactor=16, seed=0, initiator=(0.57, 0.19, 0.19, 0.05), symmetric=True, remove_loops=True):
    """GenWriteMatrix recipe: edges -> drop self loops -> (A + A^T) -> merge duplicates. Returns n, I, J (pattern)."""
    n = 1 << scale
    I, J = rmat_edges(scale, edgefactor, seed, initiator)
    if remove_loops:
        keep = I != J
        I, J = I[keep], J[keep]
    if symmetric:
        I, J = np.concatenate([I, J]), np.concatenate([J, I])
    I, J, _ = dedup(I, J, None, n)
    return n, I, J


def _thresholds(initiator):
    a, b, c, _ = initiator
    return np.array([int(round(a * 65536)), int(round((a + b) * 65536)), int(round((a + b + c) * 65536))], np.uint32)


def set_num_threads(n):
    """OpenMP threads of the C restatement (a launcher may have exported OMP_NUM_THREADS=1)"""
    lib().oracle_set_num_threads(int(n))


def rmat_matrix_fast(scale, edgefactor=16, seed=0, initiator=(0.57, 0.19, 0.19, 0.05), symmetric=True, remove_loops=True,
                     col_major=False):
    """rmat_matrix() through the C restatement (gen_oracle.c, OpenMP): the same (n, I, J), in seconds at scale 24.
    col_major=True returns the entries sorted by column then row (what SpDCCols' tuple constructor wants)."""
    thr = _thresholds(initiator)
    nnz, pi, pj = c_int64(), c_void_p(), c_void_p()
    rc = lib().oracle_gen_matrix(scale, edgefactor, seed, _p(thr), int(symmetric), int(remove_loops), int(col_major),
                                 ctypes.byref(nnz), ctypes.byref(pi), ctypes.byref(pj))
    if rc:
        raise MemoryError("oracle_gen_matrix: out of memory")
    try:
        I = np.ctypeslib.as_array(ctypes.cast(pi, POINTER(c_int64)), (max(nnz.value, 1),))[:nnz.value].copy()
        J = np.ctypeslib.as_array(ctypes.cast(pj, POINTER(c_int64)), (max(nnz.value, 1),))[:nnz.value].copy()
    finally:
        lib().oracle_free(pi)
        lib().oracle_free(pj)
    return 1 << scale, I, J


def matrix_values_fast(I, J, n, seed, dtype):
    """matrix_values() through the C restatement"""
    I = np.ascontiguousarray(I, np.int64)
    J = np.ascontiguousarray(J, np.int64)
    out = np.empty(len(I), dtype)
    lib().oracle_matrix_values(_p(I), _p(J), len(I), n, seed, CODE_OF[np.dtype(dtype)], _p(out))
    return out


def dense_columns_fast(n, k, c0, kc, seed, dtype, kind="value"):
    """columns [c0, c0+kc) of dense_operand(n, k, ...) through the C restatement, packed n x kc"""
    out = np.empty((n, kc), dtype)
    lib().oracle_dense_columns(n, k, c0, kc, seed, CODE_OF[np.dtype(dtype)], 1 if kind == "x_minplus" else 0, _p(out))
    return out


def er_matrix(scale, edgefactor=16, seed=0):
    """Erdos-Renyi = the same generator with the uniform initiator, directed, loops removed (SpMSpVBench.cpp:499-501)."""
    return rmat_matrix(scale, edgefactor, seed, (0.25, 0.25, 0.25, 0.25), symmetric=False)


def hash_values(idx, seed, dtype, kind="value"):
    """Operand values from a counter hash.
    float: (2u+1)*2^-(b+1), u = top b bits (b=23 for f32, 52 for f64) -> exact in both languages, in (0,1).
    int MinPlus/PlusTimes: 1 + h % 100; kind='x_minplus' makes ~1% of entries numeric max.
    u8/bool: top bit."""
    with np.errstate(over="ignore"):
        h = splitmix64(np.uint64(seed) * np.uint64(0x100000001B3) ^ np.asarray(idx, np.uint64))
    dt = np.dtype(dtype)
    if dt == np.float32:
        u = (h >> np.uint64(41)).astype(np.float64)
        return ((2.0 * u + 1.0) * 2.0 ** -24).astype(np.float32)
    if dt == np.float64:
        u = (h >> np.uint64(12)).astype(np.float64)
        return (2.0 * u + 1.0) * 2.0 ** -53
    if dt in (np.dtype(np.int32), np.dtype(np.int64)):
        v = (np.uint64(1) + (h >> np.uint64(8)) % np.uint64(100)).astype(dt)
        if kind == "x_minplus":
            v[(h & np.uint64(0xFF)) < np.uint64(3)] = np.iinfo(dt).max
        return v
    return (h >> np.uint64(63)).astype(np.uint8)


def dense_operand(n, k, seed, dtype, kind="value"):
    idx = np.arange(n * k, dtype=np.uint64)
    return hash_values(idx, seed, dtype, kind).reshape(n, k)


def matrix_values(I, J, n, seed, dtype):
    return hash_values(np.asarray(I, np.uint64) * np.uint64(n) + np.asarray(J, np.uint64), seed, dtype)
