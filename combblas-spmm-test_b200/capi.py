"""ctypes binding of the C ABI (include/combblas_b200.h).  Plumbing for tests/ and bench.py only:
the product's host layer is the C++ header set under include/CombBLAS/."""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, byref, c_char_p, c_double, c_float, c_int, c_int64, c_uint64, c_void_p

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
F32, F64, I32, I64, U8, PATTERN = 0, 1, 2, 3, 4, 255
PLUS_TIMES, MIN_PLUS, MAX_SEL2ND, OR_AND = 0, 1, 2, 3
NP_OF = {F32: np.float32, F64: np.float64, I32: np.int32, I64: np.int64, U8: np.uint8}
CODE_OF = {np.dtype(np.float32): F32, np.dtype(np.float64): F64, np.dtype(np.int32): I32, np.dtype(np.int64): I64,
           np.dtype(np.uint8): U8, np.dtype(np.bool_): U8}

# every symbol include/combblas_b200.h declares (tests check the .so exports all of them)
SYMBOLS = ["cb_abi_version", "cb_device_count", "cb_comm_unique_id", "cb_ctx_create", "cb_ctx_create_grid",
           "cb_ctx_destroy", "cb_ctx_grid", "cb_ctx_sync", "cb_ctx_stream", "cb_last_error", "cb_status_string",
           "cb_timer_start", "cb_timer_stop", "cb_tile_upload_csc", "cb_tile_upload_coo", "cb_tile_from_device_coo",
           "cb_tile_free", "cb_tile_info", "cb_tile_pattern_view", "cb_tile_download_csr", "cb_dense_alloc", "cb_dense_wrap", "cb_dense_free",
           "cb_dense_upload", "cb_dense_download", "cb_dense_fill", "cb_dense_info", "cb_semiring_id", "cb_spmm_local",
           "cb_spmm_summa", "cb_summa_times", "cb_summa_plan", "cb_summa_cache_a", "cb_comm_allreduce_i64", "cb_spmm_host", "cb_launch_count", "cb_profile_enable", "cb_profile_read", "cb_gen_rmat_tile", "cb_gen_dense",
           "cb_spmm_hub_config", "cb_spmm_hub_info", "cb_hub_select_host", "cb_spmm_ring_config", "cb_tile_filter_columns", "cb_spmm_k2_config", "cb_tile_row_lengths", "cb_tile_download_rows", "cb_dense_download_rows", "cb_spmm_summa_host", "cb_spmm_k2_pipe", "cb_spmm_k2_l2", "cb_spgemm_local", "cb_spgemm_summa", "cb_coo_info", "cb_coo_download", "cb_coo_free", "cb_spmv_grid", "cb_gen_graph500_edges", "cb_gen_graph500_tile", "cb_tile_from_distributed_coo", "cb_tile_from_mm_text"]


class CBError(RuntimeError):
    def __init__(self, status, text):
        super().__init__(f"combblas_b200 status {status}: {text}")
        self.status = status


def library_path() -> str:
    # CB_LIB lets tuning experiments point at a variant build; the product path is lib/libcombblas_b200.so
    return os.environ.get("CB_LIB") or os.path.join(HERE, "lib", "libcombblas_b200.so")


def build_library(jobs: int = 8) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... (csrc/Makefile), in-tree."""
    subprocess.check_call(["make", "-s", f"-j{jobs}", "-C", os.path.join(HERE, "csrc")])
    return library_path()


_lib = None


def lib():
    """Loads the CUDA library.  Missing library = hard error: there is no other implementation."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise CBError(-1, f"{path} is missing: build it with build_library() / __graft_entry__.build(); "
                              "the product has no CPU fallback")
        L = ctypes.CDLL(path)
        L.cb_last_error.restype = c_char_p
        L.cb_last_error.argtypes = [c_void_p]
        L.cb_status_string.restype = c_char_p
        L.cb_ctx_stream.restype = c_void_p
        L.cb_ctx_stream.argtypes = [c_void_p]
        L.cb_launch_count.restype = c_int64
        L.cb_launch_count.argtypes = [c_void_p]
        L.cb_ctx_create.argtypes = [c_int, POINTER(c_void_p)]
        L.cb_ctx_create_grid.argtypes = [c_int, c_int, c_int, c_int, c_int, c_void_p, POINTER(c_void_p)]
        L.cb_ctx_destroy.argtypes = [c_void_p]
        L.cb_ctx_sync.argtypes = [c_void_p]
        L.cb_ctx_grid.argtypes = [c_void_p] + [POINTER(c_int)] * 5
        L.cb_comm_unique_id.argtypes = [c_void_p]
        L.cb_timer_start.argtypes = [c_void_p]
        L.cb_timer_stop.argtypes = [c_void_p, POINTER(c_float)]
        L.cb_tile_upload_csc.argtypes = [c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_int, c_int, POINTER(c_void_p)]
        L.cb_tile_upload_coo.argtypes = [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                         POINTER(c_void_p)]
        L.cb_tile_from_device_coo.argtypes = [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int,
                                              POINTER(c_void_p)]
        L.cb_tile_free.argtypes = [c_void_p]
        L.cb_tile_info.argtypes = [c_void_p, POINTER(c_int64)]
        L.cb_tile_pattern_view.argtypes = [c_void_p, POINTER(c_void_p)]
        L.cb_tile_download_csr.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p]
        L.cb_dense_alloc.argtypes = [c_void_p, c_int64, c_int64, c_int, POINTER(c_void_p)]
        L.cb_dense_wrap.argtypes = [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, POINTER(c_void_p)]
        L.cb_dense_free.argtypes = [c_void_p]
        L.cb_dense_upload.argtypes = [c_void_p, c_void_p, c_int64]
        L.cb_dense_download.argtypes = [c_void_p, c_void_p, c_int64]
        L.cb_dense_fill.argtypes = [c_void_p, c_void_p]
        L.cb_dense_info.argtypes = [c_void_p, POINTER(c_int64), POINTER(c_int64), POINTER(c_int64), POINTER(c_int),
                                    POINTER(c_void_p)]
        L.cb_semiring_id.argtypes = [c_int, c_int, c_void_p]
        L.cb_spmm_local.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int]
        L.cb_spmm_summa.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int64, c_int64]
        L.cb_summa_times.argtypes = [c_void_p, POINTER(c_float)]
        L.cb_summa_plan.argtypes = [c_int, c_int, c_int64, POINTER(c_int64), POINTER(c_int), POINTER(c_int), POINTER(c_int)]
        L.cb_summa_cache_a.argtypes = [c_void_p, c_int]
        L.cb_comm_allreduce_i64.argtypes = [c_void_p, c_int, c_int, POINTER(c_int64), c_int]
        L.cb_spmm_host.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int]
        L.cb_profile_enable.argtypes = [c_void_p, c_int]
        L.cb_profile_read.argtypes = [c_void_p, POINTER(c_double), POINTER(c_int64)]
        L.cb_gen_rmat_tile.argtypes = [c_void_p, c_int, c_int, c_uint64, POINTER(c_double), c_int, c_int64, c_int64,
                                       c_int64, c_int64, c_int, c_uint64, POINTER(c_void_p)]
        L.cb_gen_dense.argtypes = [c_void_p, c_uint64, c_int64, c_int64, c_int64, c_int]
        L.cb_spmm_hub_config.argtypes = [c_void_p, c_int, c_int, c_int]
        L.cb_spmm_hub_info.argtypes = [c_void_p, POINTER(c_int64)]
        L.cb_spmm_ring_config.argtypes = [c_void_p, c_int]
        L.cb_spmm_k2_config.argtypes = [c_void_p, c_int, c_int]
        L.cb_spmm_k2_pipe.argtypes = [c_void_p, c_int]
        L.cb_spmm_k2_l2.argtypes = [c_void_p, c_int]
        L.cb_spgemm_local.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, POINTER(c_void_p)]
        L.cb_spgemm_summa.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int64, c_int64, c_int64, POINTER(c_void_p)]
        L.cb_coo_info.argtypes = [c_void_p, POINTER(c_int64), POINTER(c_int64), POINTER(c_int64), POINTER(c_int)]
        L.cb_coo_download.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p]
        L.cb_coo_free.argtypes = [c_void_p]
        L.cb_tile_from_distributed_coo.argtypes = [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_int, POINTER(c_void_p)]
        L.cb_tile_from_mm_text.argtypes = [c_void_p, c_int64, c_int64, ctypes.c_char_p, c_int64, c_int, c_int, c_int, POINTER(c_void_p)]
        L.cb_spmv_grid.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_int64, c_int64]
        L.cb_gen_graph500_edges.argtypes = [c_void_p, c_int, c_uint64, c_uint64, c_int64, c_int64, c_void_p, c_void_p]
        L.cb_gen_graph500_tile.argtypes = [c_void_p, c_int, c_int, c_uint64, c_uint64, c_int, c_int, c_int64, c_int64, c_int64, c_int64, c_int,
                                           c_uint64, POINTER(c_void_p)]
        L.cb_tile_row_lengths.argtypes = [c_void_p, c_void_p]
        L.cb_tile_download_rows.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]
        L.cb_dense_download_rows.argtypes = [c_void_p, c_int64, c_void_p, c_void_p]
        L.cb_spmm_summa_host.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int64, c_int64, c_int64, c_int]
        L.cb_tile_filter_columns.argtypes = [c_void_p, c_void_p, c_void_p, POINTER(c_void_p)]
        L.cb_hub_select_host.argtypes = [c_void_p, c_int64, c_int, c_void_p, c_void_p]
        _lib = L
    return _lib


def _check(status, ctx=None):
    if status != 0:
        text = lib().cb_last_error(ctx) or b""
        raise CBError(status, text.decode() or lib().cb_status_string(status).decode())


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(c_void_p)
    return c_void_p(int(a))


def summa_plan(pr, pc, gn):
    """-> (seg boundaries [ns+1], a_owner_col [ns], x_owner_row [ns]); host arithmetic only."""
    seg = (c_int64 * (pr + pc + 1))()
    ac, xr, ns = (c_int * (pr + pc))(), (c_int * (pr + pc))(), c_int()
    _check(lib().cb_summa_plan(pr, pc, gn, seg, ac, xr, byref(ns)))
    return list(seg[:ns.value + 1]), list(ac[:ns.value]), list(xr[:ns.value])


def hub_select(counts, max_hubs):
    """The hub selection rule of the hub variant (cb_hub.cu): -> (hub columns by rank, cumulative nonzero counts); host arithmetic only."""
    counts = np.ascontiguousarray(counts, np.int32)
    cols = np.empty(max(max_hubs, 1), np.int32)
    cum = np.empty(max(max_hubs, 1), np.int64)
    n = lib().cb_hub_select_host(_ptr(counts), len(counts), max_hubs, _ptr(cols), _ptr(cum))
    if n < 0:
        raise CBError(n, "cb_hub_select_host: bad arguments")
    return cols[:n].copy(), cum[:n].copy()


def block_range(total, nb, b):
    """Owner rule of the reference (SpParMat.cpp:5066-5096): floor division, last block takes the remainder."""
    per = total // nb
    start = b * per
    return start, (total - start if b == nb - 1 else per)


def unique_id() -> bytes:
    buf = ctypes.create_string_buffer(128)
    _check(lib().cb_comm_unique_id(buf))
    return buf.raw


class Context:
    """One device + its place in the pr x pc process grid (CommGrid, reference CommGrid.h:44-166)."""

    def __init__(self, device=0, rank=0, nranks=1, pr=1, pc=1, uid: bytes | None = None):
        self.h = c_void_p()
        if nranks == 1:
            _check(lib().cb_ctx_create(device, byref(self.h)))
        else:
            buf = ctypes.create_string_buffer(uid, 128)
            _check(lib().cb_ctx_create_grid(device, rank, nranks, pr, pc, buf, byref(self.h)))
        self.rank, self.nranks, self.pr, self.pc = rank, nranks, pr, pc
        self.myprocrow, self.myproccol = rank // pc, rank % pc

    def close(self):
        if self.h:
            lib().cb_ctx_destroy(self.h)
            self.h = c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def sync(self):
        _check(lib().cb_ctx_sync(self.h), self.h)

    @property
    def stream(self):
        return lib().cb_ctx_stream(self.h)

    @property
    def launches(self):
        return lib().cb_launch_count(self.h)

    def profile(self, on=True):
        _check(lib().cb_profile_enable(self.h, int(on)), self.h)

    def profile_read(self):
        """-> ({'fill','spmm','fixup'} -> summed ms, same -> launches) since profile(True)."""
        ms, n = (c_double * 3)(), (c_int64 * 3)()
        _check(lib().cb_profile_read(self.h, ms, n), self.h)
        names = ("fill", "spmm", "fixup")
        return dict(zip(names, ms)), dict(zip(names, n))

    def timer_start(self):
        _check(lib().cb_timer_start(self.h), self.h)

    def timer_stop(self) -> float:
        ms = c_float()
        _check(lib().cb_timer_stop(self.h, byref(ms)), self.h)
        return ms.value

    # ---- tiles
    def tile_from_csc(self, m, n, cp, ir, numx=None, jc=None, pattern=None):
        """Column-compressed arrays as the reference ships them (SpDCCols::GetArrays): jc=None means plain CSC."""
        idx = np.int64 if np.asarray(cp).dtype == np.int64 else np.int32
        cp = np.ascontiguousarray(cp, idx)
        ir = np.ascontiguousarray(ir, idx)
        jc_ = None if jc is None else np.ascontiguousarray(jc, idx)
        vd, vals = _vals(numx)
        t = c_void_p()
        _check(lib().cb_tile_upload_csc(self.h, m, n, len(ir), 0 if jc_ is None else len(jc_), _ptr(cp), _ptr(jc_),
                                        _ptr(ir), _ptr(vals), CODE_OF[np.dtype(idx)], vd, byref(t)), self.h)
        return Tile(self, t)

    def tile_from_coo(self, m, n, rows, cols, vals=None):
        idx = np.int64 if np.asarray(rows).dtype == np.int64 else np.int32
        rows = np.ascontiguousarray(rows, idx)
        cols = np.ascontiguousarray(cols, idx)
        vd, v = _vals(vals)
        t = c_void_p()
        _check(lib().cb_tile_upload_coo(self.h, m, n, len(rows), _ptr(rows), _ptr(cols), _ptr(v), CODE_OF[np.dtype(idx)],
                                        vd, byref(t)), self.h)
        return Tile(self, t)

    def tile_from_distributed_coo(self, gm, gn, rows, cols, vals=None, dup_op=2):
        """collective: every rank hands in triples with GLOBAL coordinates (any owner); they are routed to their owners between the
        GPUs, duplicates merged (0 first, 1 sum, 2 max, 3 min) and this rank's tile built on the device"""
        rows = np.ascontiguousarray(rows, np.int64)
        cols = np.ascontiguousarray(cols, np.int64)
        vd, v = _vals(vals)
        t = c_void_p()
        _check(lib().cb_tile_from_distributed_coo(self.h, gm, gn, len(rows), _ptr(rows), _ptr(cols), _ptr(v), vd, dup_op, byref(t)), self.h)
        return Tile(self, t)

    def tile_from_mm_text(self, gm, gn, text: bytes, onebased=True, pattern=False, symmetric=False, val_dtype=PATTERN, dup_op=2):
        """collective: this rank's share of the data lines of a Matrix Market file, parsed, routed and merged on the GPUs"""
        flags = (1 if onebased else 0) | (2 if pattern else 0) | (4 if symmetric else 0)
        t = c_void_p()
        _check(lib().cb_tile_from_mm_text(self.h, gm, gn, text, len(text), flags, val_dtype, dup_op, byref(t)), self.h)
        return Tile(self, t)

    def tile_from_device_coo(self, m, n, nz, d_rows, d_cols, d_vals=None, val_dtype=PATTERN):
        t = c_void_p()
        _check(lib().cb_tile_from_device_coo(self.h, m, n, nz, _ptr(d_rows), _ptr(d_cols), _ptr(d_vals), val_dtype,
                                             byref(t)), self.h)
        return Tile(self, t)

    def gen_rmat_tile(self, scale, edgefactor=16, seed=0, initiator=(0.57, 0.19, 0.19, 0.05), symmetric=True,
                      row0=0, m=None, col0=0, n=None, val_dtype=PATTERN, val_seed=1):
        N = 1 << scale
        m = N if m is None else m
        n = N if n is None else n
        init = (c_double * 4)(*initiator)
        t = c_void_p()
        _check(lib().cb_gen_rmat_tile(self.h, scale, edgefactor, seed, init, int(symmetric), row0, m, col0, n, val_dtype,
                                      val_seed, byref(t)), self.h)
        return Tile(self, t)

    def graph500_edges(self, log_numverts, first, count, userseed=(0, 0)):
        """edges [first, first+count) of the reference's packed Graph500 stream (RefGen21) -> (src, dst)"""
        src, dst = np.empty(count, np.int64), np.empty(count, np.int64)
        _check(lib().cb_gen_graph500_edges(self.h, log_numverts, userseed[0], userseed[1], first, count, _ptr(src), _ptr(dst)), self.h)
        return src, dst

    def gen_graph500_tile(self, scale, edgefactor=16, symmetric=True, remove_loops=True, row0=0, m=None, col0=0, n=None, val_dtype=PATTERN,
                          val_seed=0, userseed=(0, 0)):
        """the block of the matrix ReleaseTests/GenWriteMatrix.cpp builds (val_seed 0: values are edge multiplicities)"""
        N = 1 << scale
        m = N if m is None else m
        n = N if n is None else n
        t = c_void_p()
        _check(lib().cb_gen_graph500_tile(self.h, scale, edgefactor, userseed[0], userseed[1], int(symmetric), int(remove_loops), row0, m, col0, n,
                                          val_dtype, val_seed, byref(t)), self.h)
        return Tile(self, t)

    def spmv_grid(self, tile, x_piece, x_off, y_off, y_len, semiring, gm, gn):
        """y = A (x).(+) x with the vector exchange on the device; collective.  -> this rank's piece of y"""
        x_piece = np.ascontiguousarray(x_piece)
        y = np.empty(y_len, x_piece.dtype)
        _check(lib().cb_spmv_grid(self.h, tile.h, _ptr(x_piece), x_off, len(x_piece), _ptr(y), y_off, y_len, semiring, CODE_OF[x_piece.dtype], gm, gn), self.h)
        return y

    # ---- dense panels
    def dense(self, rows, cols, dtype):
        d = c_void_p()
        code = dtype if isinstance(dtype, int) else CODE_OF[np.dtype(dtype)]
        _check(lib().cb_dense_alloc(self.h, rows, cols, code, byref(d)), self.h)
        return Dense(self, d, rows, cols, code)

    def dense_from(self, host: np.ndarray):
        host = np.ascontiguousarray(host)
        if host.dtype == np.bool_:
            host = host.view(np.uint8)
        d = self.dense(host.shape[0], host.shape[1], host.dtype)
        d.upload(host)
        return d

    def wrap(self, device_ptr, rows, cols, ld, dtype):
        d = c_void_p()
        code = dtype if isinstance(dtype, int) else CODE_OF[np.dtype(dtype)]
        _check(lib().cb_dense_wrap(self.h, c_void_p(device_ptr), rows, cols, ld, code, byref(d)), self.h)
        return Dense(self, d, rows, cols, code)

    # ---- multiply
    def spmm_local(self, tile, X, Y, semiring, accumulate=False):
        _check(lib().cb_spmm_local(self.h, tile.h, X.h, Y.h, semiring, int(accumulate)), self.h)

    def spmm_summa(self, tile, X, Y, semiring, gm, gn, gk):
        _check(lib().cb_spmm_summa(self.h, tile.h, X.h, Y.h, semiring, gm, gn, gk), self.h)

    def _coo_out(self, c):
        nnz, m, k, dt = c_int64(), c_int64(), c_int64(), c_int()
        _check(lib().cb_coo_info(c, byref(nnz), byref(m), byref(k), byref(dt)), self.h)
        rows, cols = np.empty(nnz.value, np.int64), np.empty(nnz.value, np.int64)
        vals = np.empty(nnz.value, NP_OF[dt.value])
        _check(lib().cb_coo_download(c, _ptr(rows), _ptr(cols), _ptr(vals)), self.h)
        lib().cb_coo_free(c)
        return rows, cols, vals

    def spgemm_local(self, A, B, semiring, dtype):
        """C = A (x).(+) B for a sparse B on this GPU -> (rows, cols, vals), column-major sorted, merged."""
        c = c_void_p()
        _check(lib().cb_spgemm_local(self.h, A.h, B.h, semiring, CODE_OF[np.dtype(dtype)], byref(c)), self.h)
        return self._coo_out(c)

    def spgemm_summa(self, A, B, semiring, dtype, gm, gn, gk):
        """the same over the process grid (collective): this rank's block of C with local indices"""
        c = c_void_p()
        _check(lib().cb_spgemm_summa(self.h, A.h, B.h, semiring, CODE_OF[np.dtype(dtype)], gm, gn, gk, byref(c)), self.h)
        return self._coo_out(c)

    def spmm_summa_host(self, tile, X: np.ndarray, Y: np.ndarray, semiring, gm, gn, gk):
        """This rank's X tile in host memory in, its Y tile in host memory out; collective over the grid."""
        assert X.flags.c_contiguous and Y.flags.c_contiguous and X.dtype == Y.dtype
        _check(lib().cb_spmm_summa_host(self.h, tile.h, _ptr(X), X.shape[1] if X.ndim == 2 else 0, _ptr(Y), Y.shape[1] if Y.ndim == 2 else 0,
                                        semiring, gm, gn, gk, CODE_OF[np.dtype(np.uint8) if X.dtype == np.bool_ else X.dtype]), self.h)
        return Y

    def hub_config(self, enable, cluster=0, slab_bytes=0):
        """Opt into the hub variant of the local multiply (K2H); enable=-1 follows CB_SPMM_HUB."""
        _check(lib().cb_spmm_hub_config(self.h, int(enable), int(cluster), int(slab_bytes)), self.h)

    def ring_config(self, depth):
        """Opt into the ring-pipelined variant of the local multiply (K2R): depth 8 on, 0 off, -1 follows CB_SPMM_RING."""
        _check(lib().cb_spmm_ring_config(self.h, int(depth)), self.h)

    def k2_config(self, slab_bytes=0, point=-1):
        """Shape of the default local multiply: column-slab width in bytes (0 automatic) and operating point (-1 automatic)."""
        _check(lib().cb_spmm_k2_config(self.h, int(slab_bytes), int(point)), self.h)

    def k2_l2(self, budget_mb=-1):
        """L2 residency hints of K2P: MB of most-used X rows gathered with evict_last (0 off, -1 default)."""
        _check(lib().cb_spmm_k2_l2(self.h, int(budget_mb)), self.h)

    def k2_pipe(self, depth=-1):
        """Variant of the local multiply: 0 K2, 1 K2 + entry prefetch, 16 K2T (bulk-copy ring), 32 K2W (hub panel under a persisting L2 window, budget k2_l2), 64 K2 with persistent warps, 4 / 8 K2P ring depth, -1 default."""
        _check(lib().cb_spmm_k2_pipe(self.h, int(depth)), self.h)

    def summa_cache_a(self, on=True):
        _check(lib().cb_summa_cache_a(self.h, int(on)), self.h)

    def summa_times(self):
        ms = (c_float * 4)()
        _check(lib().cb_summa_times(self.h, ms), self.h)
        return list(ms)

    def spmm_host(self, tile, X: np.ndarray, semiring, Y: np.ndarray | None = None):
        """Host panels in, host panel out (upload, multiply, download): the e2e path."""
        Xc = np.ascontiguousarray(X)
        if Xc.dtype == np.bool_:
            Xc = Xc.view(np.uint8)
        k = Xc.shape[1]
        if Y is None:
            Y = np.empty((tile.m, k), Xc.dtype)
        _check(lib().cb_spmm_host(self.h, tile.h, _ptr(Xc), k, _ptr(Y), k, k, CODE_OF[Xc.dtype], semiring), self.h)
        return Y


def _vals(v):
    if v is None:
        return PATTERN, None
    v = np.ascontiguousarray(v)
    if v.dtype == np.bool_:
        v = v.view(np.uint8)
    return CODE_OF[v.dtype], v


class Tile:
    def __init__(self, ctx, h):
        self.ctx, self.h = ctx, h
        info = (c_int64 * 8)()
        _check(lib().cb_tile_info(h, info))
        (self.nnz, self.m, self.n, self.nzr, self.nzc, self.nchunks, self.nsplit, self.bytes) = list(info)

    def free(self):
        if self.h:
            lib().cb_tile_free(self.h)
            self.h = c_void_p()

    def filter_columns(self, keep_cols):
        """new tile with only the nonzeros whose column is marked in keep_cols (uint8 per column)"""
        keep = np.ascontiguousarray(keep_cols, np.uint8)
        assert keep.shape == (self.n,)
        v = c_void_p()
        _check(lib().cb_tile_filter_columns(self.ctx.h, self.h, _ptr(keep), byref(v)), self.ctx.h)
        return Tile(self.ctx, v)

    def hub_info(self):
        """{built, hub columns known, hub rows resident in the last multiply, share of nonzeros they serve}"""
        info = (c_int64 * 4)()
        _check(lib().cb_spmm_hub_info(self.h, info))
        return {"built": bool(info[0]), "known": info[1], "resident": info[2], "cover": info[3] / 1e6}

    def pattern_view(self):
        v = c_void_p()
        _check(lib().cb_tile_pattern_view(self.h, byref(v)), self.ctx.h)
        return Tile(self.ctx, v)

    def row_lengths(self):
        """number of nonzeros of every row"""
        out = np.empty(self.m, np.int64)
        _check(lib().cb_tile_row_lengths(self.h, _ptr(out)), self.ctx.h)
        return out

    def rows(self, rows, lengths, val_dtype=None):
        """-> (offsets [len(rows)+1], columns, values or None) of the selected rows; lengths = row_lengths()"""
        rows = np.ascontiguousarray(rows, np.int64)
        off = np.concatenate([[0], np.cumsum(lengths[rows])]).astype(np.int64)
        cols = np.empty(max(int(off[-1]), 1), np.int64)
        vals = None if val_dtype is None else np.empty(max(int(off[-1]), 1), val_dtype)
        _check(lib().cb_tile_download_rows(self.h, len(rows), _ptr(rows), _ptr(cols), _ptr(vals)), self.ctx.h)
        return off, cols[:off[-1]], (None if vals is None else vals[:off[-1]])

    def to_csr(self, val_dtype=None):
        rowptr = np.empty(self.m + 1, np.int64)
        col = np.empty(self.nnz, np.int64)
        vals = None if val_dtype is None else np.empty(self.nnz, val_dtype)
        _check(lib().cb_tile_download_csr(self.h, _ptr(rowptr), _ptr(col), _ptr(vals)), self.ctx.h)
        return rowptr, col, vals


class Dense:
    def __init__(self, ctx, h, rows, cols, code):
        self.ctx, self.h, self.rows, self.cols, self.code = ctx, h, rows, cols, code

    def free(self):
        if self.h:
            lib().cb_dense_free(self.h)
            self.h = c_void_p()

    def info(self):
        r, c, ld, dt, p = c_int64(), c_int64(), c_int64(), c_int(), c_void_p()
        lib().cb_dense_info(self.h, byref(r), byref(c), byref(ld), byref(dt), byref(p))
        return r.value, c.value, ld.value, dt.value, p.value

    def upload(self, host: np.ndarray, ld=None):
        host = np.ascontiguousarray(host) if ld is None else host
        _check(lib().cb_dense_upload(self.h, _ptr(host), host.shape[1] if ld is None else ld), self.ctx.h)

    def download(self, out: np.ndarray | None = None):
        if out is None:
            out = np.empty((self.rows, self.cols), NP_OF[self.code])
        _check(lib().cb_dense_download(self.h, _ptr(out), out.shape[1] if out.ndim == 2 else self.cols), self.ctx.h)
        return out

    def download_rows(self, rows):
        """selected rows of the panel, packed"""
        rows = np.ascontiguousarray(rows, np.int64)
        out = np.empty((len(rows), self.cols), NP_OF[self.code])
        _check(lib().cb_dense_download_rows(self.h, len(rows), _ptr(rows), _ptr(out)), self.ctx.h)
        return out

    def fill(self, value):
        v = np.array([value], NP_OF[self.code])
        _check(lib().cb_dense_fill(self.h, _ptr(v)), self.ctx.h)

    def generate(self, seed, row0=0, col0=0, gk=None, kind=0):
        _check(lib().cb_gen_dense(self.h, seed, row0, col0, self.cols if gk is None else gk, kind), self.ctx.h)
