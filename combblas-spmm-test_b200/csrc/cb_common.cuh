// Internal definitions shared by the translation units of libcombblas_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "combblas_b200.h"

#define CB_WARP 32
#define CB_MAX_SLABS 8

struct cb_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t compute = nullptr;   // local kernels, panel uploads
    cudaStream_t comm = nullptr;      // NCCL broadcasts of SUMMA stages
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    // process grid (CommGrid.h:106-110): rank = myprocrow * pc + myproccol
    int rank = 0, nranks = 1, pr = 1, pc = 1, myprocrow = 0, myproccol = 0;
    void* nccl_world = nullptr;       // ncclComm_t
    void* nccl_row = nullptr;         // ranks with the same myprocrow ("RowWorld", src/CommGrid.cpp:66)
    void* nccl_col = nullptr;         // ranks with the same myproccol ("ColWorld", src/CommGrid.cpp:67)
    int64_t launches = 0;
    bool summa_cache_a = true;        // keep received A parts on the device between multiplies (cb_summa_cache_a)
    bool summa_merge = true;          // fuse the resident parts of a block-row per X owner (CB_SUMMA_MERGE=0 disables)
    bool summa_p2p = true;            // dense panels travel by copy-engine pushes into peer memory (cb_p2p.cu), not NCCL
    void* p2p_state = nullptr;
    // optional per-kernel device timing (bench.py's roofline): event pairs around K3 / K2 / fix-up launches
    bool profiling = false;
    std::vector<cudaEvent_t> prof_events[3];     // [kind] -> begin,end,begin,end,...
    std::vector<cudaEvent_t> prof_pool;
    // device panels reused by cb_spmm_host (host operands): grown, never shrunk
    void* ws_x = nullptr; size_t ws_x_bytes = 0;
    void* ws_y = nullptr; size_t ws_y_bytes = 0;
    cudaStream_t h2d = nullptr, d2h = nullptr;               // column slabs of host panels flow up / down on these
    cudaEvent_t slab_up[8] = {nullptr}, slab_done[8] = {nullptr}, host_begin = nullptr;
    void* summa_state = nullptr;      // receive buffers + events, owned by cb_summa.cu
    // hub variant of K2 (cb_hub.cu): -1 / 0 = follow the CB_SPMM_HUB* environment, otherwise set by cb_spmm_hub_config
    int hub_enable = -1, hub_cluster = 0, hub_slab_bytes = 0;
    int ring_depth = -1;              // K2R ring depth: -1 = follow CB_SPMM_RING, 0 = off (cb_spmm_ring_config)
    int k2_l2_mb = -1;                // K2P L2 residency hints: budget in MB for the rows kept with evict_last; 0 off, -1 default
    int k2_pipe = -1;                 // variant of the local multiply: -1 default, 0 K2, 1 prefetch, 4 / 8 K2P ring depth, 16 K2T, 32 K2W
    unsigned* k2_counter = nullptr;   // chunk counter of the persistent-warp form of K2 (cb_spmm_k2_pipe 64)
    void* win_panel = nullptr;        // K2W: packed rows of the most used columns, kept in L2 by a persisting access-policy window
    size_t win_panel_bytes = 0;
    size_t win_l2_limit = 0;          // persisting L2 set-aside currently configured on the device
    int k2_slab_bytes = 0, k2_point = -1;   // plain K2: column-slab width (0 = automatic) and operating point (cb_spmm_k2_config)
    std::string err;
};

// Everything a receiver needs to lay out a tile that arrives as one message (the "essentials" of
// SpDCCols::GetEssentials, SpDCCols.cpp:787-795, extended by the work partition).
struct cb_tile_meta {
    int64_t m, n, nnz, nzr, nzc, nchunks, nsplit;
    int32_t chunk_len, val_dtype;
    int32_t layout_dtype;     // value type the slab was laid out for (differs from val_dtype for a pattern view)
    int32_t reserved;
};
struct cb_tile_layout { size_t colflag, vals, nzrows, rowptr, emptyrows, chunk_start, chunk_row, split_row, total; };
static inline cb_tile_layout cb_layout(const cb_tile_meta& t) {
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    size_t vs = 0;
    switch (t.layout_dtype) { case CB_F32: case CB_I32: vs = 4; break; case CB_F64: case CB_I64: vs = 8; break; case CB_U8: vs = 1; break; default: vs = 0; }
    cb_tile_layout L;
    size_t o = 0;
    L.colflag = o;     o = up(o + (size_t)t.nnz * 4);
    L.vals = o;        o = up(o + (size_t)t.nnz * vs);
    L.nzrows = o;      o = up(o + (size_t)t.nzr * 4);
    L.rowptr = o;      o = up(o + (size_t)(t.nzr + 1) * 4);
    L.emptyrows = o;   o = up(o + (size_t)(t.m - t.nzr) * 4);
    L.chunk_start = o; o = up(o + (size_t)(t.nchunks + 1) * 4);
    L.chunk_row = o;   o = up(o + (size_t)t.nchunks * 4);
    L.split_row = o;   o = up(o + (size_t)t.nsplit * 4);
    L.total = o ? o : 256;
    return L;
}

// Device tile: doubly compressed rows (the row-major mirror of Dcsc, dcsc.h:124-131) plus a work partition.
struct cb_hub;
struct cb_tile {
    cb_ctx* ctx = nullptr;
    uint64_t uid = 0;         // unique per built tile (0 for views onto received buffers)
    int64_t m = 0, n = 0, nnz = 0;
    int64_t nzr = 0;          // rows with at least one nonzero
    int64_t nzc = 0;          // columns with at least one nonzero (n_used of the traffic model)
    int val_dtype = CB_PATTERN;
    int layout_dtype = CB_PATTERN;   // value type the slab is laid out for (== val_dtype except for pattern views)
    char* slab = nullptr;     // the one allocation holding every array below (cb_tile_layout)
    size_t slab_bytes = 0;
    bool owns_slab = false;
    // per nonzero, row-major, ascending column inside a row
    int32_t* colflag = nullptr;   // column index | (last nonzero of its row ? 0x80000000 : 0)
    void* vals = nullptr;         // nnz values of val_dtype, or NULL for CB_PATTERN
    // per nonempty row
    int32_t* nzrows = nullptr;    // [nzr] row ids, ascending
    int32_t* rowptr = nullptr;    // [nzr+1] offsets into colflag/vals
    int32_t* emptyrows = nullptr; // [m - nzr] ids of rows without nonzeros (receive SR::id())
    // work partition: chunk c covers nonzeros [chunk_start[c], chunk_start[c+1]); chunk_row[c] is the index into
    // nzrows of the row owning its first nonzero, top bit set when that row began in an earlier chunk
    int64_t nchunks = 0;
    int32_t chunk_len = 0;        // nominal L
    int32_t* chunk_start = nullptr;  // [nchunks+1]
    int32_t* chunk_row = nullptr;    // [nchunks]
    // rows longer than L are cut at multiples of L; their pieces are combined by the fix-up kernel
    int64_t nsplit = 0;
    int32_t* split_row = nullptr;    // [nsplit] index into nzrows
    // scratch for partial rows, sized at first use for the widest panel seen: 2 slots (head, tail) per chunk
    void* carry = nullptr;
    size_t carry_bytes = 0;
    // column slices of this tile, one per SUMMA stage it roots (built at the first cb_spmm_summa, cb_summa.cu)
    std::vector<cb_tile*> summa_parts;
    int64_t summa_key[3] = {-1, -1, -1};   // (grid id, gn, stages) the parts were cut for
    // parts of the other ranks of my processor row, kept after the first cb_spmm_summa with this tile so that later
    // multiplies with the same (immutable) matrix move only the dense panels (cb_summa_cache_a)
    std::vector<cb_tile*> summa_remote;
    // my whole block-row of A regrouped by the X row block it multiplies (one tile per processor row), built from the
    // own + cached parts at the second multiply; entries stay NULL where a single existing part already is that tile
    std::vector<cb_tile*> summa_merged;
    std::vector<cb_tile*> spgemm_remote;   // sparse x sparse stage loop: A parts of my row neighbours kept after the first product (cb_summa_cache_a)
    std::vector<char> spgemm_sent;         // ... and, for the parts I root, whether the neighbours already hold them
    // most frequent columns and the per-nonzero hub ranks, built at the first hub multiply (cb_hub.cu); owned tiles only
    cb_hub* hub = nullptr;
};

static inline cb_tile_meta cb_tile_get_meta(const cb_tile* t) {
    cb_tile_meta m;
    memset(&m, 0, sizeof m);
    m.m = t->m; m.n = t->n; m.nnz = t->nnz; m.nzr = t->nzr; m.nzc = t->nzc; m.nchunks = t->nchunks; m.nsplit = t->nsplit;
    m.chunk_len = t->chunk_len; m.val_dtype = t->val_dtype; m.layout_dtype = t->layout_dtype;
    return m;
}
static inline void cb_tile_bind(cb_tile* t, const cb_tile_meta& m, char* slab) {
    const cb_tile_layout L = cb_layout(m);
    t->m = m.m; t->n = m.n; t->nnz = m.nnz; t->nzr = m.nzr; t->nzc = m.nzc; t->nchunks = m.nchunks; t->nsplit = m.nsplit;
    t->chunk_len = m.chunk_len; t->val_dtype = m.val_dtype; t->layout_dtype = m.layout_dtype;
    t->slab = slab;
    t->colflag = (int32_t*)(slab + L.colflag);
    t->vals = (m.val_dtype == CB_PATTERN || m.nnz == 0) ? nullptr : (void*)(slab + L.vals);
    t->nzrows = (int32_t*)(slab + L.nzrows);
    t->rowptr = (int32_t*)(slab + L.rowptr);
    t->emptyrows = (int32_t*)(slab + L.emptyrows);
    t->chunk_start = (int32_t*)(slab + L.chunk_start);
    t->chunk_row = (int32_t*)(slab + L.chunk_row);
    t->split_row = (int32_t*)(slab + L.split_row);
}

struct cb_dense {
    cb_ctx* ctx = nullptr;
    int64_t rows = 0, cols = 0, ld = 0;   // ld in elements
    int dtype = CB_F32;
    void* ptr = nullptr;
    bool owned = true;
};

static inline size_t cb_dtype_size(int dt) {
    switch (dt) {
        case CB_F32: case CB_I32: return 4;
        case CB_F64: case CB_I64: return 8;
        case CB_U8: return 1;
        default: return 0;
    }
}

// thread-local last error for calls without a ctx
std::string& cb_tls_error();
int cb_fail(cb_ctx* ctx, int status, const char* fmt, ...);

#define CB_CUDA(ctx, expr)                                                                                 \
    do {                                                                                                   \
        cudaError_t e__ = (expr);                                                                          \
        if (e__ != cudaSuccess)                                                                            \
            return cb_fail((ctx), (e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver) ? CB_ERR_NO_DEVICE : CB_ERR_CUDA, \
                           "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__);   \
    } while (0)

#define CB_TRY(expr)                       \
    do {                                   \
        int s__ = (expr);                  \
        if (s__ != CB_OK) return s__;      \
    } while (0)

#define CB_LAUNCHED(ctx) ((ctx)->launches++)

enum { CB_PROF_FILL = 0, CB_PROF_SPMM = 1, CB_PROF_FIXUP = 2 };
// brackets one kernel launch with events when profiling is on (no-ops otherwise)
struct cb_prof_scope {
    cb_ctx* c; cudaStream_t st; int kind;
    cb_prof_scope(cb_ctx* ctx, cudaStream_t stream, int k);
    ~cb_prof_scope();
};

struct cb_scratch {            // frees device temporaries on scope exit
    std::vector<void*> ptrs;
    ~cb_scratch() { for (void* p : ptrs) cudaFree(p); }
    template <typename T> cudaError_t alloc(T** p, size_t n) {
        cudaError_t e = cudaMalloc((void**)p, (n ? n : 1) * sizeof(T));
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
};

// internal entry points across translation units
int cb_tile_build_from_keys(cb_ctx* ctx, int64_t m, int64_t n, int64_t nz, uint64_t* d_keys, const void* d_vals,
                            int val_dtype, bool presorted, cb_scratch& sc, cb_tile** out);
int cb_spmm_launch(cb_ctx* ctx, cudaStream_t stream, const cb_tile* t, const void* X, int64_t ldx, void* Y, int64_t ldy,
                   int64_t k, int dtype, int semiring, int accumulate);
int cb_nccl_init(cb_ctx* ctx, const void* id128);
void cb_nccl_destroy(cb_ctx* ctx);
void cb_summa_release(cb_ctx* ctx);
int cb_p2p_prepare(cb_ctx* ctx, size_t need_bytes);
char* cb_p2p_xfull(cb_ctx* ctx);
int cb_p2p_begin(cb_ctx* ctx);
int cb_p2p_push(cb_ctx* ctx, cudaEvent_t operands_ready, int stage, size_t dst_off, const void* src, size_t bytes);
int cb_p2p_wait_stage(cb_ctx* ctx, cudaStream_t stream, int stage);
int cb_p2p_finish(cb_ctx* ctx, cudaStream_t compute);
void cb_p2p_release(cb_ctx* ctx);
int cb_p2p_debug_times(cb_ctx* ctx, cudaEvent_t origin, float* t_start, float* t_end);
int cb_nccl_allgather_col(cb_ctx* ctx, const void* send_host, void* recv_host, size_t bytes);
int cb_ingest_device_coo(cb_ctx* ctx, int64_t gm, int64_t gn, int64_t nz, const int64_t* d_rows, const int64_t* d_cols, const void* d_vals,
                         int val_dtype, int dup_op, cb_tile** out);
