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
    // optional per-kernel device timing (bench.py's roofline): event pairs around K3 / K2 / fix-up launches
    bool profiling = false;
    std::vector<cudaEvent_t> prof_events[3];     // [kind] -> begin,end,begin,end,...
    std::vector<cudaEvent_t> prof_pool;
    // device panels reused by cb_spmm_host (host operands): grown, never shrunk
    void* ws_x = nullptr; size_t ws_x_bytes = 0;
    void* ws_y = nullptr; size_t ws_y_bytes = 0;
    float summa_ms[4] = {0, 0, 0, 0};
    void* summa_state = nullptr;      // receive buffers + events, owned by cb_summa.cu
    std::string err;
};

// Device tile: doubly compressed rows (the row-major mirror of Dcsc, dcsc.h:124-131) plus a work partition.
struct cb_tile {
    cb_ctx* ctx = nullptr;
    int64_t m = 0, n = 0, nnz = 0;
    int64_t nzr = 0;          // rows with at least one nonzero
    int64_t nzc = 0;          // columns with at least one nonzero (n_used of the traffic model)
    int val_dtype = CB_PATTERN;
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
    int32_t* split_first = nullptr;  // [nsplit] first chunk holding a piece
    int32_t* split_last = nullptr;   // [nsplit] last chunk holding a piece
    // scratch for partial rows, sized at first use for the widest panel seen: 2 slots (head, tail) per chunk
    void* carry = nullptr;
    size_t carry_bytes = 0;
    size_t bytes = 0;
};

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
