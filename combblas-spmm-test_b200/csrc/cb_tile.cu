// K1: build the device tile (doubly compressed rows + nnz-balanced work partition) from the
// reference's column-compressed wire format or from COO triples.
//
// Replaces, on the device, what the reference does on the host when a tile is materialised:
// SpDCCols(const SpTuples&, bool) / the threaded tuple ctor (include/CombBLAS/SpDCCols.cpp:108-184,
// :197-304) and Transpose (:853-868).  Sorting uses CUB's device radix sort (set-up, not the hot loop).
#include <atomic>
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>
#include <algorithm>
#include "cb_common.cuh"
#include "cb_hub.cuh"

namespace {


struct ToI64 { __host__ __device__ int64_t operator()(uint8_t v) const { return (int64_t)v; } };

__host__ __device__ inline int bits_for(int64_t v) { int b = 1; while (b < 63 && (int64_t(1) << b) < v) ++b; return b; }

template <typename IT>
__global__ void expand_csc_kernel(const IT* __restrict__ cp, const IT* __restrict__ jc, const IT* __restrict__ ir,
                                  int64_t ncp /* entries in cp minus 1 */, int64_t nz, uint64_t* __restrict__ keys) {
    // one thread per nonzero: find its compressed column by binary search in cp, emit key = row<<32 | col
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nz; p += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = ncp;           // largest c with cp[c] <= p
        while (hi - lo > 1) { int64_t mid = (lo + hi) >> 1; if ((int64_t)cp[mid] <= p) lo = mid; else hi = mid; }
        const uint64_t col = jc ? (uint64_t)jc[lo] : (uint64_t)lo;
        keys[p] = ((uint64_t)ir[p] << 32) | col;
    }
}

template <typename IT>
__global__ void coo_keys_kernel(const IT* __restrict__ rows, const IT* __restrict__ cols, int64_t nz, uint64_t* __restrict__ keys) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nz; p += (int64_t)gridDim.x * blockDim.x)
        keys[p] = ((uint64_t)rows[p] >> 32 || (uint64_t)cols[p] >> 32) ? ~0ull          // negative or >= 2^32: caught by validate_keys_kernel
                                                                        : (((uint64_t)rows[p] << 32) | (uint64_t)cols[p]);
}

// ingestion check: a nonzero outside the m x n tile would write outside the column / row arrays further down
__global__ void validate_keys_kernel(const uint64_t* __restrict__ keys, int64_t nz, uint32_t m, uint32_t n, unsigned int* __restrict__ bad) {
    unsigned int mine = 0;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nz; p += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t k = keys[p];
        if ((uint32_t)(k >> 32) >= m || (uint32_t)k >= n) mine = 1;
    }
    if (__any_sync(0xffffffffu, mine) && (threadIdx.x & 31) == 0) atomicOr(bad, 1u);
}

__global__ void iota_kernel(uint32_t* p, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = (uint32_t)i;
}

// per sorted nonzero: row-start flag and column presence
__global__ void mark_entries_kernel(const uint64_t* __restrict__ keys, int64_t nz, uint8_t* __restrict__ rowstart,
                                    uint8_t* __restrict__ colseen) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nz; p += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t kcur = keys[p];
        const uint32_t row = (uint32_t)(kcur >> 32), col = (uint32_t)kcur;
        rowstart[p] = ((p == 0) || ((uint32_t)(keys[p - 1] >> 32) != row)) ? 1 : 0;
        colseen[col] = 1;
    }
}

// per sorted nonzero: column index with the end-of-row flag in the top bit
__global__ void colflag_kernel(const uint64_t* __restrict__ keys, int64_t nz, int32_t* __restrict__ colflag) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nz; p += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t kcur = keys[p];
        const uint32_t row = (uint32_t)(kcur >> 32), col = (uint32_t)kcur;
        const bool last = (p == nz - 1) || ((uint32_t)(keys[p + 1] >> 32) != row);
        colflag[p] = (int32_t)(col | (last ? 0x80000000u : 0u));
    }
}

__global__ void row_ids_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ rowptr, int64_t nzr, int64_t nz,
                               int32_t* __restrict__ nzrows, int32_t* __restrict__ rowptr_end) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nzr; i += (int64_t)gridDim.x * blockDim.x)
        nzrows[i] = (int32_t)(keys[rowptr[i]] >> 32);
    if (blockIdx.x == 0 && threadIdx.x == 0) *rowptr_end = (int32_t)nz;
}

template <typename V>
__global__ void gather_vals_kernel(const V* __restrict__ in, const uint32_t* __restrict__ perm, int64_t nz, V* __restrict__ out) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nz; p += (int64_t)gridDim.x * blockDim.x) out[p] = in[perm[p]];
}

// rows without nonzeros, ascending: row r is empty iff it is not in nzrows; its slot is r - (#nonempty rows below r)
__global__ void empty_rows_kernel(const int32_t* __restrict__ nzrows, int64_t nzr, int64_t m, int32_t* __restrict__ emptyrows) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < m; r += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = nzr;           // first index with nzrows[idx] >= r
        while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (nzrows[mid] < r) lo = mid + 1; else hi = mid; }
        if (lo == nzr || nzrows[lo] != r) emptyrows[r - lo] = (int32_t)r;
    }
}

// Work partition.  Chunk g nominally starts at nonzero g*L.  If that nonzero lies in a row of at most L nonzeros the
// start is moved back to the row's first nonzero (short rows are never cut); otherwise the long row is cut right there.
__global__ void chunk_kernel(const int32_t* __restrict__ rowptr, int64_t nzr, int64_t nz, int32_t L, int64_t nchunks,
                             int32_t* __restrict__ chunk_start, int32_t* __restrict__ chunk_row) {
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g <= nchunks; g += (int64_t)gridDim.x * blockDim.x) {
        if (g == nchunks) { chunk_start[g] = (int32_t)nz; continue; }
        const int64_t pos = g * L;
        int64_t lo = 0, hi = nzr;           // largest ridx with rowptr[ridx] <= pos
        while (hi - lo > 1) { int64_t mid = (lo + hi) >> 1; if (rowptr[mid] <= pos) lo = mid; else hi = mid; }
        const int32_t rs = rowptr[lo], re = rowptr[lo + 1];
        if (re - rs > L) {
            chunk_start[g] = (int32_t)pos;
            chunk_row[g] = (int32_t)((uint32_t)lo | (pos > rs ? 0x80000000u : 0u));
        } else {
            chunk_start[g] = rs;
            chunk_row[g] = (int32_t)lo;
        }
    }
}

__global__ void split_rows_kernel(const int32_t* __restrict__ starts, int64_t nzr, int64_t nz, int32_t L, int32_t* __restrict__ out,
                                  unsigned long long* __restrict__ count) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nzr; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t end = (i + 1 < nzr) ? starts[i + 1] : (int32_t)nz;
        if (end - starts[i] > L) {
            const unsigned long long slot = atomicAdd(count, 1ULL);
            if (out) out[slot] = (int32_t)i;
        }
    }
}

inline int grid_for(int64_t n, int sm) {
    int64_t b = (n + 255) / 256;
    int64_t cap = (int64_t)sm * 32;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

int pick_chunk_len(int64_t nnz) {
    // enough chunks for ~4 waves of resident virtual warps on 148 SMs, but long enough that the carry traffic of
    // split rows (2 panel rows per chunk) stays ~1/L of the gather traffic
    int64_t L = nnz / 32768;
    int p = 32;
    while (p * 2 <= L && p < 512) p *= 2;
    return p;
}

// keys (row<<32|col) on the device, unsorted -> tile.  `vals` (device, nz elements of val_dtype) may be NULL.
}  // namespace

// keys (row<<32|col) on the device -> tile.  `d_vals` (device, nz elements of val_dtype, in the order of d_keys) may be
// NULL.  presorted: keys are already ascending and unique (the generators sort while deduplicating).
// All arrays of the finished tile live in ONE device allocation (cb_tile_layout), so a tile travels as one message.
int cb_tile_build_from_keys(cb_ctx* ctx, int64_t m, int64_t n, int64_t nz, uint64_t* d_keys, const void* d_vals,
                            int val_dtype, bool presorted, cb_scratch& sc, cb_tile** out) {
    const int sm = ctx->sm_count;
    cudaStream_t st = ctx->compute;
    static std::atomic<uint64_t> next_uid{1};
    cb_tile* t = new cb_tile();
    t->ctx = ctx;
    t->uid = next_uid.fetch_add(1);
    *out = nullptr;
    cb_tile_meta meta;
    memset(&meta, 0, sizeof meta);
    meta.m = m; meta.n = n; meta.nnz = nz; meta.val_dtype = val_dtype; meta.layout_dtype = val_dtype; meta.chunk_len = 32;
    auto fail = [&](int s) { cb_tile_free(t); return s; };
#define T_CUDA(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { cb_fail(ctx, e__ == cudaErrorMemoryAllocation ? CB_ERR_ALLOC : CB_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); return fail(e__ == cudaErrorMemoryAllocation ? CB_ERR_ALLOC : CB_ERR_CUDA); } } while (0)
    const size_t vs = cb_dtype_size(val_dtype);
    if (nz == 0) {
        const cb_tile_layout L = cb_layout(meta);
        T_CUDA(cudaMalloc((void**)&t->slab, L.total));
        t->slab_bytes = L.total; t->owns_slab = true;
        cb_tile_bind(t, meta, t->slab);
        if (m > 0) { empty_rows_kernel<<<grid_for(m, sm), 256, 0, st>>>(nullptr, 0, m, t->emptyrows); CB_LAUNCHED(ctx); }
        T_CUDA(cudaStreamSynchronize(st));
        *out = t;
        return CB_OK;
    }
    // 0. every index inside the tile?  (malformed COO / CSC input must not turn into an out-of-bounds device write)
    {
        unsigned int* d_bad = nullptr;
        T_CUDA(sc.alloc(&d_bad, 1));
        T_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(unsigned int), st));
        validate_keys_kernel<<<grid_for(nz, sm), 256, 0, st>>>(d_keys, nz, (uint32_t)m, (uint32_t)n, d_bad);
        CB_LAUNCHED(ctx);
        unsigned int bad = 0;
        T_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof bad, cudaMemcpyDeviceToHost, st));
        T_CUDA(cudaStreamSynchronize(st));
        if (bad) { cb_fail(ctx, CB_ERR_INVALIDPARAMS, "tile input holds a row index >= %lld or a column index >= %lld (or a negative one)", (long long)m, (long long)n); return fail(CB_ERR_INVALIDPARAMS); }
    }
    // 1. stable radix sort by (row, col)
    uint64_t* keys_sorted = d_keys; uint32_t *perm = nullptr, *perm_sorted = nullptr;
    const int end_bit = 32 + bits_for(m);
    size_t tmp_bytes = 0;
    void* tmp = nullptr;
    if (presorted) {
        // nothing to do
    } else if (d_vals) {
        T_CUDA(sc.alloc(&keys_sorted, (size_t)nz));
        T_CUDA(sc.alloc(&perm, (size_t)nz));
        T_CUDA(sc.alloc(&perm_sorted, (size_t)nz));
        iota_kernel<<<grid_for(nz, sm), 256, 0, st>>>(perm, nz); CB_LAUNCHED(ctx);
        T_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, keys_sorted, perm, perm_sorted, (int)nz, 0, end_bit, st));
        T_CUDA(sc.alloc((char**)&tmp, tmp_bytes));
        T_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, d_keys, keys_sorted, perm, perm_sorted, (int)nz, 0, end_bit, st));
        ctx->launches += 4;
    } else {
        T_CUDA(sc.alloc(&keys_sorted, (size_t)nz));
        T_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, d_keys, keys_sorted, (int)nz, 0, end_bit, st));
        T_CUDA(sc.alloc((char**)&tmp, tmp_bytes));
        T_CUDA(cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, d_keys, keys_sorted, (int)nz, 0, end_bit, st));
        ctx->launches += 4;
    }
    // 2. count nonempty rows / columns, find the row starts
    uint8_t *rowstart, *colseen;
    T_CUDA(sc.alloc(&rowstart, (size_t)nz));
    T_CUDA(sc.alloc(&colseen, (size_t)(n > 0 ? n : 1)));
    T_CUDA(cudaMemsetAsync(colseen, 0, (size_t)(n > 0 ? n : 1), st));
    mark_entries_kernel<<<grid_for(nz, sm), 256, 0, st>>>(keys_sorted, nz, rowstart, colseen); CB_LAUNCHED(ctx);
    int64_t* d_counts;                 // [0] = nzr, [1] = nzc
    T_CUDA(sc.alloc(&d_counts, 2));
    int32_t* starts_tmp;
    T_CUDA(sc.alloc(&starts_tmp, (size_t)nz));
    {
        size_t b = 0;
        thrust::counting_iterator<int32_t> iota(0);
        T_CUDA(cub::DeviceSelect::Flagged(nullptr, b, iota, rowstart, starts_tmp, d_counts, (int)nz, st));
        void* tmp2; T_CUDA(sc.alloc((char**)&tmp2, b));
        T_CUDA(cub::DeviceSelect::Flagged(tmp2, b, iota, rowstart, starts_tmp, d_counts, (int)nz, st));
        size_t b2 = 0;
        auto seen64 = thrust::make_transform_iterator((const uint8_t*)colseen, ToI64());
        T_CUDA(cub::DeviceReduce::Sum(nullptr, b2, seen64, d_counts + 1, (int)n, st));
        void* tmp3; T_CUDA(sc.alloc((char**)&tmp3, b2));
        T_CUDA(cub::DeviceReduce::Sum(tmp3, b2, seen64, d_counts + 1, (int)n, st));
        ctx->launches += 3;
    }
    int64_t h_counts[2];
    T_CUDA(cudaMemcpyAsync(h_counts, d_counts, sizeof h_counts, cudaMemcpyDeviceToHost, st));
    T_CUDA(cudaStreamSynchronize(st));
    meta.nzr = h_counts[0];
    meta.nzc = h_counts[1];
    // 3. size of the work partition
    meta.chunk_len = pick_chunk_len(nz);
    meta.nchunks = (nz + meta.chunk_len - 1) / meta.chunk_len;
    unsigned long long* d_nsplit;
    T_CUDA(sc.alloc(&d_nsplit, 1));
    T_CUDA(cudaMemsetAsync(d_nsplit, 0, sizeof(unsigned long long), st));
    split_rows_kernel<<<grid_for(meta.nzr, sm), 256, 0, st>>>(starts_tmp, meta.nzr, nz, meta.chunk_len, nullptr, d_nsplit); CB_LAUNCHED(ctx);
    unsigned long long h_nsplit = 0;
    T_CUDA(cudaMemcpyAsync(&h_nsplit, d_nsplit, sizeof h_nsplit, cudaMemcpyDeviceToHost, st));
    T_CUDA(cudaStreamSynchronize(st));
    meta.nsplit = (int64_t)h_nsplit;
    // 4. one allocation for the whole tile, then fill it
    const cb_tile_layout L = cb_layout(meta);
    T_CUDA(cudaMalloc((void**)&t->slab, L.total));
    t->slab_bytes = L.total; t->owns_slab = true;
    cb_tile_bind(t, meta, t->slab);
    colflag_kernel<<<grid_for(nz, sm), 256, 0, st>>>(keys_sorted, nz, t->colflag); CB_LAUNCHED(ctx);
    if (d_vals && presorted) {
        T_CUDA(cudaMemcpyAsync(t->vals, d_vals, vs * (size_t)nz, cudaMemcpyDeviceToDevice, st));
    } else if (d_vals) {
        switch (vs) {
            case 1: gather_vals_kernel<uint8_t><<<grid_for(nz, sm), 256, 0, st>>>((const uint8_t*)d_vals, perm_sorted, nz, (uint8_t*)t->vals); break;
            case 4: gather_vals_kernel<uint32_t><<<grid_for(nz, sm), 256, 0, st>>>((const uint32_t*)d_vals, perm_sorted, nz, (uint32_t*)t->vals); break;
            default: gather_vals_kernel<uint64_t><<<grid_for(nz, sm), 256, 0, st>>>((const uint64_t*)d_vals, perm_sorted, nz, (uint64_t*)t->vals); break;
        }
        CB_LAUNCHED(ctx);
    }
    T_CUDA(cudaMemcpyAsync(t->rowptr, starts_tmp, sizeof(int32_t) * (size_t)t->nzr, cudaMemcpyDeviceToDevice, st));
    row_ids_kernel<<<grid_for(t->nzr, sm), 256, 0, st>>>(keys_sorted, t->rowptr, t->nzr, nz, t->nzrows, t->rowptr + t->nzr); CB_LAUNCHED(ctx);
    if (m > t->nzr) { empty_rows_kernel<<<grid_for(m, sm), 256, 0, st>>>(t->nzrows, t->nzr, m, t->emptyrows); CB_LAUNCHED(ctx); }
    chunk_kernel<<<grid_for(t->nchunks + 1, sm), 256, 0, st>>>(t->rowptr, t->nzr, nz, t->chunk_len, t->nchunks, t->chunk_start, t->chunk_row); CB_LAUNCHED(ctx);
    if (t->nsplit) {
        T_CUDA(cudaMemsetAsync(d_nsplit, 0, sizeof(unsigned long long), st));
        split_rows_kernel<<<grid_for(t->nzr, sm), 256, 0, st>>>(starts_tmp, t->nzr, nz, t->chunk_len, t->split_row, d_nsplit); CB_LAUNCHED(ctx);
    }
    T_CUDA(cudaStreamSynchronize(st));
    *out = t;
    return CB_OK;
#undef T_CUDA
}

namespace {

__global__ void keys_to_csr_kernel(const int32_t* __restrict__ colflag, int64_t nz, int64_t* __restrict__ colidx) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nz; i += (int64_t)gridDim.x * blockDim.x) colidx[i] = colflag[i] & 0x7fffffff;
}

int check_sizes(cb_ctx* ctx, int64_t m, int64_t n, int64_t nz, int idx_dtype, int val_dtype, const void* vals) {
    if (!ctx) return cb_fail(nullptr, CB_ERR_INVALIDPARAMS, "null ctx");
    if (m < 0 || n < 0 || nz < 0) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "negative tile dimension");
    if (idx_dtype != CB_I32 && idx_dtype != CB_I64) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "index dtype must be CB_I32 or CB_I64");
    if (val_dtype != CB_PATTERN && !cb_dtype_size(val_dtype)) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "value dtype %d", val_dtype);
    if ((val_dtype == CB_PATTERN) != (vals == nullptr)) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "CB_PATTERN tiles carry no value array and every other dtype needs one");
    if (m >= (int64_t(1) << 31) || n >= (int64_t(1) << 31) || nz >= (int64_t(1) << 31))
        return cb_fail(ctx, CB_ERR_TOO_LARGE, "tile %lld x %lld with %lld nonzeros: local arrays are limited to 2^31-1 elements", (long long)m, (long long)n, (long long)nz);
    return CB_OK;
}

}  // namespace

extern "C" {

int cb_tile_upload_csc(cb_ctx* ctx, int64_t m, int64_t n, int64_t nz, int64_t nzc, const void* cp, const void* jc,
                       const void* ir, const void* numx, int idx_dtype, int val_dtype, cb_tile** out) {
    *out = nullptr;
    CB_TRY(check_sizes(ctx, m, n, nz, idx_dtype, val_dtype, numx));
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cb_scratch sc;
    const size_t is = cb_dtype_size(idx_dtype), vs = cb_dtype_size(val_dtype);
    const int64_t ncp = jc ? nzc : n;
    cudaStream_t st = ctx->compute;
    void *d_cp = nullptr, *d_jc = nullptr, *d_ir = nullptr, *d_vals = nullptr;
    uint64_t* d_keys = nullptr;
    CB_CUDA(ctx, sc.alloc((char**)&d_cp, is * (size_t)(ncp + 1)));
    CB_CUDA(ctx, sc.alloc((char**)&d_ir, is * (size_t)nz));
    CB_CUDA(ctx, sc.alloc(&d_keys, (size_t)nz));
    CB_CUDA(ctx, cudaMemcpyAsync(d_cp, cp, is * (size_t)(ncp + 1), cudaMemcpyHostToDevice, st));
    if (nz) CB_CUDA(ctx, cudaMemcpyAsync(d_ir, ir, is * (size_t)nz, cudaMemcpyHostToDevice, st));
    if (jc) {
        CB_CUDA(ctx, sc.alloc((char**)&d_jc, is * (size_t)nzc));
        if (nzc) CB_CUDA(ctx, cudaMemcpyAsync(d_jc, jc, is * (size_t)nzc, cudaMemcpyHostToDevice, st));
    }
    if (numx && nz) {
        CB_CUDA(ctx, sc.alloc((char**)&d_vals, vs * (size_t)nz));
        CB_CUDA(ctx, cudaMemcpyAsync(d_vals, numx, vs * (size_t)nz, cudaMemcpyHostToDevice, st));
    }
    if (nz) {
        if (idx_dtype == CB_I32) expand_csc_kernel<int32_t><<<grid_for(nz, ctx->sm_count), 256, 0, st>>>((const int32_t*)d_cp, (const int32_t*)d_jc, (const int32_t*)d_ir, ncp, nz, d_keys);
        else expand_csc_kernel<int64_t><<<grid_for(nz, ctx->sm_count), 256, 0, st>>>((const int64_t*)d_cp, (const int64_t*)d_jc, (const int64_t*)d_ir, ncp, nz, d_keys);
        CB_LAUNCHED(ctx);
        CB_CUDA(ctx, cudaGetLastError());
    }
    return cb_tile_build_from_keys(ctx, m, n, nz, d_keys, nz ? d_vals : nullptr, val_dtype, false, sc, out);
}

int cb_tile_upload_coo(cb_ctx* ctx, int64_t m, int64_t n, int64_t nz, const void* rows, const void* cols, const void* vals,
                       int idx_dtype, int val_dtype, cb_tile** out) {
    *out = nullptr;
    CB_TRY(check_sizes(ctx, m, n, nz, idx_dtype, val_dtype, vals));
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cb_scratch sc;
    const size_t is = cb_dtype_size(idx_dtype), vs = cb_dtype_size(val_dtype);
    cudaStream_t st = ctx->compute;
    void *d_r = nullptr, *d_c = nullptr, *d_vals = nullptr;
    uint64_t* d_keys = nullptr;
    CB_CUDA(ctx, sc.alloc((char**)&d_r, is * (size_t)nz));
    CB_CUDA(ctx, sc.alloc((char**)&d_c, is * (size_t)nz));
    CB_CUDA(ctx, sc.alloc(&d_keys, (size_t)nz));
    if (nz) {
        CB_CUDA(ctx, cudaMemcpyAsync(d_r, rows, is * (size_t)nz, cudaMemcpyHostToDevice, st));
        CB_CUDA(ctx, cudaMemcpyAsync(d_c, cols, is * (size_t)nz, cudaMemcpyHostToDevice, st));
        if (vals) {
            CB_CUDA(ctx, sc.alloc((char**)&d_vals, vs * (size_t)nz));
            CB_CUDA(ctx, cudaMemcpyAsync(d_vals, vals, vs * (size_t)nz, cudaMemcpyHostToDevice, st));
        }
        if (idx_dtype == CB_I32) coo_keys_kernel<int32_t><<<grid_for(nz, ctx->sm_count), 256, 0, st>>>((const int32_t*)d_r, (const int32_t*)d_c, nz, d_keys);
        else coo_keys_kernel<int64_t><<<grid_for(nz, ctx->sm_count), 256, 0, st>>>((const int64_t*)d_r, (const int64_t*)d_c, nz, d_keys);
        CB_LAUNCHED(ctx);
        CB_CUDA(ctx, cudaGetLastError());
    }
    return cb_tile_build_from_keys(ctx, m, n, nz, d_keys, d_vals, val_dtype, false, sc, out);
}

int cb_tile_from_device_coo(cb_ctx* ctx, int64_t m, int64_t n, int64_t nz, const int64_t* d_rows, const int64_t* d_cols,
                            const void* d_vals, int val_dtype, cb_tile** out) {
    *out = nullptr;
    CB_TRY(check_sizes(ctx, m, n, nz, CB_I64, val_dtype, d_vals));
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cb_scratch sc;
    uint64_t* d_keys = nullptr;
    CB_CUDA(ctx, sc.alloc(&d_keys, (size_t)nz));
    if (nz) {
        coo_keys_kernel<int64_t><<<grid_for(nz, ctx->sm_count), 256, 0, ctx->compute>>>(d_rows, d_cols, nz, d_keys);
        CB_LAUNCHED(ctx);
        CB_CUDA(ctx, cudaGetLastError());
    }
    return cb_tile_build_from_keys(ctx, m, n, nz, d_keys, d_vals, val_dtype, false, sc, out);
}

int cb_tile_free(cb_tile* t) {
    if (!t) return CB_OK;
    if (t->ctx) { cudaSetDevice(t->ctx->device); cudaStreamSynchronize(t->ctx->compute); cudaStreamSynchronize(t->ctx->comm); }
    if (t->owns_slab) cudaFree(t->slab);
    cudaFree(t->carry);
    cb_hub_release(t);
    for (cb_tile* sub : t->summa_parts) cb_tile_free(sub);
    for (cb_tile* sub : t->summa_remote) cb_tile_free(sub);
    for (cb_tile* sub : t->summa_merged) cb_tile_free(sub);
    for (cb_tile* sub : t->spgemm_remote) cb_tile_free(sub);
    delete t;
    return CB_OK;
}

int cb_tile_pattern_view(const cb_tile* t, cb_tile** view) {
    cb_tile* v = new cb_tile();
    v->ctx = t->ctx; v->uid = 0;
    v->m = t->m; v->n = t->n; v->nnz = t->nnz; v->nzr = t->nzr; v->nzc = t->nzc;
    v->val_dtype = CB_PATTERN;
    v->layout_dtype = t->layout_dtype;          // the shared slab keeps its value array; the view just does not look at it
    v->slab = t->slab; v->slab_bytes = t->slab_bytes; v->owns_slab = false;
    v->colflag = t->colflag; v->vals = nullptr; v->nzrows = t->nzrows; v->rowptr = t->rowptr; v->emptyrows = t->emptyrows;
    v->nchunks = t->nchunks; v->chunk_len = t->chunk_len; v->chunk_start = t->chunk_start; v->chunk_row = t->chunk_row;
    v->nsplit = t->nsplit; v->split_row = t->split_row;
    *view = v;
    return CB_OK;
}

int cb_tile_info(const cb_tile* t, int64_t info[8]) {
    info[0] = t->nnz; info[1] = t->m; info[2] = t->n; info[3] = t->nzr; info[4] = t->nzc;
    info[5] = t->nchunks; info[6] = t->nsplit; info[7] = (int64_t)t->slab_bytes;
    return CB_OK;
}

int cb_tile_download_csr(cb_tile* t, int64_t* rowptr, int64_t* colidx, void* vals) {
    cb_ctx* ctx = t->ctx;
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->compute;
    if (rowptr) {
        std::vector<int32_t> rp((size_t)t->nzr + 1, 0), rows((size_t)t->nzr);
        if (t->nzr) {
            CB_CUDA(ctx, cudaMemcpyAsync(rp.data(), t->rowptr, sizeof(int32_t) * (size_t)(t->nzr + 1), cudaMemcpyDeviceToHost, st));
            CB_CUDA(ctx, cudaMemcpyAsync(rows.data(), t->nzrows, sizeof(int32_t) * (size_t)t->nzr, cudaMemcpyDeviceToHost, st));
            CB_CUDA(ctx, cudaStreamSynchronize(st));
        }
        for (int64_t r = 0; r <= t->m; ++r) rowptr[r] = 0;
        for (int64_t i = 0; i < t->nzr; ++i) rowptr[rows[(size_t)i] + 1] = rp[(size_t)i + 1] - rp[(size_t)i];
        for (int64_t r = 0; r < t->m; ++r) rowptr[r + 1] += rowptr[r];
    }
    if (colidx && t->nnz) {
        cb_scratch sc;
        int64_t* d;
        CB_CUDA(ctx, sc.alloc(&d, (size_t)t->nnz));
        keys_to_csr_kernel<<<grid_for(t->nnz, ctx->sm_count), 256, 0, st>>>(t->colflag, t->nnz, d);
        CB_LAUNCHED(ctx);
        CB_CUDA(ctx, cudaMemcpyAsync(colidx, d, sizeof(int64_t) * (size_t)t->nnz, cudaMemcpyDeviceToHost, st));
        CB_CUDA(ctx, cudaStreamSynchronize(st));
    }
    if (vals && t->vals && t->nnz) {
        CB_CUDA(ctx, cudaMemcpyAsync(vals, t->vals, cb_dtype_size(t->val_dtype) * (size_t)t->nnz, cudaMemcpyDeviceToHost, st));
        CB_CUDA(ctx, cudaStreamSynchronize(st));
    }
    return CB_OK;
}


// lengths of all m rows (0 for rows without nonzeros): the row pointer differences of the reference's CSR view
int cb_tile_row_lengths(cb_tile* t, int64_t* len) {
    if (!t || !len) return cb_fail(t ? t->ctx : nullptr, CB_ERR_INVALIDPARAMS, "cb_tile_row_lengths: null argument");
    cb_ctx* ctx = t->ctx;
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<int32_t> rp((size_t)t->nzr + 1, 0), rows((size_t)t->nzr);
    if (t->nzr) {
        CB_CUDA(ctx, cudaMemcpyAsync(rp.data(), t->rowptr, sizeof(int32_t) * (size_t)(t->nzr + 1), cudaMemcpyDeviceToHost, ctx->compute));
        CB_CUDA(ctx, cudaMemcpyAsync(rows.data(), t->nzrows, sizeof(int32_t) * (size_t)t->nzr, cudaMemcpyDeviceToHost, ctx->compute));
        CB_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
    }
    for (int64_t r = 0; r < t->m; ++r) len[r] = 0;
    for (int64_t i = 0; i < t->nzr; ++i) len[rows[(size_t)i]] = rp[(size_t)i + 1] - rp[(size_t)i];
    return CB_OK;
}

// the nonzeros of selected rows, concatenated in the order of `rows` (columns ascending inside a row); the caller sizes
// cols / vals from cb_tile_row_lengths.  For parity checks at sizes where copying the whole tile back is too much.
int cb_tile_download_rows(cb_tile* t, int64_t nrows, const int64_t* rows, int64_t* cols, void* vals) {
    if (!t || (nrows > 0 && (!rows || !cols))) return cb_fail(t ? t->ctx : nullptr, CB_ERR_INVALIDPARAMS, "cb_tile_download_rows: null argument");
    cb_ctx* ctx = t->ctx;
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<int32_t> rp((size_t)t->nzr + 1, 0), nzrows((size_t)t->nzr);
    if (t->nzr) {
        CB_CUDA(ctx, cudaMemcpyAsync(rp.data(), t->rowptr, sizeof(int32_t) * (size_t)(t->nzr + 1), cudaMemcpyDeviceToHost, ctx->compute));
        CB_CUDA(ctx, cudaMemcpyAsync(nzrows.data(), t->nzrows, sizeof(int32_t) * (size_t)t->nzr, cudaMemcpyDeviceToHost, ctx->compute));
        CB_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
    }
    const size_t vs = (t->vals && vals) ? cb_dtype_size(t->val_dtype) : 0;
    std::vector<int32_t> tmp;
    int64_t out = 0;
    for (int64_t i = 0; i < nrows; ++i) {
        if (rows[i] < 0 || rows[i] >= t->m) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_tile_download_rows: row %lld of %lld", (long long)rows[i], (long long)t->m);
        auto it = std::lower_bound(nzrows.begin(), nzrows.end(), (int32_t)rows[i]);
        if (it == nzrows.end() || *it != (int32_t)rows[i]) continue;                       // empty row
        const size_t ri = (size_t)(it - nzrows.begin());
        const int64_t s = rp[ri], n = rp[ri + 1] - rp[ri];
        tmp.resize((size_t)n);
        CB_CUDA(ctx, cudaMemcpyAsync(tmp.data(), t->colflag + s, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, ctx->compute));
        if (vs) CB_CUDA(ctx, cudaMemcpyAsync((char*)vals + (size_t)out * vs, (const char*)t->vals + (size_t)s * vs, vs * (size_t)n, cudaMemcpyDeviceToHost, ctx->compute));
        CB_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
        for (int64_t j = 0; j < n; ++j) cols[out + j] = (int64_t)(tmp[(size_t)j] & 0x7fffffff);
        out += n;
    }
    return CB_OK;
}

}  // extern "C"
