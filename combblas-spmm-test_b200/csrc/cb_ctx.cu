// Context, error reporting, timers and dense panels of the C ABI (include/combblas_b200.h).
#include <cstdarg>
#include <cfloat>
#include <climits>
#include "cb_common.cuh"

std::string& cb_tls_error() {
    static thread_local std::string e;
    return e;
}

int cb_fail(cb_ctx* ctx, int status, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    cb_tls_error() = buf;
    if (ctx) ctx->err = buf;
    return status;
}

static cudaEvent_t prof_event(cb_ctx* c) {
    cudaEvent_t e = nullptr;
    if (!c->prof_pool.empty()) { e = c->prof_pool.back(); c->prof_pool.pop_back(); }
    else cudaEventCreate(&e);
    return e;
}
cb_prof_scope::cb_prof_scope(cb_ctx* ctx, cudaStream_t stream, int k) : c(ctx), st(stream), kind(k) {
    if (c->profiling) { cudaEvent_t e = prof_event(c); cudaEventRecord(e, st); c->prof_events[kind].push_back(e); }
}
cb_prof_scope::~cb_prof_scope() {
    if (c->profiling) { cudaEvent_t e = prof_event(c); cudaEventRecord(e, st); c->prof_events[kind].push_back(e); }
}

extern "C" {

int cb_abi_version(void) { return CB_ABI_VERSION; }

const char* cb_status_string(int s) {
    switch (s) {
        case CB_OK: return "ok";
        case CB_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU path)";
        case CB_ERR_CUDA: return "CUDA error";
        case CB_ERR_NCCL: return "NCCL error";
        case CB_ERR_UNSUPPORTED: return "unsupported semiring/dtype combination";
        case CB_ERR_ALLOC: return "allocation failed";
        case CB_ERR_TOO_LARGE: return "tile array has 2^31 or more elements";
        case CB_ERR_GRIDMISMATCH: return "GRIDMISMATCH";
        case CB_ERR_DIMMISMATCH: return "DIMMISMATCH";
        case CB_ERR_NOTSQUARE: return "NOTSQUARE";
        case CB_ERR_NOFILE: return "NOFILE";
        case CB_ERR_MATRIXALIAS: return "MATRIXALIAS";
        case CB_ERR_INVALIDPARAMS: return "INVALIDPARAMS";
    }
    return "unknown status";
}

const char* cb_last_error(const cb_ctx* ctx) { return ctx ? ctx->err.c_str() : cb_tls_error().c_str(); }

int cb_device_count(int* count) {
    *count = 0;
    CB_CUDA(nullptr, cudaGetDeviceCount(count));
    return CB_OK;
}

int cb_ctx_create_grid(int device, int rank, int nranks, int pr, int pc, const void* id128, cb_ctx** out) {
    *out = nullptr;
    if (nranks < 1 || pr < 1 || pc < 1 || pr * pc != nranks || rank < 0 || rank >= nranks)
        return cb_fail(nullptr, CB_ERR_INVALIDPARAMS, "grid %d x %d does not match %d ranks (rank %d)", pr, pc, nranks, rank);
    int ndev = 0;
    CB_CUDA(nullptr, cudaGetDeviceCount(&ndev));
    if (ndev == 0) return cb_fail(nullptr, CB_ERR_NO_DEVICE, "no CUDA device visible");
    if (device < 0 || device >= ndev) return cb_fail(nullptr, CB_ERR_INVALIDPARAMS, "device %d of %d", device, ndev);
    CB_CUDA(nullptr, cudaSetDevice(device));
    cb_ctx* c = new cb_ctx();
    c->device = device;
    c->rank = rank; c->nranks = nranks; c->pr = pr; c->pc = pc;
    c->myprocrow = rank / pc;          // src/CommGrid.cpp:62-63
    c->myproccol = rank % pc;
    cudaDeviceProp prop;
    CB_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    if (prop.major < 10) {
        delete c;
        return cb_fail(nullptr, CB_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    }
    CB_CUDA(nullptr, cudaStreamCreateWithFlags(&c->compute, cudaStreamNonBlocking));
    CB_CUDA(nullptr, cudaStreamCreateWithFlags(&c->comm, cudaStreamNonBlocking));
    CB_CUDA(nullptr, cudaEventCreate(&c->t0));
    CB_CUDA(nullptr, cudaEventCreate(&c->t1));
    if (const char* tr = getenv("CB_SUMMA_TRANSPORT")) c->summa_p2p = strcmp(tr, "nccl") != 0;
    if (const char* mg = getenv("CB_SUMMA_MERGE")) c->summa_merge = strcmp(mg, "0") != 0;
    if (nranks > 1) {
        if (!id128) { cb_ctx_destroy(c); return cb_fail(nullptr, CB_ERR_INVALIDPARAMS, "a unique id is required for %d ranks", nranks); }
        int s = cb_nccl_init(c, id128);
        if (s != CB_OK) { std::string keep = c->err; cb_ctx_destroy(c); cb_tls_error() = keep; return s; }
    }
    *out = c;
    return CB_OK;
}

int cb_ctx_create(int device, cb_ctx** out) { return cb_ctx_create_grid(device, 0, 1, 1, 1, nullptr, out); }

int cb_ctx_destroy(cb_ctx* c) {
    if (!c) return CB_OK;
    cudaSetDevice(c->device);
    if (c->compute) cudaStreamSynchronize(c->compute);
    if (c->comm) cudaStreamSynchronize(c->comm);
    cb_summa_release(c);
    cb_p2p_release(c);
    cb_nccl_destroy(c);
    for (int k = 0; k < 3; ++k) for (cudaEvent_t e : c->prof_events[k]) cudaEventDestroy(e);
    for (cudaEvent_t e : c->prof_pool) cudaEventDestroy(e);
    cudaFree(c->ws_x);
    cudaFree(c->ws_y);
    cudaFree(c->win_panel);
    cudaFree(c->k2_counter);
    if (c->h2d) {
        cudaStreamSynchronize(c->h2d); cudaStreamSynchronize(c->d2h);
        cudaStreamDestroy(c->h2d); cudaStreamDestroy(c->d2h);
        for (int i = 0; i < CB_MAX_SLABS; ++i) { cudaEventDestroy(c->slab_up[i]); cudaEventDestroy(c->slab_done[i]); }
        cudaEventDestroy(c->host_begin);
    }
    if (c->t0) cudaEventDestroy(c->t0);
    if (c->t1) cudaEventDestroy(c->t1);
    if (c->compute) cudaStreamDestroy(c->compute);
    if (c->comm) cudaStreamDestroy(c->comm);
    delete c;
    return CB_OK;
}

int cb_ctx_grid(const cb_ctx* c, int* rank, int* pr, int* pc, int* myprocrow, int* myproccol) {
    if (rank) *rank = c->rank;
    if (pr) *pr = c->pr;
    if (pc) *pc = c->pc;
    if (myprocrow) *myprocrow = c->myprocrow;
    if (myproccol) *myproccol = c->myproccol;
    return CB_OK;
}

int cb_ctx_sync(cb_ctx* c) {
    CB_CUDA(c, cudaSetDevice(c->device));
    CB_CUDA(c, cudaStreamSynchronize(c->comm));
    CB_CUDA(c, cudaStreamSynchronize(c->compute));
    return CB_OK;
}

void* cb_ctx_stream(cb_ctx* c) { return (void*)c->compute; }

int cb_timer_start(cb_ctx* c) {
    CB_CUDA(c, cudaEventRecord(c->t0, c->compute));
    return CB_OK;
}

int cb_timer_stop(cb_ctx* c, float* ms) {
    CB_CUDA(c, cudaEventRecord(c->t1, c->compute));
    CB_CUDA(c, cudaEventSynchronize(c->t1));
    CB_CUDA(c, cudaEventElapsedTime(ms, c->t0, c->t1));
    return CB_OK;
}

int64_t cb_launch_count(const cb_ctx* c) { return c->launches; }

int cb_profile_enable(cb_ctx* c, int on) {
    CB_CUDA(c, cudaStreamSynchronize(c->compute));
    for (int k = 0; k < 3; ++k) { for (cudaEvent_t e : c->prof_events[k]) c->prof_pool.push_back(e); c->prof_events[k].clear(); }
    c->profiling = on != 0;
    return CB_OK;
}

int cb_profile_read(cb_ctx* c, double ms[3], int64_t launches[3]) {
    CB_CUDA(c, cudaStreamSynchronize(c->compute));
    for (int k = 0; k < 3; ++k) {
        ms[k] = 0; launches[k] = (int64_t)c->prof_events[k].size() / 2;
        for (size_t i = 0; i + 1 < c->prof_events[k].size(); i += 2) {
            float t = 0;
            CB_CUDA(c, cudaEventElapsedTime(&t, c->prof_events[k][i], c->prof_events[k][i + 1]));
            ms[k] += t;
        }
    }
    return CB_OK;
}

// ---------------------------------------------------------------------------------- dense panels

int cb_dense_alloc(cb_ctx* c, int64_t rows, int64_t cols, int dtype, cb_dense** out) {
    *out = nullptr;
    size_t es = cb_dtype_size(dtype);
    if (!es || rows < 0 || cols < 0) return cb_fail(c, CB_ERR_INVALIDPARAMS, "cb_dense_alloc(%lld x %lld, dtype %d)", (long long)rows, (long long)cols, dtype);
    CB_CUDA(c, cudaSetDevice(c->device));
    cb_dense* d = new cb_dense();
    d->ctx = c; d->rows = rows; d->cols = cols; d->dtype = dtype;
    const int64_t per16 = 16 / (int64_t)es;                    // rows start on 16-byte boundaries so the
    d->ld = (cols + per16 - 1) / per16 * per16;                // kernel can use 128-bit loads for any k
    if (d->ld == 0) d->ld = per16;
    size_t bytes = (size_t)(rows > 0 ? rows : 1) * (size_t)d->ld * es;
    cudaError_t e = cudaMalloc(&d->ptr, bytes);
    if (e != cudaSuccess) { delete d; return cb_fail(c, CB_ERR_ALLOC, "cudaMalloc(%zu) for a %lld x %lld panel: %s", bytes, (long long)rows, (long long)cols, cudaGetErrorString(e)); }
    if (d->ld != cols) CB_CUDA(c, cudaMemsetAsync(d->ptr, 0, bytes, c->compute));   // defined padding
    *out = d;
    return CB_OK;
}

int cb_dense_wrap(cb_ctx* c, void* ptr, int64_t rows, int64_t cols, int64_t ld, int dtype, cb_dense** out) {
    *out = nullptr;
    if (!cb_dtype_size(dtype) || ld < cols) return cb_fail(c, CB_ERR_INVALIDPARAMS, "cb_dense_wrap: bad dtype or ld");
    cb_dense* d = new cb_dense();
    d->ctx = c; d->rows = rows; d->cols = cols; d->ld = ld; d->dtype = dtype; d->ptr = ptr; d->owned = false;
    *out = d;
    return CB_OK;
}

int cb_dense_free(cb_dense* d) {
    if (!d) return CB_OK;
    if (d->owned && d->ptr) {
        cudaSetDevice(d->ctx->device);
        cudaStreamSynchronize(d->ctx->compute);
        cudaFree(d->ptr);
    }
    delete d;
    return CB_OK;
}

int cb_dense_upload(cb_dense* d, const void* host, int64_t ld_host) {
    cb_ctx* c = d->ctx;
    size_t es = cb_dtype_size(d->dtype);
    if (ld_host < d->cols) return cb_fail(c, CB_ERR_INVALIDPARAMS, "cb_dense_upload: ld_host < cols");
    CB_CUDA(c, cudaSetDevice(c->device));
    if (d->rows == 0 || d->cols == 0) return CB_OK;
    CB_CUDA(c, cudaMemcpy2DAsync(d->ptr, (size_t)d->ld * es, host, (size_t)ld_host * es, (size_t)d->cols * es,
                                 (size_t)d->rows, cudaMemcpyHostToDevice, c->compute));
    return CB_OK;
}

int cb_dense_download(cb_dense* d, void* host, int64_t ld_host) {
    cb_ctx* c = d->ctx;
    size_t es = cb_dtype_size(d->dtype);
    if (ld_host < d->cols) return cb_fail(c, CB_ERR_INVALIDPARAMS, "cb_dense_download: ld_host < cols");
    CB_CUDA(c, cudaSetDevice(c->device));
    if (d->rows && d->cols)
        CB_CUDA(c, cudaMemcpy2DAsync(host, (size_t)ld_host * es, d->ptr, (size_t)d->ld * es, (size_t)d->cols * es,
                                     (size_t)d->rows, cudaMemcpyDeviceToHost, c->compute));
    CB_CUDA(c, cudaStreamSynchronize(c->compute));
    return CB_OK;
}


}  // extern "C"
__global__ void __launch_bounds__(256)
cb_gather_rows_kernel(const char* __restrict__ src, int64_t ld_bytes, const int64_t* __restrict__ rows, int64_t nrows, int row_bytes, char* __restrict__ dst) {
    const int64_t total = nrows * row_bytes;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / row_bytes;
        dst[i] = src[rows[r] * ld_bytes + (i - r * row_bytes)];
    }
}

extern "C" {
// selected rows of a panel, packed: host[i, 0:cols] = d[rows[i], 0:cols].  For parity checks at sizes where copying the whole
// panel back is too much (bench.py's sampled-row check, tests at BASELINE.json's full sizes).
int cb_dense_download_rows(cb_dense* d, int64_t nrows, const int64_t* rows, void* host) {
    if (!d || (nrows > 0 && (!rows || !host))) return cb_fail(d ? d->ctx : nullptr, CB_ERR_INVALIDPARAMS, "cb_dense_download_rows: null argument");
    cb_ctx* c = d->ctx;
    const size_t es = cb_dtype_size(d->dtype);
    if (nrows <= 0 || d->cols == 0) return CB_OK;
    for (int64_t i = 0; i < nrows; ++i)
        if (rows[i] < 0 || rows[i] >= d->rows) return cb_fail(c, CB_ERR_INVALIDPARAMS, "cb_dense_download_rows: row %lld of %lld", (long long)rows[i], (long long)d->rows);
    CB_CUDA(c, cudaSetDevice(c->device));
    cb_scratch sc;
    int64_t* d_rows = nullptr;
    char* d_out = nullptr;
    const size_t rb = (size_t)d->cols * es;
    CB_CUDA(c, sc.alloc(&d_rows, (size_t)nrows));
    CB_CUDA(c, sc.alloc(&d_out, (size_t)nrows * rb));
    CB_CUDA(c, cudaMemcpyAsync(d_rows, rows, sizeof(int64_t) * (size_t)nrows, cudaMemcpyHostToDevice, c->compute));
    int64_t blocks = ((int64_t)nrows * (int64_t)rb + 255) / 256;
    if (blocks > (int64_t)c->sm_count * 16) blocks = (int64_t)c->sm_count * 16;
    cb_gather_rows_kernel<<<(unsigned)blocks, 256, 0, c->compute>>>((const char*)d->ptr, (int64_t)d->ld * (int64_t)es, d_rows, nrows, (int)rb, d_out);
    CB_LAUNCHED(c);
    CB_CUDA(c, cudaGetLastError());
    CB_CUDA(c, cudaMemcpyAsync(host, d_out, (size_t)nrows * rb, cudaMemcpyDeviceToHost, c->compute));
    CB_CUDA(c, cudaStreamSynchronize(c->compute));
    return CB_OK;
}

int cb_dense_info(const cb_dense* d, int64_t* rows, int64_t* cols, int64_t* ld, int* dtype, void** ptr) {
    if (rows) *rows = d->rows;
    if (cols) *cols = d->cols;
    if (ld) *ld = d->ld;
    if (dtype) *dtype = d->dtype;
    if (ptr) *ptr = d->ptr;
    return CB_OK;
}

int cb_semiring_id(int semiring, int dtype, void* out) {
    // SR::id(): Semirings.h:215 (0), :239 (numeric max), :194 (-1)
    switch (semiring) {
        case CB_PLUS_TIMES: case CB_OR_AND:
            memset(out, 0, cb_dtype_size(dtype));
            return cb_dtype_size(dtype) ? CB_OK : CB_ERR_UNSUPPORTED;
        case CB_MIN_PLUS:
            switch (dtype) {
                case CB_F32: { float v = FLT_MAX; memcpy(out, &v, 4); return CB_OK; }
                case CB_F64: { double v = DBL_MAX; memcpy(out, &v, 8); return CB_OK; }
                case CB_I32: { int32_t v = INT32_MAX; memcpy(out, &v, 4); return CB_OK; }
                case CB_I64: { int64_t v = INT64_MAX; memcpy(out, &v, 8); return CB_OK; }
            }
            return CB_ERR_UNSUPPORTED;
        case CB_MAX_SEL2ND:
            switch (dtype) {
                case CB_F32: { float v = -1.f; memcpy(out, &v, 4); return CB_OK; }
                case CB_F64: { double v = -1.0; memcpy(out, &v, 8); return CB_OK; }
                case CB_I32: { int32_t v = -1; memcpy(out, &v, 4); return CB_OK; }
                case CB_I64: { int64_t v = -1; memcpy(out, &v, 8); return CB_OK; }
            }
            return CB_ERR_UNSUPPORTED;
    }
    return CB_ERR_UNSUPPORTED;
}

}  // extern "C"

// ---- fill: one 16-byte store per thread over the padded rows
template <typename T>
__global__ void cb_fill_kernel(T* __restrict__ p, int64_t rows, int64_t cols, int64_t ld, T v) {
    const int64_t total = rows * cols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols, c = i - r * cols;
        p[r * ld + c] = v;
    }
}

extern "C" int cb_dense_fill(cb_dense* d, const void* scalar) {
    cb_ctx* c = d->ctx;
    CB_CUDA(c, cudaSetDevice(c->device));
    const int64_t total = d->rows * d->cols;
    if (total == 0) return CB_OK;
    int blocks = (int)((total + 255) / 256 < (int64_t)c->sm_count * 16 ? (total + 255) / 256 : (int64_t)c->sm_count * 16);
    switch (cb_dtype_size(d->dtype)) {
        case 1: { uint8_t v; memcpy(&v, scalar, 1); cb_fill_kernel<uint8_t><<<blocks, 256, 0, c->compute>>>((uint8_t*)d->ptr, d->rows, d->cols, d->ld, v); break; }
        case 4: { uint32_t v; memcpy(&v, scalar, 4); cb_fill_kernel<uint32_t><<<blocks, 256, 0, c->compute>>>((uint32_t*)d->ptr, d->rows, d->cols, d->ld, v); break; }
        case 8: { uint64_t v; memcpy(&v, scalar, 8); cb_fill_kernel<uint64_t><<<blocks, 256, 0, c->compute>>>((uint64_t*)d->ptr, d->rows, d->cols, d->ld, v); break; }
        default: return cb_fail(c, CB_ERR_UNSUPPORTED, "cb_dense_fill: dtype %d", d->dtype);
    }
    CB_LAUNCHED(c);
    CB_CUDA(c, cudaGetLastError());
    return CB_OK;
}
