// cb_spmm_local / cb_spmm_host: validation, K3 (identity fill of empty rows) and dispatch into K2.
#include <algorithm>
#include "cb_spmm_dispatch.cuh"

#ifndef CB_L2_HINT_MB_DEFAULT
#define CB_L2_HINT_MB_DEFAULT 0      // budget of the evict_last rows when nothing else is asked for (0 = no hints)
#endif

// K3: rows of the tile without nonzeros receive SR::id() - the dense-output convention of the reference's
// dense SpMV (std::fill_n(localy, ysize, SR::id()), include/CombBLAS/ParFriends.h:1960-1963), restricted to
// the rows K2 will not write so Y is written exactly once.
__global__ void __launch_bounds__(256)
cb_fill_rows_kernel(const int32_t* __restrict__ rows, int64_t nrows, char* __restrict__ Y, int64_t ldy_bytes,
                    int row_bytes, uint4 pattern) {
    const int vecs = row_bytes >> 4;
    const int64_t total = nrows * vecs;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / vecs;
        const int v = (int)(i - r * vecs);
        *reinterpret_cast<uint4*>(Y + (int64_t)rows[r] * ldy_bytes + v * 16) = pattern;
    }
}

static uint4 id_pattern(int semiring, int dtype) {
    unsigned char s[8] = {0};
    cb_semiring_id(semiring, dtype, s);
    unsigned char b[16];
    const size_t es = cb_dtype_size(dtype);
    for (size_t i = 0; i < 16; ++i) b[i] = s[i % es];
    uint4 p;
    memcpy(&p, b, 16);
    return p;
}

int cb_spmm_launch(cb_ctx* ctx, cudaStream_t stream, const cb_tile* t, const void* X, int64_t ldx, void* Y, int64_t ldy,
                   int64_t k, int dtype, int semiring, int accumulate) {
    const size_t es = cb_dtype_size(dtype);
    if (!es) return cb_fail(ctx, CB_ERR_UNSUPPORTED, "cb_spmm: dtype %d", dtype);
    if (semiring == CB_PLUS_TIMES && dtype == CB_U8) semiring = CB_OR_AND;      // PlusTimesSRing<bool,bool>
    int akind;
    if (t->val_dtype == CB_PATTERN) akind = cbk::A_PATTERN;
    else if (t->val_dtype == CB_U8) akind = cbk::A_BOOL;
    else if (t->val_dtype == dtype) akind = cbk::A_SAME;
    else return cb_fail(ctx, CB_ERR_UNSUPPORTED, "cb_spmm: tile values of dtype %d with a panel of dtype %d (promote_trait has no such pair)", t->val_dtype, dtype);
    bool ok = false;
    switch (semiring) {
        case CB_PLUS_TIMES: ok = dtype != CB_U8; break;
        case CB_OR_AND: ok = dtype == CB_U8 && akind != cbk::A_SAME; break;    // a CB_U8 tile was classified A_BOOL above
        case CB_MIN_PLUS: ok = dtype != CB_U8 && akind == cbk::A_SAME; break;
        case CB_MAX_SEL2ND: ok = dtype != CB_U8 && akind != cbk::A_SAME; break;
    }
    if (!ok) return cb_fail(ctx, CB_ERR_UNSUPPORTED, "cb_spmm: semiring %d with tile dtype %d and panel dtype %d is not part of the ABI", semiring, t->val_dtype, dtype);
    const int64_t row_bytes = (k * (int64_t)es + 15) / 16 * 16;
    if (k <= 0 || t->m == 0) return CB_OK;
    if (((uintptr_t)X | (uintptr_t)Y) & 15 || (ldx * es) % 16 || (ldy * es) % 16 || row_bytes > ldx * (int64_t)es || row_bytes > ldy * (int64_t)es)
        return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_spmm: panels must be 16-byte aligned with leading dimensions padded to 16 bytes (use cb_dense_alloc)");
    if (row_bytes > (1 << 20)) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_spmm: panel row of %lld bytes", (long long)row_bytes);
    CB_CUDA(ctx, cudaSetDevice(ctx->device));

    if (!accumulate && t->m > t->nzr) {
        const int64_t total = (t->m - t->nzr) * (row_bytes / 16);
        int64_t blocks = (total + 255) / 256;
        if (blocks > (int64_t)ctx->sm_count * 8) blocks = (int64_t)ctx->sm_count * 8;
        {
            cb_prof_scope prof(ctx, stream, CB_PROF_FILL);
            cb_fill_rows_kernel<<<(unsigned)blocks, 256, 0, stream>>>(t->emptyrows, t->m - t->nzr, (char*)Y, ldy * (int64_t)es,
                                                                      (int)row_bytes, id_pattern(semiring, dtype));
        }
        CB_LAUNCHED(ctx);
        CB_CUDA(ctx, cudaGetLastError());
    }
    if (t->nnz == 0) return CB_OK;
    if (t->nsplit > 0) {
        const size_t need = (size_t)2 * (size_t)t->nchunks * (size_t)row_bytes;
        cb_tile* mt = const_cast<cb_tile*>(t);       // scratch only; grows to the widest panel seen
        if (mt->carry_bytes < need) {
            if (mt->carry) { CB_CUDA(ctx, cudaStreamSynchronize(stream)); CB_CUDA(ctx, cudaFree(mt->carry)); mt->carry = nullptr; mt->carry_bytes = 0; }
            cudaError_t e = cudaMalloc(&mt->carry, need);
            if (e != cudaSuccess) return cb_fail(ctx, CB_ERR_ALLOC, "cudaMalloc(%zu) for the split-row carry buffer: %s", need, cudaGetErrorString(e));
            mt->carry_bytes = need;
        }
    }
    cbk::LaunchParams p;
    p.ctx = ctx; p.stream = stream; p.t = t;
    p.X = X; p.ldx_bytes = ldx * (int64_t)es;
    p.Y = Y; p.ldy_bytes = ldy * (int64_t)es;
    p.total_row_bytes = (int)row_bytes;
    p.accumulate = accumulate;
    p.slab_bytes = ctx->k2_slab_bytes;
    p.point = ctx->k2_point;
    p.pipe = ctx->k2_pipe;
    {
        // L2 residency hints (K2P): worth it only when the X rows this tile touches do not fit in L2 anyway
        static const int l2_env = getenv("CB_K2_L2_MB") ? atoi(getenv("CB_K2_L2_MB")) : -1;
        const int l2_mb = ctx->k2_l2_mb >= 0 ? ctx->k2_l2_mb : (l2_env >= 0 ? l2_env : CB_L2_HINT_MB_DEFAULT);
        if (l2_mb > 0 && stream == ctx->compute && (double)t->nzc * (double)row_bytes > 1.5e6 * (double)l2_mb) {
            CB_TRY(cb_hubcls_get(ctx, t, &p.hubcls));
            int c = -1;
            while ((2LL << (c + 1)) * row_bytes <= (int64_t)l2_mb * 1000000LL) ++c;     // 2^(c+1) rows of the classes 0..c fit the budget
            p.cls_max = c;
            if (c < 0) p.hubcls = nullptr;
        }
    }
    // K2W (opt-in, cb_spmm_k2_pipe(ctx, 32)): the rows of the most used columns - as many as fit the budget (cb_spmm_k2_l2, default
    // 64 MB) - are packed into one panel that a persisting L2 access-policy window keeps on the chip for the duration of the multiply
    bool win_on = false;
    {
        static const int pipe_env = getenv("CB_K2_PIPE") ? atoi(getenv("CB_K2_PIPE")) : -1;
        const int pipe = ctx->k2_pipe >= 0 ? ctx->k2_pipe : pipe_env;
        const int mb = ctx->k2_l2_mb > 0 ? ctx->k2_l2_mb : 64;
        if (pipe == 32 && stream == ctx->compute && ldx * (int64_t)es == row_bytes && row_bytes >= 256 && (dtype == CB_F32 || dtype == CB_F64) &&
            semiring == CB_PLUS_TIMES && akind == cbk::A_SAME) {
            const int32_t* wcols = nullptr;
            int64_t h = 0;
            CB_TRY(cb_hubwin_get(ctx, t, (int64_t)mb * (1 << 20) / row_bytes, &p.win_colflag, &wcols, &h, nullptr));
            if (p.win_colflag) {
                const size_t need = (size_t)h * (size_t)row_bytes;
                if (ctx->win_panel_bytes < need) {
                    if (ctx->win_panel) { CB_CUDA(ctx, cudaStreamSynchronize(stream)); CB_CUDA(ctx, cudaFree(ctx->win_panel)); ctx->win_panel = nullptr; ctx->win_panel_bytes = 0; }
                    cudaError_t e = cudaMalloc(&ctx->win_panel, need);
                    if (e != cudaSuccess) return cb_fail(ctx, CB_ERR_ALLOC, "cudaMalloc(%zu) for the hub panel: %s", need, cudaGetErrorString(e));
                    ctx->win_panel_bytes = need;
                }
                cudaDeviceProp prop;
                static int max_persist = -1, max_window = -1;
                if (max_persist < 0) {
                    CB_CUDA(ctx, cudaGetDeviceProperties(&prop, ctx->device));
                    max_persist = prop.persistingL2CacheMaxSize;
                    max_window = prop.accessPolicyMaxWindowSize;
                }
                const size_t limit = std::min<size_t>(need, (size_t)max_persist);
                if (ctx->win_l2_limit != limit) { CB_CUDA(ctx, cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, limit)); ctx->win_l2_limit = limit; }
                CB_TRY(cb_hubwin_gather(ctx, stream, X, ldx * (int64_t)es, wcols, h, (int)row_bytes, ctx->win_panel));
                cudaStreamAttrValue attr;
                memset(&attr, 0, sizeof attr);
                attr.accessPolicyWindow.base_ptr = ctx->win_panel;
                attr.accessPolicyWindow.num_bytes = std::min<size_t>(need, (size_t)max_window);
                attr.accessPolicyWindow.hitRatio = limit >= need ? 1.0f : (float)((double)limit / (double)need);
                attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                CB_CUDA(ctx, cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &attr));
                p.win_delta = (const char*)ctx->win_panel - (const char*)X;
                win_on = true;
            }
        }
    }
    {
        static const int pipe_env2 = getenv("CB_K2_PIPE") ? atoi(getenv("CB_K2_PIPE")) : -1;
        const int pipe = ctx->k2_pipe >= 0 ? ctx->k2_pipe : pipe_env2;
        if (pipe == 64) {                         // K2 with persistent warps: the chunk counter starts at zero on this stream
            if (!ctx->k2_counter) CB_CUDA(ctx, cudaMalloc((void**)&ctx->k2_counter, sizeof(unsigned)));
            CB_CUDA(ctx, cudaMemsetAsync(ctx->k2_counter, 0, sizeof(unsigned), stream));
            p.persist_counter = ctx->k2_counter;
        }
    }
    struct WinOff {                               // the window must not outlive the multiply
        cudaStream_t s; bool on;
        ~WinOff() {
            if (!on) return;
            cudaStreamAttrValue attr;
            memset(&attr, 0, sizeof attr);
            attr.accessPolicyWindow.num_bytes = 0;
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
            cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &attr);
        }
    } win_off{stream, win_on};
    cbk::HubPlan hub_plan;                        // opt-in persistent variants K2H / K2R (cb_hub.cu); inactive -> plain K2
    CB_TRY(cb_hub_plan(ctx, t, row_bytes, stream, &hub_plan));
    if (hub_plan.active) p.hub = &hub_plan;
    switch (semiring) {
        case CB_PLUS_TIMES:
            return (dtype == CB_F32 || dtype == CB_F64) ? cb_launch_plus_times_f(dtype, akind, p) : cb_launch_plus_times_i(dtype, akind, p);
        case CB_MIN_PLUS: return cb_launch_min_plus(dtype, p);
        case CB_MAX_SEL2ND: return cb_launch_select_max(dtype, p);
        case CB_OR_AND: return cb_launch_or_and(akind, p);
    }
    return CB_ERR_UNSUPPORTED;
}

extern "C" {

int cb_spmm_k2_l2(cb_ctx* ctx, int budget_mb) {
    if (!ctx) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_spmm_k2_l2: null ctx");
    if (budget_mb < -1 || budget_mb > 4096) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_spmm_k2_l2: budget of %d MB (-1 default, 0 off)", budget_mb);
    ctx->k2_l2_mb = budget_mb;
    return CB_OK;
}

int cb_spmm_k2_pipe(cb_ctx* ctx, int depth) {
    if (!ctx) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_spmm_k2_pipe: null ctx");
    if (depth != -1 && depth != 0 && depth != 1 && depth != 4 && depth != 8 && depth != 16 && depth != 32 && depth != 64) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_spmm_k2_pipe: mode %d (-1 default, 0 round-1 walk, 1 round-1 walk with entry prefetch, 4 / 8 ring depth, 16 bulk-copy ring, 32 hub panel under an L2 window, 64 persistent warps)", depth);
    ctx->k2_pipe = depth;
    return CB_OK;
}

int cb_spmm_k2_config(cb_ctx* ctx, int slab_bytes, int point) {
    if (!ctx) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_spmm_k2_config: null ctx");
    if (slab_bytes < 0 || (slab_bytes != 0 && slab_bytes != 64 && slab_bytes != 128 && slab_bytes != 256 && slab_bytes != 512))
        return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_spmm_k2_config: slab of %d bytes (0 = automatic, 64, 128, 256 or 512)", slab_bytes);
    if (point < -1 || point > 1) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_spmm_k2_config: operating point %d (-1 automatic, 0 deep, 1 wide)", point);
    ctx->k2_slab_bytes = slab_bytes;
    ctx->k2_point = point;
    return CB_OK;
}

int cb_spmm_local(cb_ctx* ctx, const cb_tile* t, const cb_dense* X, cb_dense* Y, int semiring, int accumulate) {
    if (!ctx || !t || !X || !Y) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_spmm_local: null argument");
    // CheckSpGEMMCompliance (ParFriends.h:160-181): inner dimensions must agree
    if (X->rows != t->n || Y->rows != t->m || X->cols != Y->cols)
        return cb_fail(ctx, CB_ERR_DIMMISMATCH, "cb_spmm_local: A is %lld x %lld, X is %lld x %lld, Y is %lld x %lld",
                       (long long)t->m, (long long)t->n, (long long)X->rows, (long long)X->cols, (long long)Y->rows, (long long)Y->cols);
    if (X->dtype != Y->dtype) return cb_fail(ctx, CB_ERR_UNSUPPORTED, "cb_spmm_local: X and Y dtypes differ");
    if (X->ptr == Y->ptr) return cb_fail(ctx, CB_ERR_MATRIXALIAS, "cb_spmm_local: X and Y alias");
    return cb_spmm_launch(ctx, ctx->compute, t, X->ptr, X->ld, Y->ptr, Y->ld, X->cols, X->dtype, semiring, accumulate);
}

static int ws_reserve(cb_ctx* ctx, void** p, size_t* have, size_t need) {
    if (*have >= need) return CB_OK;
    if (*p) { CB_CUDA(ctx, cudaStreamSynchronize(ctx->compute)); CB_CUDA(ctx, cudaFree(*p)); *p = nullptr; *have = 0; }
    cudaError_t e = cudaMalloc(p, need);
    if (e != cudaSuccess) return cb_fail(ctx, CB_ERR_ALLOC, "cudaMalloc(%zu) for a host-path panel: %s", need, cudaGetErrorString(e));
    *have = need;
    return CB_OK;
}

int cb_spmm_host(cb_ctx* ctx, const cb_tile* t, const void* X_host, int64_t ldx, void* Y_host, int64_t ldy, int64_t k,
                 int dtype, int semiring) {
    // Host panels in, host panel out.  The panel is cut into column slabs that flow through three streams - slab s+1 goes up
    // (H2D) while slab s is multiplied and slab s-1 comes down (D2H) - so both PCIe directions and the kernel overlap;
    // the device panels are kept in the ctx between calls.
    if (!ctx || !t) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_spmm_host: null argument");
    const size_t es = cb_dtype_size(dtype);
    if (!es || k <= 0 || ldx < k || ldy < k) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_spmm_host: bad dtype / k / leading dimension");
    if ((t->n > 0 && !X_host) || (t->m > 0 && !Y_host)) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_spmm_host: null panel");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t per16 = 16 / (int64_t)es;
    const int64_t ld = (k + per16 - 1) / per16 * per16;
    CB_TRY(ws_reserve(ctx, &ctx->ws_x, &ctx->ws_x_bytes, (size_t)(t->n > 0 ? t->n : 1) * (size_t)ld * es));
    CB_TRY(ws_reserve(ctx, &ctx->ws_y, &ctx->ws_y_bytes, (size_t)(t->m > 0 ? t->m : 1) * (size_t)ld * es));
    if (!ctx->h2d) {
        CB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->h2d, cudaStreamNonBlocking));
        CB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->d2h, cudaStreamNonBlocking));
        for (int i = 0; i < CB_MAX_SLABS; ++i) {
            CB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->slab_up[i], cudaEventDisableTiming));
            CB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->slab_done[i], cudaEventDisableTiming));
        }
        CB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->host_begin, cudaEventDisableTiming));
    }
    // slabs of whole 16-byte vectors, at least 256 bytes of every row per slab (measured: narrower 2D copies run PCIe
    // at ~70% - 305 ms vs 257 ms per 17 GB round trip on C3)
    static const int want = getenv("CB_HOST_SLABS") ? atoi(getenv("CB_HOST_SLABS")) : 4;
    int nslab = 1;
    if (ld == k) {
        const int64_t min_cols = std::max<int64_t>(per16, 256 / (int64_t)es);
        nslab = (int)std::min<int64_t>(std::min<int64_t>(want, CB_MAX_SLABS), std::max<int64_t>(1, k / min_cols));
    }
    const int64_t slab_cols = ((k + nslab - 1) / nslab + per16 - 1) / per16 * per16;
    if (ld != k) CB_CUDA(ctx, cudaMemsetAsync(ctx->ws_x, 0, (size_t)t->n * (size_t)ld * es, ctx->compute));
    CB_CUDA(ctx, cudaEventRecord(ctx->host_begin, ctx->compute));       // earlier work on the panels has finished
    CB_CUDA(ctx, cudaStreamWaitEvent(ctx->h2d, ctx->host_begin, 0));
    CB_CUDA(ctx, cudaStreamWaitEvent(ctx->d2h, ctx->host_begin, 0));
    int used = 0;
    for (int64_t c0 = 0; c0 < k; c0 += slab_cols, ++used) {
        const int64_t cw = std::min(slab_cols, k - c0);
        char* dx = (char*)ctx->ws_x + (size_t)c0 * es;
        char* dy = (char*)ctx->ws_y + (size_t)c0 * es;
        if (t->n) CB_CUDA(ctx, cudaMemcpy2DAsync(dx, (size_t)ld * es, (const char*)X_host + (size_t)c0 * es, (size_t)ldx * es, (size_t)cw * es,
                                                 (size_t)t->n, cudaMemcpyHostToDevice, ctx->h2d));
        CB_CUDA(ctx, cudaEventRecord(ctx->slab_up[used], ctx->h2d));
        CB_CUDA(ctx, cudaStreamWaitEvent(ctx->compute, ctx->slab_up[used], 0));
        CB_TRY(cb_spmm_launch(ctx, ctx->compute, t, dx, ld, dy, ld, cw, dtype, semiring, 0));
        CB_CUDA(ctx, cudaEventRecord(ctx->slab_done[used], ctx->compute));
        CB_CUDA(ctx, cudaStreamWaitEvent(ctx->d2h, ctx->slab_done[used], 0));
        if (t->m) CB_CUDA(ctx, cudaMemcpy2DAsync((char*)Y_host + (size_t)c0 * es, (size_t)ldy * es, dy, (size_t)ld * es, (size_t)cw * es,
                                                 (size_t)t->m, cudaMemcpyDeviceToHost, ctx->d2h));
    }
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->d2h));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
    return CB_OK;
}

}  // extern "C"
