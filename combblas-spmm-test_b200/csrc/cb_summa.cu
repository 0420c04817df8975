// 2D SUMMA stage loop over a pr x pc process grid, one process per GPU, NCCL over NVLink.
//
// Replaces Mult_AnXBn_Synch / Mult_AnXBn_Overlap (reference include/CombBLAS/ParFriends.h:1004-1108, :1110-1235)
// and their transport SpParHelper::BCastMatrix / GetSetSizes (SpParHelper.cpp:582-601, :797-809):
//   * the reference ships a tile as 3-4 MPI_Bcast calls of host arrays per stage; here a tile is ONE device
//     allocation (cb_tile_layout) and travels as one ncclBroadcast on the row communicator, the dense panel as one
//     ncclBroadcast on the column communicator, both on the communication stream;
//   * receive buffers are double buffered and the broadcasts of stage s+1 are enqueued before the local multiply
//     of stage s has run, so transfer and kernel overlap (events hand buffers between the two streams);
//   * the stage partials are never materialised: K2 accumulates into the stationary Y tile
//     (the role of MultiwayMerge, MultiwayMerge.h:411-526);
//   * the inner dimension is cut at the union of A's column-block and X's row-block boundaries, so any pr x pc
//     works (the reference's ProductGrid demands a square grid, src/CommGrid.cpp:164-180).  On a square grid with
//     divisible n this is exactly the reference's `stages = grcols` loop.
// NCCL is loaded with dlopen at the first multi-rank context so single-GPU use has no NCCL dependency.
#include <dlfcn.h>
#include <algorithm>
#include <cub/cub.cuh>
#include "cb_common.cuh"
#include "cb_hub.cuh"
#include "cb_spgemm.cuh"

// ------------------------------------------------------------------------------------------ NCCL, by hand
namespace {

typedef void* ncclComm_t;
struct ncclUniqueId { char internal[128]; };
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0 };

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, void*) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string error;
};

NcclApi& nccl() {
    static NcclApi api;
    return api;
}

bool nccl_load() {
    NcclApi& a = nccl();
    if (a.handle) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (a.handle) break;
    }
    if (!a.handle) { a.error = std::string("dlopen(libnccl.so.2): ") + dlerror(); return false; }
#define SYM(field, name)                                                    \
    *(void**)(&a.field) = dlsym(a.handle, name);                            \
    if (!a.field) { a.error = std::string("missing symbol ") + name; a.handle = nullptr; return false; }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommSplit, "ncclCommSplit")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(Broadcast, "ncclBroadcast")
    SYM(AllGather, "ncclAllGather")
    SYM(AllReduce, "ncclAllReduce")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    return true;
}

#define CB_NCCL(ctx, expr)                                                                                  \
    do {                                                                                                    \
        int r__ = (expr);                                                                                   \
        if (r__ != ncclSuccess)                                                                             \
            return cb_fail((ctx), CB_ERR_NCCL, "%s failed: %s (%s:%d)", #expr, nccl().GetErrorString(r__), __FILE__, __LINE__); \
    } while (0)

// first index and length of block b of nb over `total` (reference Owner rule, SpParMat.cpp:5066-5096)
inline void block_range(int64_t total, int nb, int b, int64_t* start, int64_t* len) {
    const int64_t per = total / nb;
    *start = (int64_t)b * per;
    *len = (b == nb - 1) ? total - *start : per;
}

struct SummaState {
    int64_t gn = -1, kl = -1;
    int dtype = -1;
    int nstages = 0;
    std::vector<int64_t> seg;          // nstages+1 boundaries of the inner dimension
    std::vector<int> a_root, x_root;   // rank inside the row / column communicator that owns the stage's operand
    char* slotA[2] = {nullptr, nullptr};
    size_t slotA_bytes = 0;
    char* slotX[2] = {nullptr, nullptr};
    size_t slotX_bytes = 0;
    cb_tile* view[2] = {nullptr, nullptr};
    cudaEvent_t ready[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr}, begin = nullptr, end = nullptr;
    std::vector<cudaEvent_t> comm_ev;  // begin/end pairs per stage on the communication stream
    std::vector<char> comm_used;       // stage s used the communication stream in the last call
    // metas of every stage's A part as seen by this rank's row communicator, valid for `meta_tile`
    const cb_tile* meta_tile = nullptr;
    uint64_t meta_uid = 0;
    std::vector<cb_tile_meta> metas;
};

}  // namespace

int cb_nccl_init(cb_ctx* ctx, const void* id128) {
    if (!nccl_load()) return cb_fail(ctx, CB_ERR_NCCL, "NCCL unavailable: %s", nccl().error.c_str());
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    ncclComm_t world = nullptr, row = nullptr, col = nullptr;
    CB_NCCL(ctx, nccl().CommInitRank(&world, ctx->nranks, id, ctx->rank));
    // "RowWorld": same processor row, ordered by column; "ColWorld": same processor column (src/CommGrid.cpp:66-67)
    CB_NCCL(ctx, nccl().CommSplit(world, ctx->myprocrow, ctx->myproccol, &row, nullptr));
    CB_NCCL(ctx, nccl().CommSplit(world, ctx->myproccol, ctx->myprocrow, &col, nullptr));
    ctx->nccl_world = world; ctx->nccl_row = row; ctx->nccl_col = col;
    return CB_OK;
}

void cb_nccl_destroy(cb_ctx* ctx) {
    if (!nccl().handle) return;
    if (ctx->nccl_row) nccl().CommDestroy((ncclComm_t)ctx->nccl_row);
    if (ctx->nccl_col) nccl().CommDestroy((ncclComm_t)ctx->nccl_col);
    if (ctx->nccl_world) nccl().CommDestroy((ncclComm_t)ctx->nccl_world);
    ctx->nccl_row = ctx->nccl_col = ctx->nccl_world = nullptr;
}

void cb_summa_release(cb_ctx* ctx) {
    SummaState* s = (SummaState*)ctx->summa_state;
    if (!s) return;
    for (int i = 0; i < 2; ++i) {
        cudaFree(s->slotA[i]); cudaFree(s->slotX[i]);
        if (s->view[i]) { cudaFree(s->view[i]->carry); cb_hub_release(s->view[i]); delete s->view[i]; }
        if (s->ready[i]) cudaEventDestroy(s->ready[i]);
        if (s->done[i]) cudaEventDestroy(s->done[i]);
    }
    if (s->begin) cudaEventDestroy(s->begin);
    if (s->end) cudaEventDestroy(s->end);
    for (cudaEvent_t e : s->comm_ev) cudaEventDestroy(e);
    delete s;
    ctx->summa_state = nullptr;
}

// ------------------------------------------------------------------------------------------ column slices of a tile
namespace {

__global__ void slice_keys_kernel(const int32_t* __restrict__ colflag, const int32_t* __restrict__ rowptr,
                                  const int32_t* __restrict__ nzrows, int64_t nzr, int64_t nz, int32_t c0, int32_t c1,
                                  uint64_t* __restrict__ keys, uint8_t* __restrict__ keep) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nz; p += (int64_t)gridDim.x * blockDim.x) {
        const int32_t col = colflag[p] & 0x7fffffff;
        const bool in = col >= c0 && col < c1;
        keep[p] = in ? 1 : 0;
        if (in) {
            int64_t lo = 0, hi = nzr;           // largest ridx with rowptr[ridx] <= p
            while (hi - lo > 1) { int64_t mid = (lo + hi) >> 1; if (rowptr[mid] <= p) lo = mid; else hi = mid; }
            keys[p] = ((uint64_t)(uint32_t)nzrows[lo] << 32) | (uint64_t)(uint32_t)(col - c0);
        } else {
            keys[p] = ~0ULL;
        }
    }
}

inline int grid_for(int64_t n, int sm) {
    int64_t b = (n + 255) / 256, cap = (int64_t)sm * 32;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// columns [c0, c1) of `t` as a new tile with local column indices
int slice_cols(cb_ctx* ctx, const cb_tile* t, int64_t c0, int64_t c1, cb_tile** out) {
    cb_scratch sc;
    cudaStream_t st = ctx->compute;
    const int64_t nz = t->nnz;
    uint64_t *keys = nullptr, *keys_sel = nullptr;
    uint8_t* keep = nullptr;
    int64_t* d_n = nullptr;
    void* vals_sel = nullptr;
    const size_t vs = cb_dtype_size(t->val_dtype);
    int64_t nsel = 0;
    CB_CUDA(ctx, sc.alloc(&keys, (size_t)nz));
    CB_CUDA(ctx, sc.alloc(&keys_sel, (size_t)nz));
    if (nz > 0) {
        CB_CUDA(ctx, sc.alloc(&keep, (size_t)nz));
        CB_CUDA(ctx, sc.alloc(&d_n, 1));
        slice_keys_kernel<<<grid_for(nz, ctx->sm_count), 256, 0, st>>>(t->colflag, t->rowptr, t->nzrows, t->nzr, nz, (int32_t)c0, (int32_t)c1, keys, keep);
        CB_LAUNCHED(ctx);
        size_t b = 0;
        CB_CUDA(ctx, cub::DeviceSelect::Flagged(nullptr, b, keys, keep, keys_sel, d_n, (int)nz, st));
        void* tmp; CB_CUDA(ctx, sc.alloc((char**)&tmp, b));
        CB_CUDA(ctx, cub::DeviceSelect::Flagged(tmp, b, keys, keep, keys_sel, d_n, (int)nz, st));
        if (t->vals) {
            CB_CUDA(ctx, sc.alloc((char**)&vals_sel, vs * (size_t)nz));
            size_t b2 = 0;
            void* tmp2 = nullptr;
            switch (vs) {
                case 1:
                    CB_CUDA(ctx, cub::DeviceSelect::Flagged(nullptr, b2, (const uint8_t*)t->vals, keep, (uint8_t*)vals_sel, d_n, (int)nz, st));
                    CB_CUDA(ctx, sc.alloc((char**)&tmp2, b2));
                    CB_CUDA(ctx, cub::DeviceSelect::Flagged(tmp2, b2, (const uint8_t*)t->vals, keep, (uint8_t*)vals_sel, d_n, (int)nz, st));
                    break;
                case 4:
                    CB_CUDA(ctx, cub::DeviceSelect::Flagged(nullptr, b2, (const uint32_t*)t->vals, keep, (uint32_t*)vals_sel, d_n, (int)nz, st));
                    CB_CUDA(ctx, sc.alloc((char**)&tmp2, b2));
                    CB_CUDA(ctx, cub::DeviceSelect::Flagged(tmp2, b2, (const uint32_t*)t->vals, keep, (uint32_t*)vals_sel, d_n, (int)nz, st));
                    break;
                default:
                    CB_CUDA(ctx, cub::DeviceSelect::Flagged(nullptr, b2, (const uint64_t*)t->vals, keep, (uint64_t*)vals_sel, d_n, (int)nz, st));
                    CB_CUDA(ctx, sc.alloc((char**)&tmp2, b2));
                    CB_CUDA(ctx, cub::DeviceSelect::Flagged(tmp2, b2, (const uint64_t*)t->vals, keep, (uint64_t*)vals_sel, d_n, (int)nz, st));
            }
        }
        ctx->launches += 4;
        CB_CUDA(ctx, cudaMemcpyAsync(&nsel, d_n, sizeof nsel, cudaMemcpyDeviceToHost, st));
        CB_CUDA(ctx, cudaStreamSynchronize(st));
    }
    return cb_tile_build_from_keys(ctx, t->m, c1 - c0, nsel, keys_sel, (t->vals && nsel) ? vals_sel : nullptr, t->val_dtype, true, sc, out);
}

// the nonzeros of `t` whose column is marked in `keep_col` (one byte per column), as a new tile of the same shape: what a
// sparse right-hand side leaves of A when most rows of B are empty (cb_tile_filter_columns).  Same steps as slice_cols.
__global__ void filter_keys_kernel(const int32_t* __restrict__ colflag, const int32_t* __restrict__ rowptr,
                                   const int32_t* __restrict__ nzrows, int64_t nzr, int64_t nz, const uint8_t* __restrict__ keep_col,
                                   uint64_t* __restrict__ keys, uint8_t* __restrict__ keep) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nz; p += (int64_t)gridDim.x * blockDim.x) {
        const int32_t col = colflag[p] & 0x7fffffff;
        const bool in = keep_col[col] != 0;
        keep[p] = in ? 1 : 0;
        if (in) {
            int64_t lo = 0, hi = nzr;           // largest ridx with rowptr[ridx] <= p
            while (hi - lo > 1) { int64_t mid = (lo + hi) >> 1; if (rowptr[mid] <= p) lo = mid; else hi = mid; }
            keys[p] = ((uint64_t)(uint32_t)nzrows[lo] << 32) | (uint64_t)(uint32_t)col;
        } else {
            keys[p] = ~0ULL;
        }
    }
}

template <typename V>
int select_values(cb_ctx* ctx, cb_scratch& sc, const void* vals, const uint8_t* keep, void* out, int64_t* d_n, int64_t nz, cudaStream_t st) {
    size_t b = 0;
    void* tmp = nullptr;
    CB_CUDA(ctx, cub::DeviceSelect::Flagged(nullptr, b, (const V*)vals, keep, (V*)out, d_n, (int)nz, st));
    CB_CUDA(ctx, sc.alloc((char**)&tmp, b));
    CB_CUDA(ctx, cub::DeviceSelect::Flagged(tmp, b, (const V*)vals, keep, (V*)out, d_n, (int)nz, st));
    return CB_OK;
}

int filter_cols(cb_ctx* ctx, const cb_tile* t, const uint8_t* keep_col_host, cb_tile** out) {
    cb_scratch sc;
    cudaStream_t st = ctx->compute;
    const int64_t nz = t->nnz;
    uint64_t *keys = nullptr, *keys_sel = nullptr;
    uint8_t *keep = nullptr, *keep_col = nullptr;
    int64_t* d_n = nullptr;
    void* vals_sel = nullptr;
    const size_t vs = cb_dtype_size(t->val_dtype);
    int64_t nsel = 0;
    CB_CUDA(ctx, sc.alloc(&keys, (size_t)nz));
    CB_CUDA(ctx, sc.alloc(&keys_sel, (size_t)nz));
    if (nz > 0) {
        CB_CUDA(ctx, sc.alloc(&keep, (size_t)nz));
        CB_CUDA(ctx, sc.alloc(&keep_col, (size_t)t->n));
        CB_CUDA(ctx, sc.alloc(&d_n, 1));
        CB_CUDA(ctx, cudaMemcpyAsync(keep_col, keep_col_host, (size_t)t->n, cudaMemcpyHostToDevice, st));
        filter_keys_kernel<<<grid_for(nz, ctx->sm_count), 256, 0, st>>>(t->colflag, t->rowptr, t->nzrows, t->nzr, nz, keep_col, keys, keep);
        CB_LAUNCHED(ctx);
        CB_CUDA(ctx, cudaGetLastError());
        CB_TRY(select_values<uint64_t>(ctx, sc, keys, keep, keys_sel, d_n, nz, st));
        if (t->vals) {
            CB_CUDA(ctx, sc.alloc((char**)&vals_sel, vs * (size_t)nz));
            if (vs == 1) CB_TRY(select_values<uint8_t>(ctx, sc, t->vals, keep, vals_sel, d_n, nz, st));
            else if (vs == 4) CB_TRY(select_values<uint32_t>(ctx, sc, t->vals, keep, vals_sel, d_n, nz, st));
            else CB_TRY(select_values<uint64_t>(ctx, sc, t->vals, keep, vals_sel, d_n, nz, st));
        }
        ctx->launches += 4;
        CB_CUDA(ctx, cudaMemcpyAsync(&nsel, d_n, sizeof nsel, cudaMemcpyDeviceToHost, st));
        CB_CUDA(ctx, cudaStreamSynchronize(st));
    }
    return cb_tile_build_from_keys(ctx, t->m, t->n, nsel, keys_sel, (t->vals && nsel) ? vals_sel : nullptr, t->val_dtype, true, sc, out);
}

// all nonzeros of `t` as keys row<<32 | (col + coloff), for concatenating column-disjoint parts of one block-row
__global__ void emit_keys_kernel(const int32_t* __restrict__ colflag, const int32_t* __restrict__ rowptr,
                                 const int32_t* __restrict__ nzrows, int64_t nzr, int64_t nz, int32_t coloff,
                                 uint64_t* __restrict__ keys) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nz; p += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = nzr;           // largest ridx with rowptr[ridx] <= p
        while (hi - lo > 1) { int64_t mid = (lo + hi) >> 1; if (rowptr[mid] <= p) lo = mid; else hi = mid; }
        keys[p] = ((uint64_t)(uint32_t)nzrows[lo] << 32) | (uint64_t)(uint32_t)((colflag[p] & 0x7fffffff) + coloff);
    }
}

// one tile from column-disjoint parts of the same rows; part i's columns are shifted by coloff[i]
int merge_parts(cb_ctx* ctx, const std::vector<const cb_tile*>& parts, const std::vector<int64_t>& coloff, int64_t m, int64_t n,
                int val_dtype, cb_tile** out) {
    cb_scratch sc;
    cudaStream_t st = ctx->compute;
    int64_t nz = 0;
    for (const cb_tile* p : parts) nz += p->nnz;
    if (nz >= (int64_t(1) << 31)) return cb_fail(ctx, CB_ERR_TOO_LARGE, "merged block-row part has %lld nonzeros", (long long)nz);
    const size_t vs = cb_dtype_size(val_dtype);
    uint64_t* keys = nullptr;
    char* vals = nullptr;
    CB_CUDA(ctx, sc.alloc(&keys, (size_t)nz));
    if (vs) CB_CUDA(ctx, sc.alloc(&vals, vs * (size_t)nz));
    int64_t off = 0;
    for (size_t i = 0; i < parts.size(); ++i) {
        const cb_tile* p = parts[i];
        if (p->nnz == 0) continue;
        emit_keys_kernel<<<grid_for(p->nnz, ctx->sm_count), 256, 0, st>>>(p->colflag, p->rowptr, p->nzrows, p->nzr, p->nnz, (int32_t)coloff[i], keys + off);
        CB_LAUNCHED(ctx);
        if (vs) CB_CUDA(ctx, cudaMemcpyAsync(vals + vs * (size_t)off, p->vals, vs * (size_t)p->nnz, cudaMemcpyDeviceToDevice, st));
        off += p->nnz;
    }
    CB_CUDA(ctx, cudaGetLastError());
    return cb_tile_build_from_keys(ctx, m, n, nz, keys, (vs && nz) ? vals : nullptr, val_dtype, false, sc, out);
}

}  // namespace

extern "C" {

int cb_tile_filter_columns(cb_ctx* ctx, const cb_tile* tile, const uint8_t* keep_cols, cb_tile** out) {
    if (!ctx || !tile || !keep_cols || !out) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_tile_filter_columns: null argument");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    return filter_cols(ctx, tile, keep_cols, out);
}

int cb_comm_unique_id(void* id128) {
    if (!nccl_load()) return cb_fail(nullptr, CB_ERR_NCCL, "NCCL unavailable: %s", nccl().error.c_str());
    ncclUniqueId id;
    CB_NCCL(nullptr, nccl().GetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
    return CB_OK;
}

// Stage plan, pure host arithmetic (exported so the host logic is testable without a GPU).
// Cuts [0, gn) at every boundary of A's column blocks (pc of them) and of X's row blocks (pr of them).
// seg has room for pr + pc entries; a_owner_col[s] / x_owner_row[s] say which block owns stage s.
int cb_summa_plan(int pr, int pc, int64_t gn, int64_t* seg, int* a_owner_col, int* x_owner_row, int* nstages) {
    if (pr < 1 || pc < 1 || gn < 0) return CB_ERR_INVALIDPARAMS;
    std::vector<int64_t> cuts;
    for (int b = 0; b < pc; ++b) { int64_t s, l; block_range(gn, pc, b, &s, &l); cuts.push_back(s); }
    for (int b = 0; b < pr; ++b) { int64_t s, l; block_range(gn, pr, b, &s, &l); cuts.push_back(s); }
    cuts.push_back(gn);
    std::sort(cuts.begin(), cuts.end());
    cuts.erase(std::unique(cuts.begin(), cuts.end()), cuts.end());
    int ns = 0;
    for (size_t i = 0; i + 1 < cuts.size(); ++i) {
        const int64_t a = cuts[i], b = cuts[i + 1];
        if (b <= a) continue;
        // owner of global inner index a under the floor rule: min(a / per, nb - 1), all to the last block if per == 0
        const int64_t perc = gn / pc, perr = gn / pr;
        const int oc = perc ? (int)std::min<int64_t>(a / perc, pc - 1) : pc - 1;
        const int orow = perr ? (int)std::min<int64_t>(a / perr, pr - 1) : pr - 1;
        seg[ns] = a;
        a_owner_col[ns] = oc;
        x_owner_row[ns] = orow;
        ++ns;
    }
    seg[ns] = gn;
    *nstages = ns;
    return CB_OK;
}

int cb_spmm_summa(cb_ctx* ctx, const cb_tile* tile, const cb_dense* X, cb_dense* Y, int semiring, int64_t gm, int64_t gn, int64_t gk) {
    if (!ctx || !tile || !X || !Y) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_spmm_summa: null argument");
    const int pr = ctx->pr, pc = ctx->pc;
    int64_t r0, rl, c0, cl, x0, xl, k0, kl;
    block_range(gm, pr, ctx->myprocrow, &r0, &rl);
    block_range(gn, pc, ctx->myproccol, &c0, &cl);
    block_range(gn, pr, ctx->myprocrow, &x0, &xl);
    block_range(gk >= 0 ? gk : 0, pc, ctx->myproccol, &k0, &kl);
    if (gk < 0) kl = X->cols;        // internal (cb_spmm_summa_host): a column slab; the ranks of a processor column agree on its width
    // CheckSpGEMMCompliance (ParFriends.h:160-181) on the local blocks
    if (tile->m != rl || tile->n != cl || X->rows != xl || X->cols != kl || Y->rows != rl || Y->cols != kl)
        return cb_fail(ctx, CB_ERR_DIMMISMATCH,
                       "cb_spmm_summa rank %d (%d,%d) of %dx%d: local A %lldx%lld (want %lldx%lld), X %lldx%lld (want %lldx%lld), Y %lldx%lld (want %lldx%lld)",
                       ctx->rank, ctx->myprocrow, ctx->myproccol, pr, pc, (long long)tile->m, (long long)tile->n, (long long)rl, (long long)cl,
                       (long long)X->rows, (long long)X->cols, (long long)xl, (long long)kl, (long long)Y->rows, (long long)Y->cols, (long long)rl, (long long)kl);
    if (X->dtype != Y->dtype) return cb_fail(ctx, CB_ERR_UNSUPPORTED, "cb_spmm_summa: X and Y dtypes differ");
    if (X->ptr == Y->ptr) return cb_fail(ctx, CB_ERR_MATRIXALIAS, "cb_spmm_summa: X and Y alias");
    if (ctx->nranks == 1) return cb_spmm_local(ctx, tile, X, Y, semiring, 0);
    if (gn == 0) {                         // empty inner dimension: no stage runs, the product is SR::id() everywhere (as on one rank)
        unsigned char idv[8] = {0};
        if (cb_semiring_id(semiring, X->dtype, idv) != CB_OK) return cb_fail(ctx, CB_ERR_UNSUPPORTED, "cb_spmm_summa: semiring %d with dtype %d", semiring, X->dtype);
        return cb_dense_fill(Y, idv);
    }
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t es = cb_dtype_size(X->dtype);

    SummaState* S = (SummaState*)ctx->summa_state;
    if (!S) {
        S = new SummaState();
        ctx->summa_state = S;
        for (int i = 0; i < 2; ++i) {
            CB_CUDA(ctx, cudaEventCreateWithFlags(&S->ready[i], cudaEventDisableTiming));
            CB_CUDA(ctx, cudaEventCreateWithFlags(&S->done[i], cudaEventDisableTiming));
            S->view[i] = new cb_tile();
            S->view[i]->ctx = ctx;
        }
        CB_CUDA(ctx, cudaEventCreate(&S->begin));
        CB_CUDA(ctx, cudaEventCreate(&S->end));
    }
    // ---- plan
    if (S->gn != gn) {
        S->seg.assign(pr + pc + 1, 0); S->a_root.assign(pr + pc, 0); S->x_root.assign(pr + pc, 0);
        cb_summa_plan(pr, pc, gn, S->seg.data(), S->a_root.data(), S->x_root.data(), &S->nstages);
        S->gn = gn;
        S->meta_tile = nullptr;
        while ((int)S->comm_ev.size() < 2 * S->nstages) { cudaEvent_t e; CB_CUDA(ctx, cudaEventCreate(&e)); S->comm_ev.push_back(e); }
    }
    const int ns = S->nstages;
    S->comm_used.assign(ns, 0);
    // ---- my parts of A: one column slice per stage this rank roots (cached on the tile)
    cb_tile* mt = const_cast<cb_tile*>(tile);
    const int64_t key[3] = {(int64_t)pr * 1000 + pc, gn, ns};
    if (mt->summa_key[0] != key[0] || mt->summa_key[1] != key[1] || mt->summa_key[2] != key[2]) {
        for (cb_tile* p : mt->summa_parts) cb_tile_free(p);
        for (cb_tile* p : mt->summa_remote) cb_tile_free(p);
        mt->summa_remote.clear();
        for (cb_tile* p : mt->summa_merged) cb_tile_free(p);
        mt->summa_merged.clear();
        for (cb_tile* p : mt->spgemm_remote) cb_tile_free(p);
        mt->spgemm_remote.clear();
        mt->spgemm_sent.clear();
        mt->summa_parts.assign(ns, nullptr);
        for (int s = 0; s < ns; ++s) {
            if (S->a_root[s] != ctx->myproccol) continue;
            const int64_t a = S->seg[s] - c0, b = S->seg[s + 1] - c0;
            if (a == 0 && b == cl) continue;                       // the whole tile: no copy
            CB_TRY(slice_cols(ctx, tile, a, b, &mt->summa_parts[s]));
        }
        mt->summa_key[0] = key[0]; mt->summa_key[1] = key[1]; mt->summa_key[2] = key[2];
        S->meta_tile = nullptr;
    }
    auto my_part = [&](int s) -> const cb_tile* { return mt->summa_parts[s] ? mt->summa_parts[s] : tile; };
    // ---- sizes of every stage's A part in my processor row (GetSetSizes, SpParHelper.cpp:797-809): one allgather
    if (S->meta_tile != tile || S->meta_uid != tile->uid) {
        S->metas.assign(ns, cb_tile_meta());
        if (pc > 1) {
            std::vector<cb_tile_meta> mine(ns);
            memset(mine.data(), 0, sizeof(cb_tile_meta) * ns);
            for (int s = 0; s < ns; ++s) if (S->a_root[s] == ctx->myproccol) mine[s] = cb_tile_get_meta(my_part(s));
            cb_scratch sc;
            char *d_send, *d_recv;
            const size_t bytes = sizeof(cb_tile_meta) * (size_t)ns;
            CB_CUDA(ctx, sc.alloc(&d_send, bytes));
            CB_CUDA(ctx, sc.alloc(&d_recv, bytes * pc));
            CB_CUDA(ctx, cudaMemcpyAsync(d_send, mine.data(), bytes, cudaMemcpyHostToDevice, ctx->comm));
            CB_NCCL(ctx, nccl().AllGather(d_send, d_recv, bytes, ncclInt8, (ncclComm_t)ctx->nccl_row, ctx->comm));
            std::vector<cb_tile_meta> all((size_t)ns * pc);
            CB_CUDA(ctx, cudaMemcpyAsync(all.data(), d_recv, bytes * pc, cudaMemcpyDeviceToHost, ctx->comm));
            CB_CUDA(ctx, cudaStreamSynchronize(ctx->comm));
            for (int s = 0; s < ns; ++s) S->metas[s] = all[(size_t)S->a_root[s] * ns + s];
        } else {
            for (int s = 0; s < ns; ++s) S->metas[s] = cb_tile_get_meta(my_part(s));
        }
        S->meta_tile = tile;
        S->meta_uid = tile->uid;
    }
    // ---- A parts of my row neighbours: kept on the device after the first multiply with this tile (tiles are immutable)
    const bool cache = ctx->summa_cache_a && pc > 1;
    bool cached = cache && (int)mt->summa_remote.size() == ns;
    if (cache && !cached) {
        mt->summa_remote.assign(ns, nullptr);
        for (int s = 0; s < ns; ++s) {
            if (S->a_root[s] == ctx->myproccol) continue;
            cb_tile* v = new cb_tile();
            v->ctx = ctx;
            const size_t bytes = cb_layout(S->metas[s]).total;
            if (cudaMalloc((void**)&v->slab, bytes) != cudaSuccess) {
                delete v;
                for (cb_tile* p : mt->summa_remote) cb_tile_free(p);
                mt->summa_remote.clear();
                return cb_fail(ctx, CB_ERR_ALLOC, "cudaMalloc(%zu) for a cached SUMMA tile part", bytes);
            }
            v->slab_bytes = bytes; v->owns_slab = true;
            cb_tile_bind(v, S->metas[s], v->slab);
            mt->summa_remote[s] = v;
        }
    }
    // ---- panel transport: copy-engine pushes into peer memory (cb_p2p.cu) unless CB_SUMMA_TRANSPORT=nccl
    if (pr > 1 && ctx->summa_p2p) CB_TRY(cb_p2p_prepare(ctx, (size_t)gn * (size_t)X->ld * es));    // may switch the column to NCCL
    const bool p2p = pr > 1 && ctx->summa_p2p;

    // ---- steady state with a resident block-row: once every part of my block-row of A is on this GPU (second multiply with
    // the same tile onwards) the parts that meet the same X row block are fused into one tile, so a multiply is pr kernel
    // launches over full-length rows instead of one launch per stage over row fragments.
    if (cached && ctx->summa_merge && (p2p || pr == 1)) {
        if ((int)mt->summa_merged.size() != pr) {
            for (cb_tile* p : mt->summa_merged) cb_tile_free(p);
            mt->summa_merged.assign(pr, nullptr);
            for (int r = 0; r < pr; ++r) {
                int64_t b0, bl;
                block_range(gn, pr, r, &b0, &bl);
                std::vector<const cb_tile*> parts;
                std::vector<int64_t> offs;
                for (int s = 0; s < ns; ++s) {
                    if (S->x_root[s] != r) continue;
                    parts.push_back(S->a_root[s] == ctx->myproccol ? my_part(s) : mt->summa_remote[s]);
                    offs.push_back(S->seg[s] - b0);
                }
                if (parts.size() == 1 && offs[0] == 0 && parts[0]->n == bl) continue;      // a single part is already the tile
                CB_TRY(merge_parts(ctx, parts, offs, rl, bl, tile->val_dtype, &mt->summa_merged[r]));
            }
        }
        CB_CUDA(ctx, cudaEventRecord(S->begin, ctx->compute));
        if (p2p) {
            CB_TRY(cb_p2p_begin(ctx));
            for (int s = 0; s < ns; ++s) {
                const int64_t seg_len = S->seg[s + 1] - S->seg[s];
                if (S->x_root[s] != ctx->myprocrow || seg_len <= 0 || kl <= 0) continue;
                CB_TRY(cb_p2p_push(ctx, S->begin, s, (size_t)S->seg[s] * (size_t)X->ld * es,
                                   (const char*)X->ptr + (size_t)(S->seg[s] - x0) * (size_t)X->ld * es, (size_t)seg_len * (size_t)X->ld * es));
            }
        }
        bool wrote_m = false;
        for (int oi = 0; oi < pr; ++oi) {
            const int r = oi == 0 ? ctx->myprocrow : (oi <= ctx->myprocrow ? oi - 1 : oi);     // my own X block first
            int64_t b0, bl;
            block_range(gn, pr, r, &b0, &bl);
            const cb_tile* part = mt->summa_merged[r];
            if (!part) for (int s = 0; s < ns; ++s) if (S->x_root[s] == r) part = S->a_root[s] == ctx->myproccol ? my_part(s) : mt->summa_remote[s];
            if (r != ctx->myprocrow && kl > 0)
                for (int s = 0; s < ns; ++s) if (S->x_root[s] == r && S->seg[s + 1] > S->seg[s]) CB_TRY(cb_p2p_wait_stage(ctx, ctx->compute, s));
            const char* xsrc = r == ctx->myprocrow ? (const char*)X->ptr : cb_p2p_xfull(ctx) + (size_t)b0 * (size_t)X->ld * es;
            if (!part) continue;
            if (part->nnz > 0 || !wrote_m) {
                CB_TRY(cb_spmm_launch(ctx, ctx->compute, part, xsrc, X->ld, Y->ptr, Y->ld, kl, X->dtype, semiring, wrote_m ? 1 : 0));
                wrote_m = true;
            }
        }
        if (p2p) CB_TRY(cb_p2p_finish(ctx, ctx->compute));
        CB_CUDA(ctx, cudaEventRecord(S->end, ctx->compute));
        return CB_OK;
    }
    // ---- receive buffers of the NCCL paths
    size_t needA = 0, needX = 0;
    for (int s = 0; s < ns; ++s) {
        if (!cache && S->a_root[s] != ctx->myproccol) needA = std::max(needA, cb_layout(S->metas[s]).total);
        if (!p2p && S->x_root[s] != ctx->myprocrow) needX = std::max(needX, (size_t)(S->seg[s + 1] - S->seg[s]) * (size_t)X->ld * es);
    }
    if (needA > S->slotA_bytes || needX > S->slotX_bytes) {
        CB_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
        CB_CUDA(ctx, cudaStreamSynchronize(ctx->comm));
        for (int i = 0; i < 2; ++i) {
            if (needA > S->slotA_bytes) {
                cudaFree(S->slotA[i]); S->slotA[i] = nullptr;
                if (cudaMalloc((void**)&S->slotA[i], needA) != cudaSuccess) return cb_fail(ctx, CB_ERR_ALLOC, "cudaMalloc(%zu) for a SUMMA tile slot", needA);
            }
            if (needX > S->slotX_bytes) {
                cudaFree(S->slotX[i]); S->slotX[i] = nullptr;
                if (cudaMalloc((void**)&S->slotX[i], needX) != cudaSuccess) return cb_fail(ctx, CB_ERR_ALLOC, "cudaMalloc(%zu) for a SUMMA panel slot", needX);
            }
        }
        S->slotA_bytes = std::max(S->slotA_bytes, needA);
        S->slotX_bytes = std::max(S->slotX_bytes, needX);
    }

    // ---- stage loop
    CB_CUDA(ctx, cudaEventRecord(S->begin, ctx->compute));
    CB_CUDA(ctx, cudaStreamWaitEvent(ctx->comm, S->begin, 0));      // operands produced on the compute stream are ready
    if (p2p) {
        // all panels this rank owns leave right away, one DMA stream per column peer
        CB_TRY(cb_p2p_begin(ctx));
        for (int s = 0; s < ns; ++s) {
            const int64_t seg_len = S->seg[s + 1] - S->seg[s];
            if (S->x_root[s] != ctx->myprocrow || seg_len <= 0 || kl <= 0) continue;
            CB_TRY(cb_p2p_push(ctx, S->begin, s, (size_t)S->seg[s] * (size_t)X->ld * es,
                               (const char*)X->ptr + (size_t)(S->seg[s] - x0) * (size_t)X->ld * es, (size_t)seg_len * (size_t)X->ld * es));
        }
    }
    // Stage order.  With the peer transport nothing about X is collective, so a rank multiplies the stages whose panel it
    // owns first (no waiting) while the other panels are still in flight; ranks of one processor row share the order, which
    // keeps the row-communicator broadcasts of A matched.  (The merge is order independent: (+) is commutative; for
    // floating point the order is still fixed per rank, so results are reproducible.)
    std::vector<int> order;
    for (int s = 0; s < ns; ++s) if (p2p && S->x_root[s] == ctx->myprocrow) order.push_back(s);
    for (int s = 0; s < ns; ++s) if (!(p2p && S->x_root[s] == ctx->myprocrow)) order.push_back(s);
    bool wrote = false;
    for (int oi = 0; oi < ns; ++oi) {
        const int s = order[oi];
        const int slot = oi & 1;
        const int64_t seg_a = S->seg[s], seg_len = S->seg[s + 1] - S->seg[s];
        const bool a_mine = S->a_root[s] == ctx->myproccol, x_mine = S->x_root[s] == ctx->myprocrow;
        const cb_tile_meta& meta = S->metas[s];
        const bool bcast_a = pc > 1 && !cached;                    // also for an empty part: the receiver needs its empty-row list
        const bool bcast_x = !p2p && pr > 1 && seg_len > 0 && kl > 0;
        const cb_tile* part = a_mine ? my_part(s) : (cache ? mt->summa_remote[s] : S->view[slot]);
        const char* xsrc = x_mine ? (const char*)X->ptr + (size_t)(seg_a - x0) * (size_t)X->ld * es
                                  : (p2p ? cb_p2p_xfull(ctx) + (size_t)seg_a * (size_t)X->ld * es : S->slotX[slot]);
        if (bcast_a || bcast_x) {
            if (oi >= 2) CB_CUDA(ctx, cudaStreamWaitEvent(ctx->comm, S->done[slot], 0));     // slot free again
            S->comm_used[s] = 1;
            CB_CUDA(ctx, cudaEventRecord(S->comm_ev[2 * s], ctx->comm));
            CB_NCCL(ctx, nccl().GroupStart());
            if (bcast_a) {
                const size_t bytes = cb_layout(meta).total;
                char* buf = a_mine ? my_part(s)->slab : (cache ? mt->summa_remote[s]->slab : S->slotA[slot]);
                CB_NCCL(ctx, nccl().Broadcast(buf, buf, bytes, ncclInt8, S->a_root[s], (ncclComm_t)ctx->nccl_row, ctx->comm));
            }
            if (bcast_x) {
                const size_t bytes = (size_t)seg_len * (size_t)X->ld * es;
                char* buf = const_cast<char*>(xsrc);
                CB_NCCL(ctx, nccl().Broadcast(buf, buf, bytes, ncclInt8, S->x_root[s], (ncclComm_t)ctx->nccl_col, ctx->comm));
            }
            CB_NCCL(ctx, nccl().GroupEnd());
            CB_CUDA(ctx, cudaEventRecord(S->comm_ev[2 * s + 1], ctx->comm));
            CB_CUDA(ctx, cudaEventRecord(S->ready[slot], ctx->comm));
            CB_CUDA(ctx, cudaStreamWaitEvent(ctx->compute, S->ready[slot], 0));
        }
        if (p2p && !x_mine && seg_len > 0 && kl > 0) CB_TRY(cb_p2p_wait_stage(ctx, ctx->compute, s));
        if (!a_mine && !cache) {
            cb_tile* v = S->view[slot];
            void* keep_carry = v->carry; size_t keep_bytes = v->carry_bytes;
            cb_tile_bind(v, meta, S->slotA[slot]);
            v->carry = keep_carry; v->carry_bytes = keep_bytes;
        }
        static const bool skip_compute = getenv("CB_SUMMA_SKIP_COMPUTE") != nullptr;     // transport-only timing (debug)
        if (!skip_compute && (meta.nnz > 0 || !wrote)) {
            // a stage whose A part is empty still has to give Y its identity fill if nothing was written yet
            CB_TRY(cb_spmm_launch(ctx, ctx->compute, part, xsrc, X->ld, Y->ptr, Y->ld, kl, X->dtype, semiring, wrote ? 1 : 0));
            wrote = true;
        }
        CB_CUDA(ctx, cudaEventRecord(S->done[slot], ctx->compute));
    }
    if (p2p) CB_TRY(cb_p2p_finish(ctx, ctx->compute));
    CB_CUDA(ctx, cudaEventRecord(S->end, ctx->compute));
    return CB_OK;
}


// cb_spmm_summa with HOST panels: this rank's X tile comes from host memory and its Y tile goes back to host memory.
// The k-block is cut into column slabs that flow through three streams - slab s+1 goes up (H2D) while slab s runs the
// stage loop and slab s-1 comes down (D2H) - so both PCIe directions and the multiply overlap.  Every slab is its own
// compact device panel (leading dimension = slab width), because the panels that travel between column peers are whole
// contiguous blocks.  Collective: every rank of the grid calls it with its tiles of the same global product.
int cb_spmm_summa_host(cb_ctx* ctx, const cb_tile* tile, const void* X_host, int64_t ldx, void* Y_host, int64_t ldy, int semiring,
                       int64_t gm, int64_t gn, int64_t gk, int dtype) {
    if (!ctx || !tile) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_spmm_summa_host: null argument");
    const size_t es = cb_dtype_size(dtype);
    if (!es) return cb_fail(ctx, CB_ERR_UNSUPPORTED, "cb_spmm_summa_host: dtype %d", dtype);
    const int pr = ctx->pr, pc = ctx->pc;
    int64_t r0, rl, x0, xl, k0, kl;
    block_range(gm, pr, ctx->myprocrow, &r0, &rl);
    block_range(gn, pr, ctx->myprocrow, &x0, &xl);
    block_range(gk, pc, ctx->myproccol, &k0, &kl);
    if ((kl > 0 && ((xl > 0 && !X_host) || (rl > 0 && !Y_host))) || ldx < kl || ldy < kl)
        return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_spmm_summa_host: null panel or leading dimension below the local width %lld", (long long)kl);
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->h2d) {
        CB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->h2d, cudaStreamNonBlocking));
        CB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->d2h, cudaStreamNonBlocking));
        for (int i = 0; i < CB_MAX_SLABS; ++i) {
            CB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->slab_up[i], cudaEventDisableTiming));
            CB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->slab_done[i], cudaEventDisableTiming));
        }
        CB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->host_begin, cudaEventDisableTiming));
    }
    // Slab widths must be the same on every rank of a processor column (they exchange the panels) and a multiple of 16
    // bytes; every rank of the grid must run the same NUMBER of slabs (the stage loop is collective over processor rows
    // too).  Both follow from the floor rule: all ranks but the last processor column have per = gk / pc columns.
    const int64_t per16 = 16 / (int64_t)es;
    const int64_t per = pc > 1 ? gk / pc : gk;                       // width of every k-block but the last
    static const int want = getenv("CB_HOST_SLABS") ? atoi(getenv("CB_HOST_SLABS")) : 4;
    // slab width: 256 bytes of every row when the k-block has at least two of them, else 128 (profiles/r02_pcie_probe.jsonl: with
    // both directions busy 2D copies of 256-byte rows run at 48 / 50 GB/s, of 128-byte rows at 35 / 41, of 64-byte rows at 27 / 18)
    const int64_t min_cols = std::max<int64_t>(per16, (per * (int64_t)es >= 512 ? 256 : 128) / (int64_t)es);
    int nslab = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(want, CB_MAX_SLABS), per / min_cols));
    const int64_t slab_cols = per > 0 ? ((per + nslab - 1) / nslab + per16 - 1) / per16 * per16 : 0;
    if (slab_cols > 0) nslab = (int)((per + slab_cols - 1) / slab_cols);
    if (per == 0) nslab = 1;
    // slab i covers local columns [i*slab_cols, ...); the last slab takes whatever is left of THIS rank's block (the last
    // processor column holds the remainder of gk)
    const size_t xbytes = (size_t)std::max<int64_t>(xl, 1) * (size_t)((std::max<int64_t>(kl, 1) + per16 - 1) / per16 * per16 + (int64_t)nslab * per16) * es;
    const size_t ybytes = (size_t)std::max<int64_t>(rl, 1) * (size_t)((std::max<int64_t>(kl, 1) + per16 - 1) / per16 * per16 + (int64_t)nslab * per16) * es;
    auto reserve = [&](void** p, size_t* have, size_t need) -> int {
        if (*have >= need) return CB_OK;
        if (*p) { CB_CUDA(ctx, cudaStreamSynchronize(ctx->compute)); CB_CUDA(ctx, cudaFree(*p)); *p = nullptr; *have = 0; }
        cudaError_t e = cudaMalloc(p, need);
        if (e != cudaSuccess) return cb_fail(ctx, CB_ERR_ALLOC, "cudaMalloc(%zu) for a host-path panel: %s", need, cudaGetErrorString(e));
        *have = need;
        return CB_OK;
    };
    CB_TRY(reserve(&ctx->ws_x, &ctx->ws_x_bytes, xbytes));
    CB_TRY(reserve(&ctx->ws_y, &ctx->ws_y_bytes, ybytes));
    CB_CUDA(ctx, cudaEventRecord(ctx->host_begin, ctx->compute));
    CB_CUDA(ctx, cudaStreamWaitEvent(ctx->h2d, ctx->host_begin, 0));
    CB_CUDA(ctx, cudaStreamWaitEvent(ctx->d2h, ctx->host_begin, 0));
    size_t xoff = 0, yoff = 0;
    for (int i = 0; i < nslab; ++i) {
        const int64_t c0 = (int64_t)i * slab_cols;
        const int64_t cw = i == nslab - 1 ? std::max<int64_t>(kl - c0, 0) : std::min<int64_t>(slab_cols, std::max<int64_t>(kl - c0, 0));
        const int64_t ld = (cw + per16 - 1) / per16 * per16;
        char* dx = (char*)ctx->ws_x + xoff;
        char* dy = (char*)ctx->ws_y + yoff;
        xoff += ((size_t)xl * (size_t)ld * es + 255) & ~(size_t)255;
        yoff += ((size_t)rl * (size_t)ld * es + 255) & ~(size_t)255;
        if (cw > 0 && xl > 0) {
            if (ld != cw) CB_CUDA(ctx, cudaMemsetAsync(dx, 0, (size_t)xl * (size_t)ld * es, ctx->h2d));
            CB_CUDA(ctx, cudaMemcpy2DAsync(dx, (size_t)ld * es, (const char*)X_host + (size_t)c0 * es, (size_t)ldx * es, (size_t)cw * es,
                                           (size_t)xl, cudaMemcpyHostToDevice, ctx->h2d));
        }
        CB_CUDA(ctx, cudaEventRecord(ctx->slab_up[i], ctx->h2d));
        CB_CUDA(ctx, cudaStreamWaitEvent(ctx->compute, ctx->slab_up[i], 0));
        cb_dense X, Y;
        X.ctx = ctx; X.rows = xl; X.cols = cw; X.ld = ld; X.dtype = dtype; X.ptr = dx; X.owned = false;
        Y.ctx = ctx; Y.rows = rl; Y.cols = cw; Y.ld = ld; Y.dtype = dtype; Y.ptr = dy; Y.owned = false;
        CB_TRY(cb_spmm_summa(ctx, tile, &X, &Y, semiring, gm, gn, -1));
        CB_CUDA(ctx, cudaEventRecord(ctx->slab_done[i], ctx->compute));
        CB_CUDA(ctx, cudaStreamWaitEvent(ctx->d2h, ctx->slab_done[i], 0));
        if (cw > 0 && rl > 0)
            CB_CUDA(ctx, cudaMemcpy2DAsync((char*)Y_host + (size_t)c0 * es, (size_t)ldy * es, dy, (size_t)ld * es, (size_t)cw * es,
                                           (size_t)rl, cudaMemcpyDeviceToHost, ctx->d2h));
    }
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->d2h));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
    return CB_OK;
}

// ---- sparse x sparse on the grid (cb_spgemm.cu has the local kernels)
}  // extern "C"
namespace {
// one tile travels from `root` to every rank of `comm`: its essentials first (GetSetSizes, SpParHelper.cpp:797-809), then the
// slab as one message (BCastMatrix, :582-601).  The root passes its tile; the others get a view onto a fresh receive buffer.
int bcast_tile(cb_ctx* ctx, ncclComm_t comm, int root, bool mine, const cb_tile* src, cb_tile* view, char** recvbuf) {
    cb_scratch sc;
    cb_tile_meta meta;
    memset(&meta, 0, sizeof meta);
    if (mine) meta = cb_tile_get_meta(src);
    char* d_meta = nullptr;
    CB_CUDA(ctx, sc.alloc(&d_meta, sizeof meta));
    CB_CUDA(ctx, cudaMemcpyAsync(d_meta, &meta, sizeof meta, cudaMemcpyHostToDevice, ctx->comm));
    CB_NCCL(ctx, nccl().Broadcast(d_meta, d_meta, sizeof meta, ncclInt8, root, comm, ctx->comm));
    CB_CUDA(ctx, cudaMemcpyAsync(&meta, d_meta, sizeof meta, cudaMemcpyDeviceToHost, ctx->comm));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->comm));
    const size_t bytes = cb_layout(meta).total;
    char* buf = mine ? src->slab : nullptr;
    *recvbuf = nullptr;
    if (!mine) {
        cudaError_t e = cudaMalloc((void**)recvbuf, bytes);
        if (e != cudaSuccess) return cb_fail(ctx, CB_ERR_ALLOC, "cudaMalloc(%zu) for a received tile: %s", bytes, cudaGetErrorString(e));
        buf = *recvbuf;
    }
    CB_NCCL(ctx, nccl().Broadcast(buf, buf, bytes, ncclInt8, root, comm, ctx->comm));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->comm));
    if (!mine) { view->ctx = ctx; view->owns_slab = false; cb_tile_bind(view, meta, buf); }
    return CB_OK;
}
}  // namespace
extern "C" {

// C = A (x).(+) B for a SPARSE right-hand side, collective over the grid: the stage loop of Mult_AnXBn_Synch
// (ParFriends.h:1036-1083) with the tile pair of every stage multiplied on the GPU by expansion, and one sort + merge of all
// stages' partial products at the end (the role of MultiwayMerge, :1096).  A: rows of block-row myprocrow x columns of
// block-column myproccol of gn.  B: rows of block-row myprocrow of gn x columns of block myproccol of gk, values already of
// the product's type `dtype` (or a pattern).  C: this rank's block of the product as merged triples, column-major, local indices.
int cb_spgemm_summa(cb_ctx* ctx, const cb_tile* A, const cb_tile* B, int semiring, int dtype, int64_t gm, int64_t gn, int64_t gk, cb_coo** C) {
    if (!ctx || !A || !B || !C) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_spgemm_summa: null argument");
    const int pr = ctx->pr, pc = ctx->pc;
    int64_t r0, rl, c0, cl, x0, xl, k0, kl;
    block_range(gm, pr, ctx->myprocrow, &r0, &rl);
    block_range(gn, pc, ctx->myproccol, &c0, &cl);
    block_range(gn, pr, ctx->myprocrow, &x0, &xl);
    block_range(gk, pc, ctx->myproccol, &k0, &kl);
    if (A->m != rl || A->n != cl || B->m != xl || B->n != kl)
        return cb_fail(ctx, CB_ERR_DIMMISMATCH, "cb_spgemm_summa rank %d: local A %lldx%lld (want %lldx%lld), local B %lldx%lld (want %lldx%lld)", ctx->rank,
                       (long long)A->m, (long long)A->n, (long long)rl, (long long)cl, (long long)B->m, (long long)B->n, (long long)xl, (long long)kl);
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->nranks == 1) return cb_spgemm_local(ctx, A, B, semiring, dtype, C);
    std::vector<int64_t> seg(pr + pc + 1, 0);
    std::vector<int> a_root(pr + pc, 0), x_root(pr + pc, 0);
    int ns = 0;
    cb_summa_plan(pr, pc, gn, seg.data(), a_root.data(), x_root.data(), &ns);
    // my parts of A, one column slice per stage this rank roots: the cache of the dense stage loop (same key)
    cb_tile* mt = const_cast<cb_tile*>(A);
    const int64_t key[3] = {(int64_t)pr * 1000 + pc, gn, ns};
    if (mt->summa_key[0] != key[0] || mt->summa_key[1] != key[1] || mt->summa_key[2] != key[2]) {
        for (cb_tile* p : mt->summa_parts) cb_tile_free(p);
        for (cb_tile* p : mt->summa_remote) cb_tile_free(p);
        mt->summa_remote.clear();
        for (cb_tile* p : mt->summa_merged) cb_tile_free(p);
        mt->summa_merged.clear();
        for (cb_tile* p : mt->spgemm_remote) cb_tile_free(p);
        mt->spgemm_remote.clear();
        mt->spgemm_sent.clear();
        mt->summa_parts.assign(ns, nullptr);
        for (int s = 0; s < ns; ++s) {
            if (a_root[s] != ctx->myproccol) continue;
            const int64_t a = seg[s] - c0, b = seg[s + 1] - c0;
            if (a == 0 && b == cl) continue;
            CB_TRY(slice_cols(ctx, A, a, b, &mt->summa_parts[s]));
        }
        mt->summa_key[0] = key[0]; mt->summa_key[1] = key[1]; mt->summa_key[2] = key[2];
        if (ctx->summa_state) ((SummaState*)ctx->summa_state)->meta_tile = nullptr;
    }
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
    // A parts of my row neighbours stay on this GPU after the first product with this tile (tiles are immutable), like in the dense
    // stage loop; cb_summa_cache_a(ctx, 0) re-sends them every call, which is what the reference's broadcasts do.  Every rank of a
    // processor row has the same call history, so root and receivers agree on which parts have travelled.
    const bool cache = ctx->summa_cache_a && pc > 1;
    if (!cache || (int)mt->spgemm_remote.size() != ns) {
        for (cb_tile* p : mt->spgemm_remote) cb_tile_free(p);
        mt->spgemm_remote.assign(cache ? ns : 0, nullptr);
        mt->spgemm_sent.assign(cache ? ns : 0, 0);
    }
    cb_spgemm_acc acc;
    int status = CB_OK;
    for (int r = 0; r < pr && status == CB_OK; ++r) {
        int64_t b0, bl;
        block_range(gn, pr, r, &b0, &bl);
        cb_tile bview;
        char* brecv = nullptr;
        const bool b_mine = r == ctx->myprocrow;
        status = pr > 1 ? bcast_tile(ctx, (ncclComm_t)ctx->nccl_col, r, b_mine, B, &bview, &brecv) : CB_OK;
        const cb_tile* bt = b_mine ? B : &bview;
        for (int s = 0; s < ns && status == CB_OK; ++s) {
            if (x_root[s] != r || seg[s + 1] <= seg[s]) continue;
            cb_tile aview;
            char* arecv = nullptr;
            const bool a_mine = a_root[s] == ctx->myproccol;
            const cb_tile* mypart = a_mine ? (mt->summa_parts[s] ? mt->summa_parts[s] : A) : nullptr;
            const cb_tile* apart = a_mine ? mypart : &aview;
            const bool have = cache && (a_mine ? mt->spgemm_sent[s] != 0 : mt->spgemm_remote[s] != nullptr);
            if (have) {
                if (!a_mine) apart = mt->spgemm_remote[s];
            } else if (pc > 1) {
                status = bcast_tile(ctx, (ncclComm_t)ctx->nccl_row, a_root[s], a_mine, mypart, &aview, &arecv);
                if (status == CB_OK && cache) {
                    if (a_mine) mt->spgemm_sent[s] = 1;
                    else {                                      // keep the received part: the buffer changes owner
                        cb_tile* v = new cb_tile();
                        const cb_tile_meta meta = cb_tile_get_meta(&aview);
                        v->ctx = ctx;
                        v->slab = arecv; v->slab_bytes = cb_layout(meta).total; v->owns_slab = true;
                        cb_tile_bind(v, meta, v->slab);
                        mt->spgemm_remote[s] = v;
                        arecv = nullptr;
                        apart = v;
                    }
                }
            }
            if (status == CB_OK) status = cb_spgemm_expand(ctx, apart, bt, seg[s] - b0, semiring, dtype, &acc);
            cudaFree(arecv);
        }
        cudaFree(brecv);
    }
    if (status != CB_OK) { cb_spgemm_acc_release(&acc); return status; }
    return cb_spgemm_finish(ctx, &acc, semiring, dtype, rl, kl, C);
}

// ---- dense-vector multiply with the vector exchange on the device
}  // extern "C"
namespace {
// compact vector <-> one-column panel (rows padded to 16 bytes, the layout every panel of K2 has)
__global__ void __launch_bounds__(256)
vec_to_panel_kernel(const char* __restrict__ vec, int64_t n, int es, char* __restrict__ panel) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        for (int b = 0; b < es; ++b) panel[i * 16 + b] = vec[i * es + b];
}
// panel -> compact vector, folding in SR::id() the way the reference's y does (it starts as id(): ParFriends.h:1960-1963; this
// matters for SelectMax, where max(id, v) clips values below the identity)
template <typename T>
__global__ void __launch_bounds__(256)
panel_to_vec_kernel(const char* __restrict__ panel, int64_t n, int semiring, T* __restrict__ vec) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        T v = *reinterpret_cast<const T*>(panel + i * 16);
        if (semiring == CB_MAX_SEL2ND && v < T(-1)) v = T(-1);
        vec[i] = v;
    }
}
template <>
__global__ void __launch_bounds__(256)
panel_to_vec_kernel<uint8_t>(const char* __restrict__ panel, int64_t n, int, uint8_t* __restrict__ vec) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) vec[i] = (uint8_t)panel[i * 16];
}
}  // namespace
extern "C" {

// y = A (x).(+) x for a distributed dense vector, collective over the grid.  Replaces SpMV<SR>(SpParMat, FullyDistVec)
// (include/CombBLAS/ParFriends.h:1924-1996): the reference moves x to the transposed process, MPI_Allgatherv's it along the
// processor column, runs dcsc_gespmv on an id()-filled y (Friends.h:63-78) and MPI_Reduce's y along the processor row with
// SR::mpi_op.  Here the pieces of x go up once, are gathered between the GPUs (one NCCL broadcast per piece, grouped), the local
// multiply is K2 on a one-column panel, the partial results are combined along the processor row with ncclAllReduce
// (sum / min / max = SR::mpi_op) and every rank takes its piece of y down.  No vector travels through host memory in between.
//   x_piece / y_piece : this rank's pieces (host), at global offsets x_off / y_off with x_len / y_len elements; the pieces of x
//                       tile [0, gn) in rank order, y_piece must lie inside this rank's row block of gm
int cb_spmv_grid(cb_ctx* ctx, const cb_tile* tile, const void* x_piece, int64_t x_off, int64_t x_len, void* y_piece, int64_t y_off, int64_t y_len,
                 int semiring, int dtype, int64_t gm, int64_t gn) {
    if (!ctx || !tile) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_spmv_grid: null argument");
    const size_t es = cb_dtype_size(dtype);
    if (!es) return cb_fail(ctx, CB_ERR_UNSUPPORTED, "cb_spmv_grid: dtype %d", dtype);
    const int pr = ctx->pr, pc = ctx->pc;
    int64_t r0, rl, c0, cl;
    block_range(gm, pr, ctx->myprocrow, &r0, &rl);
    block_range(gn, pc, ctx->myproccol, &c0, &cl);
    if (tile->m != rl || tile->n != cl) return cb_fail(ctx, CB_ERR_DIMMISMATCH, "cb_spmv_grid: local A %lldx%lld, want %lldx%lld", (long long)tile->m, (long long)tile->n, (long long)rl, (long long)cl);
    if (x_off < 0 || x_len < 0 || x_off + x_len > gn || y_off < r0 || y_len < 0 || y_off + y_len > r0 + rl || (x_len > 0 && !x_piece) || (y_len > 0 && !y_piece))
        return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_spmv_grid: vector piece outside its range");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->compute;
    cb_scratch sc;
    char *xfull = nullptr, *xpanel = nullptr, *ypanel = nullptr, *yvec = nullptr;
    CB_CUDA(ctx, sc.alloc(&xfull, (size_t)std::max<int64_t>(gn, 1) * es));
    CB_CUDA(ctx, sc.alloc(&xpanel, (size_t)std::max<int64_t>(cl, 1) * 16));
    CB_CUDA(ctx, sc.alloc(&ypanel, (size_t)std::max<int64_t>(rl, 1) * 16));
    CB_CUDA(ctx, sc.alloc(&yvec, (size_t)std::max<int64_t>(rl, 1) * es));
    if (x_len > 0) CB_CUDA(ctx, cudaMemcpyAsync(xfull + (size_t)x_off * es, x_piece, (size_t)x_len * es, cudaMemcpyHostToDevice, st));
    if (ctx->nranks > 1) {
        // where every rank's piece sits: one small allgather, then one broadcast per piece in a group (an allgatherv)
        int64_t mine[2] = {x_off, x_len};
        int64_t* d_meta = nullptr;
        CB_CUDA(ctx, sc.alloc(&d_meta, (size_t)2 * (ctx->nranks + 1)));
        CB_CUDA(ctx, cudaMemcpyAsync(d_meta, mine, sizeof mine, cudaMemcpyHostToDevice, st));
        CB_NCCL(ctx, nccl().AllGather(d_meta, d_meta + 2, sizeof mine, ncclInt8, (ncclComm_t)ctx->nccl_world, st));
        std::vector<int64_t> all((size_t)2 * ctx->nranks);
        CB_CUDA(ctx, cudaMemcpyAsync(all.data(), d_meta + 2, sizeof(int64_t) * all.size(), cudaMemcpyDeviceToHost, st));
        CB_CUDA(ctx, cudaStreamSynchronize(st));
        CB_NCCL(ctx, nccl().GroupStart());
        for (int q = 0; q < ctx->nranks; ++q) {
            const int64_t off = all[(size_t)2 * q], len = all[(size_t)2 * q + 1];
            if (off < 0 || len < 0 || off + len > gn) { nccl().GroupEnd(); return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_spmv_grid: rank %d announced a piece outside [0, %lld)", q, (long long)gn); }
            if (len > 0) CB_NCCL(ctx, nccl().Broadcast(xfull + (size_t)off * es, xfull + (size_t)off * es, (size_t)len * es, ncclInt8, q, (ncclComm_t)ctx->nccl_world, st));
        }
        CB_NCCL(ctx, nccl().GroupEnd());
    }
    if (cl > 0) {
        vec_to_panel_kernel<<<grid_for(cl, ctx->sm_count), 256, 0, st>>>(xfull + (size_t)c0 * es, cl, (int)es, xpanel);
        CB_LAUNCHED(ctx);
    }
    const int64_t ld = 16 / (int64_t)es;
    CB_TRY(cb_spmm_launch(ctx, st, tile, xpanel, ld, ypanel, ld, 1, dtype, semiring, 0));
    if (rl > 0) {
        const int g = grid_for(rl, ctx->sm_count);
        switch (dtype) {
            case CB_F32: panel_to_vec_kernel<float><<<g, 256, 0, st>>>(ypanel, rl, semiring, (float*)yvec); break;
            case CB_F64: panel_to_vec_kernel<double><<<g, 256, 0, st>>>(ypanel, rl, semiring, (double*)yvec); break;
            case CB_I32: panel_to_vec_kernel<int32_t><<<g, 256, 0, st>>>(ypanel, rl, semiring, (int32_t*)yvec); break;
            case CB_I64: panel_to_vec_kernel<int64_t><<<g, 256, 0, st>>>(ypanel, rl, semiring, (int64_t*)yvec); break;
            default: panel_to_vec_kernel<uint8_t><<<g, 256, 0, st>>>(ypanel, rl, semiring, (uint8_t*)yvec); break;
        }
        CB_LAUNCHED(ctx);
        CB_CUDA(ctx, cudaGetLastError());
        if (pc > 1) {
            // MPI_Reduce with SR::mpi_op on RowWorld (ParFriends.h:1985-1993): sum, min or max; OR of booleans is max of bytes
            const int op = (semiring == CB_MIN_PLUS) ? 3 /*ncclMin*/ : (semiring == CB_PLUS_TIMES && dtype != CB_U8) ? 0 /*ncclSum*/ : 2 /*ncclMax*/;
            const int nt = dtype == CB_F32 ? 7 : dtype == CB_F64 ? 8 : dtype == CB_I32 ? 2 : dtype == CB_I64 ? 4 : 1 /*ncclUint8*/;
            CB_NCCL(ctx, nccl().AllReduce(yvec, yvec, (size_t)rl, nt, op, (ncclComm_t)ctx->nccl_row, st));
        }
    }
    if (y_len > 0) CB_CUDA(ctx, cudaMemcpyAsync(y_piece, yvec + (size_t)(y_off - r0) * es, (size_t)y_len * es, cudaMemcpyDeviceToHost, st));
    CB_CUDA(ctx, cudaStreamSynchronize(st));
    return CB_OK;
}

// ---- distributed ingestion: triples parsed anywhere travel to the rank that owns them, between the GPUs
}  // extern "C"
namespace {
// owner rank and local key of every triple (Owner rule of SpParMat.cpp:5066-5096: floor blocks, the last one takes the remainder)
__global__ void __launch_bounds__(256)
owner_kernel(const int64_t* __restrict__ rows, const int64_t* __restrict__ cols, int64_t nz, int64_t gm, int64_t gn, int pr, int pc,
             uint32_t* __restrict__ owner, uint64_t* __restrict__ key, unsigned int* __restrict__ bad) {
    const int64_t mper = gm / pr, nper = gn / pc;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nz; p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = rows[p], c = cols[p];
        if (r < 0 || r >= gm || c < 0 || c >= gn) { atomicOr(bad, 1u); owner[p] = 0; key[p] = 0; continue; }
        const int orow = mper ? (int)min((int64_t)(r / mper), (int64_t)(pr - 1)) : pr - 1;
        const int ocol = nper ? (int)min((int64_t)(c / nper), (int64_t)(pc - 1)) : pc - 1;
        owner[p] = (uint32_t)(orow * pc + ocol);
        key[p] = ((uint64_t)(r - (int64_t)orow * mper) << 32) | (uint64_t)(c - (int64_t)ocol * nper);
    }
}
template <typename V>
__global__ void __launch_bounds__(256)
permute_kernel(const uint64_t* __restrict__ key, const V* __restrict__ val, const uint32_t* __restrict__ perm, int64_t nz, uint64_t* __restrict__ key_out, V* __restrict__ val_out) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nz; p += (int64_t)gridDim.x * blockDim.x) {
        key_out[p] = key[perm[p]];
        if (val) val_out[p] = val[perm[p]];
    }
}
__global__ void __launch_bounds__(256)
iota32_kernel(uint32_t* p, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = (uint32_t)i;
}
__global__ void __launch_bounds__(256)
count_owner_kernel(const uint32_t* __restrict__ owner_sorted, int64_t nz, int nranks, int64_t* __restrict__ start) {
    // owner_sorted is ascending: start[q] = first position with owner >= q (start[nranks] = nz)
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q <= nranks; q += gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = nz;
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (owner_sorted[mid] < (uint32_t)q) lo = mid + 1; else hi = mid; }
        start[q] = lo;
    }
}
// duplicates merged the way SpTuples::RemoveDuplicates(BinOp) does it (SpParMat.cpp:2962-2967); op: 0 keep the first, 1 sum, 2 max, 3 min
template <typename T>
struct DupOp {
    int op;
    __host__ __device__ T operator()(const T& a, const T& b) const { return op == 1 ? (T)(a + b) : op == 2 ? (a < b ? b : a) : op == 3 ? (b < a ? b : a) : a; }
};
template <typename T>
int merge_duplicates(cb_ctx* ctx, cb_scratch& sc, uint64_t* keys_sorted, T* vals_sorted, int64_t n, int op, uint64_t* keys_out, T* vals_out, int64_t* n_out) {
    int* d_runs = nullptr;
    CB_CUDA(ctx, sc.alloc(&d_runs, 1));
    size_t b = 0;
    DupOp<T> f{op};
    CB_CUDA(ctx, cub::DeviceReduce::ReduceByKey(nullptr, b, keys_sorted, keys_out, vals_sorted, vals_out, d_runs, f, (int)n, ctx->compute));
    char* tmp = nullptr;
    CB_CUDA(ctx, sc.alloc(&tmp, b));
    CB_CUDA(ctx, cub::DeviceReduce::ReduceByKey(tmp, b, keys_sorted, keys_out, vals_sorted, vals_out, d_runs, f, (int)n, ctx->compute));
    int runs = 0;
    CB_CUDA(ctx, cudaMemcpyAsync(&runs, d_runs, sizeof runs, cudaMemcpyDeviceToHost, ctx->compute));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
    *n_out = runs;
    return CB_OK;
}
}  // namespace
extern "C" {

// Replaces the distribution half of SpParMat::ParallelReadMM / SparseCommon (include/CombBLAS/SpParMat.cpp:3978-4115, :2891-2968):
// every rank hands in the triples IT parsed from its share of the input (global 0-based coordinates, owned by anybody); they are
// routed to their owners between the GPUs (the reference's MPI_Alltoallv: one grouped ncclSend / ncclRecv exchange), duplicates are
// merged with dup_op (0 keep the first, 1 sum, 2 max, 3 min: SpTuples::RemoveDuplicates(BinOp)), and this rank's tile is built on
// the device.  rows / cols / vals are host arrays; val_dtype CB_PATTERN takes no values.  Collective over the grid.
int cb_tile_from_distributed_coo(cb_ctx* ctx, int64_t gm, int64_t gn, int64_t nz, const int64_t* rows, const int64_t* cols, const void* vals,
                                 int val_dtype, int dup_op, cb_tile** out) {
    if (!ctx || !out || nz < 0 || (nz > 0 && (!rows || !cols))) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_tile_from_distributed_coo: null argument");
    *out = nullptr;
    const size_t vs = val_dtype == CB_PATTERN ? 0 : cb_dtype_size(val_dtype);
    if (val_dtype != CB_PATTERN && !vs) return cb_fail(ctx, CB_ERR_UNSUPPORTED, "cb_tile_from_distributed_coo: value dtype %d", val_dtype);
    if (vs && nz > 0 && !vals) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_tile_from_distributed_coo: null values");
    if (nz >= (int64_t(1) << 31)) return cb_fail(ctx, CB_ERR_TOO_LARGE, "cb_tile_from_distributed_coo: %lld triples on one rank", (long long)nz);
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->compute;
    cb_scratch sc;
    const int64_t cap = std::max<int64_t>(nz, 1);
    int64_t *d_rows = nullptr, *d_cols = nullptr;
    char* d_vals = nullptr;
    CB_CUDA(ctx, sc.alloc(&d_rows, (size_t)cap)); CB_CUDA(ctx, sc.alloc(&d_cols, (size_t)cap));
    if (vs) CB_CUDA(ctx, sc.alloc(&d_vals, (size_t)cap * vs));
    if (nz > 0) {
        CB_CUDA(ctx, cudaMemcpyAsync(d_rows, rows, sizeof(int64_t) * (size_t)nz, cudaMemcpyHostToDevice, st));
        CB_CUDA(ctx, cudaMemcpyAsync(d_cols, cols, sizeof(int64_t) * (size_t)nz, cudaMemcpyHostToDevice, st));
        if (vs) CB_CUDA(ctx, cudaMemcpyAsync(d_vals, vals, vs * (size_t)nz, cudaMemcpyHostToDevice, st));
    }
    return cb_ingest_device_coo(ctx, gm, gn, nz, d_rows, d_cols, d_vals, val_dtype, dup_op, out);
}

}  // extern "C"
// the same with the triples already on the device (cb_tile_from_mm_text parses them there); d_vals may be NULL for CB_PATTERN
int cb_ingest_device_coo(cb_ctx* ctx, int64_t gm, int64_t gn, int64_t nz, const int64_t* d_rows, const int64_t* d_cols, const void* d_vals_in,
                         int val_dtype, int dup_op, cb_tile** out) {
    *out = nullptr;
    const size_t vs = val_dtype == CB_PATTERN ? 0 : cb_dtype_size(val_dtype);
    if (nz >= (int64_t(1) << 31)) return cb_fail(ctx, CB_ERR_TOO_LARGE, "distributed ingestion: %lld triples on one rank", (long long)nz);
    if (dup_op < 0 || dup_op > 3) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "distributed ingestion: duplicate rule %d", dup_op);
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->compute;
    const int pr = ctx->pr, pc = ctx->pc, np = ctx->nranks, sm = ctx->sm_count;
    int64_t r0, rl, c0, cl;
    block_range(gm, pr, ctx->myprocrow, &r0, &rl);
    block_range(gn, pc, ctx->myproccol, &c0, &cl);
    cb_scratch sc;
    const int64_t cap = std::max<int64_t>(nz, 1);
    const char* d_vals = (const char*)d_vals_in;
    int64_t* d_start = nullptr;
    char* d_vals_routed = nullptr;
    uint32_t *d_owner = nullptr, *d_owner_sorted = nullptr, *d_perm = nullptr, *d_perm_sorted = nullptr;
    uint64_t *d_key = nullptr, *d_key_routed = nullptr;
    unsigned int* d_bad = nullptr;
    CB_CUDA(ctx, sc.alloc(&d_owner, (size_t)cap)); CB_CUDA(ctx, sc.alloc(&d_owner_sorted, (size_t)cap));
    CB_CUDA(ctx, sc.alloc(&d_perm, (size_t)cap)); CB_CUDA(ctx, sc.alloc(&d_perm_sorted, (size_t)cap));
    CB_CUDA(ctx, sc.alloc(&d_key, (size_t)cap)); CB_CUDA(ctx, sc.alloc(&d_key_routed, (size_t)cap));
    CB_CUDA(ctx, sc.alloc(&d_start, (size_t)np + 1)); CB_CUDA(ctx, sc.alloc(&d_bad, 1));
    if (vs) CB_CUDA(ctx, sc.alloc(&d_vals_routed, (size_t)cap * vs));
    CB_CUDA(ctx, cudaMemsetAsync(d_bad, 0, sizeof(unsigned int), st));
    if (nz > 0) {
        owner_kernel<<<grid_for(nz, sm), 256, 0, st>>>(d_rows, d_cols, nz, gm, gn, pr, pc, d_owner, d_key, d_bad);
        iota32_kernel<<<grid_for(nz, sm), 256, 0, st>>>(d_perm, nz);
        CB_LAUNCHED(ctx); CB_LAUNCHED(ctx);
        size_t b = 0;
        int obits = 1;
        while ((1 << obits) < np) ++obits;
        CB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, b, d_owner, d_owner_sorted, d_perm, d_perm_sorted, (int)nz, 0, obits, st));
        char* tmp = nullptr;
        CB_CUDA(ctx, sc.alloc(&tmp, b));
        CB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(tmp, b, d_owner, d_owner_sorted, d_perm, d_perm_sorted, (int)nz, 0, obits, st));
        switch (vs) {
            case 0: permute_kernel<uint8_t><<<grid_for(nz, sm), 256, 0, st>>>(d_key, nullptr, d_perm_sorted, nz, d_key_routed, nullptr); break;
            case 1: permute_kernel<uint8_t><<<grid_for(nz, sm), 256, 0, st>>>(d_key, (const uint8_t*)d_vals, d_perm_sorted, nz, d_key_routed, (uint8_t*)d_vals_routed); break;
            case 4: permute_kernel<uint32_t><<<grid_for(nz, sm), 256, 0, st>>>(d_key, (const uint32_t*)d_vals, d_perm_sorted, nz, d_key_routed, (uint32_t*)d_vals_routed); break;
            default: permute_kernel<uint64_t><<<grid_for(nz, sm), 256, 0, st>>>(d_key, (const uint64_t*)d_vals, d_perm_sorted, nz, d_key_routed, (uint64_t*)d_vals_routed); break;
        }
        CB_LAUNCHED(ctx); CB_LAUNCHED(ctx);
    }
    count_owner_kernel<<<1, 256, 0, st>>>(d_owner_sorted, nz, np, d_start);
    CB_LAUNCHED(ctx);
    CB_CUDA(ctx, cudaGetLastError());
    std::vector<int64_t> start((size_t)np + 1, 0);
    unsigned int bad = 0;
    CB_CUDA(ctx, cudaMemcpyAsync(start.data(), d_start, sizeof(int64_t) * (size_t)(np + 1), cudaMemcpyDeviceToHost, st));
    CB_CUDA(ctx, cudaMemcpyAsync(&bad, d_bad, sizeof bad, cudaMemcpyDeviceToHost, st));
    CB_CUDA(ctx, cudaStreamSynchronize(st));
    // every rank must learn about a bad triple anywhere before anyone blocks in the exchange
    int64_t anybad = bad;
    CB_TRY(cb_comm_allreduce_i64(ctx, 0, 1, &anybad, 1));
    if (anybad) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "distributed ingestion: a triple lies outside the %lld x %lld matrix", (long long)gm, (long long)gn);
    // counts[q * np + r] = triples rank q holds for rank r
    std::vector<int64_t> counts((size_t)np * np, 0);
    for (int q = 0; q < np; ++q) counts[(size_t)ctx->rank * np + q] = start[(size_t)q + 1] - start[(size_t)q];
    if (np > 1) {
        int64_t *d_c = nullptr, *d_call = nullptr;
        CB_CUDA(ctx, sc.alloc(&d_c, (size_t)np));
        CB_CUDA(ctx, sc.alloc(&d_call, (size_t)np * np));
        CB_CUDA(ctx, cudaMemcpyAsync(d_c, counts.data() + (size_t)ctx->rank * np, sizeof(int64_t) * (size_t)np, cudaMemcpyHostToDevice, st));
        CB_NCCL(ctx, nccl().AllGather(d_c, d_call, sizeof(int64_t) * (size_t)np, ncclInt8, (ncclComm_t)ctx->nccl_world, st));
        CB_CUDA(ctx, cudaMemcpyAsync(counts.data(), d_call, sizeof(int64_t) * (size_t)np * np, cudaMemcpyDeviceToHost, st));
        CB_CUDA(ctx, cudaStreamSynchronize(st));
    }
    int64_t nrecv = 0;
    std::vector<int64_t> roff((size_t)np + 1, 0);
    for (int q = 0; q < np; ++q) { roff[(size_t)q] = nrecv; nrecv += counts[(size_t)q * np + ctx->rank]; }
    if (nrecv >= (int64_t(1) << 31)) return cb_fail(ctx, CB_ERR_TOO_LARGE, "distributed ingestion: %lld triples for one tile", (long long)nrecv);
    const int64_t rcap = std::max<int64_t>(nrecv, 1);
    uint64_t *d_rkey = nullptr, *d_rkey_sorted = nullptr;
    char *d_rval = nullptr, *d_rval_sorted = nullptr;
    CB_CUDA(ctx, sc.alloc(&d_rkey, (size_t)rcap)); CB_CUDA(ctx, sc.alloc(&d_rkey_sorted, (size_t)rcap));
    if (vs) { CB_CUDA(ctx, sc.alloc(&d_rval, (size_t)rcap * vs)); CB_CUDA(ctx, sc.alloc(&d_rval_sorted, (size_t)rcap * vs)); }
    // the exchange: what I hold for q goes to q, what q holds for me arrives at roff[q]
    if (np > 1) CB_NCCL(ctx, nccl().GroupStart());
    for (int q = 0; q < np; ++q) {
        const int64_t sn = counts[(size_t)ctx->rank * np + q], rn = counts[(size_t)q * np + ctx->rank];
        if (q == ctx->rank) {
            if (sn) {
                CB_CUDA(ctx, cudaMemcpyAsync(d_rkey + roff[(size_t)q], d_key_routed + start[(size_t)q], sizeof(uint64_t) * (size_t)sn, cudaMemcpyDeviceToDevice, st));
                if (vs) CB_CUDA(ctx, cudaMemcpyAsync(d_rval + (size_t)roff[(size_t)q] * vs, d_vals_routed + (size_t)start[(size_t)q] * vs, vs * (size_t)sn, cudaMemcpyDeviceToDevice, st));
            }
            continue;
        }
        if (sn) {
            CB_NCCL(ctx, nccl().Send(d_key_routed + start[(size_t)q], sizeof(uint64_t) * (size_t)sn, ncclInt8, q, (ncclComm_t)ctx->nccl_world, st));
            if (vs) CB_NCCL(ctx, nccl().Send(d_vals_routed + (size_t)start[(size_t)q] * vs, vs * (size_t)sn, ncclInt8, q, (ncclComm_t)ctx->nccl_world, st));
        }
        if (rn) {
            CB_NCCL(ctx, nccl().Recv(d_rkey + roff[(size_t)q], sizeof(uint64_t) * (size_t)rn, ncclInt8, q, (ncclComm_t)ctx->nccl_world, st));
            if (vs) CB_NCCL(ctx, nccl().Recv(d_rval + (size_t)roff[(size_t)q] * vs, vs * (size_t)rn, ncclInt8, q, (ncclComm_t)ctx->nccl_world, st));
        }
    }
    if (np > 1) CB_NCCL(ctx, nccl().GroupEnd());
    if (nrecv == 0) return cb_tile_build_from_keys(ctx, rl, cl, 0, d_rkey, nullptr, val_dtype, true, sc, out);
    // sort by (row, column) - stable, so "keep the first" means first in rank order, then in the order handed in - and merge
    {
        size_t b = 0;
        char* tmp = nullptr;
        const int end_bit = 64;
        if (vs == 0) {
            CB_CUDA(ctx, cub::DeviceRadixSort::SortKeys(nullptr, b, d_rkey, d_rkey_sorted, (int)nrecv, 0, end_bit, st));
            CB_CUDA(ctx, sc.alloc(&tmp, b));
            CB_CUDA(ctx, cub::DeviceRadixSort::SortKeys(tmp, b, d_rkey, d_rkey_sorted, (int)nrecv, 0, end_bit, st));
            int64_t* d_n = nullptr;
            CB_CUDA(ctx, sc.alloc(&d_n, 1));
            size_t b2 = 0;
            CB_CUDA(ctx, cub::DeviceSelect::Unique(nullptr, b2, d_rkey_sorted, d_rkey, d_n, (int)nrecv, st));
            char* tmp2 = nullptr;
            CB_CUDA(ctx, sc.alloc(&tmp2, b2));
            CB_CUDA(ctx, cub::DeviceSelect::Unique(tmp2, b2, d_rkey_sorted, d_rkey, d_n, (int)nrecv, st));
            int64_t nu = 0;
            CB_CUDA(ctx, cudaMemcpyAsync(&nu, d_n, sizeof nu, cudaMemcpyDeviceToHost, st));
            CB_CUDA(ctx, cudaStreamSynchronize(st));
            ctx->launches += 4;
            return cb_tile_build_from_keys(ctx, rl, cl, nu, d_rkey, nullptr, val_dtype, true, sc, out);
        }
        int64_t nu = 0;
#define SORT_MERGE(T)                                                                                                                                   \
        CB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, b, d_rkey, d_rkey_sorted, (const T*)d_rval, (T*)d_rval_sorted, (int)nrecv, 0, end_bit, st)); \
        CB_CUDA(ctx, sc.alloc(&tmp, b));                                                                                                                \
        CB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(tmp, b, d_rkey, d_rkey_sorted, (const T*)d_rval, (T*)d_rval_sorted, (int)nrecv, 0, end_bit, st));     \
        CB_TRY(merge_duplicates<T>(ctx, sc, d_rkey_sorted, (T*)d_rval_sorted, nrecv, dup_op, d_rkey, (T*)d_rval, &nu));
        switch (val_dtype) {
            case CB_F32: { SORT_MERGE(float) break; }
            case CB_F64: { SORT_MERGE(double) break; }
            case CB_I32: { SORT_MERGE(int32_t) break; }
            case CB_I64: { SORT_MERGE(int64_t) break; }
            default: { SORT_MERGE(uint8_t) break; }
        }
#undef SORT_MERGE
        ctx->launches += 4;
        return cb_tile_build_from_keys(ctx, rl, cl, nu, d_rkey, d_rval, val_dtype, true, sc, out);
    }
}

// byte allgather over the processor column with host buffers (set-up traffic of the peer transport)
int cb_nccl_allgather_col(cb_ctx* ctx, const void* send_host, void* recv_host, size_t bytes) {
    const int np = ctx->pr;
    if (np == 1) { memcpy(recv_host, send_host, bytes); return CB_OK; }
    cb_scratch sc;
    char *d_send, *d_recv;
    CB_CUDA(ctx, sc.alloc(&d_send, bytes));
    CB_CUDA(ctx, sc.alloc(&d_recv, bytes * (size_t)np));
    CB_CUDA(ctx, cudaMemcpyAsync(d_send, send_host, bytes, cudaMemcpyHostToDevice, ctx->comm));
    CB_NCCL(ctx, nccl().AllGather(d_send, d_recv, bytes, ncclInt8, (ncclComm_t)ctx->nccl_col, ctx->comm));
    CB_CUDA(ctx, cudaMemcpyAsync(recv_host, d_recv, bytes * (size_t)np, cudaMemcpyDeviceToHost, ctx->comm));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->comm));
    return CB_OK;
}
extern "C" {

// Small host-side reductions over the grid's communicators: what SpParMat::getnnz/getnrow/getncol do with
// MPI_Allreduce (reference include/CombBLAS/SpParMat.cpp:773-797).  which: 0 world, 1 row, 2 column; op: 0 sum, 1 max, 2 min.
int cb_comm_allreduce_i64(cb_ctx* ctx, int which, int op, int64_t* inout, int count) {
    if (ctx->nranks == 1 || count <= 0) return CB_OK;
    ncclComm_t comm = (ncclComm_t)(which == 0 ? ctx->nccl_world : which == 1 ? ctx->nccl_row : ctx->nccl_col);
    if (!comm) return cb_fail(ctx, CB_ERR_NCCL, "cb_comm_allreduce_i64: no communicator %d", which);
    if (op < 0 || op > 2) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_comm_allreduce_i64: op %d", op);
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cb_scratch sc;
    int64_t* d;
    CB_CUDA(ctx, sc.alloc(&d, (size_t)count));
    CB_CUDA(ctx, cudaMemcpyAsync(d, inout, sizeof(int64_t) * (size_t)count, cudaMemcpyHostToDevice, ctx->comm));
    const int nccl_op = op == 0 ? 0 /*ncclSum*/ : op == 1 ? 2 /*ncclMax*/ : 3 /*ncclMin*/;
    CB_NCCL(ctx, nccl().AllReduce(d, d, (size_t)count, 4 /*ncclInt64*/, nccl_op, comm, ctx->comm));
    CB_CUDA(ctx, cudaMemcpyAsync(inout, d, sizeof(int64_t) * (size_t)count, cudaMemcpyDeviceToHost, ctx->comm));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->comm));
    return CB_OK;
}

int cb_summa_cache_a(cb_ctx* ctx, int on) {
    ctx->summa_cache_a = on != 0;
    return CB_OK;
}

int cb_summa_times(cb_ctx* ctx, float ms[4]) {
    for (int i = 0; i < 4; ++i) ms[i] = 0;
    SummaState* S = (SummaState*)ctx->summa_state;
    if (!S || !S->nstages) return CB_OK;
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->comm));
    CB_CUDA(ctx, cudaEventElapsedTime(&ms[0], S->begin, S->end));
    for (int s = 0; s < S->nstages; ++s) {
        float t = 0;
        if (S->comm_used[s] && cudaEventQuery(S->comm_ev[2 * s + 1]) == cudaSuccess &&
            cudaEventElapsedTime(&t, S->comm_ev[2 * s], S->comm_ev[2 * s + 1]) == cudaSuccess && t > 0) ms[1] += t;
    }
    cudaGetLastError();
    ms[3] = (float)S->nstages;
    if (getenv("CB_SUMMA_DEBUG")) {
        float a = -1, b = -1;
        cb_p2p_debug_times(ctx, S->begin, &a, &b);
        fprintf(stderr, "[summa rank %d] stage loop %.3f ms; pushes to first peer started at %.3f ms, done at %.3f ms after begin\n", ctx->rank, ms[0], a, b);
    }
    return CB_OK;
}

}  // extern "C"
