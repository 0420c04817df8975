// Hub columns of a tile, for the hub variant of K2 (cb_spmm_hub_kernel.cuh): which columns are used most, and for every
// nonzero the rank of its column among them.  Built lazily at the first hub multiply with a tile; the default K2 path
// never touches any of this.  Opt-in: cb_spmm_hub_config(ctx, 1, ...) or CB_SPMM_HUB=1.
#include <algorithm>
#include <numeric>
#include <cub/cub.cuh>
#include "cb_hub.cuh"

struct cb_hub {
    uint16_t* hubslot = nullptr;      // [nnz] device
    int32_t* hubcols = nullptr;       // [nhub_max] device, rank -> column
    unsigned* counters = nullptr;     // [CB_HUB_MAX_SLABS] device, dynamic chunk counters of a launch
    bool built = false;               // hub columns selected (a ring-only launch creates the struct for its counters alone)
    int nhub_max = 0;
    std::vector<int64_t> cum;         // [nhub_max] nonzeros in the columns of rank <= r
    int last_nhub = 0;
    double last_cover = 0;
    uint8_t* hubcls = nullptr;        // [nnz] device: floor(log2(rank + 1)) of the nonzero's column by descending use, 255 = used once
    bool cls_built = false;
    // K2W (hub panel under a persisting L2 window): the column stream with the columns of the win_h most used columns replaced
    // by bit 30 + their rank, and the columns themselves by rank
    int32_t* win_colflag = nullptr;   // [nnz] device
    int32_t* win_cols = nullptr;      // [win_h] device
    int64_t win_h = 0, win_want = -1;
    double win_cover = 0;
};

void cb_hub_release(cb_tile* t) {
    if (!t || !t->hub) return;
    cudaFree(t->hub->hubslot);
    cudaFree(t->hub->hubcols);
    cudaFree(t->hub->counters);
    cudaFree(t->hub->hubcls);
    cudaFree(t->hub->win_colflag);
    cudaFree(t->hub->win_cols);
    delete t->hub;
    t->hub = nullptr;
}

__global__ void __launch_bounds__(256)
cb_hub_count_kernel(const int32_t* __restrict__ colflag, int64_t nnz, int* __restrict__ counts) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(counts + (colflag[p] & 0x7fffffff), 1);
}

__global__ void __launch_bounds__(256)
cb_hub_slot_kernel(const int32_t* __restrict__ colflag, int64_t nnz, const uint16_t* __restrict__ rank_of_col, uint16_t* __restrict__ hubslot) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * blockDim.x)
        hubslot[p] = rank_of_col[colflag[p] & 0x7fffffff];
}

// ---- use classes for the L2 residency hints of K2P
__global__ void __launch_bounds__(256)
cb_iota_kernel(int32_t* p, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = (int32_t)i;
}
__global__ void __launch_bounds__(256)
cb_class_of_col_kernel(const int* __restrict__ counts_sorted, const int32_t* __restrict__ cols_sorted, int64_t n, uint8_t* __restrict__ class_of_col) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x)
        class_of_col[cols_sorted[r]] = counts_sorted[r] >= 2 ? (uint8_t)(63 - __clzll((long long)(r + 1))) : (uint8_t)255;
}
__global__ void __launch_bounds__(256)
cb_hubcls_kernel(const int32_t* __restrict__ colflag, int64_t nnz, const uint8_t* __restrict__ class_of_col, uint8_t* __restrict__ hubcls) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * blockDim.x)
        hubcls[p] = class_of_col[colflag[p] & 0x7fffffff];
}

int cb_hubcls_get(cb_ctx* ctx, const cb_tile* tile, const uint8_t** cls) {
    *cls = nullptr;
    if (!tile->owns_slab || tile->nnz == 0 || tile->n >= (1LL << 30)) return CB_OK;
    cb_tile* t = const_cast<cb_tile*>(tile);
    if (!t->hub) t->hub = new cb_hub();
    cb_hub* h = t->hub;
    if (!h->cls_built) {
        cb_scratch sc;
        int *d_counts = nullptr, *d_counts_sorted = nullptr;
        int32_t *d_cols = nullptr, *d_cols_sorted = nullptr;
        uint8_t* d_class = nullptr;
        CB_CUDA(ctx, sc.alloc(&d_counts, (size_t)t->n));
        CB_CUDA(ctx, sc.alloc(&d_counts_sorted, (size_t)t->n));
        CB_CUDA(ctx, sc.alloc(&d_cols, (size_t)t->n));
        CB_CUDA(ctx, sc.alloc(&d_cols_sorted, (size_t)t->n));
        CB_CUDA(ctx, sc.alloc(&d_class, (size_t)t->n));
        CB_CUDA(ctx, cudaMemsetAsync(d_counts, 0, (size_t)t->n * sizeof(int), ctx->compute));
        const unsigned blocks = (unsigned)std::min<int64_t>((t->nnz + 255) / 256, (int64_t)ctx->sm_count * 16);
        const unsigned nblocks = (unsigned)std::min<int64_t>((t->n + 255) / 256, (int64_t)ctx->sm_count * 16);
        cb_hub_count_kernel<<<blocks, 256, 0, ctx->compute>>>(t->colflag, t->nnz, d_counts);
        cb_iota_kernel<<<nblocks, 256, 0, ctx->compute>>>(d_cols, t->n);
        CB_LAUNCHED(ctx); CB_LAUNCHED(ctx);
        size_t tmp_bytes = 0;
        CB_CUDA(ctx, cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, d_counts, d_counts_sorted, d_cols, d_cols_sorted, (int)t->n, 0, 32, ctx->compute));
        char* tmp = nullptr;
        CB_CUDA(ctx, sc.alloc(&tmp, tmp_bytes));
        CB_CUDA(ctx, cub::DeviceRadixSort::SortPairsDescending(tmp, tmp_bytes, d_counts, d_counts_sorted, d_cols, d_cols_sorted, (int)t->n, 0, 32, ctx->compute));
        cb_class_of_col_kernel<<<nblocks, 256, 0, ctx->compute>>>(d_counts_sorted, d_cols_sorted, t->n, d_class);
        cudaError_t e = cudaMalloc((void**)&h->hubcls, (size_t)t->nnz);
        if (e != cudaSuccess) return cb_fail(ctx, CB_ERR_ALLOC, "cudaMalloc(%lld) for the column use classes: %s", (long long)t->nnz, cudaGetErrorString(e));
        cb_hubcls_kernel<<<blocks, 256, 0, ctx->compute>>>(t->colflag, t->nnz, d_class, h->hubcls);
        CB_LAUNCHED(ctx); CB_LAUNCHED(ctx);
        CB_CUDA(ctx, cudaGetLastError());
        CB_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
        h->cls_built = true;
    }
    *cls = h->hubcls;
    return CB_OK;
}

// ---- K2W: hub rows in a compact panel that a persisting L2 access-policy window keeps on the chip
__global__ void __launch_bounds__(256)
cb_win_rank_kernel(const int32_t* __restrict__ cols_sorted, int64_t h, int32_t* __restrict__ rank_of_col) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < h; r += (int64_t)gridDim.x * blockDim.x) rank_of_col[cols_sorted[r]] = (int32_t)r;
}
__global__ void __launch_bounds__(256)
cb_win_remap_kernel(const int32_t* __restrict__ colflag, int64_t nnz, const int32_t* __restrict__ rank_of_col, int32_t* __restrict__ out,
                    unsigned long long* __restrict__ covered) {
    unsigned long long mine = 0;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * blockDim.x) {
        const int32_t cf = colflag[p];
        const int32_t r = rank_of_col[cf & 0x7fffffff];
        out[p] = r >= 0 ? (int32_t)(((uint32_t)cf & 0x80000000u) | 0x40000000u | (uint32_t)r) : cf;
        mine += r >= 0;
    }
    mine = __reduce_add_sync(0xffffffffu, (unsigned)mine);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(covered, mine);
}
// rows of the hub columns, packed: panel[r] = X[cols[r]] (row_vecs 16-byte vectors each)
__global__ void __launch_bounds__(256)
cb_win_gather_kernel(const char* __restrict__ X, int64_t ldx_bytes, const int32_t* __restrict__ cols, int64_t h, int row_vecs, char* __restrict__ panel) {
    const int64_t total = h * row_vecs;
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < total; v += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = v / row_vecs;
        const int q = (int)(v - r * row_vecs);
        reinterpret_cast<uint4*>(panel + r * ldx_bytes)[q] = __ldg(reinterpret_cast<const uint4*>(X + (int64_t)cols[r] * ldx_bytes) + q);
    }
}

// the remapped column stream for the `want` most used columns (fewer when the tile has fewer columns used twice or more);
// *h == 0: nothing to keep.  Built lazily, rebuilt when `want` changes; owned tiles with n < 2^30 only.
int cb_hubwin_get(cb_ctx* ctx, const cb_tile* tile, int64_t want, const int32_t** colflag_w, const int32_t** cols, int64_t* h_out, double* cover) {
    *colflag_w = nullptr; *cols = nullptr; *h_out = 0;
    if (cover) *cover = 0;
    if (!tile->owns_slab || tile->nnz == 0 || tile->n >= (1LL << 30) || want <= 0) return CB_OK;
    cb_tile* t = const_cast<cb_tile*>(tile);
    if (!t->hub) t->hub = new cb_hub();
    cb_hub* h = t->hub;
    if (h->win_want != want) {
        cudaFree(h->win_cols); h->win_cols = nullptr;
        h->win_h = 0; h->win_want = want;
        cb_scratch sc;
        int *d_counts = nullptr, *d_counts_sorted = nullptr;
        int32_t *d_cols = nullptr, *d_cols_sorted = nullptr, *d_rank = nullptr;
        unsigned long long* d_cov = nullptr;
        CB_CUDA(ctx, sc.alloc(&d_counts, (size_t)t->n)); CB_CUDA(ctx, sc.alloc(&d_counts_sorted, (size_t)t->n));
        CB_CUDA(ctx, sc.alloc(&d_cols, (size_t)t->n)); CB_CUDA(ctx, sc.alloc(&d_cols_sorted, (size_t)t->n));
        CB_CUDA(ctx, sc.alloc(&d_rank, (size_t)t->n)); CB_CUDA(ctx, sc.alloc(&d_cov, 1));
        CB_CUDA(ctx, cudaMemsetAsync(d_counts, 0, (size_t)t->n * sizeof(int), ctx->compute));
        CB_CUDA(ctx, cudaMemsetAsync(d_rank, 0xff, (size_t)t->n * sizeof(int32_t), ctx->compute));
        CB_CUDA(ctx, cudaMemsetAsync(d_cov, 0, sizeof(unsigned long long), ctx->compute));
        const unsigned blocks = (unsigned)std::min<int64_t>((t->nnz + 255) / 256, (int64_t)ctx->sm_count * 16);
        const unsigned nblocks = (unsigned)std::min<int64_t>((t->n + 255) / 256, (int64_t)ctx->sm_count * 16);
        cb_hub_count_kernel<<<blocks, 256, 0, ctx->compute>>>(t->colflag, t->nnz, d_counts);
        cb_iota_kernel<<<nblocks, 256, 0, ctx->compute>>>(d_cols, t->n);
        CB_LAUNCHED(ctx); CB_LAUNCHED(ctx);
        size_t tmp_bytes = 0;
        CB_CUDA(ctx, cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, d_counts, d_counts_sorted, d_cols, d_cols_sorted, (int)t->n, 0, 32, ctx->compute));
        char* tmp = nullptr;
        CB_CUDA(ctx, sc.alloc(&tmp, tmp_bytes));
        CB_CUDA(ctx, cub::DeviceRadixSort::SortPairsDescending(tmp, tmp_bytes, d_counts, d_counts_sorted, d_cols, d_cols_sorted, (int)t->n, 0, 32, ctx->compute));
        // columns used once have nothing to share: cut the list where the counts drop below two
        int64_t hh = std::min<int64_t>(want, t->n);
        std::vector<int> top((size_t)hh);
        CB_CUDA(ctx, cudaMemcpyAsync(top.data(), d_counts_sorted, (size_t)hh * sizeof(int), cudaMemcpyDeviceToHost, ctx->compute));
        CB_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
        hh = std::partition_point(top.begin(), top.end(), [](int c) { return c >= 2; }) - top.begin();
        if (hh > 0) {
            if (!h->win_colflag) {
                cudaError_t e = cudaMalloc((void**)&h->win_colflag, (size_t)t->nnz * sizeof(int32_t));
                if (e != cudaSuccess) return cb_fail(ctx, CB_ERR_ALLOC, "cudaMalloc(%lld) for the remapped column stream: %s", (long long)t->nnz * 4, cudaGetErrorString(e));
            }
            CB_CUDA(ctx, cudaMalloc((void**)&h->win_cols, (size_t)hh * sizeof(int32_t)));
            CB_CUDA(ctx, cudaMemcpyAsync(h->win_cols, d_cols_sorted, (size_t)hh * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->compute));
            cb_win_rank_kernel<<<(unsigned)std::min<int64_t>((hh + 255) / 256, (int64_t)ctx->sm_count * 16), 256, 0, ctx->compute>>>(d_cols_sorted, hh, d_rank);
            cb_win_remap_kernel<<<blocks, 256, 0, ctx->compute>>>(t->colflag, t->nnz, d_rank, h->win_colflag, d_cov);
            CB_LAUNCHED(ctx); CB_LAUNCHED(ctx);
            CB_CUDA(ctx, cudaGetLastError());
            unsigned long long cov = 0;
            CB_CUDA(ctx, cudaMemcpyAsync(&cov, d_cov, sizeof cov, cudaMemcpyDeviceToHost, ctx->compute));
            CB_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
            h->win_cover = (double)cov / (double)t->nnz;
        }
        h->win_h = hh;
    }
    if (h->win_h > 0) { *colflag_w = h->win_colflag; *cols = h->win_cols; *h_out = h->win_h; if (cover) *cover = h->win_cover; }
    return CB_OK;
}
int cb_hubwin_gather(cb_ctx* ctx, cudaStream_t stream, const void* X, int64_t ldx_bytes, const int32_t* cols, int64_t h, int row_bytes, void* panel) {
    const int64_t total = h * (row_bytes / 16);
    const unsigned blocks = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)ctx->sm_count * 16);
    cb_win_gather_kernel<<<blocks, 256, 0, stream>>>((const char*)X, ldx_bytes, cols, h, row_bytes / 16, (char*)panel);
    CB_LAUNCHED(ctx);
    CB_CUDA(ctx, cudaGetLastError());
    return CB_OK;
}

extern "C" int cb_hub_select_host(const int32_t* counts, int64_t n, int max_hubs, int32_t* hubcols, int64_t* cum) {
    // the max_hubs most frequent columns, most frequent first, ties by ascending column; columns used once or never are
    // not hubs (nothing to share).  Pure host arithmetic: tested on CPU.
    if (!counts || n < 0 || max_hubs < 0 || (max_hubs > 0 && (!hubcols || !cum))) return -1;
    if (max_hubs > CB_HUB_MAX_RANKS) max_hubs = CB_HUB_MAX_RANKS;
    std::vector<int32_t> cand;
    for (int64_t c = 0; c < n; ++c)
        if (counts[c] >= 2) cand.push_back((int32_t)c);
    const size_t take = std::min<size_t>(cand.size(), (size_t)max_hubs);
    auto before = [&](int32_t x, int32_t y) { return counts[x] != counts[y] ? counts[x] > counts[y] : x < y; };
    std::partial_sort(cand.begin(), cand.begin() + take, cand.end(), before);
    int64_t run = 0;
    for (size_t r = 0; r < take; ++r) {
        hubcols[r] = cand[r];
        run += counts[cand[r]];
        cum[r] = run;
    }
    return (int)take;
}

static int hub_build(cb_ctx* ctx, cb_tile* t) {
    cb_hub* h = new cb_hub();
    t->hub = h;
    h->built = true;
    cb_scratch sc;
    int* d_counts = nullptr;
    uint16_t* d_rank = nullptr;
    CB_CUDA(ctx, sc.alloc(&d_counts, (size_t)t->n));
    CB_CUDA(ctx, sc.alloc(&d_rank, (size_t)t->n));
    CB_CUDA(ctx, cudaMemsetAsync(d_counts, 0, (size_t)t->n * sizeof(int), ctx->compute));
    const unsigned blocks = (unsigned)std::min<int64_t>((t->nnz + 255) / 256, (int64_t)ctx->sm_count * 16);
    cb_hub_count_kernel<<<blocks, 256, 0, ctx->compute>>>(t->colflag, t->nnz, d_counts);
    CB_LAUNCHED(ctx);
    CB_CUDA(ctx, cudaGetLastError());
    std::vector<int32_t> counts((size_t)t->n);
    CB_CUDA(ctx, cudaMemcpyAsync(counts.data(), d_counts, (size_t)t->n * sizeof(int), cudaMemcpyDeviceToHost, ctx->compute));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
    std::vector<int32_t> hubcols(CB_HUB_MAX_RANKS);
    h->cum.assign(CB_HUB_MAX_RANKS, 0);
    h->nhub_max = cb_hub_select_host(counts.data(), t->n, CB_HUB_MAX_RANKS, hubcols.data(), h->cum.data());
    if (h->nhub_max < 0) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_hub_select_host failed");
    h->cum.resize((size_t)h->nhub_max);
    if (h->nhub_max == 0) return CB_OK;
    std::vector<uint16_t> rank_of((size_t)t->n, (uint16_t)0xffff);
    for (int r = 0; r < h->nhub_max; ++r) rank_of[(size_t)hubcols[r]] = (uint16_t)r;
    CB_CUDA(ctx, cudaMalloc(&h->hubslot, (size_t)t->nnz * sizeof(uint16_t)));
    CB_CUDA(ctx, cudaMalloc(&h->hubcols, (size_t)h->nhub_max * sizeof(int32_t)));
    CB_CUDA(ctx, cudaMemcpyAsync(d_rank, rank_of.data(), (size_t)t->n * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->compute));
    CB_CUDA(ctx, cudaMemcpyAsync(h->hubcols, hubcols.data(), (size_t)h->nhub_max * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->compute));
    cb_hub_slot_kernel<<<blocks, 256, 0, ctx->compute>>>(t->colflag, t->nnz, d_rank, h->hubslot);
    CB_LAUNCHED(ctx);
    CB_CUDA(ctx, cudaGetLastError());
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->compute));        // the host staging vectors go out of scope
    return CB_OK;
}

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

// Decide whether this multiply runs a persistent variant (hub rows resident: K2H; gathers through a shared-memory ring:
// K2R; or both) and with which shape.  CB_OK with plan->active == false means "use K2".
int cb_hub_plan(cb_ctx* ctx, const cb_tile* tile, int64_t row_bytes, cudaStream_t stream, cbk::HubPlan* plan) {
    plan->active = false;
    plan->nhub = 0;
    static const int env_on = env_int("CB_SPMM_HUB", 0), env_cluster = env_int("CB_SPMM_HUB_CLUSTER", 4), env_slab = env_int("CB_SPMM_HUB_SLAB", 0),
                     env_smem_kb = env_int("CB_SPMM_HUB_SMEM_KB", 200), env_cover_pct = env_int("CB_SPMM_HUB_MIN_COVER_PCT", 15),
                     env_ring = env_int("CB_SPMM_RING", 0);
    const int hub_on = ctx->hub_enable >= 0 ? ctx->hub_enable : env_on;
    const int ring = ctx->ring_depth >= 0 ? ctx->ring_depth : env_ring;
    if (ring != 0 && ring != CB_RING_D) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "ring variant: depth %d (this build has 0 or %d)", ring, CB_RING_D);
    if ((!hub_on && !ring) || tile->nnz == 0 || tile->n >= (1LL << 31) || row_bytes < 128 || stream != ctx->compute) return CB_OK;
    int cluster = ctx->hub_cluster > 0 ? ctx->hub_cluster : env_cluster;
    if (cluster != 1 && cluster != 2 && cluster != 4 && cluster != 8) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "hub variant: cluster size %d (1, 2, 4 or 8)", cluster);
    int slab = ctx->hub_slab_bytes > 0 ? ctx->hub_slab_bytes : env_slab;
    if (slab == 0) slab = row_bytes <= 128 ? 128 : row_bytes <= 256 ? 256 : 512;
    if (slab != 128 && slab != 256 && slab != 512) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "hub variant: slab of %d bytes (128, 256 or 512)", slab);
    if ((row_bytes + slab - 1) / slab > CB_HUB_MAX_SLABS) return CB_OK;
    const int smem_kb = std::min(std::max(env_smem_kb, 16), 224);
    const size_t ring_bytes = (size_t)CB_HUB_BT * (size_t)ring * 16;                  // one ring of `ring` slots per virtual warp
    if (ring_bytes + (size_t)slab > (size_t)smem_kb * 1024) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "ring variant: %zu bytes of rings do not fit in %d KB", ring_bytes, smem_kb);
    int nhub = 0;
    cb_hub* h = nullptr;
    // hub rows: derived data of an immutable, owned tile, built once (views onto receive buffers are rebound per stage)
    if (hub_on && tile->owns_slab) {
        cb_tile* mt = const_cast<cb_tile*>(tile);
        if (!mt->hub || !mt->hub->built) { cb_hub_release(mt); CB_TRY(hub_build(ctx, mt)); }
        h = mt->hub;
        if (h->nhub_max > 0) {
            const int64_t slots = ((int64_t)smem_kb * 1024 - (int64_t)ring_bytes) / slab;
            nhub = (int)std::min<int64_t>(h->nhub_max, slots * cluster);
            const double cover = nhub > 0 ? (double)h->cum[(size_t)nhub - 1] / (double)tile->nnz : 0.0;
            h->last_cover = cover;
            if (cover * 100.0 < (double)env_cover_pct) nhub = 0;      // too few nonzeros would be served on the SMs
            h->last_nhub = nhub;
        }
    }
    if (nhub == 0 && !ring) return CB_OK;                              // nothing to gain over K2
    cb_tile* mt = const_cast<cb_tile*>(tile);
    if (!mt->hub) { mt->hub = new cb_hub(); }                          // ring without hub data still needs the chunk counters
    h = mt->hub;
    if (!h->counters) CB_CUDA(ctx, cudaMalloc(&h->counters, (size_t)CB_HUB_MAX_SLABS * sizeof(unsigned)));
    plan->active = true;
    plan->cluster = nhub > 0 ? cluster : 1;
    plan->slab_bytes = slab;
    plan->nhub = nhub;
    plan->ring = ring;
    plan->smem_bytes = (size_t)((nhub + plan->cluster - 1) / plan->cluster) * (size_t)slab + ring_bytes;
    if (plan->smem_bytes == 0) plan->smem_bytes = 16;
    plan->hubslot = nhub > 0 ? h->hubslot : nullptr;
    plan->hubcols = nhub > 0 ? h->hubcols : nullptr;
    plan->counters = h->counters;
    return CB_OK;
}

extern "C" {

int cb_spmm_hub_config(cb_ctx* ctx, int enable, int cluster, int slab_bytes) {
    if (!ctx) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_spmm_hub_config: null ctx");
    if (cluster != 0 && cluster != 1 && cluster != 2 && cluster != 4 && cluster != 8) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_spmm_hub_config: cluster size %d (0 = default, 1, 2, 4 or 8)", cluster);
    if (slab_bytes != 0 && slab_bytes != 128 && slab_bytes != 256 && slab_bytes != 512) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_spmm_hub_config: slab of %d bytes (0 = automatic, 128, 256 or 512)", slab_bytes);
    ctx->hub_enable = enable < 0 ? -1 : (enable ? 1 : 0);
    ctx->hub_cluster = cluster;
    ctx->hub_slab_bytes = slab_bytes;
    return CB_OK;
}

int cb_spmm_ring_config(cb_ctx* ctx, int depth) {
    if (!ctx) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_spmm_ring_config: null ctx");
    if (depth > 0 && depth != CB_RING_D) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_spmm_ring_config: depth %d (this build has 0 or %d)", depth, CB_RING_D);
    ctx->ring_depth = depth < 0 ? -1 : depth;
    return CB_OK;
}

int cb_spmm_hub_info(const cb_tile* tile, int64_t info[4]) {
    if (!tile || !info) return cb_fail(nullptr, CB_ERR_INVALIDPARAMS, "cb_spmm_hub_info: null argument");
    info[0] = tile->hub && tile->hub->built ? 1 : 0;
    info[1] = tile->hub ? tile->hub->nhub_max : 0;
    info[2] = tile->hub ? tile->hub->last_nhub : 0;
    info[3] = tile->hub ? (int64_t)(tile->hub->last_cover * 1e6) : 0;
    return CB_OK;
}

}  // extern "C"
