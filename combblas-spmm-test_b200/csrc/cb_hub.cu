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
};

void cb_hub_release(cb_tile* t) {
    if (!t || !t->hub) return;
    cudaFree(t->hub->hubslot);
    cudaFree(t->hub->hubcols);
    cudaFree(t->hub->counters);
    cudaFree(t->hub->hubcls);
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
