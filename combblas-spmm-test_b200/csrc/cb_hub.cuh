// Shared declarations of the hub variant of K2 (cb_hub.cu, cb_spmm_hub_kernel.cuh, cb_spmm_dispatch.cuh).
#pragma once
#include "cb_common.cuh"

#define CB_HUB_MAX_RANKS 65535      // ranks are 16-bit, 0xffff = "not a hub"
#define CB_HUB_MAX_SLABS 8192
// persistent variants: one CTA of CB_HUB_BT threads per SM; CB_RING_D row copies in flight per lane in the ring variant
#ifndef CB_HUB_BT
#define CB_HUB_BT 1024
#endif
#ifndef CB_RING_D
#define CB_RING_D 8
#endif

namespace cbk {
// shape of one launch of the persistent variants, decided by cb_hub_plan; active == false means "run plain K2"
struct HubPlan {
    int cluster = 1;                  // CTAs per cluster pooling their shared memory
    int slab_bytes = 0;               // bytes of a panel row handled per column slab: 128, 256 or 512
    bool active = false;              // run a persistent variant (K2H / K2R) instead of K2
    int nhub = 0;                     // hub ranks resident on the SMs (0 = none: the ring variant on a tile without hubs)
    int ring = 0;                     // ring depth of the pipelined variant K2R (0 = K2H, gathers through registers)
    size_t smem_bytes = 0;            // dynamic shared memory per CTA
    const uint16_t* hubslot = nullptr;
    const int32_t* hubcols = nullptr;
    unsigned* counters = nullptr;
};
}  // namespace cbk

// per-nonzero use class of its column, for the L2 residency hints of K2P (built lazily, owned tiles only); *cls == nullptr when
// the tile cannot have one
int cb_hubcls_get(cb_ctx* ctx, const cb_tile* tile, const uint8_t** cls);
// K2W: the column stream with the `want` most used columns replaced by bit 30 + rank, those columns by rank, and the packing of their X rows
int cb_hubwin_get(cb_ctx* ctx, const cb_tile* tile, int64_t want, const int32_t** colflag_w, const int32_t** cols, int64_t* h, double* cover);
int cb_hubwin_gather(cb_ctx* ctx, cudaStream_t stream, const void* X, int64_t ldx_bytes, const int32_t* cols, int64_t h, int row_bytes, void* panel);
int cb_hub_plan(cb_ctx* ctx, const cb_tile* tile, int64_t row_bytes, cudaStream_t stream, cbk::HubPlan* plan);
void cb_hub_release(cb_tile* tile);
extern "C" int cb_hub_select_host(const int32_t* counts, int64_t n, int max_hubs, int32_t* hubcols, int64_t* cum);
