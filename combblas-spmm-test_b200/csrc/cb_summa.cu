// placeholder until the NCCL stage loop lands (next commit)
#include "cb_common.cuh"
int cb_nccl_init(cb_ctx* ctx, const void*) { return cb_fail(ctx, CB_ERR_NCCL, "multi-rank grids are not built yet"); }
void cb_nccl_destroy(cb_ctx*) {}
void cb_summa_release(cb_ctx*) {}
extern "C" {
int cb_comm_unique_id(void*) { return cb_fail(nullptr, CB_ERR_NCCL, "multi-rank grids are not built yet"); }
int cb_spmm_summa(cb_ctx* ctx, const cb_tile* t, const cb_dense* X, cb_dense* Y, int semiring, int64_t, int64_t, int64_t) {
    if (ctx->nranks == 1) return cb_spmm_local(ctx, t, X, Y, semiring, 0);
    return cb_fail(ctx, CB_ERR_NCCL, "multi-rank grids are not built yet");
}
int cb_summa_times(cb_ctx* ctx, float ms[4]) { for (int i = 0; i < 4; ++i) ms[i] = ctx->summa_ms[i]; return CB_OK; }
}
