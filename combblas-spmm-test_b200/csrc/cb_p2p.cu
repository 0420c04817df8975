// Copy-engine transport for the dense panels of the SUMMA stage loop: peer memory over NVLink, no SMs, no NCCL kernels.
//
// Replaces the column-communicator MPI_Bcast of the B operand (reference include/CombBLAS/ParFriends.h:1055-1068,
// SpParHelper.cpp:582-601) for ranks that are processes on one NVSwitch box:
//   * every rank owns a receive buffer `xfull` (all rows of its k-block of X) and a small mailbox; both are exported with
//     CUDA IPC once and mapped by the other ranks of the processor column;
//   * the owner of a stage's panel PUSHES it into each peer's xfull with cudaMemcpyAsync on a per-peer stream (DMA copy
//     engines over NVLink 5) and then raises that stage's flag in the peer's mailbox with a stream-ordered 32-bit write;
//   * the consumer's compute stream waits on the flag with a stream-ordered wait (cuStreamWaitValue32) right before the
//     stage's kernel - no host involvement, no spinning kernel;
//   * flow control is one ack per multiply: after its stage loop a rank writes the epoch into every peer's ack slot; the
//     receive buffer is double buffered by epoch parity, so a producer only waits for ack >= epoch-2 and the ranks of a
//     column never fall into lock step (with a single buffer the pushes of the two directions serialised).
// All panels of a multiply are in flight from the start, so transfers overlap the kernels of earlier stages completely.
// Epochs and flags are monotonic 32-bit counters compared cyclically (GEQ), never reset.
#include <cuda.h>
#include "cb_common.cuh"

namespace {

enum { MAXS = 64, MAXP = 64 };

struct Mailbox {
    uint32_t flags[MAXS];   // flags[s] = last epoch whose stage-s panel has fully arrived
    uint32_t acks[MAXP];    // acks[q]  = last epoch column-peer q has finished reading what I pushed to it
};

typedef CUresult (*WaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
typedef CUresult (*WriteValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

struct P2P {
    int np = 1, me = 0;
    char* xfull = nullptr;
    size_t xfull_bytes = 0;
    Mailbox* mailbox = nullptr;
    std::vector<char*> peer_xfull;
    std::vector<Mailbox*> peer_mailbox;
    std::vector<cudaStream_t> push;
    std::vector<cudaEvent_t> push_done, push_start;
    std::vector<char> used;             // peer stream touched in the current multiply
    uint32_t epoch = 0;
    WaitValue32Fn wait32 = nullptr;
    WriteValue32Fn write32 = nullptr;
    bool mailbox_mapped = false;
};

struct Handles {
    cudaIpcMemHandle_t xfull;
    cudaIpcMemHandle_t mailbox;
};

#define CB_CU(ctx, expr)                                                                                         \
    do {                                                                                                         \
        CUresult r__ = (expr);                                                                                   \
        if (r__ != CUDA_SUCCESS) return cb_fail((ctx), CB_ERR_CUDA, "%s failed: CUresult %d (%s:%d)", #expr, (int)r__, __FILE__, __LINE__); \
    } while (0)

}  // namespace


static P2P* get(cb_ctx* ctx) { return (P2P*)ctx->p2p_state; }

// Collective over the processor column.  Makes sure every rank's xfull holds `need` bytes and that all peers are mapped.
// Anything that can fail on one rank only (driver entry points, allocation, IPC) is recorded and carried through the
// collective steps, and the column agrees on the outcome at the end: if any rank could not set the transport up, every
// rank of the column switches to the NCCL broadcast path (ctx->summa_p2p = false) instead of failing the multiply.
int cb_p2p_prepare(cb_ctx* ctx, size_t need) {
    if (ctx->pr <= 1) return CB_OK;
    P2P* P = get(ctx);
    bool ok = true;
    const bool first = P == nullptr;      // every rank of the grid sets the transport up in its first stage loop: a grid-wide decision
    std::string why;
    auto soft = [&](cudaError_t e, const char* what) { if (e != cudaSuccess && ok) { ok = false; why = std::string(what) + ": " + cudaGetErrorString(e); } cudaGetLastError(); return e == cudaSuccess; };
    if (!P) {
        P = new P2P();
        P->np = ctx->pr;
        P->me = ctx->myprocrow;
        // size everything before the first step that can fail, so release() can always walk the vectors
        P->peer_xfull.assign(P->np, nullptr);
        P->peer_mailbox.assign(P->np, nullptr);
        P->push.assign(P->np, nullptr);
        P->push_done.assign(P->np, nullptr);
        P->push_start.assign(P->np, nullptr);
        P->used.assign(P->np, 0);
        ctx->p2p_state = P;
        if (P->np > MAXP) { ok = false; why = "more ranks per processor column than the peer transport supports"; }
        cudaDriverEntryPointQueryResult q;
        void *w = nullptr, *wr = nullptr;
        soft(cudaGetDriverEntryPoint("cuStreamWaitValue32", &w, cudaEnableDefault, &q), "cuStreamWaitValue32");
        soft(cudaGetDriverEntryPoint("cuStreamWriteValue32", &wr, cudaEnableDefault, &q), "cuStreamWriteValue32");
        if ((!w || !wr) && ok) { ok = false; why = "stream memory operations are not available in this driver"; }
        P->wait32 = (WaitValue32Fn)w;
        P->write32 = (WriteValue32Fn)wr;
        if (soft(cudaMalloc((void**)&P->mailbox, sizeof(Mailbox)), "cudaMalloc(mailbox)")) soft(cudaMemset(P->mailbox, 0, sizeof(Mailbox)), "cudaMemset(mailbox)");
        for (int q2 = 0; q2 < P->np && ok; ++q2) {
            if (q2 == P->me) continue;
            soft(cudaStreamCreateWithFlags(&P->push[q2], cudaStreamNonBlocking), "cudaStreamCreate");
            soft(cudaEventCreate(&P->push_done[q2]), "cudaEventCreate");
            soft(cudaEventCreate(&P->push_start[q2]), "cudaEventCreate");
        }
    }
    if (need <= P->xfull_bytes && P->mailbox_mapped) return CB_OK;
    // (re)allocate: nobody may still be writing into the old buffer.  Every rank drains its own pushes, then the
    // allgather below is the barrier after which all pushes of all ranks are known to be complete.
    for (int q = 0; q < P->np; ++q) if (P->push[q]) soft(cudaStreamSynchronize(P->push[q]), "cudaStreamSynchronize");
    soft(cudaStreamSynchronize(ctx->compute), "cudaStreamSynchronize");
    auto column_barrier = [&]() -> int { char token = 0; std::vector<char> all((size_t)P->np); return cb_nccl_allgather_col(ctx, &token, all.data(), 1); };
    CB_TRY(column_barrier());
    for (int q = 0; q < P->np; ++q)
        if (P->peer_xfull[q]) { soft(cudaIpcCloseMemHandle(P->peer_xfull[q]), "cudaIpcCloseMemHandle"); P->peer_xfull[q] = nullptr; }
    // an exporter may free its buffer only after every importer has closed its mapping (cudaFree before the importer's
    // cudaIpcCloseMemHandle is undefined behaviour): meet the column once more between closing and freeing
    CB_TRY(column_barrier());
    if (need > P->xfull_bytes) {
        if (P->xfull) soft(cudaFree(P->xfull), "cudaFree");
        P->xfull = nullptr;
        P->xfull_bytes = 0;
        // two buffers (even / odd epochs) in one allocation, with head room so a slightly larger panel does not remap
        const size_t bytes = (need + need / 4 + 4096 + 255) & ~(size_t)255;
        if (ok && soft(cudaMalloc((void**)&P->xfull, 2 * bytes), "cudaMalloc(peer panel buffers)")) P->xfull_bytes = bytes;
    }
    Handles mine;
    memset(&mine, 0, sizeof mine);
    if (ok && P->xfull) soft(cudaIpcGetMemHandle(&mine.xfull, P->xfull), "cudaIpcGetMemHandle");
    if (ok && P->mailbox) soft(cudaIpcGetMemHandle(&mine.mailbox, P->mailbox), "cudaIpcGetMemHandle");
    std::vector<Handles> all((size_t)P->np);
    CB_TRY(cb_nccl_allgather_col(ctx, &mine, all.data(), sizeof(Handles)));
    // did every rank get this far in one piece?  Only then may anyone open a peer's handles.
    std::vector<char> oks((size_t)P->np);
    {
        char mine_ok = ok ? 1 : 0;
        CB_TRY(cb_nccl_allgather_col(ctx, &mine_ok, oks.data(), 1));
    }
    bool all_ok = true;
    for (char c : oks) all_ok = all_ok && c;
    for (int q = 0; q < P->np && all_ok; ++q) {
        if (q == P->me) continue;
        soft(cudaIpcOpenMemHandle((void**)&P->peer_xfull[q], all[(size_t)q].xfull, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
        if (!P->peer_mailbox[q])
            soft(cudaIpcOpenMemHandle((void**)&P->peer_mailbox[q], all[(size_t)q].mailbox, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
    }
    {
        char mine_ok = ok ? 1 : 0;
        CB_TRY(cb_nccl_allgather_col(ctx, &mine_ok, oks.data(), 1));
    }
    for (char c : oks) all_ok = all_ok && c;
    if (first) {
        // the stage order depends on the transport and must be the same along a processor row (the broadcasts of A have to
        // match), so the first set-up is agreed on by the whole grid
        int64_t v = all_ok ? 1 : 0;
        CB_TRY(cb_comm_allreduce_i64(ctx, 0, 2, &v, 1));
        all_ok = v != 0;
    } else if (!all_ok) {
        return cb_fail(ctx, CB_ERR_CUDA, "growing the peer panel buffers failed on a rank of this processor column (%s)", ok ? "a peer" : why.c_str());
    }
    if (!all_ok) {
        // the column falls back to NCCL broadcasts together; what was mapped is closed again (a barrier apart from any free)
        for (int q = 0; q < P->np; ++q)
            if (P->peer_xfull[q]) { cudaIpcCloseMemHandle(P->peer_xfull[q]); P->peer_xfull[q] = nullptr; }
        cudaGetLastError();
        CB_TRY(column_barrier());
        P->mailbox_mapped = false;
        ctx->summa_p2p = false;
        if (!ok) fprintf(stderr, "combblas_b200: rank %d: peer panel transport unavailable (%s); this processor column uses NCCL broadcasts\n", ctx->rank, why.c_str());
        return CB_OK;
    }
    P->mailbox_mapped = true;
    return CB_OK;
}

char* cb_p2p_xfull(cb_ctx* ctx) { P2P* P = get(ctx); return P->xfull + (size_t)(P->epoch & 1u) * P->xfull_bytes; }

int cb_p2p_begin(cb_ctx* ctx) {
    P2P* P = get(ctx);
    ++P->epoch;
    std::fill(P->used.begin(), P->used.end(), 0);
    return CB_OK;
}

// producer: copy `bytes` from my X into every column peer's xfull at dst_off, then raise the peers' flag of `stage`
int cb_p2p_push(cb_ctx* ctx, cudaEvent_t operands_ready, int stage, size_t dst_off, const void* src, size_t bytes) {
    P2P* P = get(ctx);
    if (stage >= MAXS) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "peer transport supports up to %d stages", (int)MAXS);
    for (int q = 0; q < P->np; ++q) {
        if (q == P->me) continue;
        cudaStream_t st = P->push[q];
        if (!P->used[q]) {
            P->used[q] = 1;
            CB_CUDA(ctx, cudaStreamWaitEvent(st, operands_ready, 0));
            // the buffer of this parity was last used two multiplies ago: peer q must have finished reading that one.
            // (Receive buffers alternate between even and odd epochs so ranks do not fall into lock step.)
            CB_CU(ctx, P->wait32((CUstream)st, (CUdeviceptr)&P->mailbox->acks[q], P->epoch - 2, CU_STREAM_WAIT_VALUE_GEQ));
            CB_CUDA(ctx, cudaEventRecord(P->push_start[q], st));
        }
        if (bytes) CB_CUDA(ctx, cudaMemcpyAsync(P->peer_xfull[q] + (size_t)(P->epoch & 1u) * P->xfull_bytes + dst_off, src, bytes, cudaMemcpyDeviceToDevice, st));
        CB_CU(ctx, P->write32((CUstream)st, (CUdeviceptr)&P->peer_mailbox[q]->flags[stage], P->epoch, CU_STREAM_WRITE_VALUE_DEFAULT));
    }
    return CB_OK;
}

// consumer: the stage's panel must have landed before anything later on `stream` runs
int cb_p2p_wait_stage(cb_ctx* ctx, cudaStream_t stream, int stage) {
    P2P* P = get(ctx);
    CB_CU(ctx, P->wait32((CUstream)stream, (CUdeviceptr)&P->mailbox->flags[stage], P->epoch, CU_STREAM_WAIT_VALUE_GEQ));
    return CB_OK;
}

// end of the multiply on this rank: tell every peer its pushes have been consumed; X may be overwritten by the caller
// only after my own pushes have drained, so the compute stream waits for them
int cb_p2p_finish(cb_ctx* ctx, cudaStream_t compute) {
    P2P* P = get(ctx);
    for (int q = 0; q < P->np; ++q) {
        if (q == P->me) continue;
        CB_CU(ctx, P->write32((CUstream)compute, (CUdeviceptr)&P->peer_mailbox[q]->acks[P->me], P->epoch, CU_STREAM_WRITE_VALUE_DEFAULT));
        if (P->used[q]) {
            CB_CUDA(ctx, cudaEventRecord(P->push_done[q], P->push[q]));
            CB_CUDA(ctx, cudaStreamWaitEvent(compute, P->push_done[q], 0));
        }
    }
    return CB_OK;
}

// debug: when did the pushes of the last multiply start / end, in ms after `origin` (first peer)
int cb_p2p_debug_times(cb_ctx* ctx, cudaEvent_t origin, float* t_start, float* t_end) {
    P2P* P = get(ctx);
    *t_start = *t_end = -1;
    if (!P) return CB_OK;
    for (int q = 0; q < P->np; ++q) {
        if (q == P->me || !P->used[q]) continue;
        cudaEventSynchronize(P->push_done[q]);
        cudaEventElapsedTime(t_start, origin, P->push_start[q]);
        cudaEventElapsedTime(t_end, origin, P->push_done[q]);
        break;
    }
    cudaGetLastError();
    return CB_OK;
}

void cb_p2p_release(cb_ctx* ctx) {
    P2P* P = get(ctx);
    if (!P) return;
    // a peer may still be writing its last ack into my mailbox: drain my streams, then meet the column before freeing
    const int nq = (int)P->push.size();
    for (int q = 0; q < nq; ++q) if (P->push[q]) cudaStreamSynchronize(P->push[q]);
    cudaStreamSynchronize(ctx->compute);
    if (P->mailbox_mapped && ctx->nccl_col) { char token = 0; std::vector<char> all((size_t)P->np); cb_nccl_allgather_col(ctx, &token, all.data(), 1); }
    for (int q = 0; q < nq; ++q) {
        if (P->peer_xfull[q]) { cudaIpcCloseMemHandle(P->peer_xfull[q]); P->peer_xfull[q] = nullptr; }
        if (P->peer_mailbox[q]) { cudaIpcCloseMemHandle(P->peer_mailbox[q]); P->peer_mailbox[q] = nullptr; }
    }
    // importers have closed; only then may the exporters free (second meeting of the column)
    if (P->mailbox_mapped && ctx->nccl_col) { char token = 0; std::vector<char> all((size_t)P->np); cb_nccl_allgather_col(ctx, &token, all.data(), 1); }
    for (int q = 0; q < nq; ++q) {
        if (P->push[q]) { cudaStreamSynchronize(P->push[q]); cudaStreamDestroy(P->push[q]); }
        if (P->push_done[q]) cudaEventDestroy(P->push_done[q]);
        if (P->push_start[q]) cudaEventDestroy(P->push_start[q]);
        if (P->peer_xfull[q]) cudaIpcCloseMemHandle(P->peer_xfull[q]);
        if (P->peer_mailbox[q]) cudaIpcCloseMemHandle(P->peer_mailbox[q]);
    }
    cudaFree(P->xfull);
    cudaFree(P->mailbox);
    delete P;
    ctx->p2p_state = nullptr;
}
