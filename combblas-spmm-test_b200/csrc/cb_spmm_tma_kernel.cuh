// K2T: the K2 walk with the row gathers issued as BULK ASYNCHRONOUS COPIES (cp.async.bulk, the non-tensor TMA path:
// UBLKCP in SASS) into a per-warp shared-memory ring, completion counted on mbarriers.
//
// What it changes against K2: a gather is no longer 32 lanes x one 16-byte load each with a destination register, but ONE
// instruction of one lane that moves the whole row (VW*16 bytes) from global memory into a ring slot.  A warp step still
// covers 32 nonzeros (32/VW virtual warps x VW entries): every lane issues the copy of "its" entry's row, so a single warp
// instruction puts 32 rows in flight; lane 0 tells the stage's mbarrier how many bytes to expect.  S stages per warp: while
// step t is folded out of shared memory (one LDS.128 per lane and row), the copies of steps t+1 .. t+S-1 are in flight.
// Fold order, row ends, split-row pieces and the accumulate mode are K2's, so every result bit is K2's.
//
// Restrictions of this experiment: rows that fill the layout exactly (FULL), one vector per lane (R = 1), VW in {8, 16, 32},
// no column slabs.  Everything else stays on K2.
#pragma once
#include "cb_spmm_hub_kernel.cuh"      // cb_lds16

namespace cbk {

#ifndef CB_CLUSTER_INTRINSICS
__device__ __forceinline__ uint32_t cb_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cb_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void cb_mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void cb_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cb_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// one row: `bytes` (multiple of 16) from global memory to shared memory, completion reported to the mbarrier as bytes
__device__ __forceinline__ void cb_bulk_row(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
#endif

template <class Op, int VW, int S, int U>
__global__ void __launch_bounds__(384, 1) cb_spmm_tma_kernel(const SpmmArgs a) {
    typedef typename Op::T T;
    typedef typename Op::TA TA;
    constexpr int EPL = 16 / sizeof(T);
    constexpr bool HASVAL = Op::akind != A_PATTERN;
    constexpr int NV = 32 / VW;
    constexpr uint32_t ROWB = VW * 16;
    constexpr uint32_t STAGEB = 32 * ROWB;                 // one slot per lane
    extern __shared__ __align__(128) char cb_tma_smem[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    const int vl = lane & (VW - 1);
    const int vshift = lane & ~(VW - 1);
    const uint32_t ring = cb_smem_addr(cb_tma_smem) + (uint32_t)wib * S * STAGEB;
    const uint32_t bars = cb_smem_addr(cb_tma_smem) + (uint32_t)nwarps * S * STAGEB + (uint32_t)wib * S * 8;
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < S; ++q) cb_mbar_init(bars + q * 8, 1);
        cb_mbar_fence_init();
    }
    __syncthreads();

    const int64_t warp = (int64_t)blockIdx.x * nwarps + wib;
    const int64_t chunk = warp * NV + lane / VW;
    const bool live = chunk < a.nchunks;
    const char* const xbase = a.X;
    const uint64_t ldx = (uint64_t)a.ldx_bytes;
    int s = 0, e = 0, ridx = 0;
    bool head_open = false;
    if (live) {
        s = a.chunk_start[chunk];
        e = a.chunk_start[chunk + 1];
        const int cr = a.chunk_row[chunk];
        ridx = cr & 0x7fffffff;
        head_open = cr < 0;
    }
    const int maxlen = __reduce_max_sync(0xffffffffu, e - s);
    const int nsteps = (maxlen + VW - 1) / VW;
    const TA* __restrict__ vals = reinterpret_cast<const TA*>(a.vals);
    Vec16<T> acc;
#pragma unroll
    for (int q = 0; q < EPL; ++q) acc.v[q] = Op::id();
    bool first = true;
    int row = live ? a.nzrows[ridx] : 0;
    char* const carry_head = a.carry + (2 * chunk) * a.carry_stride;

    auto fold = [&](const TA av, const Vec16<T>& x) {
#pragma unroll
        for (int q = 0; q < EPL; ++q) {
            const T prod = Op::mul(av, x.v[q]);
            acc.v[q] = (Op::first_touch && first) ? prod : Op::add(prod, acc.v[q]);      // mtSpGEMM.h:403-414, as in K2
        }
        first = false;
    };
    auto flush = [&](bool more) {
        char* dst;
        bool rmw = false;
        if (head_open) { dst = carry_head; head_open = false; }
        else { dst = a.Y + (int64_t)row * a.ldy_bytes; rmw = a.accumulate != 0; }
        char* d = dst + vl * 16;
        if (rmw) {
            const Vec16<T> y = ld16<T>(d);
#pragma unroll
            for (int q = 0; q < EPL; ++q) acc.v[q] = Op::add(y.v[q], acc.v[q]);
        }
        st16_stream<T>(d, acc);
#pragma unroll
        for (int q = 0; q < EPL; ++q) acc.v[q] = Op::id();
        first = true;
        ++ridx;
        if (more) row = a.nzrows[ridx];
    };

    int cfq[S];
    TA avq[S];
    // load the entries of step t and put the copies of their rows in flight into stage q
    auto issue = [&](const int t, const int q) {
        const int rem = e - (s + t * VW);
        const bool valid = vl < rem;
        int cf = 0;
        TA av = TA();
        if (valid) {
            cf = ld_stream(a.colflag + s + t * VW + vl);
            if (HASVAL) av = ld_stream_val<TA>(vals + s + t * VW + vl);
        }
        cfq[q] = cf;
        avq[q] = av;
        const uint32_t nvalid = __popc(__ballot_sync(0xffffffffu, valid));
        const uint32_t bar = bars + q * 8;
        if (lane == 0) cb_mbar_expect_tx(bar, nvalid * ROWB);
        __syncwarp();
        if (valid) cb_bulk_row(ring + q * STAGEB + lane * ROWB, xbase + (uint64_t)((uint32_t)cf & 0x7fffffffu) * ldx, ROWB, bar);
    };
    auto consume = [&](const int t, const int q, const uint32_t parity) {
        cb_mbar_wait(bars + q * 8, parity);
        const int rem = e - (s + t * VW);
        const int cf = cfq[q];
        const TA av = avq[q];
        const uint32_t fm = (__ballot_sync(0xffffffffu, cf < 0) >> vshift) & (VW == 32 ? 0xffffffffu : ((1u << VW) - 1u));
        const bool fullwarp = __all_sync(0xffffffffu, rem >= VW);
        const uint32_t src = ring + q * STAGEB + vshift * ROWB + vl * 16;
#pragma unroll
        for (int j0 = 0; j0 < VW; j0 += U) {
            Vec16<T> x[U];
            TA avu[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (fullwarp || j0 + u < rem) *reinterpret_cast<uint4*>(&x[u]) = cb_lds16(src + (j0 + u) * ROWB);
                avu[u] = HASVAL ? (TA)__shfl_sync(0xffffffffu, av, j0 + u, VW) : TA();
            }
            const uint32_t bits = (fm >> j0) & ((1u << U) - 1u);
            if (fullwarp && bits == 0) {
#pragma unroll
                for (int u = 0; u < U; ++u) fold(avu[u], x[u]);
            } else {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (j0 + u < rem) {
                        fold(avu[u], x[u]);
                        if ((bits >> u) & 1u) flush(j0 + u + 1 < rem);
                    }
                }
            }
        }
        __syncwarp();                                      // every lane is done with the stage before it is filled again
    };

#pragma unroll
    for (int q = 0; q < S - 1; ++q)
        if (q < nsteps) issue(q, q);
    for (int t0 = 0; t0 < nsteps; t0 += S) {
        const uint32_t parity = (uint32_t)(t0 / S) & 1u;
#pragma unroll
        for (int q = 0; q < S; ++q) {
            const int t = t0 + q;
            if (t < nsteps) {                              // warp-uniform
                if (t + S - 1 < nsteps) issue(t + S - 1, (q + S - 1) % S);
                consume(t, q, parity);
            }
        }
    }
    if (live && !first) st16<T>(carry_head + (head_open ? 0 : a.carry_stride) + vl * 16, acc);
}

}  // namespace cbk
