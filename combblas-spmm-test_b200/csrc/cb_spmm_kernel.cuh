// K2: local SpMM on a doubly-compressed-row tile, one virtual warp per work chunk.
//
//   Y[r, 0:k] (+)= (+)_p  A.val[p] (x) X[A.col[p], 0:k]          for the nonzeros p of row r
//
// Replaces LocalHybridSpGEMM restricted to a dense right-hand side (reference
// include/CombBLAS/mtSpGEMM.h:213-460) and dcsc_gespmv generalised to k columns
// (include/CombBLAS/Friends.h:63-78).  No tensor cores: this is a gather-bound contraction; the
// design goal is bytes in flight, not flops.
//
// Mapping
//   * A row of X is k*sizeof(T) bytes = nvec 16-byte vectors.  A "virtual warp" (vw) of VW lanes
//     (VW = 4..32, power of two >= nvec/R) owns one row at a time; lane l holds R vectors.  A
//     hardware warp therefore runs 32/VW independent vws: a 256-byte row (k=64 fp32) uses
//     half-warps, a 128-byte row (k=32 int32) quarter... so every load instruction moves full
//     128-byte lines regardless of k.
//   * Work is nnz-balanced, not row-balanced: the tile builder cuts the nonzero stream into chunks
//     of ~L nonzeros at row boundaries and cuts rows longer than L at multiples of L
//     (cb_tile.cu).  One vw walks one chunk front to back.  Per step it loads VW (col,val) pairs
//     with one coalesced load, then for groups of U pairs issues all U row gathers back to back
//     (U*R 128-bit loads in flight per lane) before folding them into the accumulator in order.
//   * The last nonzero of each row carries a flag in the top bit of its column index, so the walk
//     never touches the row-pointer array; row ids come from the compressed nonempty-row list.
//   * The semiring is a functor pair inlined into the loop; the first product of a row is stored,
//     later ones are folded with add(product, acc) in ascending column order - the order of the
//     reference's hash accumulator (mtSpGEMM.h:395-423), so unsplit rows are bit-identical for
//     floating point as well.
//   * Pieces of split rows go to a carry buffer (2 slots per chunk) and are combined in chunk order
//     by cb_fixup_kernel: deterministic, no atomics.
#pragma once
#include "cb_common.cuh"

// Tuning (profiles/r01_tuning.md).  Two operating points per layout:
//   deep  : U=8 row gathers in flight per lane, 4 CTAs/SM (64 regs)  - best when the gathers mostly miss L2 (DRAM latency)
//   wide  : U=4, 5 CTAs/SM (48 regs)                                 - best when the used X rows mostly sit in L2 (more warps)
// launch_op picks by the footprint of the X rows the tile touches.

namespace cbk {

// ------------------------------------------------------------------------------ semiring functors
// T is the register type of one element.  Booleans travel as bytes 0/1 and are processed four at
// a time in a uint32_t (OR/AND are bitwise on 0/1 bytes).
enum AKind { A_SAME = 0, A_PATTERN = 1, A_BOOL = 2 };

template <typename T> struct Arith;
template <> struct Arith<float> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }   // no FMA contraction:
    static __device__ __forceinline__ float maxv() { return 3.402823466e+38f; }                 // multiply, then add
#ifdef CB_FMA      // experiment only (-DCB_FMA): one rounding instead of two; inside the 1e-5 tolerance, not bit-identical
    static __device__ __forceinline__ float madd(float a, float b, float c) { return __fmaf_rn(a, b, c); }
#else
    static __device__ __forceinline__ float madd(float a, float b, float c) { return __fadd_rn(__fmul_rn(a, b), c); }
#endif
};
template <> struct Arith<double> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double maxv() { return 1.7976931348623157e+308; }
#ifdef CB_FMA
    static __device__ __forceinline__ double madd(double a, double b, double c) { return __fma_rn(a, b, c); }
#else
    static __device__ __forceinline__ double madd(double a, double b, double c) { return __dadd_rn(__dmul_rn(a, b), c); }
#endif
};
template <> struct Arith<int32_t> {   // wrap-around like the host's two's complement arithmetic
    static __device__ __forceinline__ int32_t add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
    static __device__ __forceinline__ int32_t mul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
    static __device__ __forceinline__ int32_t maxv() { return 0x7fffffff; }
};
template <> struct Arith<int64_t> {
    static __device__ __forceinline__ int64_t add(int64_t a, int64_t b) { return (int64_t)((uint64_t)a + (uint64_t)b); }
    static __device__ __forceinline__ int64_t mul(int64_t a, int64_t b) { return (int64_t)((uint64_t)a * (uint64_t)b); }
    static __device__ __forceinline__ int64_t maxv() { return 0x7fffffffffffffffLL; }
};

// PlusTimesSRing<TA,T>  (Semirings.h:212-232)
template <typename T_, int AK>
struct PlusTimes {
    typedef T_ T;
    typedef typename std::conditional<AK == A_SAME, T_, uint8_t>::type TA;
    static constexpr int akind = AK;
    static constexpr bool first_touch = false;
    static __device__ __forceinline__ T id() { return T(0); }
    static __device__ __forceinline__ T add(T a, T b) { return Arith<T>::add(a, b); }
    static __device__ __forceinline__ T mul(TA a, T x) {
        if (AK == A_PATTERN) return x;                                   // static_cast<T>(true) * x
        if (AK == A_BOOL) return Arith<T>::mul((T)(a != 0), x);          // static_cast<T>(a) * x
        return Arith<T>::mul((T)a, x);
    }
#ifdef CB_FMA
    static __device__ __forceinline__ T madd(TA a, T x, T acc) {
        if constexpr (AK == A_SAME && (std::is_same<T, float>::value || std::is_same<T, double>::value)) return Arith<T>::madd((T)a, x, acc);
        else return add(mul(a, x), acc);
    }
#endif
};
// MinPlusSRing<T,T>  (Semirings.h:235-255, inf_plus :40-47)
template <typename T_>
struct MinPlus {
    typedef T_ T;
    typedef T_ TA;
    static constexpr int akind = A_SAME;
    static constexpr bool first_touch = false;
    static __device__ __forceinline__ T id() { return Arith<T>::maxv(); }
    static __device__ __forceinline__ T add(T a, T b) { return b < a ? b : a; }     // std::min(a, b)
    static __device__ __forceinline__ T mul(TA a, T x) {
        const T inf = Arith<T>::maxv();
        return (a == inf || x == inf) ? inf : Arith<T>::add(a, x);
    }
#ifdef CB_FMA
    static __device__ __forceinline__ T madd(TA a, T x, T acc) { return add(mul(a, x), acc); }
#endif
};
// SelectMaxSRing<bool,T>  (Semirings.h:191-210): multiply returns its second argument
template <typename T_>
struct SelectMax {
    typedef T_ T;
    typedef uint8_t TA;
    static constexpr int akind = A_PATTERN;   // the stored boolean is never read
    static constexpr bool first_touch = true;  // max(x, -1) != x for x < -1
    static __device__ __forceinline__ T id() { return T(-1); }
    static __device__ __forceinline__ T add(T a, T b) { return a < b ? b : a; }     // std::max(a, b)
    static __device__ __forceinline__ T mul(TA, T x) { return x; }
#ifdef CB_FMA
    static __device__ __forceinline__ T madd(TA a, T x, T acc) { return add(mul(a, x), acc); }
#endif
};
// PlusTimesSRing<bool,bool> (promote.h:78): + is OR, * is AND.  T = 4 packed 0/1 bytes.
template <int AK>
struct OrAnd {
    typedef uint32_t T;
    typedef uint8_t TA;
    static constexpr int akind = AK;
    static constexpr bool first_touch = false;
    static __device__ __forceinline__ T id() { return 0u; }
    static __device__ __forceinline__ T add(T a, T b) { return a | b; }
    static __device__ __forceinline__ T mul(TA a, T x) { return AK == A_PATTERN ? x : (x & (0u - (uint32_t)(a != 0))); }
#ifdef CB_FMA
    static __device__ __forceinline__ T madd(TA a, T x, T acc) { return add(mul(a, x), acc); }
#endif
};

// ------------------------------------------------------------------------------ 16-byte vectors
template <typename T> struct alignas(16) Vec16 { T v[16 / sizeof(T)]; };

template <typename T>
__device__ __forceinline__ Vec16<T> ldg16(const void* p) {
    // read-only path; X rows are re-used across the chip, keep them in L1/L2
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    Vec16<T> r;
    *reinterpret_cast<uint4*>(&r) = u;
    return r;
}
template <typename T>
__device__ __forceinline__ Vec16<T> ld16(const void* p) {
    Vec16<T> r;
    *reinterpret_cast<uint4*>(&r) = *reinterpret_cast<const uint4*>(p);
    return r;
}
template <typename T>
__device__ __forceinline__ void st16(void* p, const Vec16<T>& v) {
    *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(&v);
}

struct SpmmArgs {
    const int32_t* __restrict__ colflag;
    const void* __restrict__ vals;
    const int32_t* __restrict__ nzrows;
    const int32_t* __restrict__ chunk_start;
    const int32_t* __restrict__ chunk_row;
    int64_t nchunks;
    const char* __restrict__ X;   // row-major, ldx_bytes between rows
    char* __restrict__ Y;
    int64_t ldx_bytes, ldy_bytes;
    int row_bytes;                // bytes of one panel row handled per slab (<= VW*R*16)
    int slab_bytes;               // VW*R*16: byte offset between column slabs (gridDim.y)
    int total_row_bytes;          // k * sizeof(element)
    char* __restrict__ carry;     // [2*nchunks] slots of carry_stride bytes
    int64_t carry_stride;
    int accumulate;
    // K2P with L2 residency hints: per nonzero, floor(log2(rank + 1)) of its column among the tile's columns by descending
    // use (255 = used once); columns of class <= cls_max are gathered with evict_last, all others with evict_first
    const uint8_t* __restrict__ hubcls;
    int cls_max;
    // K2W: entries with bit 30 set carry the RANK of a much used column instead of the column; their rows sit in a packed panel
    // (kept in L2 by a persisting access-policy window) at X + win_delta, same row stride
    int64_t win_delta;
};

// One lane's share of a panel row: R vectors at byte offsets (vl + r*VW)*16.
template <class Op, int R>
struct RowFrag {
    Vec16<typename Op::T> v[R];
};

// streaming (evict-first) access for data touched once: the nonzero stream of A and the rows of Y.  Keeps L2 for X.
__device__ __forceinline__ int32_t ld_stream(const int32_t* p) { return __ldcs(p); }
template <typename TA> __device__ __forceinline__ TA ld_stream_val(const TA* p) { return __ldcs(p); }
template <> __device__ __forceinline__ uint8_t ld_stream_val<uint8_t>(const uint8_t* p) { return (uint8_t)__ldcs((const unsigned char*)p); }
template <typename T>
__device__ __forceinline__ void st16_stream(void* p, const Vec16<T>& v) {
    __stcs(reinterpret_cast<uint4*>(p), *reinterpret_cast<const uint4*>(&v));
}

#ifndef CB_CLUSTER_INTRINSICS     // the CPU warp emulator (tests/emul) supplies host versions of these
// L2 eviction priorities for the row gathers (createpolicy + ld.global.nc.L2::cache_hint): rows of the most used columns are
// kept (evict_last) while rows that are used once or twice pass through (evict_first) instead of pushing them out
__device__ __forceinline__ uint64_t cb_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t cb_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// pull a line of a stream into L2 ahead of its use (a hint: no register, no fault)
__device__ __forceinline__ void cb_prefetch_l2(const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }
__device__ __forceinline__ uint4 cb_ldg16_hint(const void* ptr, uint64_t policy) {
    uint4 u;
    asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(ptr), "l"(policy));
    return u;
}
// 16 bytes from shared memory of CTA `cta` of this cluster (distributed shared memory); addr is a shared-window address
__device__ __forceinline__ uint4 cb_ld_cluster16(uint32_t addr, uint32_t cta) {
    uint32_t remote;
    uint4 u;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(addr), "r"(cta));
    asm volatile("ld.shared::cluster.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(remote));
    return u;
}
#endif

// Hub-row source of the hub variant (cb_spmm_hub_kernel.cuh); the plain kernel passes an empty one and compiles none of it.
struct HubSrc {
    const uint16_t* __restrict__ hubslot;   // per nonzero: rank of its column among the tile's hub columns, 0xffff = not a hub
    int nhub;                               // ranks < nhub are resident in (distributed) shared memory
    uint32_t smem;                          // shared-window address of this CTA's slot 0, plus this lane's vector offset
    uint32_t slot_bytes;                    // bytes between slots
    uint32_t cta_mask, cta_shift;           // rank -> (CTA of the cluster, local slot): rank & mask, rank >> shift
};

template <typename T>
__device__ __forceinline__ Vec16<T> ld_hub16(const HubSrc& hub, uint32_t rank, int byte_off) {
    Vec16<T> r;
    *reinterpret_cast<uint4*>(&r) = cb_ld_cluster16(hub.smem + (rank >> hub.cta_shift) * hub.slot_bytes + (uint32_t)byte_off, rank & hub.cta_mask);
    return r;
}

// One virtual warp walks chunk `chunk` of the tile front to back (all virtual warps of the hardware warp in lock step).
// FULL: the panel row is exactly VW*R vectors wide (every lane owns real columns) - no lane predicates are generated.
// HUB : rows of hub columns come from shared memory of the cluster instead of global memory.
// PF  : the (column, value) entries of step s+1 are loaded while step s is processed, and the lines of both streams that will
//       be read PF_AHEAD entries later are pulled into L2 now - the profile of the round-1 kernel showed 27 % of the stall
//       samples on exactly these two loads (profiles/r02_*): they stream from DRAM once and nothing else hides them.
// WIN : K2W - entries flagged with bit 30 are gathered from the packed hub panel at X + win_delta (SpmmArgs).
template <class Op, int VW, int R, int U, bool FULL, bool HUB, bool PF = false, bool WIN = false>
__device__ __forceinline__ void cb_spmm_walk(const SpmmArgs& a, const int64_t chunk, const HubSrc& hub) {
    typedef typename Op::T T;
    typedef typename Op::TA TA;
    constexpr int EPL = 16 / sizeof(T);
    // CB_EXP_* (tools/gpu_r02_m.sh): timing-only builds that take one piece of the walk away at a time, to see what separates K2
    // from the arithmetic-free gather probe on L2-resident panels.  Their results are wrong by construction; never shipped.
#ifdef CB_EXP_NOVAL
    constexpr bool HASVAL = false;
#else
    constexpr bool HASVAL = Op::akind != A_PATTERN;
#endif
    const int lane = threadIdx.x & 31;
    const int vl = lane & (VW - 1);
    const int vshift = lane & ~(VW - 1);              // first lane of this virtual warp
    const bool live = chunk < a.nchunks;

    // this slab's columns
    const int slab_off = blockIdx.y * a.slab_bytes;
    const int slab_row_bytes = min(a.row_bytes, a.total_row_bytes - slab_off);
    bool lane_on[R];
#pragma unroll
    for (int r = 0; r < R; ++r) lane_on[r] = FULL || (vl + r * VW) * 16 < slab_row_bytes;
    const char* const xbase = a.X + slab_off + vl * 16;
    const uint32_t ldx = (uint32_t)a.ldx_bytes;

    int s = 0, e = 0, ridx = 0;
    bool head_open = false;
    if (live) {
        s = a.chunk_start[chunk];
        e = a.chunk_start[chunk + 1];
        const int cr = a.chunk_row[chunk];
        ridx = cr & 0x7fffffff;
        head_open = cr < 0;
    }
    // every lane of the warp runs the same number of steps so the shuffles stay convergent
    const int maxlen = __reduce_max_sync(0xffffffffu, e - s);

    const TA* __restrict__ vals = reinterpret_cast<const TA*>(a.vals);
    RowFrag<Op, R> acc;
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int q = 0; q < EPL; ++q) acc.v[r].v[q] = Op::id();
    bool first = true;                 // nothing folded into acc since the last row end
    int row = live ? a.nzrows[ridx] : 0;
    char* const carry_head = a.carry + (2 * chunk) * a.carry_stride + slab_off;

    // fold one product row into the accumulator
    auto fold = [&](const TA av, const RowFrag<Op, R>& x) {
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int q = 0; q < EPL; ++q) {
#ifdef CB_EXP_NOFP
                if (sizeof(T) == 4) { uint32_t u = *reinterpret_cast<const uint32_t*>(&x.v[r].v[q]) ^ *reinterpret_cast<uint32_t*>(&acc.v[r].v[q]); acc.v[r].v[q] = *reinterpret_cast<T*>(&u); }
                (void)av;
                continue;
#endif
#ifdef CB_FMA
                acc.v[r].v[q] = (Op::first_touch && first) ? Op::mul(av, x.v[r].v[q]) : Op::madd(av, x.v[r].v[q], acc.v[r].v[q]);
#else
                const T prod = Op::mul(av, x.v[r].v[q]);
                // the reference stores the first product of an output entry and folds later ones with
                // add(product, acc) (mtSpGEMM.h:403-414).  add(product, id) == product for every semiring here
                // except SelectMax with values below its identity, which keeps the explicit first-touch select.
                acc.v[r].v[q] = (Op::first_touch && first) ? prod : Op::add(prod, acc.v[r].v[q]);
#endif
            }
        first = false;
    };
    // last nonzero of a row seen: write the row out and start the next one
    auto flush = [&](bool more) {
        char* dst;
        bool rmw = false;
        if (head_open) { dst = carry_head; head_open = false; }            // piece of a split row
        else { dst = a.Y + (int64_t)row * a.ldy_bytes + slab_off; rmw = a.accumulate != 0; }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (lane_on[r]) {
                char* d = dst + (vl + r * VW) * 16;
                if (rmw) {
                    const Vec16<T> y = ld16<T>(d);
#pragma unroll
                    for (int q = 0; q < EPL; ++q) acc.v[r].v[q] = Op::add(y.v[q], acc.v[r].v[q]);
                }
#ifndef CB_EXP_NOFLUSH
                st16_stream<T>(d, acc.v[r]);
#endif
#pragma unroll
                for (int q = 0; q < EPL; ++q) acc.v[r].v[q] = Op::id();
            }
        }
        first = true;
        ++ridx;
#ifndef CB_EXP_NOFLUSH
        if (more) row = a.nzrows[ridx];       // (reading the ids two rows ahead was measured: slower, profiles/r02_k2_rowid_lookahead.jsonl)
#endif
    };

    constexpr int PF_AHEAD = 512;                        // entries: 2 KB of the column stream
    int cf_nxt = 0;
    TA av_nxt = TA();
    if (PF && vl < e - s) {
        cf_nxt = ld_stream(a.colflag + s + vl);
        if (HASVAL) av_nxt = ld_stream_val<TA>(vals + s + vl);
    }
    for (int base = 0; base < maxlen; base += VW) {
        const int rem = e - (s + base);                 // nonzeros this virtual warp still owns (may be <= 0)
        int cf = 0;
        TA av = TA();
        int hs = 0xffff;
        if (PF) {
            cf = cf_nxt;
            av = av_nxt;
            cf_nxt = 0;
            av_nxt = TA();
            if (VW + vl < rem) {
                cf_nxt = ld_stream(a.colflag + s + base + VW + vl);
                if (HASVAL) av_nxt = ld_stream_val<TA>(vals + s + base + VW + vl);
            }
            // one lane per virtual warp touches the line PF_AHEAD entries on (a 128-byte line holds 32 columns)
            if (vl == 0 && PF_AHEAD < rem && (VW >= 32 || ((base / VW) & (32 / VW - 1)) == 0)) {
                cb_prefetch_l2(a.colflag + s + base + PF_AHEAD);
                if (HASVAL) cb_prefetch_l2(vals + s + base + PF_AHEAD);
            }
        } else if (vl < rem) {
            cf = ld_stream(a.colflag + s + base + vl);
            if (HASVAL) av = ld_stream_val<TA>(vals + s + base + vl);
            if (HUB) hs = (int)__ldcs(hub.hubslot + s + base + vl);
        }
        // end-of-row flags of this virtual warp's VW entries, one bit each
#ifdef CB_EXP_NOROWEND
        const uint32_t fm = 0;
#else
        const uint32_t fm = (__ballot_sync(0xffffffffu, cf < 0) >> vshift) & (VW == 32 ? 0xffffffffu : ((1u << VW) - 1u));
#endif
        const uint32_t cm = (uint32_t)cf & 0x7fffffffu;
        // warp-uniform (every lane takes the same side of the branches below, so the shuffles inside them stay convergent):
        // all virtual warps of this warp still own a full block, hence no bounds predicates on the entries
        const bool fullwarp = __all_sync(0xffffffffu, rem >= VW);
#pragma unroll
        for (int j0 = 0; j0 < VW; j0 += U) {
            RowFrag<Op, R> x[U];
            TA avu[U];
            if (fullwarp) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint32_t c = __shfl_sync(0xffffffffu, cm, j0 + u, VW);
                    const char* xr = xbase + (uint64_t)(WIN ? (c & 0x3fffffffu) : c) * ldx;
                    if (WIN && (c & 0x40000000u)) xr += a.win_delta;
                    const int h = HUB ? __shfl_sync(0xffffffffu, hs, j0 + u, VW) : 0xffff;
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        if (lane_on[r]) {                                            // lane-constant predicate (panel narrower than VW*R vectors)
                            if (HUB && h < hub.nhub) x[u].v[r] = ld_hub16<T>(hub, (uint32_t)h, r * VW * 16);
                            else x[u].v[r] = ldg16<T>(xr + r * VW * 16);
                        }
                }
            } else {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint32_t c = __shfl_sync(0xffffffffu, cm, j0 + u, VW);
                    const char* xr = xbase + (uint64_t)(WIN ? (c & 0x3fffffffu) : c) * ldx;
                    if (WIN && (c & 0x40000000u)) xr += a.win_delta;
                    const int h = HUB ? __shfl_sync(0xffffffffu, hs, j0 + u, VW) : 0xffff;
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        if (j0 + u < rem && lane_on[r]) {
                            if (HUB && h < hub.nhub) x[u].v[r] = ld_hub16<T>(hub, (uint32_t)h, r * VW * 16);
                            else x[u].v[r] = ldg16<T>(xr + r * VW * 16);
                        }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) avu[u] = HASVAL ? (TA)__shfl_sync(0xffffffffu, av, j0 + u, VW) : TA();
            const uint32_t bits = (fm >> j0) & ((1u << U) - 1u);
            if (fullwarp && bits == 0) {
                // common case on long rows: U products, no row ends, no bounds
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
    }
    // a row still open at the end of the chunk continues in the next chunk: park the piece
    if (live && !first) {
        char* dst = carry_head + (head_open ? 0 : a.carry_stride);
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (lane_on[r]) st16<T>(dst + (vl + r * VW) * 16, acc.v[r]);
    }
}

// K2: one chunk per virtual warp, grid sized to the chunk count
template <class Op, int VW, int R, int U, int MINB, bool FULL, bool PF = false, bool WIN = false>
__global__ void __launch_bounds__(256, MINB)
cb_spmm_kernel(const SpmmArgs a) {
    constexpr int NV = 32 / VW;                       // virtual warps per warp
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    cb_spmm_walk<Op, VW, R, U, FULL, false, PF, WIN>(a, warp * NV + ((threadIdx.x & 31) / VW), HubSrc());
}

// K2 with PERSISTENT warps (opt-in, cb_spmm_k2_pipe(ctx, 64)): as many CTAs as fit the chip at once; every warp takes the next group
// of 32/VW chunks from a counter until none is left.  The plain kernel retires a CTA when its slowest warp is done and pays the
// launch of 15 000 - 130 000 CTAs of ~16 steps each: ncu shows 37 - 42 % of the warp slots occupied where the register limit allows
// 50 - 62 % (profiles/r02_k2w_s24f32_full.md, r01_c2_spmm_kernel_full.md).  Same walk, same bits.
template <class Op, int VW, int R, int U, int MINB, bool FULL>
__global__ void __launch_bounds__(256, MINB)
cb_spmm_persist_kernel(const SpmmArgs a, unsigned* __restrict__ counter) {
    constexpr int NV = 32 / VW;
    for (;;) {
        unsigned base = 0;
        if ((threadIdx.x & 31) == 0) base = atomicAdd(counter, (unsigned)NV);
        base = __shfl_sync(0xffffffffu, base, 0);
        if ((int64_t)base >= a.nchunks) break;                                 // warp-uniform
        cb_spmm_walk<Op, VW, R, U, FULL, false>(a, (int64_t)base + (threadIdx.x & 31) / VW, HubSrc());
    }
}

// ------------------------------------------------------------------------------------------------------------------
// K2P: the same walk with the row gathers PIPELINED THROUGH A REGISTER RING (opt-in, built only with -DCB_BUILD_K2P: measured slower than K2).
//
// What the profile of K2 said (profiles/r02_*): on L2-resident panels a warp spends only ~45 % of its time with gathers in
// flight - per step of VW nonzeros it waits once for the (column, value) entries and once per group of U gathers, and folds
// in between - so the SM sustains about half of what a pure gather loop reaches on the same hardware (tools/gather_probe.cu:
// 20 TB/s from L2, 7.3-7.8 TB/s from DRAM).  Here
//   * the entries of step s+1 are loaded while step s is processed (one step of prefetch, two registers);
//   * every lane keeps D row gathers in flight ALL the time: slot j mod D of a register ring holds the row of nonzero j;
//     right after nonzero j has been folded out of its slot, the gather of nonzero j+D is issued into the same slot - also
//     across step boundaries, since the next step's columns are already there.  The loop over the VW entries of a step is
//     fully unrolled, so the slots are static registers;
//   * three copies of the unrolled body, chosen per step by warp-uniform tests: no row end anywhere in the step and all
//     entries in range (straight-line fold / gather code), row ends but in range, and the general predicated one for the
//     last steps of a chunk.
// Fold order, row ends, split-row pieces and the accumulate mode are K2's, so every result bit is K2's.
template <class Op, int VW, int R, int D, bool FULL, bool POL>
__device__ __forceinline__ void cb_spmm_walk_pipe(const SpmmArgs& a, const int64_t chunk) {
    static_assert(D >= 1 && D <= VW && (VW % D) == 0, "ring depth must divide the virtual warp width");
    // POL: bit 30 of an entry's column word says "row of a much used column" (set when the entry is loaded, from hubcls);
    // it travels with the column through the shuffles and picks the L2 eviction priority of the gather (needs n < 2^30)
    constexpr uint32_t COLMASK = POL ? 0x3fffffffu : 0x7fffffffu;
    uint64_t pol_last = 0, pol_first = 0;
    if (POL) { pol_last = cb_policy_evict_last(); pol_first = cb_policy_evict_first(); }
    typedef typename Op::T T;
    typedef typename Op::TA TA;
    constexpr int EPL = 16 / sizeof(T);
    constexpr bool HASVAL = Op::akind != A_PATTERN;
    const int lane = threadIdx.x & 31;
    const int vl = lane & (VW - 1);
    const int vshift = lane & ~(VW - 1);
    const bool live = chunk < a.nchunks;

    const int slab_off = blockIdx.y * a.slab_bytes;
    const int slab_row_bytes = min(a.row_bytes, a.total_row_bytes - slab_off);
    bool lane_on[R];
#pragma unroll
    for (int r = 0; r < R; ++r) lane_on[r] = FULL || (vl + r * VW) * 16 < slab_row_bytes;
    const char* const xbase = a.X + slab_off + vl * 16;
    const uint32_t ldx = (uint32_t)a.ldx_bytes;

    int s = 0, e = 0, ridx = 0;
    bool head_open = false;
    if (live) {
        s = a.chunk_start[chunk];
        e = a.chunk_start[chunk + 1];
        const int cr = a.chunk_row[chunk];
        ridx = cr & 0x7fffffff;
        head_open = cr < 0;
    }
    const int len = e - s;
    const int maxlen = __reduce_max_sync(0xffffffffu, len);       // every lane of the warp runs the same number of steps

    const int32_t* const cfp = a.colflag + s + vl;
    const TA* const avp = reinterpret_cast<const TA*>(a.vals) + s + vl;
    RowFrag<Op, R> acc;
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int q = 0; q < EPL; ++q) acc.v[r].v[q] = Op::id();
    bool first = true;
    int row = live ? a.nzrows[ridx] : 0;
    char* const carry_head = a.carry + (2 * chunk) * a.carry_stride + slab_off;

    auto fold = [&](const TA av, const RowFrag<Op, R>& x) {
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int q = 0; q < EPL; ++q) {
#ifdef CB_FMA
                acc.v[r].v[q] = (Op::first_touch && first) ? Op::mul(av, x.v[r].v[q]) : Op::madd(av, x.v[r].v[q], acc.v[r].v[q]);
#else
                const T prod = Op::mul(av, x.v[r].v[q]);
                acc.v[r].v[q] = (Op::first_touch && first) ? prod : Op::add(prod, acc.v[r].v[q]);     // see cb_spmm_walk
#endif
            }
        first = false;
    };
    auto flush = [&](bool more) {
        char* dst;
        bool rmw = false;
        if (head_open) { dst = carry_head; head_open = false; }            // piece of a split row
        else { dst = a.Y + (int64_t)row * a.ldy_bytes + slab_off; rmw = a.accumulate != 0; }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (lane_on[r]) {
                char* d = dst + (vl + r * VW) * 16;
                if (rmw) {
                    const Vec16<T> y = ld16<T>(d);
#pragma unroll
                    for (int q = 0; q < EPL; ++q) acc.v[r].v[q] = Op::add(y.v[q], acc.v[r].v[q]);
                }
                st16_stream<T>(d, acc.v[r]);
#pragma unroll
                for (int q = 0; q < EPL; ++q) acc.v[r].v[q] = Op::id();
            }
        }
        first = true;
        ++ridx;
        if (more) row = a.nzrows[ridx];
    };
    auto gather = [&](RowFrag<Op, R>& dst, const uint32_t cw) {
        const char* xr = xbase + (uint64_t)(cw & COLMASK) * ldx;
        if (POL) {
            const uint64_t pol = (cw & 0x40000000u) ? pol_last : pol_first;
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (lane_on[r]) *reinterpret_cast<uint4*>(&dst.v[r]) = cb_ldg16_hint(xr + r * VW * 16, pol);
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (lane_on[r]) dst.v[r] = ldg16<T>(xr + r * VW * 16);
        }
    };
    // one entry of the nonzero stream: column | row-end flag (bit 31) | with POL: much-used-column mark (bit 30)
    const uint8_t* const hcp = POL ? a.hubcls + s + vl : nullptr;
    auto load_entry = [&](const int off) -> int {
        int cf = ld_stream(cfp + off);
        if (POL) cf |= ((int)__ldcs(hcp + off) <= a.cls_max) ? 0x40000000 : 0;
        return cf;
    };

    // entries of this step and of the next one: lane vl holds nonzero base + vl of the chunk
    int cf_cur = 0, cf_nxt = 0;
    TA av_cur = TA(), av_nxt = TA();
    if (vl < len) { cf_cur = load_entry(0); if (HASVAL) av_cur = ld_stream_val<TA>(avp); }
    if (VW + vl < len) { cf_nxt = load_entry(VW); if (HASVAL) av_nxt = ld_stream_val<TA>(avp + VW); }

    RowFrag<Op, R> x[D];
#pragma unroll
    for (int u = 0; u < D; ++u) {   // fill the ring: rows of the first D nonzeros
        const uint32_t c = (uint32_t)__shfl_sync(0xffffffffu, cf_cur, u, VW) & 0x7fffffffu;
        if (u < len) gather(x[u], c);
    }

    for (int base = 0; base < maxlen; base += VW) {
        const int rem = len - base;                         // nonzeros this virtual warp still owns (may be <= 0)
        const uint32_t fm = (__ballot_sync(0xffffffffu, cf_cur < 0) >> vshift) & (VW == 32 ? 0xffffffffu : ((1u << VW) - 1u));
        cf_cur &= 0x7fffffff;                               // the row-end flags now live in fm; what is left is the column
        // column of nonzero jn >= j + D of this step, or of the next step's entries (their flag is still in place)
        auto col_of = [&](const int jn) -> uint32_t {
            return jn < VW ? (uint32_t)__shfl_sync(0xffffffffu, cf_cur, jn, VW) : ((uint32_t)__shfl_sync(0xffffffffu, cf_nxt, jn - VW, VW) & 0x7fffffffu);
        };
        // warp-uniform: every virtual warp of the warp has this step and the whole next one in range / no row ends in this step
        const bool deep = __all_sync(0xffffffffu, rem >= 2 * VW);
        const bool ends = __any_sync(0xffffffffu, fm != 0u);
        if (deep && !ends) {
#pragma unroll
            for (int j = 0; j < VW; ++j) {
                const TA av = HASVAL ? (TA)__shfl_sync(0xffffffffu, av_cur, j, VW) : TA();
                fold(av, x[j % D]);
                gather(x[j % D], col_of(j + D));
            }
        } else if (deep) {
#pragma unroll
            for (int j = 0; j < VW; ++j) {
                const TA av = HASVAL ? (TA)__shfl_sync(0xffffffffu, av_cur, j, VW) : TA();
                fold(av, x[j % D]);
                if ((fm >> j) & 1u) flush(true);
                gather(x[j % D], col_of(j + D));
            }
        } else {
#pragma unroll
            for (int j = 0; j < VW; ++j) {
                const TA av = HASVAL ? (TA)__shfl_sync(0xffffffffu, av_cur, j, VW) : TA();
                if (j < rem) {
                    fold(av, x[j % D]);
                    if ((fm >> j) & 1u) flush(j + 1 < rem);
                }
                const uint32_t c = col_of(j + D);
                if (j + D < rem) gather(x[j % D], c);
            }
        }
        cf_cur = cf_nxt;
        av_cur = av_nxt;
        cf_nxt = 0;
        av_nxt = TA();
        if (base + 2 * VW + vl < len) {
            cf_nxt = load_entry(base + 2 * VW);
            if (HASVAL) av_nxt = ld_stream_val<TA>(avp + base + 2 * VW);
        }
    }
    // a row still open at the end of the chunk continues in the next chunk: park the piece
    if (live && !first) {
        char* dst = carry_head + (head_open ? 0 : a.carry_stride);
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (lane_on[r]) st16<T>(dst + (vl + r * VW) * 16, acc.v[r]);
    }
}

template <class Op, int VW, int R, int D, int MINB, bool FULL, bool POL = false>
__global__ void __launch_bounds__(256, MINB)
cb_spmm_pipe_kernel(const SpmmArgs a) {
    constexpr int NV = 32 / VW;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    cb_spmm_walk_pipe<Op, VW, R, D, FULL, POL>(a, warp * NV + ((threadIdx.x & 31) / VW));
}

// Combine the pieces of split rows in chunk order: Y[row] = tail[c0] (+) head[c0+1] (+) ... (+) head[c1].
// One virtual warp of 32 lanes per split row, looping over the row's vectors.
struct FixupArgs {
    const int32_t* __restrict__ split_row;   // indices into nzrows
    int64_t nsplit;
    const int32_t* __restrict__ nzrows;
    const int32_t* __restrict__ rowptr;
    int chunk_len;
    const char* __restrict__ carry;
    int64_t carry_stride;
    char* __restrict__ Y;
    int64_t ldy_bytes;
    int total_row_bytes;
    int accumulate;
};

template <class Op>
__global__ void __launch_bounds__(256)
cb_fixup_kernel(const FixupArgs a) {
    typedef typename Op::T T;
    constexpr int EPL = 16 / sizeof(T);
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= a.nsplit) return;
    const int ridx = a.split_row[w];
    const int rs = a.rowptr[ridx], re = a.rowptr[ridx + 1];
    const int c0 = rs / a.chunk_len, c1 = (re - 1) / a.chunk_len;
    char* yrow = a.Y + (int64_t)a.nzrows[ridx] * a.ldy_bytes;
    for (int b = lane * 16; b < a.total_row_bytes; b += 32 * 16) {
        Vec16<T> acc = ld16<T>(a.carry + (2 * (int64_t)c0 + 1) * a.carry_stride + b);
        for (int c = c0 + 1; c <= c1; ++c) {
            const Vec16<T> h = ld16<T>(a.carry + (2 * (int64_t)c) * a.carry_stride + b);
#pragma unroll
            for (int q = 0; q < EPL; ++q) acc.v[q] = Op::add(h.v[q], acc.v[q]);
        }
        if (a.accumulate) {
            const Vec16<T> y = ld16<T>(yrow + b);
#pragma unroll
            for (int q = 0; q < EPL; ++q) acc.v[q] = Op::add(y.v[q], acc.v[q]);
        }
        st16<T>(yrow + b, acc);
    }
}

}  // namespace cbk
