// Synthetic operands generated on the device: Kronecker / R-MAT and Erdos-Renyi matrices, hash-valued panels.
//
// Replaces the host pipeline DistEdgeList::GenGraph500Data (reference include/CombBLAS/DistEdgeList.cpp:223-)
// -> SpParMat(const DistEdgeList&, bool) (SpParMat.cpp:3138-3254) with the recipe of
// ReleaseTests/GenWriteMatrix.cpp:96-131 (initiator .57/.19/.19/.05, edge factor 16, scrambled vertex ids, loops
// removed, A += A^T, duplicates merged).  The bit stream is our own counter-based one: edge e is a pure function of
// (seed, e), entry values of (seed, i, j), so every rank of any grid can make exactly its own block.  The identical
// integer recipe exists in numpy for the tests.
#include <cub/cub.cuh>
#include "cb_common.cuh"
#include "cb_gen500.cuh"

namespace {

__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

struct ScrambleKey { uint64_t m1, a1, m2, a2; };

__device__ inline uint64_t scramble(uint64_t v, int scale, const ScrambleKey& k) {
    const uint64_t mask = (scale >= 64) ? ~0ULL : ((1ULL << scale) - 1);
    v = (v * k.m1 + k.a1) & mask;
    v = __brevll(v) >> (64 - scale);
    v = (v * k.m2 + k.a2) & mask;
    return v;
}

__global__ void rmat_keys_kernel(int scale, int64_t nedges, uint64_t seed, uint32_t t1, uint32_t t2, uint32_t t3,
                                 ScrambleKey sk, int symmetric, int64_t row0, int64_t m, int64_t col0, int64_t n,
                                 uint64_t* __restrict__ keys) {
    const uint64_t base = seed << 40;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nedges; e += (int64_t)gridDim.x * blockDim.x) {
        uint64_t i = 0, j = 0, h = 0;
        for (int lvl = 0; lvl < scale; ++lvl) {
            if ((lvl & 3) == 0) h = splitmix64(base ^ ((uint64_t)e * 8 + (uint64_t)(lvl >> 2)));
            const uint32_t u = (uint32_t)(h >> (16 * (lvl & 3))) & 0xFFFFu;
            const uint64_t ib = u >= t2;                                   // quadrants c,d
            const uint64_t jb = ((u >= t1) && (u < t2)) || (u >= t3);      // quadrants b,d
            i = (i << 1) | ib;
            j = (j << 1) | jb;
        }
        i = scramble(i, scale, sk);
        j = scramble(j, scale, sk);
        const bool loop = (i == j);
        uint64_t k0 = ~0ULL, k1 = ~0ULL;
        const int64_t li = (int64_t)i - row0, lj = (int64_t)j - col0;
        if (!loop && li >= 0 && li < m && lj >= 0 && lj < n) k0 = ((uint64_t)li << 32) | (uint64_t)lj;
        keys[e] = k0;
        if (symmetric) {
            const int64_t ti = (int64_t)j - row0, tj = (int64_t)i - col0;
            if (!loop && ti >= 0 && ti < m && tj >= 0 && tj < n) k1 = ((uint64_t)ti << 32) | (uint64_t)tj;
            keys[nedges + e] = k1;
        }
    }
}

template <typename T> __device__ inline T hash_value(uint64_t h, int kind);
template <> __device__ inline float hash_value<float>(uint64_t h, int) { return (float)(2 * (h >> 41) + 1) * 5.9604644775390625e-08f; }        // 2^-24
template <> __device__ inline double hash_value<double>(uint64_t h, int) { return (double)(2 * (h >> 12) + 1) * 1.1102230246251565404e-16; }   // 2^-53
template <> __device__ inline int32_t hash_value<int32_t>(uint64_t h, int kind) {
    if (kind == 1 && (h & 0xFF) < 3) return 0x7fffffff;
    return (int32_t)(1 + (h >> 8) % 100);
}
template <> __device__ inline int64_t hash_value<int64_t>(uint64_t h, int kind) {
    if (kind == 1 && (h & 0xFF) < 3) return 0x7fffffffffffffffLL;
    return (int64_t)(1 + (h >> 8) % 100);
}
template <> __device__ inline uint8_t hash_value<uint8_t>(uint64_t h, int) { return (uint8_t)(h >> 63); }

template <typename T>
__global__ void tile_values_kernel(const uint64_t* __restrict__ keys, int64_t nz, uint64_t seedmul, int64_t row0, int64_t col0,
                                   uint64_t gn, T* __restrict__ vals) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nz; p += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t key = keys[p];
        const uint64_t gi = (key >> 32) + (uint64_t)row0, gj = (key & 0xFFFFFFFFu) + (uint64_t)col0;
        vals[p] = hash_value<T>(splitmix64(seedmul ^ (gi * gn + gj)), 0);
    }
}

template <typename T>
__global__ void dense_values_kernel(T* __restrict__ p, int64_t rows, int64_t cols, int64_t ld, uint64_t seedmul, int64_t row0,
                                    int64_t col0, int64_t gk, int kind) {
    const int64_t total = rows * cols;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = q / cols, c = q - r * cols;
        const uint64_t idx = (uint64_t)(row0 + r) * (uint64_t)gk + (uint64_t)(col0 + c);
        p[r * ld + c] = hash_value<T>(splitmix64(seedmul ^ idx), kind);
    }
}

inline int grid_for(int64_t n, int sm) {
    int64_t b = (n + 255) / 256, cap = (int64_t)sm * 32;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------------
// The reference's own edge stream: the Graph500 2.1 Kronecker generator as CombBLAS drives it with packed = true
// (include/CombBLAS/RefGen21.h:88-301, DistEdgeList::GenGraph500Data DistEdgeList.cpp:223-236) - what
// ReleaseTests/GenWriteMatrix.cpp builds its matrices from.  Restated from the published algorithm:
//   * random numbers: the multiple recursive generator z(n) = 107374182 z(n-1) + 104480 z(n-5) mod 2^31-1
//     (graph500-1.2/generator/splittable_mrg.c: mrg_orig_step); edge e starts from the seed state advanced by e * 2^64
//     steps (RefGen21.h:261 mrg_skip(&state, 0, ei, 0)), so the stream does not depend on how many processes generate it;
//   * skipping ahead is a matrix power: the state is a vector of Z_p^5 and one step the companion matrix A.  The reference
//     ships a generated table of A^(256^b v); here the powers needed (b = 8..11 for the edge index, the one jump of the
//     scramble values) are computed once on the host by repeated squaring and kept on the device (5 x 5 matrices mod p);
//   * one edge: lgN draws pick a quadrant each with the initiator .57 / .19 / .19 / .05 in units of 1/10000 without modulo
//     bias (RefGen21.h:104-134), edges are clipped and flipped into the upper triangle (:213-221), and both end points are
//     scrambled with two multiply / bit-reverse rounds keyed by two 64-bit values drawn from the same generator (:183-196, :227-240).
// Checked bit for bit against the reference's generator (tests/golden/graph500_ref.npz, made by the unmodified RefGen21).
namespace g500 {

__global__ void __launch_bounds__(256)
edges_kernel(const Tables* __restrict__ t, int lgN, int64_t first, int64_t count, int64_t* __restrict__ src, int64_t* __restrict__ dst) {
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < count; q += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t ei = (uint64_t)(first + q);
        State s = t->seed;
        for (int b = 0; b < 4; ++b) {
            const unsigned v = (unsigned)((ei >> (8 * b)) & 0xFF);
            if (v) s = apply(t->edge[b][v], s);
        }
        uint64_t a, c;
        one_edge(s, lgN, t->val0, t->val1, &a, &c);
        src[q] = (int64_t)a;
        dst[q] = (int64_t)c;
    }
}

// candidate keys of the tile block, as rmat_keys_kernel makes them: loops dropped on request, A + A^T on request
__global__ void __launch_bounds__(256)
keys_kernel(const Tables* __restrict__ t, int lgN, int64_t nedges, int symmetric, int remove_loops, int64_t row0, int64_t m, int64_t col0, int64_t n,
            uint64_t* __restrict__ keys) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nedges; e += (int64_t)gridDim.x * blockDim.x) {
        State s = t->seed;
        for (int b = 0; b < 4; ++b) {
            const unsigned v = (unsigned)(((uint64_t)e >> (8 * b)) & 0xFF);
            if (v) s = apply(t->edge[b][v], s);
        }
        uint64_t i, j;
        one_edge(s, lgN, t->val0, t->val1, &i, &j);
        const bool drop = remove_loops && i == j;
        uint64_t k0 = ~0ULL, k1 = ~0ULL;
        const int64_t li = (int64_t)i - row0, lj = (int64_t)j - col0;
        if (!drop && li >= 0 && li < m && lj >= 0 && lj < n) k0 = ((uint64_t)li << 32) | (uint64_t)lj;
        keys[e] = k0;
        if (symmetric) {
            const int64_t ti = (int64_t)j - row0, tj = (int64_t)i - col0;
            if (!drop && ti >= 0 && ti < m && tj >= 0 && tj < n) k1 = ((uint64_t)ti << 32) | (uint64_t)tj;
            keys[nedges + e] = k1;
        }
    }
}

static int upload_tables(cb_ctx* ctx, cb_scratch& sc, uint64_t u1, uint64_t u2, Tables** d_t) {
    static Tables host;                       // 100 KB; rebuilt when the seed changes
    static uint64_t have1 = ~0ull, have2 = ~0ull;
    if (have1 != u1 || have2 != u2) { build_tables(u1, u2, &host); have1 = u1; have2 = u2; }
    CB_CUDA(ctx, sc.alloc(d_t, 1));
    CB_CUDA(ctx, cudaMemcpyAsync(*d_t, &host, sizeof host, cudaMemcpyHostToDevice, ctx->compute));
    return CB_OK;
}

}  // namespace g500

// candidate keys (row << 32 | column, local indices; ~0 = dropped) -> sorted, merged -> values -> tile
template <typename T>
__global__ void counts_to_values_kernel(const int* __restrict__ counts, int64_t nz, T* __restrict__ vals) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nz; p += (int64_t)gridDim.x * blockDim.x) vals[p] = (T)counts[p];
}

// multiplicity != 0: the value of an entry is how many candidates merged into it (what SpParMat(DistEdgeList) + "A += A^T"
// give: duplicates summed, SpParMat.cpp:3138-3254) instead of a hash of its coordinates
static int keys_to_tile(cb_ctx* ctx, cb_scratch& sc, uint64_t* keys, uint64_t* keys_sorted, int64_t* d_count, int64_t ncand, int scale,
                        int64_t row0, int64_t m, int64_t col0, int64_t n, int val_dtype, uint64_t val_seed, cb_tile** out, int multiplicity = 0) {
    cudaStream_t st = ctx->compute;
    uint64_t* keys_unique;
    if (multiplicity && val_dtype != CB_PATTERN) {
        size_t b = 0;
        CB_CUDA(ctx, cub::DeviceRadixSort::SortKeys(nullptr, b, keys, keys_sorted, (int)ncand, 0, 64, st));
        void* tmp;
        CB_CUDA(ctx, sc.alloc((char**)&tmp, b));
        CB_CUDA(ctx, cub::DeviceRadixSort::SortKeys(tmp, b, keys, keys_sorted, (int)ncand, 0, 64, st));
        keys_unique = keys;
        int *counts = nullptr, *d_runs = nullptr;
        CB_CUDA(ctx, sc.alloc(&counts, (size_t)ncand));
        CB_CUDA(ctx, sc.alloc(&d_runs, 1));
        size_t b2 = 0;
        CB_CUDA(ctx, cub::DeviceRunLengthEncode::Encode(nullptr, b2, keys_sorted, keys_unique, counts, d_runs, (int)ncand, st));
        void* tmp2;
        CB_CUDA(ctx, sc.alloc((char**)&tmp2, b2));
        CB_CUDA(ctx, cub::DeviceRunLengthEncode::Encode(tmp2, b2, keys_sorted, keys_unique, counts, d_runs, (int)ncand, st));
        ctx->launches += 4;
        int runs = 0;
        uint64_t lastkey = 0;
        CB_CUDA(ctx, cudaMemcpyAsync(&runs, d_runs, sizeof runs, cudaMemcpyDeviceToHost, st));
        CB_CUDA(ctx, cudaStreamSynchronize(st));
        int64_t nuniq = runs;
        if (nuniq > 0) {
            CB_CUDA(ctx, cudaMemcpyAsync(&lastkey, keys_unique + (nuniq - 1), sizeof lastkey, cudaMemcpyDeviceToHost, st));
            CB_CUDA(ctx, cudaStreamSynchronize(st));
            if (lastkey == ~0ULL) --nuniq;
        }
        void* d_vals = nullptr;
        if (nuniq > 0) {
            CB_CUDA(ctx, sc.alloc((char**)&d_vals, cb_dtype_size(val_dtype) * (size_t)nuniq));
            const int g = grid_for(nuniq, ctx->sm_count);
            switch (val_dtype) {
                case CB_F32: counts_to_values_kernel<float><<<g, 256, 0, st>>>(counts, nuniq, (float*)d_vals); break;
                case CB_F64: counts_to_values_kernel<double><<<g, 256, 0, st>>>(counts, nuniq, (double*)d_vals); break;
                case CB_I32: counts_to_values_kernel<int32_t><<<g, 256, 0, st>>>(counts, nuniq, (int32_t*)d_vals); break;
                case CB_I64: counts_to_values_kernel<int64_t><<<g, 256, 0, st>>>(counts, nuniq, (int64_t*)d_vals); break;
                case CB_U8: counts_to_values_kernel<uint8_t><<<g, 256, 0, st>>>(counts, nuniq, (uint8_t*)d_vals); break;
            }
            CB_LAUNCHED(ctx);
            CB_CUDA(ctx, cudaGetLastError());
        }
        return cb_tile_build_from_keys(ctx, m, n, nuniq, keys_unique, d_vals, val_dtype, true, sc, out);
    }
    size_t b = 0;
    CB_CUDA(ctx, cub::DeviceRadixSort::SortKeys(nullptr, b, keys, keys_sorted, (int)ncand, 0, 64, st));
    void* tmp;
    CB_CUDA(ctx, sc.alloc((char**)&tmp, b));
    CB_CUDA(ctx, cub::DeviceRadixSort::SortKeys(tmp, b, keys, keys_sorted, (int)ncand, 0, 64, st));
    keys_unique = keys;           // reuse the unsorted buffer for the deduplicated stream
    size_t b2 = 0;
    CB_CUDA(ctx, cub::DeviceSelect::Unique(nullptr, b2, keys_sorted, keys_unique, d_count, (int)ncand, st));
    void* tmp2;
    CB_CUDA(ctx, sc.alloc((char**)&tmp2, b2));
    CB_CUDA(ctx, cub::DeviceSelect::Unique(tmp2, b2, keys_sorted, keys_unique, d_count, (int)ncand, st));
    ctx->launches += 4;
    int64_t nuniq = 0;
    uint64_t lastkey = 0;
    CB_CUDA(ctx, cudaMemcpyAsync(&nuniq, d_count, sizeof nuniq, cudaMemcpyDeviceToHost, st));
    CB_CUDA(ctx, cudaStreamSynchronize(st));
    if (nuniq > 0) {
        CB_CUDA(ctx, cudaMemcpyAsync(&lastkey, keys_unique + (nuniq - 1), sizeof lastkey, cudaMemcpyDeviceToHost, st));
        CB_CUDA(ctx, cudaStreamSynchronize(st));
        if (lastkey == ~0ULL) --nuniq;         // the bucket of dropped candidates (loops, other ranks' blocks)
    }
    void* d_vals = nullptr;
    if (val_dtype != CB_PATTERN && nuniq > 0) {
        CB_CUDA(ctx, sc.alloc((char**)&d_vals, cb_dtype_size(val_dtype) * (size_t)nuniq));
        const uint64_t seedmul = val_seed * 0x100000001B3ULL, gn = 1ULL << scale;
        const int g = grid_for(nuniq, ctx->sm_count);
        switch (val_dtype) {
            case CB_F32: tile_values_kernel<float><<<g, 256, 0, st>>>(keys_unique, nuniq, seedmul, row0, col0, gn, (float*)d_vals); break;
            case CB_F64: tile_values_kernel<double><<<g, 256, 0, st>>>(keys_unique, nuniq, seedmul, row0, col0, gn, (double*)d_vals); break;
            case CB_I32: tile_values_kernel<int32_t><<<g, 256, 0, st>>>(keys_unique, nuniq, seedmul, row0, col0, gn, (int32_t*)d_vals); break;
            case CB_I64: tile_values_kernel<int64_t><<<g, 256, 0, st>>>(keys_unique, nuniq, seedmul, row0, col0, gn, (int64_t*)d_vals); break;
            case CB_U8: tile_values_kernel<uint8_t><<<g, 256, 0, st>>>(keys_unique, nuniq, seedmul, row0, col0, gn, (uint8_t*)d_vals); break;
        }
        CB_LAUNCHED(ctx);
        CB_CUDA(ctx, cudaGetLastError());
    }
    return cb_tile_build_from_keys(ctx, m, n, nuniq, keys_unique, d_vals, val_dtype, true, sc, out);
}

extern "C" {

int cb_gen_rmat_tile(cb_ctx* ctx, int scale, int edgefactor, uint64_t seed, const double initiator[4], int symmetric,
                     int64_t row0, int64_t m, int64_t col0, int64_t n, int val_dtype, uint64_t val_seed, cb_tile** out) {
    *out = nullptr;
    if (scale < 1 || scale > 31 || edgefactor < 1 || m < 0 || n < 0 || row0 < 0 || col0 < 0)
        return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_gen_rmat_tile: scale must be 1..31 and the block inside the matrix");
    if (val_dtype != CB_PATTERN && !cb_dtype_size(val_dtype)) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_gen_rmat_tile: value dtype %d", val_dtype);
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->compute;
    const int64_t nedges = (int64_t)edgefactor << scale;
    const int64_t ncand = symmetric ? 2 * nedges : nedges;
    if (ncand >= (int64_t(1) << 31)) return cb_fail(ctx, CB_ERR_TOO_LARGE, "cb_gen_rmat_tile: %lld candidate edges", (long long)ncand);
    const uint32_t t1 = (uint32_t)llround(initiator[0] * 65536.0), t2 = (uint32_t)llround((initiator[0] + initiator[1]) * 65536.0),
                   t3 = (uint32_t)llround((initiator[0] + initiator[1] + initiator[2]) * 65536.0);
    ScrambleKey sk;
    const uint64_t s1 = splitmix64(seed ^ 0x5CA1AB1EULL), s2 = splitmix64(s1);
    sk.m1 = s1 | 1; sk.a1 = s1 >> 32; sk.m2 = s2 | 1; sk.a2 = s2 >> 32;

    cb_scratch sc;
    uint64_t *keys, *keys_sorted;
    int64_t* d_count;
    CB_CUDA(ctx, sc.alloc(&keys, (size_t)ncand));
    CB_CUDA(ctx, sc.alloc(&keys_sorted, (size_t)ncand));
    CB_CUDA(ctx, sc.alloc(&d_count, 1));
    rmat_keys_kernel<<<grid_for(nedges, ctx->sm_count), 256, 0, st>>>(scale, nedges, seed, t1, t2, t3, sk, symmetric, row0, m, col0, n, keys);
    CB_LAUNCHED(ctx);
    CB_CUDA(ctx, cudaGetLastError());
    return keys_to_tile(ctx, sc, keys, keys_sorted, d_count, ncand, scale, row0, m, col0, n, val_dtype, val_seed, out);
}

// Edges [first, first + count) of the Graph500 2.1 stream the reference generates (RefGen21::generate_kronecker_range), as they
// leave the generator: global vertex ids, duplicates and loops included, source <= target before scrambling.
int cb_gen_graph500_edges(cb_ctx* ctx, int log_numverts, uint64_t userseed1, uint64_t userseed2, int64_t first, int64_t count, int64_t* src_host, int64_t* dst_host) {
    if (!ctx || log_numverts < 1 || log_numverts > 40 || first < 0 || count < 0 || first + count > (int64_t(1) << 32) || (count > 0 && (!src_host || !dst_host)))
        return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_gen_graph500_edges: bad arguments (edge indices below 2^32)");
    if (count == 0) return CB_OK;
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cb_scratch sc;
    g500::Tables* d_t = nullptr;
    CB_TRY(g500::upload_tables(ctx, sc, userseed1, userseed2, &d_t));
    int64_t *d_src = nullptr, *d_dst = nullptr;
    CB_CUDA(ctx, sc.alloc(&d_src, (size_t)count));
    CB_CUDA(ctx, sc.alloc(&d_dst, (size_t)count));
    g500::edges_kernel<<<grid_for(count, ctx->sm_count), 256, 0, ctx->compute>>>(d_t, log_numverts, first, count, d_src, d_dst);
    CB_LAUNCHED(ctx);
    CB_CUDA(ctx, cudaGetLastError());
    CB_CUDA(ctx, cudaMemcpyAsync(src_host, d_src, sizeof(int64_t) * (size_t)count, cudaMemcpyDeviceToHost, ctx->compute));
    CB_CUDA(ctx, cudaMemcpyAsync(dst_host, d_dst, sizeof(int64_t) * (size_t)count, cudaMemcpyDeviceToHost, ctx->compute));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
    return CB_OK;
}

// The block [row0, row0+m) x [col0, col0+n) of the matrix ReleaseTests/GenWriteMatrix.cpp:96-131 builds from that stream:
// edgefactor * 2^scale edges, optional loop removal, optional A += A^T, duplicates merged.  userseed 0, 0 is the reference's
// -DDETERMINISTIC stream (RefGen21.h:274-275).
int cb_gen_graph500_tile(cb_ctx* ctx, int scale, int edgefactor, uint64_t userseed1, uint64_t userseed2, int symmetric, int remove_loops,
                         int64_t row0, int64_t m, int64_t col0, int64_t n, int val_dtype, uint64_t val_seed, cb_tile** out) {
    // val_seed == 0: values are multiplicities (the reference's SpParMat(DistEdgeList) sums duplicates); otherwise hashed weights
    if (!out) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_gen_graph500_tile: null argument");
    *out = nullptr;
    if (scale < 1 || scale > 31 || edgefactor < 1 || m < 0 || n < 0 || row0 < 0 || col0 < 0)
        return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_gen_graph500_tile: scale must be 1..31 and the block inside the matrix");
    if (val_dtype != CB_PATTERN && !cb_dtype_size(val_dtype)) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_gen_graph500_tile: value dtype %d", val_dtype);
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t nedges = (int64_t)edgefactor << scale;
    const int64_t ncand = symmetric ? 2 * nedges : nedges;
    if (ncand >= (int64_t(1) << 31)) return cb_fail(ctx, CB_ERR_TOO_LARGE, "cb_gen_graph500_tile: %lld candidate edges", (long long)ncand);
    cb_scratch sc;
    g500::Tables* d_t = nullptr;
    CB_TRY(g500::upload_tables(ctx, sc, userseed1, userseed2, &d_t));
    uint64_t *keys, *keys_sorted;
    int64_t* d_count;
    CB_CUDA(ctx, sc.alloc(&keys, (size_t)ncand));
    CB_CUDA(ctx, sc.alloc(&keys_sorted, (size_t)ncand));
    CB_CUDA(ctx, sc.alloc(&d_count, 1));
    g500::keys_kernel<<<grid_for(nedges, ctx->sm_count), 256, 0, ctx->compute>>>(d_t, scale, nedges, symmetric, remove_loops, row0, m, col0, n, keys);
    CB_LAUNCHED(ctx);
    CB_CUDA(ctx, cudaGetLastError());
    return keys_to_tile(ctx, sc, keys, keys_sorted, d_count, ncand, scale, row0, m, col0, n, val_dtype, val_seed, out, val_seed == 0 ? 1 : 0);
}

int cb_gen_dense(cb_dense* d, uint64_t seed, int64_t row0, int64_t col0, int64_t gk, int kind) {
    cb_ctx* ctx = d->ctx;
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t total = d->rows * d->cols;
    if (total == 0) return CB_OK;
    const uint64_t seedmul = seed * 0x100000001B3ULL;
    const int g = grid_for(total, ctx->sm_count);
    cudaStream_t st = ctx->compute;
    switch (d->dtype) {
        case CB_F32: dense_values_kernel<float><<<g, 256, 0, st>>>((float*)d->ptr, d->rows, d->cols, d->ld, seedmul, row0, col0, gk, kind); break;
        case CB_F64: dense_values_kernel<double><<<g, 256, 0, st>>>((double*)d->ptr, d->rows, d->cols, d->ld, seedmul, row0, col0, gk, kind); break;
        case CB_I32: dense_values_kernel<int32_t><<<g, 256, 0, st>>>((int32_t*)d->ptr, d->rows, d->cols, d->ld, seedmul, row0, col0, gk, kind); break;
        case CB_I64: dense_values_kernel<int64_t><<<g, 256, 0, st>>>((int64_t*)d->ptr, d->rows, d->cols, d->ld, seedmul, row0, col0, gk, kind); break;
        case CB_U8: dense_values_kernel<uint8_t><<<g, 256, 0, st>>>((uint8_t*)d->ptr, d->rows, d->cols, d->ld, seedmul, row0, col0, gk, kind); break;
        default: return cb_fail(ctx, CB_ERR_UNSUPPORTED, "cb_gen_dense: dtype %d", d->dtype);
    }
    CB_LAUNCHED(ctx);
    CB_CUDA(ctx, cudaGetLastError());
    return CB_OK;
}

}  // extern "C"
