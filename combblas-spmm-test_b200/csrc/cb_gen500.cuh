// The Graph500 2.1 Kronecker edge stream as the reference drives it (include/CombBLAS/RefGen21.h:88-301): the arithmetic, usable on the
// host and on the device.  See cb_gen.cu for what it replaces; tests/hostgen/ runs the host side against the reference's own generator.
#pragma once
#include <cstdint>
#include <cstring>
#ifndef __CUDACC__
#define __host__
#define __device__
#endif

namespace g500 {

constexpr uint64_t P = 0x7fffffffull;          // 2^31 - 1
struct Mat { uint32_t a[5][5]; };
struct State { uint32_t z[5]; };               // z[0] = z1 (newest) ... z[4] = z5

__host__ __device__ inline uint32_t mulmod(uint64_t x, uint64_t y) { return (uint32_t)((x * y) % P); }

static Mat mat_mul(const Mat& x, const Mat& y) {
    Mat r;
    for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 5; ++j) {
            uint64_t acc = 0;
            for (int k = 0; k < 5; ++k) acc += mulmod(x.a[i][k], y.a[k][j]);
            r.a[i][j] = (uint32_t)(acc % P);
        }
    return r;
}
static Mat mat_identity() { Mat m; memset(&m, 0, sizeof m); for (int i = 0; i < 5; ++i) m.a[i][i] = 1; return m; }
static Mat mat_step() {                        // new state = A * old state: z1' = 107374182 z1 + 104480 z5, the others shift
    Mat m; memset(&m, 0, sizeof m);
    m.a[0][0] = 107374182u; m.a[0][4] = 104480u;
    for (int i = 1; i < 5; ++i) m.a[i][i - 1] = 1;
    return m;
}
static Mat mat_pow2k(Mat m, int k) { for (int i = 0; i < k; ++i) m = mat_mul(m, m); return m; }      // m^(2^k)
static Mat mat_pow(const Mat& m, unsigned e) { Mat r = mat_identity(), b = m; while (e) { if (e & 1) r = mat_mul(r, b); b = mat_mul(b, b); e >>= 1; } return r; }

__host__ __device__ inline State apply(const Mat& m, const State& s) {
    State r;
    for (int i = 0; i < 5; ++i) {
        uint64_t acc = 0;
        for (int k = 0; k < 5; ++k) acc += mulmod(m.a[i][k], s.z[k]);
        r.z[i] = (uint32_t)(acc % P);
    }
    return r;
}
__host__ __device__ inline uint32_t next_uint(State& s) {     // mrg_get_uint_orig
    const uint32_t n = (uint32_t)(((uint64_t)107374182u * s.z[0] + (uint64_t)104480u * s.z[4]) % P);
    s.z[4] = s.z[3]; s.z[3] = s.z[2]; s.z[2] = s.z[1]; s.z[1] = s.z[0]; s.z[0] = n;
    return n;
}
__host__ __device__ inline uint64_t bitreverse64(uint64_t x) {
#ifdef __CUDA_ARCH__
    return __brevll(x);
#else
    uint64_t r = 0;
    for (int i = 0; i < 64; ++i) r |= ((x >> i) & 1ull) << (63 - i);
    return r;
#endif
}
__host__ __device__ inline uint64_t scramble(uint64_t v, int lgN, uint64_t val0, uint64_t val1) {      // RefGen21.h:183-196
    v += val0 + val1;
    v *= (val0 | 0x4519840211493211ull);
    v = bitreverse64(v) >> (64 - lgN);
    v *= (val1 | 0x3050852102C843A5ull);
    v = bitreverse64(v) >> (64 - lgN);
    return v;
}
__host__ __device__ inline int quadrant(State& s) {            // generate_4way_bernoulli without noise, RefGen21.h:104-134
    uint32_t val = next_uint(s);
    while (val < 7295u) val = next_uint(s);                    // 0xFFFFFFFF % 10000: no modulo bias
    val %= 10000u;
    if (val < 1900u) return 1;
    val -= 1900u;
    if (val < 1900u) return 2;
    val -= 1900u;
    return val < 5700u ? 0 : 3;
}
__host__ __device__ inline void one_edge(State s, int lgN, uint64_t val0, uint64_t val1, uint64_t* src, uint64_t* tgt) {
    uint64_t nverts = 1ull << lgN, bs = 0, bt = 0;
    while (nverts > 1) {
        const int sq = quadrant(s);
        int so = sq / 2, to = sq % 2;
        if (bs == bt && so > to) { const int t = so; so = to; to = t; }      // clip-and-flip for the undirected graph
        nverts /= 2;
        bs += nverts * (uint64_t)so;
        bt += nverts * (uint64_t)to;
    }
    *src = scramble(bs, lgN, val0, val1);
    *tgt = scramble(bt, lgN, val0, val1);
}

struct Tables {
    Mat edge[4][256];            // A^(2^64 * 256^b * v): the jump for byte b of the edge index
    State seed;
    uint64_t val0, val1;
};

// make_mrg_seed (graph500-1.2/generator/utils.c:83-89)
static State make_seed(uint64_t u1, uint64_t u2) {
    State s;
    s.z[0] = (uint32_t)((u1 & 0x3FFFFFFF) + 1);
    s.z[1] = (uint32_t)(((u1 >> 30) & 0x3FFFFFFF) + 1);
    s.z[2] = (uint32_t)((u2 & 0x3FFFFFFF) + 1);
    s.z[3] = (uint32_t)(((u2 >> 30) & 0x3FFFFFFF) + 1);
    s.z[4] = (uint32_t)(((u2 >> 60) << 4) + (u1 >> 60) + 1);
    return s;
}
static void build_tables(uint64_t userseed1, uint64_t userseed2, Tables* t) {
    const Mat A = mat_step();
    Mat base = mat_pow2k(A, 64);                                  // A^(2^64)
    const Mat a64 = base;
    for (int b = 0; b < 4; ++b) {
        t->edge[b][0] = mat_identity();
        for (int v = 1; v < 256; ++v) t->edge[b][v] = mat_mul(t->edge[b][v - 1], base);
        base = mat_pow2k(base, 8);                                // ^256
    }
    t->seed = make_seed(userseed1, userseed2);
    // MakeScrambleValues (RefGen21.h:227-240): mrg_skip(&state, 50, 7, 0) = A^(2^128 * 50) A^(2^64 * 7), then four draws
    State s = apply(mat_pow(a64, 7), t->seed);
    s = apply(mat_pow(mat_pow2k(A, 128), 50), s);
    const uint64_t u0 = next_uint(s), u1 = next_uint(s), u2 = next_uint(s), u3 = next_uint(s);
    t->val0 = u0 * 0xFFFFFFFFull + u1;
    t->val1 = u2 * 0xFFFFFFFFull + u3;
}

}  // namespace g500
