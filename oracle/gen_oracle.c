/* oracle/gen_oracle.c - the synthetic-input recipes of oracle/oracle.py (rmat_edges / scramble / dedup / hash_values)
 * restated in C with OpenMP, so the CPU legs of bench.py can build BASELINE.json's full-size operands (scale 24: 268 M
 * candidate edges) in seconds instead of minutes.
 *
 * TEST INFRASTRUCTURE ONLY, like the rest of oracle/: used by bench.py's cpu_baseline / --impl reference legs and by tests/.
 * The recipe follows ReleaseTests/GenWriteMatrix.cpp:96-131 of the reference (Kronecker edges with an initiator, vertex
 * scramble, self loops removed, optional A += A^T, duplicates merged); the bit stream is this repo's own counter hash, the
 * same integers as csrc/cb_gen.cu and oracle.py produce (tests/test_oracle.py compares all three).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

static inline uint64_t bitrev(uint64_t v, int bits) {
    uint64_t r = 0;
    for (int b = 0; b < bits; ++b) r |= ((v >> b) & 1ull) << (bits - 1 - b);
    return r;
}

/* oracle.py scramble(): odd multiply, add, bit reversal, odd multiply, add, all mod 2^scale */
static inline uint64_t scramble(uint64_t v, int scale, uint64_t s1, uint64_t s2) {
    const uint64_t mask = (1ull << scale) - 1ull;
    v = (v * (s1 | 1ull) + (s1 >> 32)) & mask;
    v = bitrev(v, scale);
    v = (v * (s2 | 1ull) + (s2 >> 32)) & mask;
    return v;
}

/* LSD radix sort of 64-bit keys on their low `bits` bits, 16 bits per pass, per-thread histograms; result in keys */
static int radix_sort_u64(uint64_t* keys, uint64_t* tmp, int64_t n, int bits) {
    const int nt = omp_get_max_threads();
    const int RB = 16, NB = 1 << RB;
    int64_t* hist = (int64_t*)malloc((size_t)nt * NB * sizeof(int64_t));
    if (!hist) return -1;
    uint64_t *src = keys, *dst = tmp;
    for (int shift = 0; shift < bits; shift += RB) {
        memset(hist, 0, (size_t)nt * NB * sizeof(int64_t));
#pragma omp parallel num_threads(nt)
        {
            const int t = omp_get_thread_num();
            const int64_t lo = n * t / nt, hi = n * (t + 1) / nt;
            int64_t* h = hist + (size_t)t * NB;
            for (int64_t i = lo; i < hi; ++i) h[(src[i] >> shift) & (NB - 1)]++;
        }
        int64_t run = 0;
        for (int d = 0; d < NB; ++d)
            for (int t = 0; t < nt; ++t) {
                int64_t c = hist[(size_t)t * NB + d];
                hist[(size_t)t * NB + d] = run;
                run += c;
            }
#pragma omp parallel num_threads(nt)
        {
            const int t = omp_get_thread_num();
            const int64_t lo = n * t / nt, hi = n * (t + 1) / nt;
            int64_t* h = hist + (size_t)t * NB;
            for (int64_t i = lo; i < hi; ++i) dst[h[(src[i] >> shift) & (NB - 1)]++] = src[i];
        }
        uint64_t* s = src; src = dst; dst = s;
    }
    if (src != keys) memcpy(keys, src, (size_t)n * sizeof(uint64_t));
    free(hist);
    return 0;
}

void oracle_free(void* p) { free(p); }
/* launchers such as torch.distributed.run export OMP_NUM_THREADS=1: the CPU legs of bench.py ask for the host's cores explicitly */
void oracle_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }

/* oracle.py rmat_matrix(): n = 2^scale; edges -> drop self loops -> optional (A + A^T) -> merge duplicates.
 * thr[3] = the 16-bit quadrant thresholds round(a*65536), round((a+b)*65536), round((a+b+c)*65536) (computed by the caller
 * so both languages round the same way).  Output: *nnz unique entries, sorted row-major (col_major = 0, the order of
 * oracle.py's dedup) or column-major (col_major = 1, the order SpDCCols' tuple constructor wants, SpDCCols.cpp:186-195).
 * The arrays are malloc'ed here and released with oracle_free.  Returns 0, or -1 when out of memory. */
int oracle_gen_matrix(int scale, int edgefactor, uint64_t seed, const uint32_t thr[3], int symmetric, int remove_loops,
                      int col_major, int64_t* nnz, int64_t** I_out, int64_t** J_out) {
    const int64_t nedges = (int64_t)edgefactor << scale;
    const int64_t cap = symmetric ? 2 * nedges : nedges;
    uint64_t* keys = (uint64_t*)malloc((size_t)cap * sizeof(uint64_t));
    uint64_t* tmp = (uint64_t*)malloc((size_t)cap * sizeof(uint64_t));
    if (!keys || !tmp) { free(keys); free(tmp); return -1; }
    const uint64_t s1 = splitmix64(seed ^ 0x5CA1AB1Eull), s2 = splitmix64(s1);
    const uint64_t base = seed << 40;
    const uint64_t t1 = thr[0], t2 = thr[1], t3 = thr[2];
    const uint64_t DROP = ~0ull;
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < nedges; ++e) {
        uint64_t i = 0, j = 0, h = 0;
        for (int lvl = 0; lvl < scale; ++lvl) {
            if ((lvl & 3) == 0) h = splitmix64(base ^ ((uint64_t)e * 8ull + (uint64_t)(lvl >> 2)));
            const uint64_t u = (h >> (16 * (lvl & 3))) & 0xFFFFull;
            const uint64_t ib = u >= t2;                                  /* quadrants c, d -> row bit */
            const uint64_t jb = ((u >= t1) && (u < t2)) || (u >= t3);     /* quadrants b, d -> column bit */
            i = (i << 1) | ib;
            j = (j << 1) | jb;
        }
        i = scramble(i, scale, s1, s2);
        j = scramble(j, scale, s1, s2);
        const int loop = remove_loops && i == j;
        const uint64_t k1 = col_major ? ((j << scale) | i) : ((i << scale) | j);
        const uint64_t k2 = col_major ? ((i << scale) | j) : ((j << scale) | i);
        keys[e] = loop ? DROP : k1;
        if (symmetric) keys[nedges + e] = loop ? DROP : k2;
    }
    /* dropped entries carry all ones: sort on 2*scale+1 bits would not place them last, so compact them away first */
    int64_t live = 0;
    for (int64_t q = 0; q < cap; ++q)
        if (keys[q] != DROP) keys[live++] = keys[q];
    if (radix_sort_u64(keys, tmp, live, 2 * scale) != 0) { free(keys); free(tmp); return -1; }
    int64_t uniq = 0;
    for (int64_t q = 0; q < live; ++q)
        if (q == 0 || keys[q] != keys[q - 1]) keys[uniq++] = keys[q];
    free(tmp);
    int64_t* I = (int64_t*)malloc((size_t)(uniq ? uniq : 1) * sizeof(int64_t));
    int64_t* J = (int64_t*)malloc((size_t)(uniq ? uniq : 1) * sizeof(int64_t));
    if (!I || !J) { free(keys); free(I); free(J); return -1; }
    const uint64_t mask = (1ull << scale) - 1ull;
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < uniq; ++q) {
        const uint64_t hi = keys[q] >> scale, lo = keys[q] & mask;
        I[q] = (int64_t)(col_major ? lo : hi);
        J[q] = (int64_t)(col_major ? hi : lo);
    }
    free(keys);
    *nnz = uniq;
    *I_out = I;
    *J_out = J;
    return 0;
}

/* oracle.py hash_values(): operand values from the counter hash.  dtype: 0 f32, 1 f64, 2 i32, 3 i64, 4 u8;
 * kind 1 = MinPlus panel (about 1 % of the integer entries at numeric max). */
static inline void hash_store(uint64_t idx, uint64_t seed, int dtype, int kind, void* out, int64_t q) {
    const uint64_t h = splitmix64(seed * 0x100000001B3ull ^ idx);
    switch (dtype) {
        case 0: ((float*)out)[q] = (float)((2.0 * (double)(h >> 41) + 1.0) * 5.9604644775390625e-08); break;       /* 2^-24 */
        case 1: ((double*)out)[q] = (2.0 * (double)(h >> 12) + 1.0) * 1.1102230246251565e-16; break;               /* 2^-53 */
        case 2: ((int32_t*)out)[q] = (kind == 1 && (h & 0xFF) < 3) ? INT32_MAX : (int32_t)(1 + (h >> 8) % 100); break;
        case 3: ((int64_t*)out)[q] = (kind == 1 && (h & 0xFF) < 3) ? INT64_MAX : (int64_t)(1 + (h >> 8) % 100); break;
        default: ((uint8_t*)out)[q] = (uint8_t)(h >> 63); break;
    }
}

/* oracle.py matrix_values(): value of entry (I[q], J[q]) = hash(I*n + J) */
void oracle_matrix_values(const int64_t* I, const int64_t* J, int64_t nnz, int64_t n, uint64_t seed, int dtype, void* out) {
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < nnz; ++q) hash_store((uint64_t)I[q] * (uint64_t)n + (uint64_t)J[q], seed, dtype, 0, out, q);
}

/* columns [c0, c0+kc) of oracle.py dense_operand(n, k): X[i, j] = hash(i*k + j), packed n x kc */
void oracle_dense_columns(int64_t n, int64_t k, int64_t c0, int64_t kc, uint64_t seed, int dtype, int kind, void* out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = 0; j < kc; ++j) hash_store((uint64_t)i * (uint64_t)k + (uint64_t)(c0 + j), seed, dtype, kind, out, i * kc + j);
}
