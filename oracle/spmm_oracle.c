/* oracle/spmm_oracle.c - CPU restatement of the reference's sparse x dense-panel multiply.
 *
 * TEST INFRASTRUCTURE ONLY.  Imported/linked by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.  The product (combblas-spmm-test_b200/)
 * never calls into this file and has no CPU fallback.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against
 *   (i)  the reference's known answers in its own tree (torus G*G = 112 nnz,
 *        Applications/SpMMError.cpp:32-33,80; hep-th.mtx = 31502 nnz after expansion), and
 *   (ii) outputs of the unmodified reference itself (oracle/_ref/libcbref.so, built from
 *        /root/reference by oracle/Makefile) - live where /root/reference exists, and
 *        through the committed fixtures tests/golden/ (made by tests/golden/make_golden.py), and
 *   (iii) outputs of the unmodified reference run on 2x2 / 3x3 PROCESS grids (oracle/_ref/cbref_grid =
 *        the reference + the process-per-rank MPI stand-in oracle/mpi_multi): the emulated stage loop
 *        below and the owner rule against tests/golden/grid_ref.npz (tests/test_ref_grid.py).
 *
 * What is restated, with the reference lines each function follows:
 *   semiring functors ............ include/CombBLAS/Semirings.h:40-47 (inf_plus), :191-210
 *                                  (SelectMaxSRing<bool,T>), :212-232 (PlusTimesSRing), :235-255 (MinPlusSRing)
 *   column-by-column multiply .... include/CombBLAS/mtSpGEMM.h:292-441 (LocalHybridSpGEMM numeric phase,
 *                                  hash branch :395-423: for each nonzero B(kk,j) in ascending kk, for each
 *                                  A(i,kk): first touch stores the product, later ones do SR::add(product, acc))
 *   dense SpMV ................... include/CombBLAS/Friends.h:63-78 (dcsc_gespmv: y[row] = axpy(a, x[col], y[row]),
 *                                  y pre-filled with SR::id(), include/CombBLAS/ParFriends.h:1960-1963)
 *   2D ownership ................. include/CombBLAS/SpParMat.cpp:5066-5096 (Owner)
 *   SUMMA stage loop + merge ..... include/CombBLAS/ParFriends.h:1036-1083 (stage i: A tile (r,i) x B tile (i,c)),
 *                                  include/CombBLAS/MultiwayMerge.h:205-218 (equal (row,col) merged with SR::add)
 * The dense operand X (n x k, row-major, leading dimension ldx) plays B; the dense result Y
 * (m x k) holds SR::id() where the reference's sparse C has no entry.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include <float.h>
#include <stdbool.h>

enum { CB_F32 = 0, CB_F64 = 1, CB_I32 = 2, CB_I64 = 3, CB_U8 = 4, CB_PATTERN = 255 };
enum { CB_PLUS_TIMES = 0, CB_MIN_PLUS = 1, CB_MAX_SEL2ND = 2, CB_OR_AND = 3 };

/* ---- semiring functors (Semirings.h) ------------------------------------------------ */
/* PlusTimesSRing: id 0; add a+b; multiply (T)a*(T)b.  Integer + and * wrap (done in unsigned). */
#define PT_ID(T) ((T)0)
static inline float pt_add_f32(float a, float b) { return a + b; }
static inline double pt_add_f64(double a, double b) { return a + b; }
static inline int32_t pt_add_i32(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static inline int64_t pt_add_i64(int64_t a, int64_t b) { return (int64_t)((uint64_t)a + (uint64_t)b); }
static inline float pt_mul_f32(float a, float b) { return a * b; }
static inline double pt_mul_f64(double a, double b) { return a * b; }
static inline int32_t pt_mul_i32(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
static inline int64_t pt_mul_i64(int64_t a, int64_t b) { return (int64_t)((uint64_t)a * (uint64_t)b); }
/* MinPlusSRing: id numeric_limits<T>::max(); add min; multiply inf_plus (Semirings.h:40-47) */
static inline float mp_add_f32(float a, float b) { return b < a ? b : a; }   /* std::min(a,b) */
static inline double mp_add_f64(double a, double b) { return b < a ? b : a; }
static inline int32_t mp_add_i32(int32_t a, int32_t b) { return b < a ? b : a; }
static inline int64_t mp_add_i64(int64_t a, int64_t b) { return b < a ? b : a; }
static inline float mp_mul_f32(float a, float b) { return (a == FLT_MAX || b == FLT_MAX) ? FLT_MAX : a + b; }
static inline double mp_mul_f64(double a, double b) { return (a == DBL_MAX || b == DBL_MAX) ? DBL_MAX : a + b; }
static inline int32_t mp_mul_i32(int32_t a, int32_t b) { return (a == INT32_MAX || b == INT32_MAX) ? INT32_MAX : (int32_t)((uint32_t)a + (uint32_t)b); }
static inline int64_t mp_mul_i64(int64_t a, int64_t b) { return (a == INT64_MAX || b == INT64_MAX) ? INT64_MAX : (int64_t)((uint64_t)a + (uint64_t)b); }
/* SelectMaxSRing<bool,T>: id -1; add std::max(a,b) = (a<b)?b:a; multiply(bool,x) = x */
static inline float sm_add_f32(float a, float b) { return a < b ? b : a; }
static inline double sm_add_f64(double a, double b) { return a < b ? b : a; }
static inline int32_t sm_add_i32(int32_t a, int32_t b) { return a < b ? b : a; }
static inline int64_t sm_add_i64(int64_t a, int64_t b) { return a < b ? b : a; }

/* ---- one generic column-by-column kernel, instantiated per (semiring, A type, X type) ---
 * A is column compressed: nzc nonempty columns, ids jc[c] (jc==NULL: plain CSC, column c),
 * entries cp[c]..cp[c+1) with local row ir[p] and value numx[p].
 * `touched` marks first touch per (row) for the current output column: exactly the hash table's
 * "key not registered yet" branch.  ACCUM=1 merges into an existing Y the way MultiwayMerge does
 * for later SUMMA stages: Y = add(Y, partial).
 */
#define DEFINE_SPMM(NAME, TA, TX, TO, AVAL, MUL, ADD)                                                   \
    static void NAME(int64_t m, int64_t nzc, const int64_t* cp, const int64_t* jc, const int64_t* ir,    \
                     const void* numx_, int64_t k, const void* X_, int64_t ldx, void* Y_, int64_t ldy,   \
                     TO id, int accum) {                                                                  \
        const TA* numx = (const TA*)numx_; const TX* X = (const TX*)X_; TO* Y = (TO*)Y_;                  \
        (void)numx;                                                                                       \
        _Pragma("omp parallel")                                                                           \
        {                                                                                                 \
            TO* acc = (TO*)malloc(sizeof(TO) * (size_t)(m > 0 ? m : 1));                                  \
            unsigned char* touched = (unsigned char*)malloc((size_t)(m > 0 ? m : 1));                     \
            _Pragma("omp for schedule(dynamic,1)")                                                        \
            for (int64_t j = 0; j < k; ++j) {             /* nonempty columns of B: all k */            \
                memset(touched, 0, (size_t)m);                                                            \
                for (int64_t c = 0; c < nzc; ++c) {       /* ascending kk = jc[c] */                    \
                    const int64_t kk = jc ? jc[c] : c;                                                    \
                    const TX bval = X[kk * ldx + j];                                                      \
                    for (int64_t p = cp[c]; p < cp[c + 1]; ++p) {                                         \
                        const int64_t i = ir[p];                                                          \
                        const TO mrhs = MUL(AVAL, bval);                                                  \
                        if (touched[i]) acc[i] = ADD(mrhs, acc[i]);                                       \
                        else { touched[i] = 1; acc[i] = mrhs; }                                           \
                    }                                                                                     \
                }                                                                                         \
                for (int64_t i = 0; i < m; ++i) {                                                         \
                    if (!accum) Y[i * ldy + j] = touched[i] ? acc[i] : id;                                \
                    else if (touched[i]) Y[i * ldy + j] = ADD(Y[i * ldy + j], acc[i]);                    \
                }                                                                                         \
            }                                                                                             \
            free(acc); free(touched);                                                                     \
        }                                                                                                 \
    }

#define A_STORED numx[p]
#define SEL2ND(a, b) (b)
/* PlusTimes, stored A of the same type as X */
DEFINE_SPMM(spmm_pt_f32, float, float, float, A_STORED, pt_mul_f32, pt_add_f32)
DEFINE_SPMM(spmm_pt_f64, double, double, double, A_STORED, pt_mul_f64, pt_add_f64)
DEFINE_SPMM(spmm_pt_i32, int32_t, int32_t, int32_t, A_STORED, pt_mul_i32, pt_add_i32)
DEFINE_SPMM(spmm_pt_i64, int64_t, int64_t, int64_t, A_STORED, pt_mul_i64, pt_add_i64)
/* PlusTimes<bool,T>: multiply = static_cast<T>(a) * x; A stored as one byte per entry */
#define BOOL_AS(T) ((T)(numx[p] != 0))
DEFINE_SPMM(spmm_pt_b_f32, uint8_t, float, float, BOOL_AS(float), pt_mul_f32, pt_add_f32)
DEFINE_SPMM(spmm_pt_b_f64, uint8_t, double, double, BOOL_AS(double), pt_mul_f64, pt_add_f64)
DEFINE_SPMM(spmm_pt_b_i32, uint8_t, int32_t, int32_t, BOOL_AS(int32_t), pt_mul_i32, pt_add_i32)
DEFINE_SPMM(spmm_pt_b_i64, uint8_t, int64_t, int64_t, BOOL_AS(int64_t), pt_mul_i64, pt_add_i64)
/* pattern A (every stored entry is true): static_cast<T>(true) * x */
DEFINE_SPMM(spmm_pt_p_f32, uint8_t, float, float, 1.0f, pt_mul_f32, pt_add_f32)
DEFINE_SPMM(spmm_pt_p_f64, uint8_t, double, double, 1.0, pt_mul_f64, pt_add_f64)
DEFINE_SPMM(spmm_pt_p_i32, uint8_t, int32_t, int32_t, 1, pt_mul_i32, pt_add_i32)
DEFINE_SPMM(spmm_pt_p_i64, uint8_t, int64_t, int64_t, 1, pt_mul_i64, pt_add_i64)
/* PlusTimes<bool,bool> (promote.h:78): bool+bool -> OR, bool*bool -> AND; bytes 0/1 */
static inline uint8_t oa_add(uint8_t a, uint8_t b) { return (uint8_t)((a + b) != 0); }
static inline uint8_t oa_mul(uint8_t a, uint8_t b) { return (uint8_t)((a != 0) * (b != 0)); }
DEFINE_SPMM(spmm_oa_b, uint8_t, uint8_t, uint8_t, numx[p], oa_mul, oa_add)
DEFINE_SPMM(spmm_oa_p, uint8_t, uint8_t, uint8_t, 1, oa_mul, oa_add)
/* MinPlus */
DEFINE_SPMM(spmm_mp_f32, float, float, float, A_STORED, mp_mul_f32, mp_add_f32)
DEFINE_SPMM(spmm_mp_f64, double, double, double, A_STORED, mp_mul_f64, mp_add_f64)
DEFINE_SPMM(spmm_mp_i32, int32_t, int32_t, int32_t, A_STORED, mp_mul_i32, mp_add_i32)
DEFINE_SPMM(spmm_mp_i64, int64_t, int64_t, int64_t, A_STORED, mp_mul_i64, mp_add_i64)
/* SelectMax<bool,T>: multiply ignores A's value altogether (Semirings.h:202-205) */
DEFINE_SPMM(spmm_sm_f32, uint8_t, float, float, 0, SEL2ND, sm_add_f32)
DEFINE_SPMM(spmm_sm_f64, uint8_t, double, double, 0, SEL2ND, sm_add_f64)
DEFINE_SPMM(spmm_sm_i32, uint8_t, int32_t, int32_t, 0, SEL2ND, sm_add_i32)
DEFINE_SPMM(spmm_sm_i64, uint8_t, int64_t, int64_t, 0, SEL2ND, sm_add_i64)

/* Dispatch.  a_dtype: dtype of stored A values, or CB_PATTERN (no value array; every entry true/1).
 * x_dtype = dtype of X = dtype of Y (promote_trait<bool,T> = T, promote_trait<T,T> = T, promote.h:44-60).
 * Returns 0, or 1 if the combination does not exist. */
int oracle_spmm_colcompressed(int semiring, int a_dtype, int x_dtype, int64_t m, int64_t nzc, const int64_t* cp,
                              const int64_t* jc, const int64_t* ir, const void* numx, int64_t k, const void* X,
                              int64_t ldx, void* Y, int64_t ldy, int accum) {
#define RUN(fn, T, idv) { fn(m, nzc, cp, jc, ir, numx, k, X, ldx, Y, ldy, (T)(idv), accum); return 0; }
    const int pat = (a_dtype == CB_PATTERN), ab = (a_dtype == CB_U8);
    switch (semiring) {
    case CB_PLUS_TIMES:
        if (x_dtype == CB_F32) { if (pat) RUN(spmm_pt_p_f32, float, 0) if (ab) RUN(spmm_pt_b_f32, float, 0) if (a_dtype == CB_F32) RUN(spmm_pt_f32, float, 0) }
        if (x_dtype == CB_F64) { if (pat) RUN(spmm_pt_p_f64, double, 0) if (ab) RUN(spmm_pt_b_f64, double, 0) if (a_dtype == CB_F64) RUN(spmm_pt_f64, double, 0) }
        if (x_dtype == CB_I32) { if (pat) RUN(spmm_pt_p_i32, int32_t, 0) if (ab) RUN(spmm_pt_b_i32, int32_t, 0) if (a_dtype == CB_I32) RUN(spmm_pt_i32, int32_t, 0) }
        if (x_dtype == CB_I64) { if (pat) RUN(spmm_pt_p_i64, int64_t, 0) if (ab) RUN(spmm_pt_b_i64, int64_t, 0) if (a_dtype == CB_I64) RUN(spmm_pt_i64, int64_t, 0) }
        if (x_dtype == CB_U8) { if (pat) RUN(spmm_oa_p, uint8_t, 0) if (ab) RUN(spmm_oa_b, uint8_t, 0) }
        return 1;
    case CB_OR_AND:
        if (x_dtype == CB_U8) { if (pat) RUN(spmm_oa_p, uint8_t, 0) if (ab) RUN(spmm_oa_b, uint8_t, 0) }
        return 1;
    case CB_MIN_PLUS:
        if (x_dtype == CB_F32 && a_dtype == CB_F32) RUN(spmm_mp_f32, float, FLT_MAX)
        if (x_dtype == CB_F64 && a_dtype == CB_F64) RUN(spmm_mp_f64, double, DBL_MAX)
        if (x_dtype == CB_I32 && a_dtype == CB_I32) RUN(spmm_mp_i32, int32_t, INT32_MAX)
        if (x_dtype == CB_I64 && a_dtype == CB_I64) RUN(spmm_mp_i64, int64_t, INT64_MAX)
        return 1;
    case CB_MAX_SEL2ND:
        if (!(pat || ab)) return 1;
        if (x_dtype == CB_F32) RUN(spmm_sm_f32, float, -1)
        if (x_dtype == CB_F64) RUN(spmm_sm_f64, double, -1)
        if (x_dtype == CB_I32) RUN(spmm_sm_i32, int32_t, -1)
        if (x_dtype == CB_I64) RUN(spmm_sm_i64, int64_t, -1)
        return 1;
    }
    return 1;
#undef RUN
}

/* ---- dense SpMV restated (Friends.h:63-78): used as the k=1 cross-check ------------------ */
void oracle_spmv_pt_f64(int64_t m, int64_t nzc, const int64_t* cp, const int64_t* jc, const int64_t* ir,
                        const double* numx, const double* x, double* y) {
    for (int64_t i = 0; i < m; ++i) y[i] = 0.0;                 /* fill_n(localy, ysize, SR::id()) */
    for (int64_t c = 0; c < nzc; ++c) {
        const int64_t col = jc ? jc[c] : c;
        for (int64_t p = cp[c]; p < cp[c + 1]; ++p) y[ir[p]] += numx[p] * x[col];   /* PlusTimes axpy: y += a*x */
    }
}

/* ---- 2D block ownership (SpParMat.cpp:5066-5096) ---------------------------------------- */
int oracle_owner(int64_t total_m, int64_t total_n, int pr, int pc, int64_t grow, int64_t gcol,
                 int64_t* lrow, int64_t* lcol) {
    const int64_t m_perproc = total_m / pr, n_perproc = total_n / pc;
    int own_r = pr - 1, own_c = pc - 1;
    if (m_perproc != 0) { int64_t q = grow / m_perproc; own_r = (int)(q < pr - 1 ? q : pr - 1); }
    if (n_perproc != 0) { int64_t q = gcol / n_perproc; own_c = (int)(q < pc - 1 ? q : pc - 1); }
    *lrow = grow - own_r * m_perproc;
    *lcol = gcol - own_c * n_perproc;
    return own_r * pc + own_c;                                  /* CommGrid::GetRank(row,col), CommGrid.h:106 */
}

/* first global index and length of block b out of nb for a dimension of size total (same floor rule) */
void oracle_block_range(int64_t total, int nb, int b, int64_t* start, int64_t* len) {
    const int64_t per = total / nb;
    *start = (int64_t)b * per;
    *len = (b == nb - 1) ? total - *start : per;
}

/* ---- whole 2D SUMMA emulated on one core set ----------------------------------------------
 * A (m x n, COO global indices, already deduplicated), X (n x k row-major), grid pr x pc.
 * Inner dimension n is cut into S = lcm(pr,pc) chunks (S = pc = pr on the reference's square grids,
 * ParFriends.h:1036 "stages"); stage s multiplies A's column chunk s by X's row chunk s on every
 * rank and the partials are merged with SR::add in stage order (MultiwayMerge.h:205-218).
 * Y tile (r,c) = rows of block-row r, columns of block c of k.  The result is written to the
 * global Y (m x k).  For integer/boolean semirings this equals the 1-rank answer exactly.
 */
static int64_t gcd64(int64_t a, int64_t b) { while (b) { int64_t t = a % b; a = b; b = t; } return a; }

int oracle_spmm_summa(int semiring, int a_dtype, int x_dtype, size_t a_size, size_t x_size, int pr, int pc,
                      int64_t m, int64_t n, int64_t nnz, const int64_t* I, const int64_t* J, const void* V,
                      int64_t k, const void* X, void* Y) {
    const int S = (int)((int64_t)pr * pc / gcd64(pr, pc));
    int rc = 0;
    for (int r = 0; r < pr && !rc; ++r) {
        int64_t r0, rl; oracle_block_range(m, pr, r, &r0, &rl);
        for (int c = 0; c < pc && !rc; ++c) {
            int64_t c0, cl; oracle_block_range(k, pc, c, &c0, &cl);
            if (cl == 0 || rl == 0) continue;
            for (int s = 0; s < S && !rc; ++s) {
                int64_t s0, sl; oracle_block_range(n, S, s, &s0, &sl);
                /* collect the entries of A in rows [r0,r0+rl) x cols [s0,s0+sl), column-major */
                int64_t cnt = 0;
                for (int64_t p = 0; p < nnz; ++p) if (I[p] >= r0 && I[p] < r0 + rl && J[p] >= s0 && J[p] < s0 + sl) ++cnt;
                int64_t* cp = (int64_t*)calloc((size_t)sl + 1, sizeof(int64_t));
                int64_t* ir = (int64_t*)malloc(sizeof(int64_t) * (size_t)(cnt ? cnt : 1));
                char* nv = (char*)malloc(a_size * (size_t)(cnt ? cnt : 1) + 1);
                for (int64_t p = 0; p < nnz; ++p) if (I[p] >= r0 && I[p] < r0 + rl && J[p] >= s0 && J[p] < s0 + sl) cp[J[p] - s0 + 1]++;
                for (int64_t q = 0; q < sl; ++q) cp[q + 1] += cp[q];
                int64_t* fill = (int64_t*)malloc(sizeof(int64_t) * (size_t)(sl + 1));
                memcpy(fill, cp, sizeof(int64_t) * (size_t)(sl + 1));
                /* stable within a column only if the input is row-sorted inside columns; sort rows per column after */
                for (int64_t p = 0; p < nnz; ++p) if (I[p] >= r0 && I[p] < r0 + rl && J[p] >= s0 && J[p] < s0 + sl) {
                    int64_t q = fill[J[p] - s0]++;
                    ir[q] = I[p] - r0;
                    if (V && a_size) memcpy(nv + (size_t)q * a_size, (const char*)V + (size_t)p * a_size, a_size);
                }
                free(fill);
                rc = oracle_spmm_colcompressed(semiring, a_dtype, x_dtype, rl, sl, cp, NULL, ir, (V && a_size) ? nv : NULL, cl,
                                               (const char*)X + ((size_t)s0 * (size_t)k + (size_t)c0) * x_size, k,
                                               (char*)Y + ((size_t)r0 * (size_t)k + (size_t)c0) * x_size, k, s > 0);
                free(cp); free(ir); free(nv);
            }
        }
    }
    return rc;
}
