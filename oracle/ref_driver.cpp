// oracle/_ref driver: runs the UNMODIFIED reference multiply (headers + src/*.cpp compiled
// where they lie under /root/reference, against oracle/mpi_stub/mpi.h, 1 rank x OpenMP threads).
// TEST INFRASTRUCTURE ONLY - the product never links or loads this.
//
// What is executed is the reference's own code:
//   * Mult_AnXBn_Synch<SR,NUO,UDERO>   include/CombBLAS/ParFriends.h:1004-1108
//       -> LocalHybridSpGEMM           include/CombBLAS/mtSpGEMM.h:213-460
//       -> MultiwayMerge               include/CombBLAS/MultiwayMerge.h:411-526
//   * SpMV<SR>(SpParMat, FullyDistVec) include/CombBLAS/ParFriends.h:1924-1996
//   * SpParMat::ParallelReadMM         include/CombBLAS/SpParMat.cpp:3978-4115
// The dense operand X (n x k) is handed to Mult_AnXBn_Synch as a fully populated
// SpDCCols; C's DCSC is scattered into Y, absent entries = SR::id().
#include <mpi.h>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <tuple>
#include <algorithm>
#include <memory>
#include <omp.h>
#include "CombBLAS/CombBLAS.h"

using namespace combblas;

// globals the reference expects the application to define (CombBLAS.h:76-102)
int cblas_splits = 1;
double cblas_alltoalltime, cblas_allgathertime, cblas_mergeconttime, cblas_transvectime, cblas_localspmvtime;

namespace {

template <class T> struct ident { typedef T type; };

std::shared_ptr<CommGrid> grid1() {
    static std::shared_ptr<CommGrid> g;
    if (!g) g.reset(new CommGrid(MPI_COMM_WORLD, 1, 1));
    return g;
}

template <class NT>
SpDCCols<int64_t, NT>* tile_from_coo(int64_t m, int64_t n, int64_t nnz, const int64_t* I, const int64_t* J, const void* Vv, bool colmajor_sorted = false) {
    typedef std::tuple<int64_t, int64_t, NT> Tup;
    const NT* V = static_cast<const NT*>(Vv);
    Tup* t = new Tup[nnz > 0 ? nnz : 1];
#pragma omp parallel for
    for (int64_t p = 0; p < nnz; ++p) t[p] = Tup(I[p], J[p], V ? V[p] : NT(1));
    // column-major order, as the SpDCCols tuple-array ctor requires (SpDCCols.cpp:186-195)
    if (!colmajor_sorted)
    std::sort(t, t + nnz, [](const Tup& a, const Tup& b) {
        return std::get<1>(a) != std::get<1>(b) ? std::get<1>(a) < std::get<1>(b) : std::get<0>(a) < std::get<0>(b);
    });
    SpDCCols<int64_t, NT>* d = new SpDCCols<int64_t, NT>(m, n, nnz, t, false);
    delete[] t;
    return d;
}

template <class NT>
SpDCCols<int64_t, NT>* dense_as_tile(int64_t n, int64_t k, const NT* X /* row-major n x k */) {
    typedef std::tuple<int64_t, int64_t, NT> Tup;
    Tup* t = new Tup[n * k];
#pragma omp parallel for
    for (int64_t j = 0; j < k; ++j)
        for (int64_t i = 0; i < n; ++i) t[j * n + i] = Tup(i, j, X[i * k + j]);
    SpDCCols<int64_t, NT>* d = new SpDCCols<int64_t, NT>(n, k, n * k, t, false);
    delete[] t;
    return d;
}

// Y = A (x).(+) X through the reference SpGEMM, in column panels of `panel` to bound memory.
template <class SR, class NA, class NX>
int run_synch(int64_t m, int64_t n, int64_t nnz, const int64_t* I, const int64_t* J, const void* V,
              int64_t k, const void* Xv, void* Yv, int64_t panel, double* seconds) {
    typedef typename SR::T_promote NO;
    typedef SpDCCols<int64_t, NA> DA;
    typedef SpDCCols<int64_t, NX> DX;
    typedef SpDCCols<int64_t, NO> DO;
    const NX* X = static_cast<const NX*>(Xv);
    NO* Y = static_cast<NO*>(Yv);
    SpParMat<int64_t, NA, DA> A(tile_from_coo<NA>(m, n, nnz, I, J, V), grid1());
    const NO id = SR::id();
    for (int64_t q = 0; q < m * k; ++q) Y[q] = id;
    if (panel <= 0 || panel > k) panel = k;
    double total = 0;
    std::unique_ptr<NX[]> xp(new NX[n * panel > 0 ? n * panel : 1]);
    for (int64_t c0 = 0; c0 < k; c0 += panel) {
        int64_t kc = std::min(panel, k - c0);
        for (int64_t i = 0; i < n; ++i)
            for (int64_t j = 0; j < kc; ++j) xp[i * kc + j] = X[i * k + c0 + j];
        SpParMat<int64_t, NX, DX> Xs(dense_as_tile<NX>(n, kc, xp.get()), grid1());
        double t0 = MPI_Wtime();
        SpParMat<int64_t, NO, DO> C = Mult_AnXBn_Synch<SR, NO, DO>(A, Xs);
        total += MPI_Wtime() - t0;
        Dcsc<int64_t, NO>* d = C.seq().GetDCSC();
        if (d) {
            for (int64_t c = 0; c < d->nzc; ++c) {
                int64_t col = d->jc[c];
                for (int64_t p = d->cp[c]; p < d->cp[c + 1]; ++p) Y[d->ir[p] * k + c0 + col] = d->numx[p];
            }
        }
    }
    if (seconds) *seconds = total;
    return 0;
}

// Second, independent reference path: k dense SpMVs (ParFriends.h:1924-1996).
template <class SR, class NA, class NX>
int run_spmv(int64_t m, int64_t n, int64_t nnz, const int64_t* I, const int64_t* J, const void* V,
             int64_t k, const void* Xv, void* Yv, double* seconds) {
    typedef typename SR::T_promote NO;
    typedef SpDCCols<int64_t, NA> DA;
    const NX* X = static_cast<const NX*>(Xv);
    NO* Y = static_cast<NO*>(Yv);
    SpParMat<int64_t, NA, DA> A(tile_from_coo<NA>(m, n, nnz, I, J, V), grid1());
    double total = 0;
    for (int64_t j = 0; j < k; ++j) {
        FullyDistVec<int64_t, NX> x(grid1(), n, NX());
        for (int64_t i = 0; i < n; ++i) x.SetElement(i, X[i * k + j]);
        double t0 = MPI_Wtime();
        FullyDistVec<int64_t, NO> y = SpMV<SR>(A, x);
        total += MPI_Wtime() - t0;
        for (int64_t i = 0; i < m; ++i) Y[i * k + j] = y.GetElement(i);
    }
    if (seconds) *seconds = total;
    return 0;
}

// FullyDistVec<IT,bool> does not exist in the reference (FullyDistSpVec.h:73 disables it)
template <class SR>
int run_spmv_bool(double*) { fprintf(stderr, "cbref_spmm: SpMV path has no bool vectors in the reference\n"); return 3; }
template <>
int run_spmv<PlusTimesSRing<bool, bool>, bool, bool>(int64_t, int64_t, int64_t, const int64_t*, const int64_t*, const void*,
                                                     int64_t, const void*, void*, double* s) { return run_spmv_bool<void>(s); }

template <class SR, class NA, class NX>
int run(int via, int64_t m, int64_t n, int64_t nnz, const int64_t* I, const int64_t* J, const void* V,
        int64_t k, const void* X, void* Y, int64_t panel, double* seconds) {
    if (via == 1) return run_spmv<SR, NA, NX>(m, n, nnz, I, J, V, k, X, Y, seconds);
    return run_synch<SR, NA, NX>(m, n, nnz, I, J, V, k, X, Y, panel, seconds);
}

// A resident reference matrix for repeated multiplies (bench.py's CPU legs at BASELINE.json's full sizes: the SpParMat is
// built once, every timed step is one Mult_AnXBn_Synch with a fresh panel).
struct RefMatBase {
    virtual ~RefMatBase() {}
    virtual int mult(int64_t k, const void* X, void* Y, double* seconds, int64_t* nnzC) = 0;
};
template <class SR, class NA, class NX>
struct RefMat : RefMatBase {
    typedef typename SR::T_promote NO;
    typedef SpDCCols<int64_t, NA> DA;
    typedef SpDCCols<int64_t, NX> DX;
    typedef SpDCCols<int64_t, NO> DO;
    int64_t m, n;
    SpParMat<int64_t, NA, DA> A;
    RefMat(int64_t m_, int64_t n_, int64_t nnz, const int64_t* I, const int64_t* J, const void* V, bool sorted)
        : m(m_), n(n_), A(tile_from_coo<NA>(m_, n_, nnz, I, J, V, sorted), grid1()) {}
    int mult(int64_t k, const void* Xv, void* Yv, double* seconds, int64_t* nnzC) override {
        SpParMat<int64_t, NX, DX> Xs(dense_as_tile<NX>(n, k, static_cast<const NX*>(Xv)), grid1());
        double t0 = MPI_Wtime();
        SpParMat<int64_t, NO, DO> C = Mult_AnXBn_Synch<SR, NO, DO>(A, Xs);
        if (seconds) *seconds = MPI_Wtime() - t0;
        if (nnzC) *nnzC = C.getnnz();
        if (Yv) {
            NO* Y = static_cast<NO*>(Yv);
            const NO id = SR::id();
            for (int64_t q = 0; q < m * k; ++q) Y[q] = id;
            Dcsc<int64_t, NO>* d = C.seq().GetDCSC();
            if (d)
                for (int64_t c = 0; c < d->nzc; ++c)
                    for (int64_t p = d->cp[c]; p < d->cp[c + 1]; ++p) Y[d->ir[p] * k + d->jc[c]] = d->numx[p];
        }
        return 0;
    }
};

}  // namespace

extern "C" {

// key as in cbref_spmm; colmajor_sorted != 0 promises (I, J) sorted by column then row (skips the driver's own sort)
void* cbref_matrix_create(const char* key, int64_t m, int64_t n, int64_t nnz, const int64_t* I, const int64_t* J, const void* V, int colmajor_sorted) {
    std::string s(key);
#define CASE(name, SR, NA, NX) if (s == name) return new RefMat<SR<NA, NX>, NA, NX>(m, n, nnz, I, J, V, colmajor_sorted != 0);
    CASE("plus_times:f64:f64", PlusTimesSRing, double, double)
    CASE("plus_times:f32:f32", PlusTimesSRing, float, float)
    CASE("plus_times:bool:i32", PlusTimesSRing, bool, int32_t)
    CASE("plus_times:bool:bool", PlusTimesSRing, bool, bool)
    CASE("min_plus:i32:i32", MinPlusSRing, int32_t, int32_t)
    CASE("select_max:bool:i32", SelectMaxSRing, bool, int32_t)
#undef CASE
    fprintf(stderr, "cbref_matrix_create: unknown key %s\n", key);
    return nullptr;
}
// one Mult_AnXBn_Synch of the resident matrix with the n x k row-major panel X; Y (m x k, may be NULL) receives the product
int cbref_matrix_mult(void* h, int64_t k, const void* X, void* Y, double* seconds, int64_t* nnzC) {
    return h ? static_cast<RefMatBase*>(h)->mult(k, X, Y, seconds, nnzC) : 1;
}
void cbref_matrix_free(void* h) { delete static_cast<RefMatBase*>(h); }

// The matrix ReleaseTests/GenWriteMatrix.cpp writes, built by the reference's own classes with the calls of its lines 96-124:
// packed Graph500 edges, SpParMat(DEL, false) (keeps loops, sums duplicates: values are multiplicities), RemoveLoops, and for
// symmetric != 0 the program's own Symmetricize (A += A^T).  Call with I == NULL for the size.
int cbref_genwrite(int scale, int edgefactor, int symmetric, int64_t* nnz, int64_t* I, int64_t* J, int32_t* V) {
    typedef SpParMat<int64_t, int, SpDCCols<int32_t, int>> PSpMat;
    double initiator[4] = {.57, .19, .19, .05};
    DistEdgeList<int64_t>* DEL = new DistEdgeList<int64_t>();
    DEL->GenGraph500Data(initiator, scale, edgefactor, true, true);
    PSpMat G(*DEL, false);
    delete DEL;
    G.RemoveLoops();
    if (symmetric) { PSpMat GT = G; GT.Transpose(); G += GT; }
    *nnz = G.getnnz();
    if (I) {
        Dcsc<int32_t, int>* d = G.seq().GetDCSC();
        int64_t q = 0;
        if (d)
            for (int32_t c = 0; c < d->nzc; ++c)
                for (int32_t p = d->cp[c]; p < d->cp[c + 1]; ++p, ++q) { I[q] = d->ir[p]; J[q] = d->jc[c]; V[q] = d->numx[p]; }
    }
    return 0;
}

// Edges [first, first + count) of the reference's own packed Graph500 stream (RefGen21::generate_kronecker_range,
// include/CombBLAS/RefGen21.h:246-262, seeded as make_graph does under -DDETERMINISTIC: make_mrg_seed(0, 0)), for the goldens the
// device generator csrc/cb_gen.cu (cb_gen_graph500_edges) is pinned to.
int cbref_graph500_edges(int log_numverts, int64_t first, int64_t count, int64_t* src, int64_t* dst) {
    uint_fast32_t seed[5];
    make_mrg_seed(0, 0, seed);
    std::vector<packed_edge> e((size_t)(count > 0 ? count : 1));
    RefGen21::generate_kronecker_range(seed, log_numverts, first, first + count, e.data());
    for (int64_t q = 0; q < count; ++q) { src[q] = get_v0_from_edge(&e[(size_t)q]); dst[q] = get_v1_from_edge(&e[(size_t)q]); }
    return 0;
}

int cbref_num_threads() { return omp_get_max_threads(); }
void cbref_set_num_threads(int t) { omp_set_num_threads(t); }

// semiring/dtype keys: "<semiring>:<A dtype>:<X dtype>"; A dtype "bool" with V==NULL means pattern (all true).
// via: 0 = Mult_AnXBn_Synch, 1 = k x SpMV.  X row-major n x k, Y row-major m x k of T_promote.
int cbref_spmm(const char* key, int via, int64_t m, int64_t n, int64_t nnz, const int64_t* I, const int64_t* J,
               const void* V, int64_t k, const void* X, void* Y, int64_t panel, double* seconds) {
    std::string s(key);
#define CASE(name, SR, NA, NX) \
    if (s == name) return run<SR<NA, NX>, NA, NX>(via, m, n, nnz, I, J, V, k, X, Y, panel, seconds);
    CASE("plus_times:f64:f64", PlusTimesSRing, double, double)
    CASE("plus_times:f32:f32", PlusTimesSRing, float, float)
    CASE("plus_times:i32:i32", PlusTimesSRing, int32_t, int32_t)
    CASE("plus_times:i64:i64", PlusTimesSRing, int64_t, int64_t)
    CASE("plus_times:bool:i32", PlusTimesSRing, bool, int32_t)
    CASE("plus_times:bool:i64", PlusTimesSRing, bool, int64_t)
    CASE("plus_times:bool:f32", PlusTimesSRing, bool, float)
    CASE("plus_times:bool:f64", PlusTimesSRing, bool, double)
    CASE("plus_times:bool:bool", PlusTimesSRing, bool, bool)
    CASE("min_plus:i32:i32", MinPlusSRing, int32_t, int32_t)
    CASE("min_plus:i64:i64", MinPlusSRing, int64_t, int64_t)
    CASE("min_plus:f32:f32", MinPlusSRing, float, float)
    CASE("min_plus:f64:f64", MinPlusSRing, double, double)
    CASE("select_max:bool:i32", SelectMaxSRing, bool, int32_t)
    CASE("select_max:bool:i64", SelectMaxSRing, bool, int64_t)
#undef CASE
    fprintf(stderr, "cbref_spmm: unknown key %s\n", key);
    return 2;
}

// ParallelReadMM (SpParMat.cpp:3978-4115) with maximum<double>() on duplicates, then dump triples.
// Call once with I==NULL to get sizes (nnz after symmetric expansion), then with buffers.
int cbref_read_mm(const char* path, int64_t* m, int64_t* n, int64_t* nnz, int64_t* I, int64_t* J, double* V) {
    typedef SpDCCols<int64_t, double> D;
    SpParMat<int64_t, double, D> A(grid1());
    A.ParallelReadMM(std::string(path), true, maximum<double>());
    *m = A.getnrow();
    *n = A.getncol();
    *nnz = A.getnnz();
    if (I) {
        Dcsc<int64_t, double>* d = A.seq().GetDCSC();
        int64_t q = 0;
        if (d)
            for (int64_t c = 0; c < d->nzc; ++c)
                for (int64_t p = d->cp[c]; p < d->cp[c + 1]; ++p, ++q) { I[q] = d->ir[p]; J[q] = d->jc[c]; V[q] = d->numx[p]; }
    }
    return 0;
}

// Sparse x sparse product through the same entry point (what Applications/SpMMError.cpp:83 does):
// C = A*B under PlusTimesSRing<int64,int64>; returns nnz(C) and, if buffers given, the triples column-major.
int cbref_spgemm_i64(int64_t m, int64_t kdim, int64_t n, int64_t nnzA, const int64_t* AI, const int64_t* AJ, const int64_t* AV,
                     int64_t nnzB, const int64_t* BI, const int64_t* BJ, const int64_t* BV,
                     int64_t* nnzC, int64_t* CI, int64_t* CJ, int64_t* CV) {
    typedef SpDCCols<int64_t, int64_t> D;
    SpParMat<int64_t, int64_t, D> A(tile_from_coo<int64_t>(m, kdim, nnzA, AI, AJ, AV), grid1());
    SpParMat<int64_t, int64_t, D> B(tile_from_coo<int64_t>(kdim, n, nnzB, BI, BJ, BV), grid1());
    SpParMat<int64_t, int64_t, D> C = Mult_AnXBn_Synch<PlusTimesSRing<int64_t, int64_t>, int64_t, D>(A, B);
    *nnzC = C.getnnz();
    if (CI) {
        Dcsc<int64_t, int64_t>* d = C.seq().GetDCSC();
        int64_t q = 0;
        if (d)
            for (int64_t c = 0; c < d->nzc; ++c)
                for (int64_t p = d->cp[c]; p < d->cp[c + 1]; ++p, ++q) { CI[q] = d->ir[p]; CJ[q] = d->jc[c]; CV[q] = d->numx[p]; }
    }
    return 0;
}

}
