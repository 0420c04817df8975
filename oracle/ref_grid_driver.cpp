// oracle/_ref/cbref_grid: the UNMODIFIED reference multiply on a pr x pr process grid inside one container
// (SURVEY.md section 8 row f4).  TEST INFRASTRUCTURE ONLY - the product never links, loads or runs this.
//
// Compiled from the reference sources where they lie under /root/reference against oracle/mpi_multi (one forked
// process per rank over shared memory; launcher in cbmpi.cpp, CBMPI_NP ranks).  What runs is the reference's own code:
//   * CommGrid / row and column worlds            src/CommGrid.cpp:37-75
//   * SpParMat(m,n,FullyDistVec i,j,v)            SpParMat.cpp "matlab sparse" constructor (Alltoallv redistribution)
//   * Mult_AnXBn_Synch / _DoubleBuff / _Overlap   ParFriends.h:798-1235  (BCastMatrix, LocalHybridSpGEMM, MultiwayMerge)
//   * SpMV<SR>(SpParMat, FullyDistVec)            ParFriends.h:1924-1996 (TransposeVector, Allgatherv, Reduce_scatter)
//
//   cbref_grid torus
//       the program of Applications/SpMMError.cpp (that file no longer compiles against its own headers: it omits the
//       NUO/UDERO template arguments, ParFriends.h:1004-1005); same steps with the arguments spelled out.
//   cbref_grid layout <glen>
//       the FullyDistVec distribution of a vector of that length: LengthUntil / MyLocLength per rank, Owner of every index
//   cbref_grid mm <file.mtx> <k> <dir>
//       BASELINE config C1 on a grid: ParallelReadMM by the reference itself, then x dense k columns fp64 (X from <dir>/X.bin)
//   cbref_grid spmm <key> <via> <dir>
//       reads <dir>/meta.txt ("m n nnz k hasV"), I.bin J.bin (int64), V.bin (A's value type), X.bin (n x k row-major),
//       writes <dir>/Y_<rank>.bin = int64 header {row0, col0, rows, cols} + the rank's dense block of Y (T_promote),
//       absent entries = SR::id().  via: 0 Synch, 1 k x SpMV, 2 DoubleBuff, 3 Overlap.  A is distributed twice: by the
//       owner rule restated here (SpParMat.cpp:5066-5096) and by the reference's own constructor; the two must be ==.
#include <mpi.h>
#include <omp.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <tuple>
#include <vector>
#include "CombBLAS/CombBLAS.h"

using namespace combblas;

int cblas_splits = 1;
double cblas_alltoalltime, cblas_allgathertime, cblas_mergeconttime, cblas_transvectime, cblas_localspmvtime;

namespace {

template <class T>
std::vector<T> slurp(const std::string& path, size_t count) {
    std::vector<T> v(count);
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) { std::fprintf(stderr, "cbref_grid: cannot open %s\n", path.c_str()); MPI_Abort(MPI_COMM_WORLD, NOFILE); }
    if (count && std::fread(v.data(), sizeof(T), count, f) != count) { std::fprintf(stderr, "cbref_grid: short read %s\n", path.c_str()); MPI_Abort(MPI_COMM_WORLD, NOFILE); }
    std::fclose(f);
    return v;
}
// bool vectors cannot hand out data(); keep value arrays as bytes for NT = bool
template <class NT> struct store { typedef NT type; };
template <> struct store<bool> { typedef uint8_t type; };

struct Block { int64_t lo, len; };
// block b of `total` over `parts`: floor division, the last block takes the remainder (SpParMat.cpp:5066-5096)
Block block_of(int64_t total, int parts, int b) {
    const int64_t per = total / parts;
    return Block{per * b, b == parts - 1 ? total - per * (parts - 1) : per};
}

template <class NT>
SpDCCols<int64_t, NT>* local_tile(int64_t rows, int64_t cols, std::vector<std::tuple<int64_t, int64_t, NT>>& t) {
    typedef std::tuple<int64_t, int64_t, NT> Tup;
    std::sort(t.begin(), t.end(), [](const Tup& a, const Tup& b) {
        return std::get<1>(a) != std::get<1>(b) ? std::get<1>(a) < std::get<1>(b) : std::get<0>(a) < std::get<0>(b);
    });
    Tup dummy;
    return new SpDCCols<int64_t, NT>(rows, cols, (int64_t)t.size(), t.empty() ? &dummy : t.data(), false);
}


// A through SpParMat(m, n, FullyDistVec rows, cols, vals): the reference redistributes the triples itself
template <class NA, class SA>
bool same_as_reference_ctor(SpParMat<int64_t, NA, SpDCCols<int64_t, NA>>& A, int64_t m, int64_t n, int64_t nnz, const std::vector<int64_t>& I,
                            const std::vector<int64_t>& J, const std::vector<SA>& V, bool hasV, std::shared_ptr<CommGrid> grid, std::false_type) {
    FullyDistVec<int64_t, int64_t> vi(grid, nnz, 0), vj(grid, nnz, 0);
    FullyDistVec<int64_t, NA> vv(grid, nnz, NA());
    const int64_t off = vi.LengthUntil(), len = vi.MyLocLength();
    for (int64_t q = 0; q < len; ++q) {
        vi.SetLocalElement(q, I[(size_t)(off + q)]);
        vj.SetLocalElement(q, J[(size_t)(off + q)]);
        vv.SetLocalElement(q, hasV ? (NA)V[(size_t)(off + q)] : NA(1));
    }
    SpParMat<int64_t, NA, SpDCCols<int64_t, NA>> A2(m, n, vi, vj, vv, false);
    return A2 == A;
}
// bool matrices: FullyDistVec<IT,bool> is disabled in the reference (FullyDistVec.h:61); build an int pattern and convert
template <class NA, class SA>
bool same_as_reference_ctor(SpParMat<int64_t, NA, SpDCCols<int64_t, NA>>& A, int64_t m, int64_t n, int64_t nnz, const std::vector<int64_t>& I,
                            const std::vector<int64_t>& J, const std::vector<SA>& V, bool hasV, std::shared_ptr<CommGrid> grid, std::true_type) {
    if (hasV) return true;                                        // explicit boolean values: nothing to cross-check with
    FullyDistVec<int64_t, int64_t> vi(grid, nnz, 0), vj(grid, nnz, 0);
    const int64_t off = vi.LengthUntil(), len = vi.MyLocLength();
    for (int64_t q = 0; q < len; ++q) { vi.SetLocalElement(q, I[(size_t)(off + q)]); vj.SetLocalElement(q, J[(size_t)(off + q)]); }
    SpParMat<int64_t, int, SpDCCols<int64_t, int>> A2i(m, n, vi, vj, 1, false);
    SpParMat<int64_t, NA, SpDCCols<int64_t, NA>> A2 = A2i;        // the reference's own type conversion (SpParMat.h)
    return A2 == A;
}

// k dense SpMVs with vectors in the reference's FullyDistVec distribution; writes this rank's slice of Y
template <class SR, class NA, class NX>
double spmv_columns(SpParMat<int64_t, NA, SpDCCols<int64_t, NA>>& A, const std::vector<typename store<NX>::type>& X, int64_t n, int64_t k,
                    std::shared_ptr<CommGrid> grid, const std::string& path, std::false_type) {
    typedef typename SR::T_promote NO;
    std::vector<NO> out;
    int64_t yoff = 0, ylen = 0;
    double seconds = 0;
    for (int64_t j = 0; j < k; ++j) {
        FullyDistVec<int64_t, NX> x(grid, n, NX());
        const int64_t off = x.LengthUntil(), len = x.MyLocLength();
        for (int64_t q = 0; q < len; ++q) x.SetLocalElement(q, (NX)X[(size_t)((off + q) * k + j)]);
        const double t0 = MPI_Wtime();
        FullyDistVec<int64_t, NO> y = SpMV<SR>(A, x);
        seconds += MPI_Wtime() - t0;
        yoff = y.LengthUntil(); ylen = y.MyLocLength();
        if (j == 0) out.assign((size_t)(ylen * k), NO());
        for (int64_t q = 0; q < ylen; ++q) out[(size_t)(q * k + j)] = y.GetLocArr()[q];
    }
    const int64_t hdr[4] = {yoff, 0, ylen, k};
    FILE* f = std::fopen(path.c_str(), "wb");
    std::fwrite(hdr, sizeof(int64_t), 4, f);
    std::fwrite(out.data(), sizeof(NO), out.size(), f);
    std::fclose(f);
    return seconds;
}
template <class SR, class NA, class NX>
double spmv_columns(SpParMat<int64_t, NA, SpDCCols<int64_t, NA>>&, const std::vector<typename store<NX>::type>&, int64_t, int64_t,
                    std::shared_ptr<CommGrid>, const std::string&, std::true_type) {
    std::fprintf(stderr, "cbref_grid: the SpMV path has no bool vectors in the reference\n");
    MPI_Abort(MPI_COMM_WORLD, INVALIDPARAMS);
    return 0;
}

template <class SR, class NA, class NX>
int run_spmm(int via, const std::string& dir) {
    typedef typename SR::T_promote NO;
    typedef typename store<NA>::type SA;
    typedef typename store<NX>::type SX;
    typedef typename store<NO>::type SO;
    typedef SpDCCols<int64_t, NA> DA;
    typedef SpDCCols<int64_t, NX> DX;
    typedef SpDCCols<int64_t, NO> DO;
    long long m, n, nnz, k; int hasV;
    {
        FILE* f = std::fopen((dir + "/meta.txt").c_str(), "r");
        if (!f || std::fscanf(f, "%lld %lld %lld %lld %d", &m, &n, &nnz, &k, &hasV) != 5) { std::fprintf(stderr, "cbref_grid: bad meta.txt\n"); MPI_Abort(MPI_COMM_WORLD, NOFILE); }
        std::fclose(f);
    }
    std::vector<int64_t> I = slurp<int64_t>(dir + "/I.bin", (size_t)nnz), J = slurp<int64_t>(dir + "/J.bin", (size_t)nnz);
    std::vector<SA> V = hasV ? slurp<SA>(dir + "/V.bin", (size_t)nnz) : std::vector<SA>();
    std::vector<SX> X = slurp<SX>(dir + "/X.bin", (size_t)(n * k));

    std::shared_ptr<CommGrid> grid(new CommGrid(MPI_COMM_WORLD, 0, 0));
    const int pr = grid->GetGridRows(), pc = grid->GetGridCols();
    const int myrow = grid->GetRankInProcCol(), mycol = grid->GetRankInProcRow();       // CommGrid.h:109-110
    const int rank = grid->GetRank();

    // A by the restated owner rule
    const Block ar = block_of(m, pr, myrow), ac = block_of(n, pc, mycol);
    std::vector<std::tuple<int64_t, int64_t, NA>> at;
    for (long long p = 0; p < nnz; ++p)
        if (I[p] >= ar.lo && I[p] < ar.lo + ar.len && J[p] >= ac.lo && J[p] < ac.lo + ac.len)
            at.push_back(std::make_tuple(I[p] - ar.lo, J[p] - ac.lo, hasV ? (NA)V[p] : NA(1)));
    SpParMat<int64_t, NA, DA> A(local_tile<NA>(ar.len, ac.len, at), grid);

    // A again through the reference's own distribution code; the two must agree tile for tile
    if (!std::getenv("CBREF_SKIP_DISTCHECK")) {
        const bool same = same_as_reference_ctor(A, m, n, nnz, I, J, V, hasV != 0, grid, std::is_same<NA, bool>());
        if (rank == 0) std::printf("distribution %s\n", same ? "agrees with the reference constructor" : "DIFFERS from the reference constructor");
        if (!same) return 5;
    }
    const int64_t gm = A.getnrow(), gn = A.getncol(), gnnz = A.getnnz();
    if (rank == 0) std::printf("A: %lld x %lld, %lld nonzeros on a %d x %d grid\n", (long long)gm, (long long)gn, (long long)gnnz, pr, pc);

    const NO id = SR::id();
    double seconds = 0;
    if (via == 1) {
        seconds = spmv_columns<SR, NA, NX>(A, X, n, k, grid, dir + "/Y_" + std::to_string(rank) + ".bin", std::is_same<NX, bool>());
    } else {
        const Block xr = block_of(n, pr, myrow), xc = block_of(k, pc, mycol);
        std::vector<std::tuple<int64_t, int64_t, NX>> xt;
        for (int64_t i = 0; i < xr.len; ++i)
            for (int64_t j = 0; j < xc.len; ++j) xt.push_back(std::make_tuple(i, j, (NX)X[(size_t)((xr.lo + i) * k + xc.lo + j)]));
        SpParMat<int64_t, NX, DX> Xs(local_tile<NX>(xr.len, xc.len, xt), grid);
        const double t0 = MPI_Wtime();
        SpParMat<int64_t, NO, DO> C = via == 2 ? Mult_AnXBn_DoubleBuff<SR, NO, DO>(A, Xs)
                                    : via == 3 ? Mult_AnXBn_Overlap<SR, NO, DO>(A, Xs)
                                               : Mult_AnXBn_Synch<SR, NO, DO>(A, Xs);
        seconds = MPI_Wtime() - t0;
        std::vector<SO> out((size_t)(ar.len * xc.len), (SO)id);
        Dcsc<int64_t, NO>* d = C.seq().GetDCSC();
        if (d)
            for (int64_t c = 0; c < d->nzc; ++c)
                for (int64_t p = d->cp[c]; p < d->cp[c + 1]; ++p) out[(size_t)(d->ir[p] * xc.len + d->jc[c])] = (SO)d->numx[p];
        const int64_t hdr[4] = {ar.lo, xc.lo, ar.len, xc.len};
        FILE* f = std::fopen((dir + "/Y_" + std::to_string(rank) + ".bin").c_str(), "wb");
        std::fwrite(hdr, sizeof(int64_t), 4, f);
        std::fwrite(out.data(), sizeof(SO), out.size(), f);
        std::fclose(f);
        const int64_t cnnz = C.getnnz();
        if (rank == 0) std::printf("C: %lld stored entries\n", (long long)cnnz);
    }
    double worst = 0;
    MPI_Allreduce(&seconds, &worst, 1, MPI_DOUBLE, MPI_MAX, MPI_COMM_WORLD);
    if (rank == 0) std::printf("multiply seconds %.6f (max over ranks, %d thread(s) per rank)\n", worst, omp_get_max_threads());
    return 0;
}

// BASELINE config C1 on a process grid: the reference reads the Matrix Market file itself (ParallelReadMM: every rank
// parses its byte range of the file, symmetric expansion, Alltoallv redistribution, SpParMat.cpp:3978-4115), then
// Y = A x X(k, fp64) through Mult_AnXBn_Synch.  X comes from <dir>/X.bin (n x k doubles); writes <dir>/Y_<rank>.bin.
int run_mm(const std::string& file, int64_t k, const std::string& dir) {
    typedef SpDCCols<int64_t, double> DD;
    typedef PlusTimesSRing<double, double> SR;
    std::shared_ptr<CommGrid> grid(new CommGrid(MPI_COMM_WORLD, 0, 0));
    const int pr = grid->GetGridRows(), pc = grid->GetGridCols();
    const int myrow = grid->GetRankInProcCol(), mycol = grid->GetRankInProcRow(), rank = grid->GetRank();
    SpParMat<int64_t, double, DD> A(grid);
    A.ParallelReadMM(file, true, maximum<double>());
    const int64_t m = A.getnrow(), n = A.getncol(), nnz = A.getnnz();
    std::vector<double> X = slurp<double>(dir + "/X.bin", (size_t)(n * k));
    const Block ar = block_of(m, pr, myrow), xr = block_of(n, pr, myrow), xc = block_of(k, pc, mycol);
    std::vector<std::tuple<int64_t, int64_t, double>> xt;
    for (int64_t i = 0; i < xr.len; ++i)
        for (int64_t j = 0; j < xc.len; ++j) xt.push_back(std::make_tuple(i, j, X[(size_t)((xr.lo + i) * k + xc.lo + j)]));
    SpParMat<int64_t, double, DD> Xs(local_tile<double>(xr.len, xc.len, xt), grid);
    SpParMat<int64_t, double, DD> C = Mult_AnXBn_Synch<SR, double, DD>(A, Xs);
    std::vector<double> out((size_t)(ar.len * xc.len), 0.0);
    Dcsc<int64_t, double>* d = C.seq().GetDCSC();
    if (d)
        for (int64_t c = 0; c < d->nzc; ++c)
            for (int64_t p = d->cp[c]; p < d->cp[c + 1]; ++p) out[(size_t)(d->ir[p] * xc.len + d->jc[c])] = d->numx[p];
    const int64_t hdr[4] = {ar.lo, xc.lo, ar.len, xc.len};
    FILE* f = std::fopen((dir + "/Y_" + std::to_string(rank) + ".bin").c_str(), "wb");
    std::fwrite(hdr, sizeof(int64_t), 4, f);
    std::fwrite(out.data(), sizeof(double), out.size(), f);
    std::fclose(f);
    const int64_t cnnz = C.getnnz();
    if (rank == 0) std::printf("A: %lld x %lld, %lld nonzeros on a %d x %d grid; C: %lld stored entries\n", (long long)m, (long long)n, (long long)nnz, pr, pc, (long long)cnnz);
    return 0;
}

// the FullyDistVec distribution as the reference computes it (FullyDist.h:107-260), for every rank and every index
int run_layout(int64_t glen) {
    std::shared_ptr<CommGrid> grid(new CommGrid(MPI_COMM_WORLD, 0, 0));
    FullyDistVec<int64_t, int64_t> v(grid, glen, 0);
    int rank, p;
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    MPI_Comm_size(MPI_COMM_WORLD, &p);
    long long mine[2] = {(long long)v.LengthUntil(), (long long)v.MyLocLength()};
    std::vector<long long> all((size_t)2 * p);
    MPI_Allgather(mine, 2, MPI_LONG_LONG, all.data(), 2, MPI_LONG_LONG, MPI_COMM_WORLD);
    if (rank == 0) {
        std::printf("until");
        for (int r = 0; r < p; ++r) std::printf(" %lld", all[2 * r]);
        std::printf("\nlen");
        for (int r = 0; r < p; ++r) std::printf(" %lld", all[2 * r + 1]);
        std::printf("\nowner");
        std::vector<int64_t> lind((size_t)glen);
        for (int64_t g = 0; g < glen; ++g) std::printf(" %d", v.Owner(g, lind[(size_t)g]));
        std::printf("\nlind");
        for (int64_t g = 0; g < glen; ++g) std::printf(" %lld", (long long)lind[(size_t)g]);
        std::printf("\n");
    }
    return 0;
}

int run_torus() {
    typedef int64_t ValueType;
    typedef SpDCCols<int64_t, ValueType> DCColsType;
    typedef SpParMat<int64_t, ValueType, DCColsType> MatType;
    const int tj[64] = {3,0,1,2,7,4,5,6,11,8,9,10,15,12,13,14,1,2,3,0,5,6,7,4,9,10,11,8,13,14,15,12,
                        12,13,14,15,0,1,2,3,4,5,6,7,8,9,10,11,4,5,6,7,8,9,10,11,12,13,14,15,0,1,2,3};
    int rank;
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    FullyDistVec<int64_t, ValueType> dpvi(64, 0), dpvj(64, 0), dpvv(64, 1);
    for (int i = 0; i < 64; ++i) { dpvi.SetElement(i, i % 16); dpvj.SetElement(i, tj[i]); }
    MatType G1(16, 16, dpvi, dpvj, dpvv), G2(16, 16, dpvi, dpvj, dpvv);
    MatType G3(G1);
    const int64_t n1 = G1.getnnz(), n2 = G2.getnnz(), n3 = G3.getnnz();
    MatType G12 = Mult_AnXBn_Synch<PlusTimesSRing<int64_t, ValueType>, ValueType, DCColsType>(G1, G2);
    MatType G13 = Mult_AnXBn_Synch<PlusTimesSRing<int64_t, ValueType>, ValueType, DCColsType>(G1, G3);
    MatType G23 = Mult_AnXBn_Synch<PlusTimesSRing<int64_t, ValueType>, ValueType, DCColsType>(G2, G3);
    const int64_t n12 = G12.getnnz(), n13 = G13.getnnz(), n23 = G23.getnnz();
    int64_t local[2] = {0, 0}, total[2] = {0, 0};
    Dcsc<int64_t, ValueType>* d = G12.seq().GetDCSC();
    if (d) for (int64_t p = 0; p < d->nz; ++p) { local[0] += d->numx[p] == 2; local[1] += d->numx[p] == 4; }
    MPI_Allreduce(local, total, 2, MPI_LONG_LONG, MPI_SUM, MPI_COMM_WORLD);
    const bool e13 = (G13 == G12), e23 = (G23 == G12);
    if (rank == 0)
        std::printf("torus inputs nnz %lld %lld %lld; products nnz %lld %lld %lld; G12 twos %lld fours %lld; G13==G12 %d G23==G12 %d\n",
                    (long long)n1, (long long)n2, (long long)n3, (long long)n12, (long long)n13, (long long)n23,
                    (long long)total[0], (long long)total[1], (int)e13, (int)e23);
    return 0;
}

}  // namespace

int main(int argc, char* argv[]) {
    MPI_Init(&argc, &argv);
    int rc = 2;
    {
        const std::string mode = argc > 1 ? argv[1] : "";
        if (mode == "torus") rc = run_torus();
        else if (mode == "layout" && argc >= 3) rc = run_layout(std::atoll(argv[2]));
        else if (mode == "mm" && argc >= 5) rc = run_mm(argv[2], std::atoll(argv[3]), argv[4]);
        else if (mode == "spmm" && argc >= 5) {
            const std::string s = argv[2], dir = argv[4];
            const int via = std::atoi(argv[3]);
#define CASE(name, SR, NA, NX) if (s == name) rc = run_spmm<SR<NA, NX>, NA, NX>(via, dir);
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
            if (rc == 2) std::fprintf(stderr, "cbref_grid: unknown key %s\n", s.c_str());
        } else if (!mode.empty()) std::fprintf(stderr, "usage: cbref_grid torus | layout <glen> | mm <file.mtx> <k> <dir> | spmm <key> <via> <dir>\n");
    }
    MPI_Finalize();
    return rc;
}
