// spmm_driver - a driver in the style of the reference's ReleaseTests/MultTest.cpp / MultTiming.cpp, written against
// the B200 host layer: declare the SpParMat typedefs, build a grid, read or generate A, multiply, PrintInfo, verify.
//
//   spmm_driver mtx  <file.mtx> <k> [ydump.bin [copy.mtx]]   fp64 PlusTimes on a Matrix Market file (BASELINE config C1); optional
//                                                        ParallelWriteMM -> ParallelReadMM round trip through copy.mtx
//   spmm_driver rmat <scale> <k> <pt_f32|mp_i32|sel_i64|bool> [pr pc]   Kronecker matrix generated on the GPU
//   spmm_driver torus                                    the sparse x sparse program of Applications/SpMMError.cpp
//   spmm_driver spmv <scale> [pr pc]                     dense SpMV<SR>(A, FullyDistVec) under three semirings, DenseParMat::Reduce,
//                                                        DenseParMat += SpParMat and SpParMat::EWiseScale (the steps around the multiply)
//
// Single process, or one process per GPU under a launcher that sets RANK / WORLD_SIZE / LOCAL_RANK
// (python -m torch.distributed.run --no-python ./spmm_driver ...).  Verification replays the multiply with the
// host semiring functors on the local tile for the columns it owns (exact for integers, 1e-12 / 1e-5 for fp).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <limits>
#include "CombBLAS/CombBLAS.h"

using namespace combblas;

static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
template <class T> T hashval(uint64_t h);
template <> double hashval<double>(uint64_t h) { return (double)(2 * (h >> 12) + 1) * 1.1102230246251565404e-16; }
template <> float hashval<float>(uint64_t h) { return (float)(2 * (h >> 41) + 1) * 5.9604644775390625e-08f; }
template <> int32_t hashval<int32_t>(uint64_t h) { return (int32_t)(1 + (h >> 8) % 100); }
template <> int64_t hashval<int64_t>(uint64_t h) { return (int64_t)(1 + (h >> 8) % 100); }
template <> bool hashval<bool>(uint64_t h) { return (h >> 63) != 0; }

// X[i,j] = value(seed 42, i*k + j): the same operand the tests' oracle builds
template <class IT, class NT>
DenseParMat<IT, NT> MakeX(std::shared_ptr<CommGrid> grid, IT n, IT k) {
    DenseParMat<IT, NT> X = DenseParMat<IT, NT>::Global(NT(), grid, n, k);
    IT r0, c0;
    X.GetPlaceInGlobalGrid(n, k, r0, c0);
    for (IT i = 0; i < X.getlocalrows(); ++i)
        for (IT j = 0; j < X.getlocalcols(); ++j)
            X(i, j) = hashval<NT>(splitmix64(42ULL * 0x100000001B3ULL ^ (uint64_t)((r0 + i) * k + (c0 + j))));
    return X;
}

// replay on the host with the semiring's own functors; only valid on a 1 x 1 grid (everything is local)
template <class SR, class IT, class NA, class NX>
bool VerifyLocal(const SpParMat<IT, NA, SpDCCols<IT, NA>>& A, const DenseParMat<IT, NX>& X, const DenseParMat<IT, NX>& Y, double tol) {
    const SpDCCols<IT, NA>& t = A.seq();
    const IT m = t.getnrow(), k = X.getlocalcols();
    std::vector<NX> acc((size_t)m * (size_t)k, SR::id());
    std::vector<char> touched((size_t)m * (size_t)k, 0);
    for (size_t c = 0; c < t.jc.size(); ++c)
        for (IT p = t.cp[c]; p < t.cp[c + 1]; ++p)
            for (IT j = 0; j < k; ++j) {
                const NX prod = SR::multiply((NA)t.numx[(size_t)p], (NX)X(t.jc[c], j));
                const size_t q = (size_t)t.ir[(size_t)p] * (size_t)k + (size_t)j;
                acc[q] = touched[q] ? SR::add(prod, acc[q]) : prod;
                touched[q] = 1;
            }
    size_t bad = 0;
    for (IT i = 0; i < m; ++i)
        for (IT j = 0; j < k; ++j) {
            const double a = (double)acc[(size_t)i * k + j], b = (double)Y(i, j);
            if (tol == 0 ? a != b : std::fabs(a - b) > tol * std::max(std::fabs(a), 1e-300)) ++bad;
        }
    return bad == 0;
}

template <class SR, class NA, class NX>
int RunRmat(int scale, int64_t k, int pr, int pc, double tol, bool values) {
    typedef SpParMat<int64_t, NA, SpDCCols<int64_t, NA>> PSpMat;
    std::shared_ptr<CommGrid> grid(new CommGrid(MPI_COMM_WORLD, pr, pc));
    PSpMat A(grid);
    A.GenGraph500(scale, 16, true, 0, values, 1);
    A.PrintInfo();
    DenseParMat<int64_t, NX> X = MakeX<int64_t, NX>(grid, A.getncol(), k);
    double t0 = MPI_Wtime();
    DenseParMat<int64_t, NX> Y = SpMM<SR>(A, X);
    double t1 = MPI_Wtime();
    Y = SpMM<SR>(A, X);
    double t2 = MPI_Wtime();
    float imb = A.LoadImbalance();
    if (grid->GetRank() == 0)
        std::cout << "SpMM first call " << (t1 - t0) << " s, second call " << (t2 - t1) << " s, load imbalance " << imb << std::endl;
    if (grid->GetSize() == 1 && scale <= 16) {
        if (VerifyLocal<SR>(A, X, Y, tol)) SpParHelper::Print("SpMM working correctly\n");
        else { SpParHelper::Print("ERROR in SpMM, go fix it!\n"); return 1; }
    }
    return 0;
}

int main(int argc, char* argv[]) {
    MPI_Init(&argc, &argv);
    int rc = 0;
    if (argc < 4 && !(argc >= 2 && std::string(argv[1]) == "torus") && !(argc >= 3 && std::string(argv[1]) == "spmv")) {
        SpParHelper::Print("Usage: spmm_driver mtx <file.mtx> <k> [ydump.bin] | rmat <scale> <k> <pt_f32|mp_i32|sel_i64|bool> [pr pc]\n");
        MPI_Finalize();
        return 2;
    }
    {
        const std::string mode = argv[1];
        if (mode == "torus") {
            // the program of Applications/SpMMError.cpp: G1, G2 from the built-in 16x16 torus arrays, G3 a copy,
            // three sparse x sparse products through Mult_AnXBn_Synch; "The nnz values should be 112, 112, 112" (:80)
            typedef SpDCCols<int64_t, int64_t> DCColsType;
            typedef SpParMat<int64_t, int64_t, DCColsType> MatType;
            const int64_t tj[64] = {3,0,1,2,7,4,5,6,11,8,9,10,15,12,13,14,1,2,3,0,5,6,7,4,9,10,11,8,13,14,15,12,
                                    12,13,14,15,0,1,2,3,4,5,6,7,8,9,10,11,4,5,6,7,8,9,10,11,12,13,14,15,0,1,2,3};
            std::vector<int64_t> ri(64), ci(tj, tj + 64), vv(64, 1);
            for (int i = 0; i < 64; ++i) ri[i] = i % 16;
            std::shared_ptr<CommGrid> grid(new CommGrid(MPI_COMM_WORLD, 0, 0));
            MatType G1(16, 16, ri, ci, vv, grid), G2(16, 16, ri, ci, vv, grid);
            MatType G3(G1);
            G1.PrintInfo(); G2.PrintInfo(); G3.PrintInfo();
            SpParHelper::Print("The nnz values should be 112, 112, 112:\n");
            MatType G12 = Mult_AnXBn_Synch<PlusTimesSRing<int64_t, int64_t>, int64_t, DCColsType>(G1, G2);
            G12.PrintInfo();
            MatType G13 = PSpGEMM<PlusTimesSRing<int64_t, int64_t>>(G1, G3);
            G13.PrintInfo();
            MatType G23 = Mult_AnXBn_Synch<PlusTimesSRing<int64_t, int64_t>, int64_t, DCColsType>(G2, G3);
            G23.PrintInfo();
            int64_t twos = 0, fours = 0;
            for (int64_t v : G12.seq().numx) { twos += v == 2; fours += v == 4; }
            twos = grid->SumWorld(twos); fours = grid->SumWorld(fours);
            if (G12.getnnz() == 112 && G13 == G12 && G23 == G12 && twos == 96 && fours == 16) SpParHelper::Print("SpGEMM (sparse x sparse) working correctly\n");
            else { SpParHelper::Print("ERROR in SpGEMM, go fix it!\n"); rc = 1; }
        } else if (mode == "spgemm") {
            // A (Kronecker, int64 weights) times a tall-skinny SPARSE B with ~d nonzeros per column, MinPlus semiring;
            // verified entry by entry (structure and values) against a host replay with the semiring's own functors
            typedef SpDCCols<int64_t, int64_t> DC;
            typedef SpParMat<int64_t, int64_t, DC> Mat;
            typedef MinPlusSRing<int64_t, int64_t> SRmp;
            const int scale = std::atoi(argv[2]);
            const int64_t k = std::atoll(argv[3]), d = argc > 4 ? std::atoll(argv[4]) : 3;
            std::shared_ptr<CommGrid> grid(new CommGrid(MPI_COMM_WORLD, 0, 0));
            Mat A(grid);
            A.GenGraph500(scale, 8, true, 0, true, 1);
            const int64_t n = A.getncol();
            std::vector<int64_t> bi, bj, bv;
            for (int64_t j = 0; j < k; ++j)
                for (int64_t e = 0; e < d; ++e) {
                    const uint64_t h = splitmix64((uint64_t)(j * 131 + e) ^ 0xB5ULL);
                    bi.push_back((int64_t)(h % (uint64_t)n)); bj.push_back(j); bv.push_back(1 + (int64_t)((h >> 40) % 50));
                }
            Mat B(n, k, bi, bj, bv, grid, true);
            A.PrintInfo(); B.PrintInfo();
            Mat C = Mult_AnXBn_Synch<SRmp, int64_t, DC>(A, B);
            C.PrintInfo();
            if (grid->GetSize() == 1) {
                const DC &a = A.seq(), &b = B.seq(), &c = C.seq();
                std::vector<int64_t> acc((size_t)n * (size_t)k, 0);
                std::vector<char> hit((size_t)n * (size_t)k, 0);
                std::vector<std::vector<std::pair<int64_t, int64_t>>> bcol((size_t)k);      // column j of B: (row, value)
                for (size_t cc = 0; cc < b.jc.size(); ++cc)
                    for (int64_t p = b.cp[cc]; p < b.cp[cc + 1]; ++p) bcol[(size_t)b.jc[cc]].push_back({b.ir[(size_t)p], b.numx[(size_t)p]});
                std::vector<int64_t> acol_of((size_t)n, -1);
                for (size_t cc = 0; cc < a.jc.size(); ++cc) acol_of[(size_t)a.jc[cc]] = (int64_t)cc;
                for (int64_t j = 0; j < k; ++j)
                    for (auto& e : bcol[(size_t)j]) {
                        const int64_t cc = acol_of[(size_t)e.first];
                        if (cc < 0) continue;
                        for (int64_t p = a.cp[(size_t)cc]; p < a.cp[(size_t)cc + 1]; ++p) {
                            const size_t q = (size_t)a.ir[(size_t)p] * (size_t)k + (size_t)j;
                            const int64_t prod = SRmp::multiply(a.numx[(size_t)p], e.second);
                            acc[q] = hit[q] ? SRmp::add(prod, acc[q]) : prod;
                            hit[q] = 1;
                        }
                    }
                int64_t want = 0, bad = 0;
                for (char h : hit) want += h;
                for (size_t cc = 0; cc < c.jc.size(); ++cc)
                    for (int64_t p = c.cp[cc]; p < c.cp[cc + 1]; ++p) {
                        const size_t q = (size_t)c.ir[(size_t)p] * (size_t)k + (size_t)c.jc[cc];
                        if (!hit[q] || acc[q] != c.numx[(size_t)p]) ++bad;
                    }
                if (bad == 0 && want == c.getnnz()) SpParHelper::Print("SpGEMM (sparse x sparse) working correctly\n");
                else { SpParHelper::Print("ERROR in SpGEMM, go fix it!\n"); rc = 1; }
            }
        } else if (mode == "spmv") {
            // y = A x through the dense SpMV interface (reference ParFriends.h:1924-1996), checked against k = 1 SpMM and
            // against a host replay on one process; then the DenseParMat / SpParMat epilogues of SURVEY.md section 8 row f3
            typedef SpParMat<int64_t, int64_t, SpDCCols<int64_t, int64_t>> Mat;
            typedef SpParMat<int64_t, bool, SpDCCols<int64_t, bool>> BMat;
            const int scale = std::atoi(argv[2]);
            const int pr = argc > 4 ? std::atoi(argv[3]) : 0, pc = argc > 4 ? std::atoi(argv[4]) : 0;
            std::shared_ptr<CommGrid> grid(new CommGrid(MPI_COMM_WORLD, pr, pc));
            Mat A(grid);
            A.GenGraph500(scale, 8, true, 0, true, 1);
            A.PrintInfo();
            const int64_t n = A.getncol(), m = A.getnrow();
            bool ok = true;
            // (1) MinPlus: vector of hashed weights with a few "infinite" entries
            FullyDistVec<int64_t, int64_t> x(grid, n, 0);
            for (int64_t i = 0; i < x.LocArrSize(); ++i) {
                const uint64_t h = splitmix64((uint64_t)(x.LengthUntil() + i) ^ 0xC0FFEEULL);
                x.SetLocalElement(i, h % 97 == 0 ? std::numeric_limits<int64_t>::max() : (int64_t)(1 + h % 100));
            }
            FullyDistVec<int64_t, int64_t> y = SpMV<MinPlusSRing<int64_t, int64_t>>(A, x);
            ok = ok && y.TotalLength() == m;
            // the same product through SpMM: x repeated in one panel column per processor column (every process owns one)
            DenseParMat<int64_t, int64_t> X1 = DenseParMat<int64_t, int64_t>::Global(0, grid, n, grid->GetGridCols());
            {
                const std::vector<int64_t> xw = x.Gather();
                int64_t r0, c0;
                X1.GetPlaceInGlobalGrid(n, (int64_t)grid->GetGridCols(), r0, c0);
                for (int64_t i = 0; i < X1.getlocalrows(); ++i)
                    for (int64_t j = 0; j < X1.getlocalcols(); ++j) X1(i, j) = xw[(size_t)(r0 + i)];
            }
            DenseParMat<int64_t, int64_t> Y1 = SpMM<MinPlusSRing<int64_t, int64_t>>(A, X1);
            FullyDistVec<int64_t, int64_t> yr = Y1.Reduce(Row, [](int64_t a, int64_t b) { return std::min(a, b); }, std::numeric_limits<int64_t>::max());
            ok = ok && (yr == y);
            const int64_t reached = y.Count([](int64_t v) { return v != std::numeric_limits<int64_t>::max(); });
            if (grid->GetRank() == 0) std::cout << "SpMV MinPlus: " << reached << " of " << m << " rows reached" << std::endl;
            // (2) PlusTimes with an all-ones vector = row degrees weighted by A; Reduce(Row, +) of A as dense must agree
            FullyDistVec<int64_t, int64_t> ones(grid, n, 1);
            FullyDistVec<int64_t, int64_t> deg = SpMV<PlusTimesSRing<int64_t, int64_t>>(A, ones);
            const int64_t total = deg.Reduce(std::plus<int64_t>(), (int64_t)0);
            int64_t local = 0;
            for (int64_t v : A.seq().numx) local += v;
            ok = ok && total == grid->SumWorld(local);
            // (3) SelectMax<bool,T> with values below the identity -1: SpMV starts from id(), so they are clipped (axpy semantics)
            // the boolean matrix through the reference's construction path: DistEdgeList -> SpParMat (GenWriteMatrix.cpp:96-110)
            double initiator[4] = {.57, .19, .19, .05};
            DistEdgeList<int64_t> DEL(grid);
            DEL.GenGraph500Data(initiator, scale, 8, true, true);
            BMat P(DEL, false);
            ok = ok && P.getnrow() == n && P.RemoveLoops() >= 0 && P.RemoveLoops() == 0;      // the reference's stream has self loops; none after the first call
            FullyDistVec<int64_t, int64_t> neg(grid, n, -5);
            FullyDistVec<int64_t, int64_t> sel = SpMV<SelectMaxSRing<bool, int64_t>>(P, neg);
            ok = ok && sel.Count([](int64_t v) { return v != -1; }) == 0;
            // (4) epilogues on one process grid cell: D += A, then A.EWiseScale(D) squares the stored values
            DenseParMat<int64_t, int64_t> D(0, grid, A.getlocalrows(), A.getlocalcols());
            D += A;
            FullyDistVec<int64_t, int64_t> rowsum = D.Reduce(Row, std::plus<int64_t>(), (int64_t)0);
            ok = ok && (rowsum == deg);
            FullyDistVec<int64_t, int64_t> colsum = D.Reduce(Column, std::plus<int64_t>(), (int64_t)0);
            ok = ok && colsum.TotalLength() == n && colsum.Reduce(std::plus<int64_t>(), (int64_t)0) == total;
            Mat A2(A);
            A2.EWiseScale(D);
            FullyDistVec<int64_t, int64_t> sq = SpMV<PlusTimesSRing<int64_t, int64_t>>(A2, ones);
            int64_t localsq = 0;
            for (int64_t v : A.seq().numx) localsq += v * v;
            ok = ok && sq.Reduce(std::plus<int64_t>(), (int64_t)0) == grid->SumWorld(localsq);
            if (grid->GetSize() == 1) {
                // host replay of (1) with the semiring's own functors
                const SpDCCols<int64_t, int64_t>& t = A.seq();
                std::vector<int64_t> ref((size_t)m, MinPlusSRing<int64_t, int64_t>::id());
                for (size_t c = 0; c < t.jc.size(); ++c)
                    for (int64_t p = t.cp[c]; p < t.cp[c + 1]; ++p)
                        MinPlusSRing<int64_t, int64_t>::axpy(t.numx[(size_t)p], x.GetLocArr()[t.jc[c]], ref[(size_t)t.ir[(size_t)p]]);
                for (int64_t i = 0; i < m; ++i) ok = ok && ref[(size_t)i] == y.GetLocArr()[i];
            }
            ok = grid->MinWorld(ok ? 1 : 0) == 1;
            if (ok) SpParHelper::Print("SpMV and dense epilogues working correctly\n");
            else { SpParHelper::Print("ERROR in SpMV / epilogues, go fix it!\n"); rc = 1; }
        } else if (mode == "mtx") {
            typedef SpParMat<int64_t, double, SpDCCols<int64_t, double>> PSpMat_Double;
            std::shared_ptr<CommGrid> grid(new CommGrid(MPI_COMM_WORLD, 0, 0));
            PSpMat_Double A(grid);
            A.ParallelReadMM(argv[2], true, maximum<double>());
            A.PrintInfo();
            const int64_t k = std::atoll(argv[3]);
            DenseParMat<int64_t, double> X = MakeX<int64_t, double>(grid, A.getncol(), k);
            DenseParMat<int64_t, double> Y = SpMM<PlusTimesSRing<double, double>>(A, X);
            if (argc > 5) {
                // ParallelWriteMM -> ParallelReadMM round trip (reference SpParMat.cpp:4118-4210 / :3978-4115)
                A.ParallelWriteMM(argv[5], true);
                PSpMat_Double A2(grid);
                A2.ParallelReadMM(argv[5], true, maximum<double>());
                if (A2 == A) SpParHelper::Print("Matrix Market round trip working correctly\n");
                else { SpParHelper::Print("ERROR in the Matrix Market round trip, go fix it!\n"); rc = 1; }
            }
            if (grid->GetSize() == 1) {
                if (VerifyLocal<PlusTimesSRing<double, double>>(A, X, Y, 1e-12)) SpParHelper::Print("SpMM working correctly\n");
                else { SpParHelper::Print("ERROR in SpMM, go fix it!\n"); rc = 1; }
                if (argc > 4) {
                    FILE* f = std::fopen(argv[4], "wb");
                    std::fwrite(Y.data(), sizeof(double), (size_t)Y.getlocalrows() * (size_t)Y.getlocalcols(), f);
                    std::fclose(f);
                }
            }
        } else {
            const int scale = std::atoi(argv[2]);
            const int64_t k = std::atoll(argv[3]);
            const std::string what = argc > 4 ? argv[4] : "pt_f32";
            const int pr = argc > 6 ? std::atoi(argv[5]) : 0, pc = argc > 6 ? std::atoi(argv[6]) : 0;
            if (what == "pt_f32") rc = RunRmat<PlusTimesSRing<float, float>, float, float>(scale, k, pr, pc, 1e-5, true);
            else if (what == "mp_i32") rc = RunRmat<MinPlusSRing<int32_t, int32_t>, int32_t, int32_t>(scale, k, pr, pc, 0, true);
            else if (what == "sel_i64") rc = RunRmat<SelectMaxSRing<bool, int64_t>, bool, int64_t>(scale, k, pr, pc, 0, false);
            else if (what == "bool") rc = RunRmat<PlusTimesSRing<bool, bool>, bool, bool>(scale, k, pr, pc, 0, false);
            else { SpParHelper::Print("unknown semiring key\n"); rc = 2; }
        }
    }
    MPI_Finalize();
    return rc;
}
