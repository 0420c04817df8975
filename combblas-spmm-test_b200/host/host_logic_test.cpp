// CPU-only exercise of the host layer's containers and arithmetic (no CommGrid, no GPU).  Prints one line per check in a
// "key value..." format that tests/test_host_logic_cpu.py compares with independent numpy restatements.
//   host_logic_test tile <m> <n> <triples.txt>   -> DCSC and CSC arrays of the triples
//   host_logic_test mm <file.mtx>                -> triples after Matrix Market expansion
//   host_logic_test semirings                    -> functor tables
//   host_logic_test owner <pr> <pc> <m> <n> <r> <c>
//   host_logic_test fdv <glen> <pr> <pc>         -> the FullyDistVec layout: per process (LengthUntil, MyLocLength), owners of all indices
#include <cstdio>
#include <fstream>
#include <iostream>
#include "CombBLAS/CombBLAS.h"

using namespace combblas;

template <class V>
static void dump(const char* key, const V& v) {
    std::cout << key;
    for (auto x : v) std::cout << ' ' << (long long)x;
    std::cout << '\n';
}
template <class V>
static void dumpf(const char* key, const V& v) {
    std::cout << key;
    std::cout.precision(17);
    for (auto x : v) std::cout << ' ' << (double)x;
    std::cout << '\n';
}

int main(int argc, char** argv) {
    const std::string mode = argc > 1 ? argv[1] : "";
    if (mode == "tile" && argc > 4) {
        const int64_t m = std::atoll(argv[2]), n = std::atoll(argv[3]);
        std::ifstream in(argv[4]);
        std::vector<int64_t> r, c;
        std::vector<double> v;
        long long i, j;
        double x;
        while (in >> i >> j >> x) { r.push_back(i); c.push_back(j); v.push_back(x); }
        SpTuples<int64_t, double> t(m, n, r, c, v);
        t.RemoveDuplicates(cb_sum<double>());
        std::cout << "nnz " << t.getnnz() << '\n';
        if (t.getnnz() == 0) { std::cerr << "no triples parsed\n"; return 3; }
        SpDCCols<int64_t, double> d(t, false);
        dump("ess", d.GetEssentials());
        dump("cp", d.cp); dump("jc", d.jc); dump("ir", d.ir); dumpf("numx", d.numx);
        Arr<int64_t, double> a = d.GetArrays();
        std::cout << "arrs " << a.indarrs.size() << ' ' << a.numarrs.size() << ' ' << a.indarrs[0].count << ' ' << a.indarrs[1].count << ' '
                  << a.indarrs[2].count << ' ' << a.numarrs[0].count << '\n';
        SpDCCols<int64_t, double> e;
        e.Create(d.GetEssentials());                              // receiver side of BCastMatrix: allocate from essentials, then fill
        Arr<int64_t, double> b = e.GetArrays();
        for (size_t q = 0; q < a.indarrs.size(); ++q) std::copy(a.indarrs[q].addr, a.indarrs[q].addr + a.indarrs[q].count, b.indarrs[q].addr);
        std::copy(a.numarrs[0].addr, a.numarrs[0].addr + a.numarrs[0].count, b.numarrs[0].addr);
        std::cout << "roundtrip " << (int)(d == e) << '\n';
        e.numx[0] += 0.5 * EPSILON * std::max(1.0, std::abs(e.numx[0]));      // inside ErrorTolerantEqual
        std::cout << "tolerant " << (int)(d == e) << '\n';
        e.numx[0] += 1.0 + std::abs(e.numx[0]);
        std::cout << "different " << (int)(d == e) << '\n';
        SpCCols<int64_t, double> cs(t, false);
        dump("csc_ess", cs.GetEssentials());
        dump("csc_jc", cs.jc); dump("csc_ir", cs.ir); dumpf("csc_num", cs.num);
        SpDCCols<int64_t, double> tr(t, true);                    // transposed constructor
        dump("tr_ess", tr.GetEssentials());
        SpTuples<int64_t, double> back = TilesToTuples(d);
        std::cout << "tuples_back " << back.getnnz() << ' ' << (back.getnnz() ? back.rowindex(0) : -1) << ' ' << (back.getnnz() ? back.colindex(0) : -1) << '\n';
        return 0;
    }
    if (mode == "mm" && argc > 2) {
        int64_t m = 0, n = 0;
        std::vector<int64_t> r, c;
        std::vector<double> v;
        typedef SpParMat<int64_t, double, SpDCCols<int64_t, double>> M;
        if (!M::ReadMMTriples(argv[2], true, m, n, r, c, v)) { std::cout << "nofile\n"; return 1; }
        std::cout << "dims " << m << ' ' << n << ' ' << r.size() << '\n';
        SpTuples<int64_t, double> t(m, n, r, c, v);
        t.RemoveDuplicates(maximum<double>());
        std::vector<int64_t> rr, cc;
        std::vector<double> vv;
        for (int64_t p = 0; p < t.getnnz(); ++p) { rr.push_back(t.rowindex(p)); cc.push_back(t.colindex(p)); vv.push_back(t.numvalue(p)); }
        dump("rows", rr); dump("cols", cc); dumpf("vals", vv);
        return 0;
    }
    if (mode == "semirings") {
        typedef MinPlusSRing<int32_t, int32_t> MP;
        typedef PlusTimesSRing<double, double> PT;
        typedef PlusTimesSRing<bool, int64_t> PTB;
        typedef PlusTimesSRing<bool, bool> BB;
        typedef SelectMaxSRing<bool, int64_t> SM;
        const int32_t inf = std::numeric_limits<int32_t>::max();
        std::cout << "mp " << MP::id() << ' ' << MP::add(3, 7) << ' ' << MP::multiply(3, 7) << ' ' << MP::multiply(inf, 7) << ' ' << MP::multiply(3, inf)
                  << ' ' << inf_plus<int32_t>(inf, inf) << '\n';
        int32_t y = 10; MP::axpy(2, 3, y);
        std::cout << "mp_axpy " << y << '\n';
        std::cout << "pt " << PT::id() << ' ' << PT::add(1.5, 2.25) << ' ' << PT::multiply(1.5, 2.0) << '\n';
        std::cout << "ptb " << PTB::id() << ' ' << PTB::multiply(true, 9) << ' ' << PTB::multiply(false, 9) << '\n';
        std::cout << "bb " << (int)BB::id() << ' ' << (int)BB::add(true, true) << ' ' << (int)BB::add(false, false) << ' ' << (int)BB::multiply(true, false)
                  << ' ' << (int)BB::multiply(true, true) << '\n';
        std::cout << "sm " << SM::id() << ' ' << SM::add(-7, -1) << ' ' << SM::multiply(false, 42) << ' ' << SM::multiply(true, -5) << '\n';
        std::cout << "ops " << semiring_traits<PT>::op << ' ' << semiring_traits<MP>::op << ' ' << semiring_traits<SM>::op << ' ' << semiring_traits<BB>::op << ' '
                  << semiring_traits<PTB>::op << ' ' << (int)semiring_traits<MinPlusSRing<bool, bool>>::supported << '\n';
        static_assert(std::is_same<promote_trait<bool, int64_t>::T_promote, int64_t>::value, "bool x T promotes to T");
        static_assert(std::is_same<promote_trait<int, double>::T_promote, double>::value, "int x double promotes to double");
        static_assert(std::is_same<promote_trait<bool, bool>::T_promote, bool>::value, "bool x bool stays bool");
        static_assert(cb_dtype_of<float>::value == CB_F32 && cb_dtype_of<int64_t>::value == CB_I64 && cb_dtype_of<bool>::value == CB_U8, "dtype codes");
        std::cout << "codes " << GRIDMISMATCH << ' ' << DIMMISMATCH << ' ' << NOTSQUARE << ' ' << NOFILE << ' ' << MATRIXALIAS << ' ' << INVALIDPARAMS << '\n';
        return 0;
    }
    if (mode == "owner" && argc > 7) {
        typedef SpParMat<int64_t, double, SpDCCols<int64_t, double>> M;
        int64_t lr, lc;
        const int o = M::OwnerOnGrid(std::atoi(argv[2]), std::atoi(argv[3]), std::atoll(argv[4]), std::atoll(argv[5]), std::atoll(argv[6]), std::atoll(argv[7]), lr, lc);
        std::cout << "owner " << o << ' ' << lr << ' ' << lc << '\n';
        int64_t s, l;
        M::BlockRange(std::atoll(argv[4]), std::atoi(argv[2]), std::atoi(argv[2]) - 1, s, l);
        std::cout << "lastblock " << s << ' ' << l << '\n';
        return 0;
    }
    if (mode == "fdv" && argc > 4) {
        const int64_t glen = std::atoll(argv[2]);
        const int pr = std::atoi(argv[3]), pc = std::atoi(argv[4]);
        FullyDistLayout<int64_t> L(glen, pr, pc);
        std::vector<int64_t> until, len, owner, lind;
        for (int i = 0; i < pr; ++i)
            for (int j = 0; j < pc; ++j) { until.push_back(L.LengthUntil(i, j)); len.push_back(L.LocLength(i, j)); }
        for (int64_t g = 0; g < glen; ++g) {
            int a, b;
            int64_t l;
            L.Owner(g, a, b, l);
            owner.push_back(a * pc + b);
            lind.push_back(l);
        }
        dump("until", until); dump("len", len); dump("owner", owner); dump("lind", lind);
        return 0;
    }
    std::cerr << "usage: host_logic_test tile|mm|semirings|owner|fdv ...\n";
    return 2;
}
