// tests/emul/mmparse_host.cpp - the device Matrix Market number parser (csrc/cb_mmparse.cuh) compiled as plain C++ and compared
// with strtod on the CPU: random doubles printed in the formats matrix files use, exact ties, 19-digit mantissas, both ends of
// the exponent range.  Prints "checked N mismatches M hard H".
#include <cinttypes>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include "../../combblas-spmm-test_b200/csrc/cb_mmparse.cuh"

static long checked = 0, mismatches = 0, hard = 0;

static void check(const std::string& s, bool may_be_hard) {
    uint64_t bits = 0;
    const int st = mmparse::parse_double_bits(s.data(), s.data() + s.size(), &bits);
    ++checked;
    if (st == 2) {
        ++hard;
        if (!may_be_hard) { ++mismatches; if (mismatches < 20) std::printf("unexpectedly hard: %s\n", s.c_str()); }
        return;
    }
    const double want = std::strtod(s.c_str(), nullptr);
    uint64_t wb;
    std::memcpy(&wb, &want, 8);
    if (st != 0 || wb != bits) { ++mismatches; if (mismatches < 20) std::printf("mismatch: %s -> %016" PRIx64 " want %016" PRIx64 "\n", s.c_str(), bits, wb); }
}

int main(int argc, char** argv) {
    const long n = argc > 1 ? std::atol(argv[1]) : 2000000;
    std::mt19937_64 rng(12345);
    char buf[128];
    const char* fmts[] = {"%.17g", "%.16g", "%.15g", "%.6e", "%.10E", "%+.3f", "%.9g", "%.12f", "%.18e", "%g"};
    for (long i = 0; i < n; ++i) {
        // random bit patterns of normal doubles of every magnitude, and "ordinary" magnitudes, alternating
        uint64_t b = rng();
        double v;
        if (i & 1) {
            b = (b & ~(0x7FFull << 52)) | ((uint64_t)(1023 - 40 + (int)(rng() % 80)) << 52);
        } else {
            uint64_t ex = 1 + rng() % 2045;
            b = (b & ~(0x7FFull << 52)) | (ex << 52);
        }
        std::memcpy(&v, &b, 8);
        const char* f = fmts[i % 10];
        if (i % 10 == 5 && (std::fabs(v) > 1e15 || std::fabs(v) < 1e-9)) f = "%.17g";     // %f of a huge number has > 19 digits
        if (i % 10 == 7 && (std::fabs(v) >= 1e6 || std::fabs(v) < 1e-9)) f = "%.17g";
        std::snprintf(buf, sizeof buf, f, v);
        check(buf, false);
    }
    // integer mantissas of up to 19 digits with every decimal exponent that keeps the value normal: ties and near-ties included
    for (long i = 0; i < n; ++i) {
        const int nd = 1 + (int)(rng() % 19);
        uint64_t m = rng();
        uint64_t lim = 1;
        for (int d = 0; d < nd; ++d) lim *= 10;
        m %= lim;
        const int q = -300 + (int)(rng() % 590);
        std::snprintf(buf, sizeof buf, "%" PRIu64 "e%d", m, q);
        check(buf, true);                                                          // over/underflowing combinations may be refused
    }
    // exact halfway cases between neighbouring doubles: 2^53 + odd, scaled by powers of two that stay integers below 10^19
    for (int sh = 0; sh < 10; ++sh)
        for (uint64_t k = 1; k < 2000; k += 2) {
            const uint64_t m = ((1ull << 53) + k) << sh;
            std::snprintf(buf, sizeof buf, "%" PRIu64, m);
            check(buf, false);
            std::snprintf(buf, sizeof buf, "%" PRIu64 ".000", m);
            check(buf, true);                                                      // 19 digits + zeros may pass the digit limit
        }
    const char* fixed[] = {"0", "-0", "0.0", "-0.000e5", "1", "-1", "+1", "1.", ".5", "-.5e1", "1e0", "1E+0", "1e-0", "123456789012345678", "1.7976931348623157e308",
                           "2.2250738585072014e-308", "4.9406564584124654e-324", "1e-400", "1e400", "0.1", "0.2", "0.3", "1e22", "1e23", "9007199254740993",
                           "9007199254740992", "18014398509481985", "6.02214076e23", "1.602176634e-19"};
    for (const char* s : fixed) check(s, true);
    std::printf("checked %ld mismatches %ld hard %ld\n", checked, mismatches, hard);
    return mismatches ? 1 : 0;
}
