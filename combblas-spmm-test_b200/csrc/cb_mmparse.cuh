// cb_mmparse.cuh - the number parsers of the device Matrix Market reader (cb_mmio.cu), written so that the same code also
// compiles as plain C++ (tests/emul/mmparse_host.cpp checks it against strtod on the CPU for millions of strings).
//
// Decimal -> double: the digits are collected into a 64-bit mantissa w (at most 19 significant digits) and a decimal exponent q;
// w * 10^q is converted with the Eisel-Lemire algorithm (D. Lemire, "Number parsing at a gigabyte per second", 2021): one or two
// 64 x 64 -> 128-bit multiplications by a truncated 128-bit power of five (cb_pow5_table.inc, tools/make_pow5_table.py), the top
// 55 bits rounded to nearest-even, with the round-to-even tie detected exactly.  The result is the correctly rounded double -
// what strtod, operator>> and the reference's sscanf("%lg") return.  Inputs the conversion does not decide (more than 19
// significant digits, subnormal or overflowing results, the rare inconclusive product outside the safe exponent range, inf / nan /
// hexadecimal) are reported as "hard" and the caller falls back to the host parser.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define CB_HD __host__ __device__ __forceinline__
#else
#define CB_HD inline
#endif

namespace mmparse {

#ifdef __CUDACC__
__constant__ uint64_t kPow5Dev[2 * 651] = {
#include "cb_pow5_table.inc"
};
#endif
static const uint64_t kPow5Host[2 * 651] = {
#include "cb_pow5_table.inc"
};
#if defined(__CUDA_ARCH__)
#define CB_POW5(i) kPow5Dev[i]
#else
#define CB_POW5(i) kPow5Host[i]
#endif

CB_HD bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }
CB_HD bool is_digit(char c) { return c >= '0' && c <= '9'; }

CB_HD void mul64(uint64_t a, uint64_t b, uint64_t* hi, uint64_t* lo) {
#if defined(__CUDA_ARCH__)
    *hi = __umul64hi(a, b);
    *lo = a * b;
#else
    const unsigned __int128 p = (unsigned __int128)a * b;
    *hi = (uint64_t)(p >> 64);
    *lo = (uint64_t)p;
#endif
}
CB_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return __builtin_clzll(x);
#endif
}

// w * 10^q (w != 0) -> bits of the nearest double.  Returns false when this routine does not decide the result.
CB_HD bool decimal_to_double_bits(uint64_t w, int q, uint64_t* bits) {
    if (q < -342 || q > 308) return false;
    int lz = clz64(w);
    w <<= lz;
    const uint64_t t_hi = CB_POW5(2 * (q + 342)), t_lo = CB_POW5(2 * (q + 342) + 1);
    uint64_t upper, lower;
    mul64(w, t_hi, &upper, &lower);
    if ((upper & 0x1FFull) == 0x1FFull) {              // the low 9 bits below the 55 kept could still carry: refine with the low word
        uint64_t s_hi, s_lo;
        mul64(w, t_lo, &s_hi, &s_lo);
        lower += s_hi;
        if (s_hi > lower) ++upper;
        (void)s_lo;
    }
    if (lower == 0xFFFFFFFFFFFFFFFFull && (q < -27 || q > 55)) return false;       // inconclusive truncated product
    const int upperbit = (int)(upper >> 63);
    uint64_t mantissa = upper >> (upperbit + 9);
    const int power2 = (int)(((int64_t)(152170 + 65536) * q) >> 16) + 63 + upperbit - lz + 1023;
    if (power2 <= 0) return false;                      // subnormal or zero: left to the host
    if (lower <= 1 && q >= -4 && q <= 23 && (mantissa & 3) == 1 && (mantissa << (upperbit + 9)) == upper) mantissa &= ~1ull;   // exact tie -> even
    mantissa += mantissa & 1;
    mantissa >>= 1;
    int p2 = power2;
    if (mantissa >= (2ull << 52)) { mantissa = 1ull << 52; ++p2; }
    mantissa &= ~(1ull << 52);
    if (p2 >= 0x7FF) return false;                      // overflow to infinity: left to the host
    *bits = mantissa | ((uint64_t)p2 << 52);
    return true;
}

// status of a token: 0 parsed, 1 nothing there, 2 hard (the caller's fallback decides)
CB_HD int parse_int(const char* p, const char* e, const char** next, long long* out) {
    while (p < e && is_space(*p)) ++p;
    bool neg = false;
    if (p < e && (*p == '-' || *p == '+')) { neg = *p == '-'; ++p; }
    if (p >= e || !is_digit(*p)) return 1;
    unsigned long long v = 0;
    int nd = 0;
    while (p < e && is_digit(*p)) {
        if (++nd > 18) return 2;
        v = v * 10ull + (unsigned long long)(*p - '0');
        ++p;
    }
    *out = neg ? -(long long)v : (long long)v;
    *next = p;
    return 0;
}

CB_HD int parse_double_bits(const char* p, const char* e, uint64_t* bits) {
    while (p < e && is_space(*p)) ++p;
    if (p >= e) return 1;
    bool neg = false;
    if (*p == '-' || *p == '+') { neg = *p == '-'; ++p; }
    uint64_t mant = 0;
    int sig = 0, e10 = 0, ndig = 0;
    bool nonzero = false;
    while (p < e && is_digit(*p)) {
        ++ndig;
        if (*p != '0' || nonzero) {
            nonzero = true;
            if (++sig > 19) return 2;
            mant = mant * 10ull + (uint64_t)(*p - '0');
        }
        ++p;
    }
    if (p < e && *p == '.') {
        ++p;
        while (p < e && is_digit(*p)) {
            ++ndig;
            if (*p != '0' || nonzero) {
                nonzero = true;
                if (++sig > 19) return 2;
                mant = mant * 10ull + (uint64_t)(*p - '0');
            }
            --e10;
            ++p;
        }
    }
    if (ndig == 0) return 2;                            // inf, nan, a lone sign or dot
    if (p < e && (*p == 'e' || *p == 'E')) {
        ++p;
        bool eneg = false;
        if (p < e && (*p == '-' || *p == '+')) { eneg = *p == '-'; ++p; }
        if (p >= e || !is_digit(*p)) return 2;
        int ex = 0;
        while (p < e && is_digit(*p)) {
            ex = ex * 10 + (*p - '0');
            if (ex > 10000) return 2;
            ++p;
        }
        e10 += eneg ? -ex : ex;
    }
    if (p < e && !is_space(*p)) return 2;               // characters glued to the number (hex floats, suffixes)
    const uint64_t sign = neg ? (1ull << 63) : 0;
    if (mant == 0) { *bits = sign; return 0; }
    uint64_t b;
    if (!decimal_to_double_bits(mant, e10, &b)) return 2;
    *bits = b | sign;
    return 0;
}

}  // namespace mmparse
