// Semiring structs with the reference's static-function contract (id / add / multiply / axpy / returnedSAID),
// plus the compile-time map onto the opcodes of the C ABI.
//
// Behaviour follows reference include/CombBLAS/Semirings.h: inf_plus :40-47, SelectMaxSRing<bool,T> :191-210,
// PlusTimesSRing :212-232, MinPlusSRing :235-255.  The host functions exist so user code that calls SR::add etc.
// directly keeps compiling; the multiply itself runs the same functors on the GPU (csrc/cb_spmm_kernel.cuh).
// A semiring that is not in the table below is a COMPILE-TIME error in SpMM<SR>() - there is no CPU fallback.
#ifndef CB_SEMIRINGS_H
#define CB_SEMIRINGS_H

#include <algorithm>
#include <limits>
#include "combblas_b200.h"
#include "promote.h"

namespace combblas {

template <typename T>
inline T inf_plus(const T& a, const T& b) {
    const T inf = std::numeric_limits<T>::max();
    if (a == inf || b == inf) return inf;
    return a + b;
}

template <class T1, class T2>
struct PlusTimesSRing {
    typedef typename promote_trait<T1, T2>::T_promote T_promote;
    static T_promote id() { return 0; }
    static bool returnedSAID() { return false; }
    static T_promote add(const T_promote& a, const T_promote& b) { return a + b; }
    static T_promote multiply(const T1& a, const T2& b) { return static_cast<T_promote>(a) * static_cast<T_promote>(b); }
    static void axpy(T1 a, const T2& x, T_promote& y) { y += a * x; }
};

template <class T1, class T2>
struct MinPlusSRing {
    typedef typename promote_trait<T1, T2>::T_promote T_promote;
    static T_promote id() { return std::numeric_limits<T_promote>::max(); }
    static bool returnedSAID() { return false; }
    static T_promote add(const T_promote& a, const T_promote& b) { return std::min(a, b); }
    static T_promote multiply(const T1& a, const T2& b) { return inf_plus<T_promote>(static_cast<T_promote>(a), static_cast<T_promote>(b)); }
    static void axpy(T1 a, const T2& x, T_promote& y) { y = std::min(y, multiply(a, x)); }
};

template <class T1, class T2>
struct SelectMaxSRing;            // only the <bool, T> form used for BFS-style iterations crosses the ABI
template <class T2>
struct SelectMaxSRing<bool, T2> {
    typedef T2 T_promote;
    static T_promote id() { return -1; }
    static bool returnedSAID() { return false; }
    static T_promote add(const T_promote& a, const T_promote& b) { return std::max(a, b); }
    static T_promote multiply(const bool&, const T2& b) { return b; }
    static void axpy(bool, const T2& x, T_promote& y) { y = std::max(y, x); }
};

// ---- semiring -> opcode.  value == true only for the combinations libcombblas_b200 implements.
template <class SR>
struct semiring_traits { static const bool supported = false; };

template <class T>
struct semiring_traits<PlusTimesSRing<T, T>> {
    static const bool supported = true;
    static const int op = std::is_same<T, bool>::value ? CB_OR_AND : CB_PLUS_TIMES;      // bool+bool = OR, bool*bool = AND
};
template <class T>
struct semiring_traits<PlusTimesSRing<bool, T>> { static const bool supported = true; static const int op = CB_PLUS_TIMES; };
template <>
struct semiring_traits<PlusTimesSRing<bool, bool>> { static const bool supported = true; static const int op = CB_OR_AND; };
template <class T>
struct semiring_traits<MinPlusSRing<T, T>> { static const bool supported = !std::is_same<T, bool>::value; static const int op = CB_MIN_PLUS; };
template <class T>
struct semiring_traits<SelectMaxSRing<bool, T>> { static const bool supported = !std::is_same<T, bool>::value; static const int op = CB_MAX_SEL2ND; };

}  // namespace combblas
#endif
