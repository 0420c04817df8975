// Numeric promotion table of the multiply: T_promote of (matrix value type, panel value type).
// Same pairs as reference include/CombBLAS/promote.h:37-91, restricted to the element types the C ABI carries.
#ifndef CB_PROMOTE_H
#define CB_PROMOTE_H

#include <cstdint>
#include <type_traits>

namespace combblas {

template <class T1, class T2, class Enable = void>
struct promote_trait {};

template <class NT>
struct promote_trait<NT, NT> { typedef NT T_promote; };                       // self promotion is always allowed
template <class NT>
struct promote_trait<NT, bool, typename std::enable_if<!std::is_same<NT, bool>::value>::type> { typedef NT T_promote; };
template <class NT>
struct promote_trait<bool, NT, typename std::enable_if<!std::is_same<NT, bool>::value>::type> { typedef NT T_promote; };

#define CB_DECLARE_PROMOTE(A, B, C) \
    template <> struct promote_trait<A, B> { typedef C T_promote; };
CB_DECLARE_PROMOTE(int64_t, int, int64_t)
CB_DECLARE_PROMOTE(int, int64_t, int64_t)
CB_DECLARE_PROMOTE(float, int, float)
CB_DECLARE_PROMOTE(double, int, double)
CB_DECLARE_PROMOTE(int, float, float)
CB_DECLARE_PROMOTE(int, double, double)
CB_DECLARE_PROMOTE(double, int64_t, double)
CB_DECLARE_PROMOTE(int64_t, double, double)
#undef CB_DECLARE_PROMOTE

// user-side extension point with the reference's name (promote.h:56-60; Applications/SpMMError.cpp:28-29 uses it)
#define DECLARE_PROMOTE(A, B, C) \
    template <> struct promote_trait<A, B> { typedef C T_promote; };

// element type -> cb_dtype code of the C ABI
template <class T> struct cb_dtype_of;
template <> struct cb_dtype_of<float> { static const int value = 0; };
template <> struct cb_dtype_of<double> { static const int value = 1; };
template <> struct cb_dtype_of<int32_t> { static const int value = 2; };
template <> struct cb_dtype_of<int64_t> { static const int value = 3; };
template <> struct cb_dtype_of<bool> { static const int value = 4; };
template <> struct cb_dtype_of<uint8_t> { static const int value = 4; };

// std::vector<bool> is bit packed; dense panels and value arrays of bool are stored one byte per element
template <class T> struct cb_storage { typedef T type; };
template <> struct cb_storage<bool> { typedef uint8_t type; };

}  // namespace combblas
#endif
