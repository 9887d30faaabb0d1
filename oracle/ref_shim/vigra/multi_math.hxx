// ORACLE — TEST INFRASTRUCTURE ONLY.  Stand-in for <vigra/multi_math.hxx>: the element-wise comparison
// expressions and any() of sift.cpp:366-371 (any(expr) is true iff an element is).
#ifndef REF_SHIM_VIGRA_MULTI_MATH_HXX
#define REF_SHIM_VIGRA_MULTI_MATH_HXX

#include "multi_array.hxx"

namespace vigra {
namespace multi_math {

template <class T, bool GREATER>
struct CompareWithScalar {
    MultiArrayView<2, T> view;
    T scalar;
};

template <class T>
inline CompareWithScalar<T, true> operator>(const MultiArrayView<2, T>& v, const T& s) {
    CompareWithScalar<T, true> e = {v, s};
    return e;
}
template <class T>
inline CompareWithScalar<T, false> operator<(const MultiArrayView<2, T>& v, const T& s) {
    CompareWithScalar<T, false> e = {v, s};
    return e;
}

template <class T, bool GREATER>
inline bool any(const CompareWithScalar<T, GREATER>& e) {
    bool res = false;
    for (MultiArrayIndex y = 0; y < e.view.shape(1); ++y)
        for (MultiArrayIndex x = 0; x < e.view.shape(0); ++x)
            res = res || (GREATER ? (e.view(x, y) > e.scalar) : (e.view(x, y) < e.scalar));
    return res;
}

}  // namespace multi_math
}  // namespace vigra
#endif
