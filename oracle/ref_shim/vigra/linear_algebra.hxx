// ORACLE — TEST INFRASTRUCTURE ONLY.  Stand-in for <vigra/linear_algebra.hxx>: inverse(), linearSolve()
// (default method "QR") and dot() with Vigra's signatures, backed by the restated Householder-QR routines of
// oracle/vigra_linalg.hpp.  Call sites: sift.cpp:306,311,322, algorithms.cpp:175.
// The arithmetic inside these three routines is the part of the pipeline that stays RESTATED (not compiled
// from reference-owned source): Vigra itself is not in the reference tree.
#ifndef REF_SHIM_VIGRA_LINEAR_ALGEBRA_HXX
#define REF_SHIM_VIGRA_LINEAR_ALGEBRA_HXX

#include <string>

#include "../../vigra_linalg.hpp"
#include "matrix.hxx"

namespace vigra {
namespace linalg {
namespace shim_detail {
template <class T>
inline oracle::la::MV mv(const MultiArrayView<2, T>& v) {
    static_assert(sizeof(T) == sizeof(float), "fp32 only");
    oracle::la::MV m;
    m.p = const_cast<float*>(v.data());
    m.s0 = v.stride(0);
    m.s1 = v.stride(1);
    m.n0 = v.shape(0);
    m.n1 = v.shape(1);
    return m;
}
}  // namespace shim_detail

// bool inverse(const MultiArrayView<2,T,C1>& v, MultiArrayView<2,T,C2>& res)
template <class T>
inline bool inverse(const MultiArrayView<2, T>& v, MultiArrayView<2, T>& res) {
    const MultiArrayIndex n = columnCount(v), m = rowCount(v);
    vigra_precondition(n <= m, "inverse(): Matrix must have at least as many rows as columns.");
    vigra_precondition(n == rowCount(res) && m == columnCount(res), "inverse(): shape of output matrix must be the transpose of the input matrix' shape.");
    vigra_precondition(m == n, "inverse(): the shim only provides the square case the reference uses.");
    return oracle::la::inverse(shim_detail::mv(v), shim_detail::mv(res));
}
template <class T>
inline bool inverse(const MultiArrayView<2, T>& v, MultiArray<2, T>& res) {
    MultiArrayView<2, T> w = res.writable_view();
    return inverse(v, w);
}

// bool linearSolve(const MultiArrayView<2,T,C1>& A, const MultiArrayView<2,T,C2>& b, MultiArrayView<2,T,C3> res, std::string method = "QR")
template <class T>
inline bool linearSolve(const MultiArrayView<2, T>& A, const MultiArrayView<2, T>& b, MultiArrayView<2, T> res,
                        std::string method = "QR") {
    const MultiArrayIndex n = columnCount(A), m = rowCount(A);
    vigra_precondition(n <= m, "linearSolve(): Coefficient matrix A must have at least as many rows as columns.");
    vigra_precondition(n == rowCount(res) && m == rowCount(b) && columnCount(b) == columnCount(res),
                       "linearSolve(): matrix shape mismatch.");
    vigra_precondition(method == "QR" || method == "qr", "linearSolve(): the shim only provides the default method.");
    return oracle::la::linear_solve(shim_detail::mv(A), shim_detail::mv(b), shim_detail::mv(res));
}
template <class T>
inline bool linearSolve(const MultiArrayView<2, T>& A, const MultiArrayView<2, T>& b, MultiArray<2, T>& res,
                        std::string method = "QR") {
    return linearSolve(A, b, res.writable_view(), method);
}

// NormTraits<float>::SquaredNormType dot(x, y): fp32, index order
template <class T>
inline T dot(const MultiArrayView<2, T>& x, const MultiArrayView<2, T>& y) {
    T ret = T();
    if (y.shape(1) == 1) {
        const MultiArrayIndex size = y.shape(0);
        if (x.shape(0) == 1 && x.shape(1) == size)
            for (MultiArrayIndex i = 0; i < size; ++i) ret += x(0, i) * y(i, 0);
        else if (x.shape(1) == 1 && x.shape(0) == size)
            for (MultiArrayIndex i = 0; i < size; ++i) ret += x(i, 0) * y(i, 0);
        else
            vigra_precondition(false, "dot(): wrong matrix shapes.");
    } else if (y.shape(0) == 1) {
        const MultiArrayIndex size = y.shape(1);
        if (x.shape(0) == 1 && x.shape(1) == size)
            for (MultiArrayIndex i = 0; i < size; ++i) ret += x(0, i) * y(0, i);
        else if (x.shape(1) == 1 && x.shape(0) == size)
            for (MultiArrayIndex i = 0; i < size; ++i) ret += x(i, 0) * y(0, i);
        else
            vigra_precondition(false, "dot(): wrong matrix shapes.");
    } else
        vigra_precondition(false, "dot(): wrong matrix shapes.");
    return ret;
}

}  // namespace linalg
}  // namespace vigra
#endif
