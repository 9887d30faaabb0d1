// ORACLE — TEST INFRASTRUCTURE ONLY.  Stand-in for <vigra/matrix.hxx>: vigra::linalg::Matrix<T> as the
// reference uses it (sift.cpp:300-336, algorithms.cpp:66-106): (row, column) indexing over column-major
// storage (== MultiArray(x, y) with x contiguous), implicit construction from any 2-D view, *= scalar, transpose().
#ifndef REF_SHIM_VIGRA_MATRIX_HXX
#define REF_SHIM_VIGRA_MATRIX_HXX

#include <cmath>
#include <set>
#include <vector>

#include "multi_array.hxx"

namespace vigra {
namespace linalg {

template <class T>
class Matrix : public MultiArray<2, T> {
    typedef MultiArray<2, T> base;

   public:
    Matrix() {}
    explicit Matrix(const Shape2& shape) : base(shape) {}
    Matrix(MultiArrayIndex rows, MultiArrayIndex cols) : base(Shape2(rows, cols)) {}
    Matrix(const Matrix& rhs) : base(static_cast<const base&>(rhs)) {}
    Matrix(const MultiArrayView<2, T>& rhs) : base(rhs) {}
    Matrix(const MultiArray<2, T>& rhs) : base(rhs) {}
    Matrix& operator=(const Matrix& rhs) {
        base::operator=(static_cast<const base&>(rhs));
        return *this;
    }
    Matrix& operator*=(const T& rhs) {
        base::operator*=(rhs);
        return *this;
    }
    MultiArrayView<2, T> transpose() const { return MultiArrayView<2, T>::transpose(); }
};

template <class T>
inline MultiArrayIndex rowCount(const MultiArrayView<2, T>& m) { return m.shape(0); }
template <class T>
inline MultiArrayIndex columnCount(const MultiArrayView<2, T>& m) { return m.shape(1); }

}  // namespace linalg
using linalg::Matrix;
}  // namespace vigra
#endif
