// ORACLE — TEST INFRASTRUCTURE ONLY.  Stand-in for <vigra/multi_array.hxx> (Vigra 1.11, absent from this
// image) so that the reference's own, UNMODIFIED sift.cpp / algorithms.cpp compile into oracle/_ref/.
// Only what those two files use is provided: Shape2, MultiArrayView<2,T> (strided, x = first index =
// contiguous, half-open subarray with negative-coordinate wrap), MultiArray<2,T>
// (owning, zero-initialised, deep-copying), PreconditionViolation (A.7).
//
// -DREF_SHIM_FAST (oracle/_ref/libref_fast.so) makes MultiArray copies copy-on-write; observable results
// are identical (tests/test_ref_pin.py compares the two builds), the reference's O(pixels) copies per
// candidate (sift.cpp:297-298) just stop costing time.  Without it (libref.so) every copy is deep, as in Vigra.
#ifndef REF_SHIM_VIGRA_MULTI_ARRAY_HXX
#define REF_SHIM_VIGRA_MULTI_ARRAY_HXX

#include <algorithm>
#include <array>
#include <cmath>
#include <cstddef>
#include <functional>
#include <iostream>
#include <set>
#include <vector>
#include <cstdlib>
#include <cstring>
#include <new>
#include <type_traits>
#include <exception>
#include <memory>
#include <string>

namespace vigra {

typedef std::ptrdiff_t MultiArrayIndex;

class ContractViolation : public std::exception {
   public:
    explicit ContractViolation(const std::string& m) : what_(m) {}
    ~ContractViolation() throw() {}
    const char* what() const throw() { return what_.c_str(); }

   private:
    std::string what_;
};
class PreconditionViolation : public ContractViolation {
   public:
    explicit PreconditionViolation(const std::string& m) : ContractViolation("\nPrecondition violation!\n" + m) {}
};
inline void vigra_precondition(bool ok, const char* msg) {
    if (!ok) throw PreconditionViolation(msg);
}

// TinyVector<MultiArrayIndex, 2>
struct Shape2 {
    MultiArrayIndex v[2];
    Shape2() { v[0] = v[1] = 0; }
    Shape2(MultiArrayIndex a, MultiArrayIndex b) { v[0] = a; v[1] = b; }
    MultiArrayIndex& operator[](int i) { return v[i]; }
    const MultiArrayIndex& operator[](int i) const { return v[i]; }
    bool operator==(const Shape2& o) const { return v[0] == o.v[0] && v[1] == o.v[1]; }
    bool operator!=(const Shape2& o) const { return !(*this == o); }
    Shape2 operator-(const Shape2& o) const { return Shape2(v[0] - o.v[0], v[1] - o.v[1]); }
};

template <unsigned int N, class T>
class MultiArray;

template <unsigned int N, class T>
class MultiArrayView {
    static_assert(N == 2, "the shim only provides two-dimensional arrays");

   public:
    typedef T value_type;
    typedef Shape2 difference_type;

    MultiArrayView() : m_ptr(0) {}
    MultiArrayView(const Shape2& shape, const Shape2& stride, T* ptr) : m_shape(shape), m_stride(stride), m_ptr(ptr) {}

    const Shape2& shape() const { return m_shape; }
    MultiArrayIndex shape(int i) const { return m_shape[i]; }
    const Shape2& stride() const { return m_stride; }
    MultiArrayIndex stride(int i) const { return m_stride[i]; }
    MultiArrayIndex width() const { return m_shape[0]; }
    MultiArrayIndex height() const { return m_shape[1]; }
    MultiArrayIndex size() const { return m_shape[0] * m_shape[1]; }
    bool hasData() const { return m_ptr != 0; }
    T* data() const { return m_ptr; }

    T& operator()(MultiArrayIndex x, MultiArrayIndex y) { return m_ptr[x * m_stride[0] + y * m_stride[1]]; }
    const T& operator()(MultiArrayIndex x, MultiArrayIndex y) const { return m_ptr[x * m_stride[0] + y * m_stride[1]]; }

    // half-open [p, q); negative coordinates count from the end (Vigra RelativeToAbsoluteCoordinate)
    MultiArrayView subarray(Shape2 p, Shape2 q) const {
        for (int k = 0; k < 2; ++k) {
            if (p[k] < 0) p[k] += m_shape[k];
            if (q[k] < 0) q[k] += m_shape[k];
        }
        return MultiArrayView(q - p, m_stride, m_ptr + p[0] * m_stride[0] + p[1] * m_stride[1]);
    }

    MultiArrayView transpose() const {
        return MultiArrayView(Shape2(m_shape[1], m_shape[0]), Shape2(m_stride[1], m_stride[0]), m_ptr);
    }

    MultiArrayView& operator*=(const T& rhs) {
        for (MultiArrayIndex y = 0; y < m_shape[1]; ++y)
            for (MultiArrayIndex x = 0; x < m_shape[0]; ++x) (*this)(x, y) *= rhs;
        return *this;
    }

   protected:
    Shape2 m_shape, m_stride;
    T* m_ptr;
};

template <unsigned int N, class T>
class MultiArray : public MultiArrayView<N, T> {
    typedef MultiArrayView<N, T> view_type;

   public:
    MultiArray() : leaked_(false) {}
    explicit MultiArray(const Shape2& shape) : leaked_(false) { allocate(shape, true); }
    MultiArray(MultiArrayIndex w, MultiArrayIndex h) : leaked_(false) { allocate(Shape2(w, h), true); }

    MultiArray(const MultiArray& rhs) : view_type(), leaked_(false) { take(rhs); }
    // Vigra's converting constructor is implicit: views passed where `const MultiArray&` is expected are deep-copied.
    MultiArray(const view_type& rhs) : leaked_(false) { deep_from(rhs); }

    MultiArray& operator=(const MultiArray& rhs) {
        if (this == &rhs) return *this;
        if (leaked_ && this->m_shape == rhs.shape()) {  // copyOrReshape: same shape copies in place, views stay valid
            copy_elements(rhs);
            return *this;
        }
        leaked_ = false;
        take(rhs);
        return *this;
    }
    MultiArray& operator=(const view_type& rhs) {
        if (leaked_ && this->m_shape == rhs.shape()) {
            copy_elements(rhs);
            return *this;
        }
        leaked_ = false;
        deep_from(rhs);
        return *this;
    }

    // writers un-share first; a handed-out view pins the buffer (no later sharing of it)
    T& operator()(MultiArrayIndex x, MultiArrayIndex y) {
        unshare();
        return this->m_ptr[x * this->m_stride[0] + y * this->m_stride[1]];
    }
    const T& operator()(MultiArrayIndex x, MultiArrayIndex y) const {
        return this->m_ptr[x * this->m_stride[0] + y * this->m_stride[1]];
    }
    view_type subarray(Shape2 p, Shape2 q) {
        unshare();
        leaked_ = true;
        return view_type::subarray(p, q);
    }
    view_type subarray(Shape2 p, Shape2 q) const { return view_type::subarray(p, q); }
    view_type writable_view() {
        unshare();
        return view_type(this->m_shape, this->m_stride, this->m_ptr);
    }
    MultiArray& operator*=(const T& rhs) {
        unshare();
        view_type::operator*=(rhs);
        return *this;
    }
    // O(1) adoption of another array's buffer (shim-internal: the memoised blur of REF_SHIM_FAST)
    void share_from(const MultiArray& rhs) {
        leaked_ = false;
        take(rhs);
    }

   private:
    void allocate(const Shape2& shape, bool zero) {
        const std::size_t n = (std::size_t)(shape[0] * shape[1]);
        // calloc: large zero-initialised arrays come as fresh zero pages that cost nothing until touched (the two temporaries
        // of alg::convolveWithGauss are never touched when REF_SHIM_FAST's memoised blur answers)
        static_assert(std::is_trivial<T>::value, "calloc'd storage");
        T* raw = static_cast<T*>(zero ? std::calloc(n ? n : 1, sizeof(T)) : std::malloc((n ? n : 1) * sizeof(T)));
        if (!raw) throw std::bad_alloc();
        buf_ = std::shared_ptr<T>(raw, [](T* q) { std::free(q); });
        this->m_shape = shape;
        this->m_stride = Shape2(1, shape[0]);
        this->m_ptr = buf_.get();
    }
    void copy_elements(const view_type& rhs) {
        unshare();
        for (MultiArrayIndex y = 0; y < this->m_shape[1]; ++y)
            for (MultiArrayIndex x = 0; x < this->m_shape[0]; ++x)
                this->m_ptr[x * this->m_stride[0] + y * this->m_stride[1]] = rhs(x, y);
    }
    void deep_from(const view_type& rhs) {
        if (!rhs.hasData()) {
            buf_.reset();
            this->m_shape = this->m_stride = Shape2();
            this->m_ptr = 0;
            return;
        }
        allocate(rhs.shape(), false);
        if (rhs.stride(0) == 1 && rhs.stride(1) == rhs.shape(0))
            std::memcpy(this->m_ptr, rhs.data(), sizeof(T) * (std::size_t)rhs.size());
        else
            for (MultiArrayIndex y = 0; y < this->m_shape[1]; ++y)
                for (MultiArrayIndex x = 0; x < this->m_shape[0]; ++x) this->m_ptr[x + y * this->m_shape[0]] = rhs(x, y);
    }
    void take(const MultiArray& rhs) {
#ifdef REF_SHIM_FAST
        if (!rhs.leaked_) {
            buf_ = rhs.buf_;
            this->m_shape = rhs.m_shape;
            this->m_stride = rhs.m_stride;
            this->m_ptr = rhs.m_ptr;
            return;
        }
#endif
        deep_from(rhs);
    }
    void unshare() {
#ifdef REF_SHIM_FAST
        if (buf_ && buf_.use_count() > 1) {
            view_type old(this->m_shape, this->m_stride, this->m_ptr);
            std::shared_ptr<T> keep = buf_;
            deep_from(old);
        }
#endif
    }

    std::shared_ptr<T> buf_;
    bool leaked_;
};

}  // namespace vigra
#endif
