// Host-side value types of the reference's public surface, Vigra-free.
//   fixed-width aliases   reference types.hpp:4-15 (note u32_t is `unsigned long`, 64-bit on LP64)
//   Point<T,U>            reference point.hpp:11-27
//   Image                 stands in for vigra::MultiArray<2, f32_t>: (x, y) indexing with x contiguous,
//                         width() = shape(0), zero-initialised
//   OctaveElem            reference octaveelem.hpp:12-25
//   PreconditionViolation stands in for vigra::PreconditionViolation (std::exception-derived)
#ifndef SIFT_CORE_HPP
#define SIFT_CORE_HPP

#include <cstddef>
#include <stdexcept>
#include <string>
#include <vector>

using u8_t = unsigned char;
using i8_t = char;
using u16_t = unsigned short int;
using i16_t = short int;
using u32_t = unsigned long int;
using i32_t = long int;
using u64_t = unsigned long long int;
using i64_t = long long int;
using f32_t = float;
using f64_t = double;
using f80_t = long double;

namespace sift {

template <typename T, typename U>
class Point {
   public:
    T x;
    U y;
    Point() = default;
    Point(T x_, U y_) : x(x_), y(y_) {}
};

class PreconditionViolation : public std::runtime_error {
   public:
    explicit PreconditionViolation(const std::string& what) : std::runtime_error(what) {}
};

class Image {
   public:
    Image() = default;
    Image(std::ptrdiff_t width, std::ptrdiff_t height) : w_(width), h_(height), px_((size_t)(width * height), 0.0f) {}
    Image(std::ptrdiff_t width, std::ptrdiff_t height, const f32_t* src) : w_(width), h_(height), px_(src, src + width * height) {}

    std::ptrdiff_t width() const { return w_; }
    std::ptrdiff_t height() const { return h_; }
    std::ptrdiff_t shape(int dim) const { return dim == 0 ? w_ : h_; }
    std::size_t size() const { return px_.size(); }
    f32_t& operator()(std::ptrdiff_t x, std::ptrdiff_t y) { return px_[(size_t)(y * w_ + x)]; }
    const f32_t& operator()(std::ptrdiff_t x, std::ptrdiff_t y) const { return px_[(size_t)(y * w_ + x)]; }
    f32_t* data() { return px_.data(); }
    const f32_t* data() const { return px_.data(); }

   private:
    std::ptrdiff_t w_ = 0, h_ = 0;
    std::vector<f32_t> px_;
};

class OctaveElem {
   public:
    f32_t scale;
    Image img;
    OctaveElem() = default;
};

}  // namespace sift
#endif  // SIFT_CORE_HPP
