// sift::Matrix<T> — the reference's [octave][element] container (matrix.hpp:22-103): u16 x u16 grid,
// element (x, y) stored at x*height + y, storage shared between copies (the reference keeps a
// shared_ptr<T>, so copying a Matrix aliases it).  Host-side container only.
#ifndef SIFT_MATRIX_HPP
#define SIFT_MATRIX_HPP

#include <cassert>
#include <memory>
#include <vector>

#include "core.hpp"

namespace sift {

template <typename T>
class Matrix {
   public:
    Matrix() = default;
    explicit Matrix(u16_t width, u16_t height, const T& def = T())
        : width_(width), height_(height), cells_(std::make_shared<std::vector<T>>((size_t)width * height, def)) {
        assert(width > 0 && height > 0);
    }

    u16_t width() const { return width_; }
    u16_t height() const { return height_; }

    T& operator[](const Point<u16_t, u16_t>& at) { return cell(at.x, at.y); }
    const T& operator[](const Point<u16_t, u16_t>& at) const { return cell(at.x, at.y); }
    T& operator()(u16_t x, u16_t y) { return cell(x, y); }
    const T& operator()(u16_t x, u16_t y) const { return cell(x, y); }

   private:
    T& cell(u16_t x, u16_t y) const {
        assert(x < width_ && y < height_);
        return (*cells_)[(size_t)x * height_ + y];
    }
    u16_t width_ = 0, height_ = 0;
    std::shared_ptr<std::vector<T>> cells_;
};

}  // namespace sift
#endif  // SIFT_MATRIX_HPP
