// ORACLE — TEST INFRASTRUCTURE ONLY.  Stand-in for <vigra/convolution.hxx> (which also brings in
// resizeimage.hxx): Kernel1D<float>::initGaussian, separableConvolveX/Y with BORDER_TREATMENT_REFLECT and
// resizeImageNoInterpolation, as called from algorithms.cpp:13-19,33,46.  Backed by the restated routines of
// oracle/sift_oracle.cpp — like the QR code these stay RESTATED.
//
// -DREF_SHIM_FAST adds a small per-thread memo (exact: keyed on the full source contents and the taps) so the
// reference's full-image blur per keypoint (sift.cpp:87) of an unchanged image is computed once.
#ifndef REF_SHIM_VIGRA_CONVOLUTION_HXX
#define REF_SHIM_VIGRA_CONVOLUTION_HXX

#include <vector>

#include "../../sift_oracle.hpp"
#include "multi_array.hxx"

namespace vigra {

template <class T>
class Kernel1D {
   public:
    Kernel1D() : taps_(1, T(1)), radius_(0) {}
    void initGaussian(double std_dev) {
        try {
            taps_ = oracle::gaussian_taps_d(std_dev, &radius_);
        } catch (const oracle::Precondition& e) {
            throw PreconditionViolation(e.what());
        }
    }
    int left() const { return -radius_; }
    int right() const { return radius_; }
    const std::vector<T>& taps() const { return taps_; }  // taps()[j + radius] = kernel[j]

   private:
    std::vector<T> taps_;
    int radius_;
};

namespace shim_detail {
#ifdef REF_SHIM_FAST
struct BlurMemo {
    int axis;
    std::vector<float> taps;
    MultiArray<2, float> src, dst;
};
inline std::vector<BlurMemo>& blur_memo() {
    static thread_local std::vector<BlurMemo> m;
    return m;
}
#endif

// axis 0: lines along x (separableConvolveX), axis 1: lines along y
inline void convolve_axis(const MultiArrayView<2, float>& src, MultiArray<2, float>* dst_owner, MultiArrayView<2, float> dst,
                          const Kernel1D<float>& k, int axis) {
    const MultiArrayIndex w = src.shape(0), h = src.shape(1);
    vigra_precondition(src.shape() == dst.shape(), "separableConvolve(): shape mismatch between input and output.");
    const int r = k.right();
    vigra_precondition((axis == 0 ? w : h) >= r + 1,
                       axis == 0 ? "separableConvolveX(): kernel longer than line\n" : "separableConvolveY(): kernel longer than line\n");
#ifdef REF_SHIM_FAST
    const bool contiguous = src.stride(0) == 1 && src.stride(1) == w && dst_owner != 0 && w * h >= 4096;
    if (contiguous) {
        std::vector<BlurMemo>& memo = blur_memo();
        for (std::size_t i = 0; i < memo.size(); ++i) {
            BlurMemo& e = memo[i];
            if (e.axis == axis && e.src.shape() == src.shape() && e.taps == k.taps() &&
                std::memcmp(e.src.data(), src.data(), sizeof(float) * (std::size_t)(w * h)) == 0) {
                dst_owner->share_from(e.dst);
                return;
            }
        }
    }
#endif
    oracle::convolve_lines_reflect(src.data(), src.stride(0), src.stride(1), dst.data(), dst.stride(0), dst.stride(1), (int)w,
                                   (int)h, k.taps().data(), r, axis);
#ifdef REF_SHIM_FAST
    if (contiguous) {
        std::vector<BlurMemo>& memo = blur_memo();
        if (memo.size() >= 8) memo.erase(memo.begin());
        memo.push_back(BlurMemo());
        BlurMemo& e = memo.back();
        e.axis = axis;
        e.taps = k.taps();
        e.src = MultiArray<2, float>(src);
        e.dst.share_from(*dst_owner);
    }
#endif
}
}  // namespace shim_detail

// void separableConvolveX(MultiArrayView<2,T1,S1> const& src, MultiArrayView<2,T2,S2> dest, Kernel1D<T> const& kernel)
inline void separableConvolveX(const MultiArrayView<2, float>& src, MultiArrayView<2, float> dest, const Kernel1D<float>& k) {
    shim_detail::convolve_axis(src, 0, dest, k, 0);
}
inline void separableConvolveX(const MultiArrayView<2, float>& src, MultiArray<2, float>& dest, const Kernel1D<float>& k) {
    shim_detail::convolve_axis(src, &dest, dest.writable_view(), k, 0);
}
inline void separableConvolveY(const MultiArrayView<2, float>& src, MultiArrayView<2, float> dest, const Kernel1D<float>& k) {
    shim_detail::convolve_axis(src, 0, dest, k, 1);
}
inline void separableConvolveY(const MultiArrayView<2, float>& src, MultiArray<2, float>& dest, const Kernel1D<float>& k) {
    shim_detail::convolve_axis(src, &dest, dest.writable_view(), k, 1);
}

// resizeImageNoInterpolation(src, dest): nearest neighbour, per line the accumulated-double index walk (A.3)
inline void resizeImageNoInterpolation(const MultiArrayView<2, float>& src, MultiArrayView<2, float> dest) {
    const MultiArrayIndex w = src.shape(0), h = src.shape(1), wn = dest.shape(0), hn = dest.shape(1);
    vigra_precondition(w > 1 && h > 1, "resizeImageNoInterpolation(): Source image too small.\n");
    vigra_precondition(wn > 1 && hn > 1, "resizeImageNoInterpolation(): Destination image too small.\n");
    const std::vector<int> mx = oracle::resize_index_map((int)w, (int)wn), my = oracle::resize_index_map((int)h, (int)hn);
    for (MultiArrayIndex y = 0; y < hn; ++y)
        for (MultiArrayIndex x = 0; x < wn; ++x) dest(x, y) = src(mx[(std::size_t)x], my[(std::size_t)y]);
}
inline void resizeImageNoInterpolation(const MultiArrayView<2, float>& src, MultiArray<2, float>& dest) {
    resizeImageNoInterpolation(src, dest.writable_view());
}

}  // namespace vigra
#endif
