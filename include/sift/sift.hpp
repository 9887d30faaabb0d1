// sift::Sift — the reference's pipeline class surface (sift.hpp:17-78) over the CUDA C ABI
// (include/sift_gpu.h).  Same constructor (argument order dogsPerEpoch, octaves, sigma, k, subpixel;
// same defaults), same public `const bool subpixel`, same calculate() contract:
//   * returns the interest points by value, in the reference's vector order, descriptors attached;
//   * overwrites `img` with the 2x upsampled image when subpixel (sift.cpp:21);
//   * violated preconditions surface as a std::exception (sift::PreconditionViolation, where the
//     reference lets vigra::PreconditionViolation escape; caller pattern main.cpp:43,90-92);
//   * octaves == 0 or dogsPerEpoch < 3 trip an assert, like sift.cpp:382-383;
//   * one object is not re-entrant (the reference mutates _gaussians etc., sift.hpp:46-56).
// Extensions (not in the reference): calculateBatch() for many independent images per call,
// device selection, and writeResults() = the text writer of main.cpp:78-89.
#ifndef SIFT_SIFT_HPP
#define SIFT_SIFT_HPP

#include <cmath>
#include <string>
#include <vector>

#include "core.hpp"
#include "interestpoint.hpp"
#include "matrix.hpp"

struct sift_gpu_ctx;

namespace sift {

class Sift {
   public:
    const bool subpixel;

    explicit Sift(u16_t dogsPerEpoch = 3, u16_t octaves = 3, f32_t sigma = 1.6, f32_t k = std::sqrt(2), bool subpixel = false);
    ~Sift();
    Sift(const Sift&) = delete;
    Sift& operator=(const Sift&) = delete;

    std::vector<InterestPoint> calculate(Image& img);

    // --- extensions ---
    std::vector<std::vector<InterestPoint>> calculateBatch(std::vector<Image>& imgs);
    void setDevice(int device) { device_ = device; }
    void setFlags(unsigned flags) { flags_ = flags; }  // SIFT_GPU_FLAG_* (default: STRICT); takes effect at the next context creation
    void setMaxBatch(int n) { max_batch_ = n; }

   private:
    void ensureContext(std::ptrdiff_t w, std::ptrdiff_t h, int batch);
    const f32_t _sigma;
    const f32_t _k;
    const u16_t _dogsPerEpoch;
    const u16_t _octaves;
    sift_gpu_ctx* ctx_ = nullptr;
    std::ptrdiff_t ctx_w_ = 0, ctx_h_ = 0;
    int ctx_batch_ = 0, device_ = 0, max_batch_ = 1;
    unsigned flags_ = 0x2u;  // SIFT_GPU_FLAG_STRICT: like the reference, calculate() throws where the dead 16x16 blur of sift.cpp:184 would (octaves >= 6)
};

// main.cpp:78-89: header line, then "[x, y]\tscale\torientation\t[d0, d1, ..., ]" per point with default
// ostream float formatting.  The reference's code writes "interstpoints.txt"; its README says "sift.txt".
void writeResults(const std::string& path, const std::vector<InterestPoint>& points);
std::string formatResults(const std::vector<InterestPoint>& points);

}  // namespace sift
#endif  // SIFT_SIFT_HPP
