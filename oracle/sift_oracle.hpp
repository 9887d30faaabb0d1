// ORACLE — TEST INFRASTRUCTURE ONLY (see vigra_linalg.hpp header).  PARITY PINNED: this CPU restatement of
// snowiow/SIFT's hot path (loaded image -> keypoints + descriptors) is bit-identical, stage by stage, to (1) the reference's
// own sift.cpp + algorithms.cpp compiled unmodified over Vigra stand-in headers (oracle/_ref, tests/test_ref_pin.py) and
// (2) the reference's SHIPPED EXECUTABLE bin/arch_x64/sift — the author's GCC 7 build with the real Vigra 1.11 compiled in —
// whose own functions are called in place (oracle/refbin_run.cpp, tests/test_refbin_pin.py), on every case that literal
// build finishes in minutes; digests of both are committed under tests/golden/.  The reference itself holds no tests or
// golden vectors.  Every function cites the reference lines it follows (file:line, relative to the reference
// tree) and, where the arithmetic lives in Vigra, the SURVEY.md Appendix A item.  Single-threaded, fp32 with the reference's
// few double detours, built -O3 without -march so that no FMA contraction can occur (the shipped binary uses mulss/addss).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace oracle {

// vigra::PreconditionViolation stand-in (std::exception-derived).
struct Precondition : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// vigra::MultiArray<2, float> stand-in: (x, y) -> y*w + x, x contiguous.
struct Image {
    int w = 0, h = 0;
    std::vector<float> px;
    Image() {}
    Image(int w_, int h_) : w(w_), h(h_), px((size_t)w_ * (size_t)h_, 0.0f) {}
    float& operator()(int x, int y) { return px[(size_t)y * (size_t)w + (size_t)x]; }
    float operator()(int x, int y) const { return px[(size_t)y * (size_t)w + (size_t)x]; }
};

// --- alg:: primitives -------------------------------------------------------------------
std::vector<float> gaussian_taps(float sigma, int* radius);           // Vigra Kernel1D::initGaussian
std::vector<float> gaussian_taps_d(double std_dev, int* radius);      // same, with the double argument Vigra takes
// every line of one separable pass (axis 0 = along x, 1 = along y) with Vigra's reflect border (A.2)
void convolve_lines_reflect(const float* src, long ssx, long ssy, float* dst, long dsx, long dsy, int w, int h,
                            const float* taps, int r, int axis);
Image convolve_with_gauss(const Image& img, float sigma);             // algorithms.cpp:10-22
Image resize_no_interpolation(const Image& src, int nw, int nh);      // Vigra resizeImageNoInterpolation
std::vector<int> resize_index_map(int n_old, int n_new);              // the per-line index walk
Image reduce_to_next_level(const Image& img, float sigma);            // algorithms.cpp:24-36
Image increase_to_next_level(const Image& img, float sigma);          // algorithms.cpp:38-49
Image dog(const Image& lower, const Image& higher);                   // algorithms.cpp:52-64
void fo_derivative(const Image* const d[3], int x, int y, float out[3]);        // algorithms.cpp:66-77
void so_derivative(const Image* const d[3], int x, int y, float out[3][3]);     // algorithms.cpp:79-106
float gradient_magnitude(const Image& img, int x, int y);             // algorithms.cpp:108-111
float gradient_orientation(const Image& img, int x, int y);           // algorithms.cpp:113-116
float vertex_parabola(uint16_t lx, float ly, uint16_t px, float py, uint16_t rx, float ry);  // algorithms.cpp:153-178
void normalize_vector(std::vector<float>& v);                         // algorithms.cpp:210-223
std::vector<float> find_peaks(const float histo[36]);                 // sift.cpp:220-286 (sorted, deduplicated)
bool inverse3(const float a[9], float out[9]);                        // row-major 3x3 wrappers over la::
bool linear_solve3(const float a[9], const float b[3], float out[3]);

// std::sort(cmpByFilter) permutation of a flag sequence (sift.cpp:37, :49; interestpoint.hpp:57-62).
// order[i] = original index of the element that ends up at position i.
void sort_by_filter_order(const uint8_t* filtered, size_t n, uint32_t* order);

// --- result / pipeline --------------------------------------------------------------------
struct KeyPoint {  // interestpoint.hpp:13-63
    float scale = 0.0f;
    uint16_t octave = 0, index = 0;
    bool filtered = false;
    uint16_t x = 0, y = 0;
    float orientation = 0.0f;
    std::vector<float> descriptors;
};

struct Params {  // sift.hpp:66-71 (ctor order: dogsPerEpoch, octaves, sigma, k, subpixel)
    uint16_t dogs_per_epoch = 3;
    uint16_t octaves = 3;
    float sigma = 1.6f;
    float k = 1.41421356237309504880f;  // (float)std::sqrt(2)
    bool subpixel = false;
    // oracle-only switches
    bool literal = false;  // keep the reference's per-candidate image copies, per-keypoint full blur and dead blur
    bool strict = false;   // propagate the dead-blur precondition (sift.cpp:184) like the reference does
};

struct Level {
    float scale = 0.0f;
    Image img;
};

struct Candidate {  // one emitted extremum, canonical (e, i, x, y) order
    uint16_t x, y, octave, index;
    float scale;
    bool filtered;
};

class Sift {
   public:
    explicit Sift(const Params& p) : prm(p) {}
    // sift.cpp:19-57.  Overwrites img with the 2x image when subpixel (sift.cpp:21).
    std::vector<KeyPoint> calculate(Image& img);

    // stage dumps of the last calculate()
    int octaves() const { return prm.octaves; }
    int n_gauss() const { return prm.dogs_per_epoch + 1; }
    int n_dogs() const { return prm.dogs_per_epoch; }
    const Level& gauss(int o, int i) const { return gaussians[(size_t)(o * n_gauss() + i)]; }
    const Level& dogl(int o, int i) const { return dogs[(size_t)(o * n_dogs() + i)]; }
    const std::vector<Candidate>& candidates() const { return cands; }  // after elimination
    const std::vector<KeyPoint>& survivors() const { return after_first_trim; }
    void nearest_gaussian(float scale, int* o, int* i) const;  // sift.cpp:205-218

    // stage functions, public so tests can drive them in isolation
    void create_dogs(const Image& img);                                         // sift.cpp:381-417
    void inject_dogs(const Image& d0, const Image& d1, const Image& d2) {        // test hook: one octave, three DoGs
        dogs.assign(3, Level());
        dogs[0].img = d0; dogs[1].img = d1; dogs[2].img = d2;
    }
    static void find_scale_space_extrema(const std::vector<Level>& dogs, int octaves, int n_dogs,
                                         std::vector<KeyPoint>& out);           // sift.cpp:348-379
    void eliminate_edge_responses(std::vector<KeyPoint>& pts) const;            // sift.cpp:288-346
    static void cleanup(std::vector<KeyPoint>& pts);                            // sift.cpp:37-42
    void create_gradient_pyramids();                                            // sift.cpp:130-160
    void orientation_assignment(std::vector<KeyPoint>& pts);                    // sift.cpp:163-203
    void create_descriptors(std::vector<KeyPoint>& pts);                        // sift.cpp:60-110

    Params prm;

   private:
    std::vector<Level> gaussians, dogs;
    std::vector<Image> magnitudes, orientations;
    std::vector<Candidate> cands;
    std::vector<KeyPoint> after_first_trim;
};

// main.cpp:78-89 result text.
std::string format_results(const std::vector<KeyPoint>& pts);

}  // namespace oracle
