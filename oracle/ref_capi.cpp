// ORACLE — TEST INFRASTRUCTURE ONLY.  C entry points over the REFERENCE'S OWN sift::Sift, compiled from the
// unmodified /root/reference/{sift.cpp,algorithms.cpp,*.hpp} against the Vigra stand-in headers of
// oracle/ref_shim/ (see oracle/Makefile, target `ref`).  Output: oracle/_ref/libref.so (deep copies, the
// reference's true cost) and oracle/_ref/libref_fast.so (-DREF_SHIM_FAST: copy-on-write arrays + memoised
// blur inside the shim, identical results).  The entry points carry the same names and signatures as
// oracle/capi.cpp so tests/oracle_lib.py drives either library.
//
// What this pins: the reference's 640 lines of control flow, float/double promotions, iteration orders, the
// real std::sort over real sift::InterestPoint objects, the u16 truncation — by the compiler, not by
// transcription.  What stays restated HERE: the Vigra routines behind the shim (Gaussian taps, reflect line
// convolution, nearest-neighbour resize walk, Householder QR inverse / linearSolve) — those are pinned by the
// reference's shipped executable instead (refbin_run.cpp, tests/test_refbin_pin.py) — and the result-text
// writer (main.cpp:78-89; main.cpp needs Boost/OpenCV/Vigra-impex and is not compiled).
#include <algorithm>
#include <array>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <memory>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include <vigra/convolution.hxx>
#include <vigra/linear_algebra.hxx>
#include <vigra/matrix.hxx>
#include <vigra/multi_array.hxx>
#include <vigra/multi_math.hxx>

// The stage dumps (pyramid, candidates with flags) need the reference's private stage functions and the
// `_gaussians` member.  Every header the reference includes is already in (include guards), so the keyword
// swap only touches the reference's own class definitions; sift.cpp itself is compiled untouched.
#define private public
#include "sift.hpp"
#undef private
#include "algorithms.hpp"

namespace {

typedef vigra::MultiArray<2, f32_t> Img;

struct Handle {
    int dpe, octaves;
    float sigma, k;
    bool subpixel;
    std::unique_ptr<sift::Sift> sift;       // the object calculate() ran on (keeps _gaussians)
    std::vector<sift::InterestPoint> result;
    Img input;                              // the image calculate() saw after its optional upsample
    // lazily computed stage dumps (a second run of the reference's own private stages)
    bool staged = false;
    sift::Matrix<sift::OctaveElem> dogs;
    std::vector<sift::InterestPoint> cands;      // after _eliminateEdgeResponses, emission order
    std::vector<sift::InterestPoint> survivors;  // after the first std::sort + u16 trim (sift.cpp:37-42)
    std::string error, text;
};

Img wrap(const float* p, int w, int h) {
    Img im(vigra::Shape2(w, h));
    vigra::MultiArrayView<2, f32_t> v = im.writable_view();
    std::memcpy(v.data(), p, sizeof(float) * (size_t)w * (size_t)h);
    return im;
}

void copy_out(const Img& im, float* out) {
    for (vigra::MultiArrayIndex y = 0; y < im.height(); ++y)
        for (vigra::MultiArrayIndex x = 0; x < im.width(); ++x) out[y * im.width() + x] = im(x, y);
}

void stage(Handle* H) {
    if (H->staged) return;
    sift::Sift s((u16_t)H->dpe, (u16_t)H->octaves, H->sigma, H->k, H->subpixel);
    Img img = H->input;
    H->dogs = s._createDOGs(img);
    H->cands.clear();
    s._findScaleSpaceExtrema(H->dogs, H->cands);
    s._eliminateEdgeResponses(H->cands, H->dogs);
    H->survivors = H->cands;
    std::sort(H->survivors.begin(), H->survivors.end(), sift::InterestPoint::cmpByFilter);
    auto it = std::find_if(H->survivors.begin(), H->survivors.end(), [](const sift::InterestPoint& p) { return p.filtered; });
    u16_t size = std::distance(H->survivors.begin(), it);
    H->survivors.resize(size);
    H->staged = true;
}

void export_points(const std::vector<sift::InterestPoint>& v, uint16_t* x, uint16_t* y, uint16_t* octave, uint16_t* index,
                   float* scale, float* orientation, uint8_t* filtered, float* desc, int* desc_len) {
    for (size_t n = 0; n < v.size(); ++n) {
        x[n] = v[n].loc.x; y[n] = v[n].loc.y; octave[n] = v[n].octave; index[n] = v[n].index;
        scale[n] = v[n].scale;
        if (orientation) orientation[n] = v[n].orientation;
        filtered[n] = v[n].filtered ? 1 : 0;
        if (desc_len) desc_len[n] = (int)v[n].descriptors.size();
        if (desc) {
            std::memset(desc + n * 128, 0, 128 * sizeof(float));
            for (size_t d = 0; d < v[n].descriptors.size() && d < 128; ++d) desc[n * 128 + d] = v[n].descriptors[d];
        }
    }
}

}  // namespace

extern "C" {

const char* oracle_flavour() {
#ifdef REF_SHIM_FAST
    return "reference sources + Vigra shim (copy-on-write arrays, memoised blur)";
#else
    return "reference sources + Vigra shim (deep copies)";
#endif
}

void* oracle_create(int dogs_per_epoch, int octaves, float sigma, float k, int subpixel, int /*literal*/, int /*strict*/) {
    Handle* h = new Handle();
    h->dpe = dogs_per_epoch; h->octaves = octaves; h->sigma = sigma; h->k = k; h->subpixel = subpixel != 0;
    return h;
}

void oracle_destroy(void* vh) { delete (Handle*)vh; }

const char* oracle_last_error(void* vh) { return ((Handle*)vh)->error.c_str(); }

// The reference's call sequence (main.cpp:56-57): construct, calculate.  -1: an exception derived from
// std::exception left calculate() (Vigra precondition); -2: an assert of the reference would have fired
// (sift.cpp:382-383; checked here first because a live assert aborts the process).
int oracle_calculate(void* vh, const float* img, int w, int h, int* out_w, int* out_h) {
    Handle* H = (Handle*)vh;
    H->error.clear();
    H->staged = false;
    if (!(H->octaves > 0) || !(H->dpe >= 3)) {
        H->error = "assert(_octaves > 0 && _dogsPerEpoch >= 3)";
        return -2;
    }
    try {
        H->sift.reset(new sift::Sift((u16_t)H->dpe, (u16_t)H->octaves, H->sigma, H->k, H->subpixel));
        Img im = wrap(img, w, h);
        H->result = H->sift->calculate(im);
        H->input = im;
        if (out_w) *out_w = (int)im.width();
        if (out_h) *out_h = (int)im.height();
        return (int)H->result.size();
    } catch (const vigra::PreconditionViolation& e) {
        H->error = e.what();
        return -1;
    } catch (const std::exception& e) {
        H->error = e.what();
        return -2;
    }
}

int oracle_level_dims(void* vh, int o, int* w, int* h) {
    Handle* H = (Handle*)vh;
    const Img& im = H->sift->_gaussians((u16_t)o, 0).img;
    *w = (int)im.width();
    *h = (int)im.height();
    return 0;
}

float oracle_get_gauss(void* vh, int o, int i, float* out) {
    const sift::OctaveElem& e = ((Handle*)vh)->sift->_gaussians((u16_t)o, (u16_t)i);
    if (out) copy_out(e.img, out);
    return e.scale;
}

float oracle_get_dog(void* vh, int o, int i, float* out) {
    Handle* H = (Handle*)vh;
    stage(H);
    const sift::OctaveElem& e = H->dogs((u16_t)o, (u16_t)i);
    if (out) copy_out(e.img, out);
    return e.scale;
}

void oracle_nearest_gaussian(void* vh, float scale, int* o, int* i) {
    sift::Point<u16_t, u16_t> p = ((Handle*)vh)->sift->_findNearestGaussian(scale);
    *o = p.x;
    *i = p.y;
}

int oracle_n_candidates(void* vh) {
    stage((Handle*)vh);
    return (int)((Handle*)vh)->cands.size();
}
void oracle_get_candidates(void* vh, uint16_t* x, uint16_t* y, uint16_t* octave, uint16_t* index, float* scale, uint8_t* filtered) {
    stage((Handle*)vh);
    export_points(((Handle*)vh)->cands, x, y, octave, index, scale, nullptr, filtered, nullptr, nullptr);
}
int oracle_n_survivors(void* vh) {
    stage((Handle*)vh);
    return (int)((Handle*)vh)->survivors.size();
}
void oracle_get_survivors(void* vh, uint16_t* x, uint16_t* y, uint16_t* octave, uint16_t* index, float* scale, uint8_t* filtered) {
    stage((Handle*)vh);
    export_points(((Handle*)vh)->survivors, x, y, octave, index, scale, nullptr, filtered, nullptr, nullptr);
}
int oracle_n_keypoints(void* vh) { return (int)((Handle*)vh)->result.size(); }
void oracle_get_keypoints(void* vh, uint16_t* x, uint16_t* y, uint16_t* octave, uint16_t* index, float* scale, float* orientation,
                          uint8_t* filtered, float* desc, int* desc_len) {
    export_points(((Handle*)vh)->result, x, y, octave, index, scale, orientation, filtered, desc, desc_len);
}

// main.cpp:78-89 (restated: main.cpp is not compilable here), over the reference's own InterestPoints.
long oracle_format_results(void* vh, char* buf, long cap) {
    Handle* H = (Handle*)vh;
    std::ostringstream out;
    out << "Location\tscale\torientation\tdescriptors\n";
    for (const sift::InterestPoint& p : H->result) {
        out << "[" << p.loc.x << ", " << p.loc.y << "]\t" << p.scale << "\t" << p.orientation << "\t" << "[";
        for (f32_t d : p.descriptors) out << d << ", ";
        out << "]\n";
    }
    H->text = out.str();
    if (buf && cap > 0) {
        long n = std::min<long>(cap - 1, (long)H->text.size());
        std::memcpy(buf, H->text.data(), (size_t)n);
        buf[n] = 0;
    }
    return (long)H->text.size();
}

// ---- unit entry points: the reference's alg:: functions and private stages, called directly -------------
int oracle_convolve(const float* src, int w, int h, float sigma, float* dst) {
    try {
        copy_out(sift::alg::convolveWithGauss(wrap(src, w, h), sigma), dst);
        return 0;
    } catch (const std::exception&) {
        return -1;
    }
}
int oracle_reduce(const float* src, int w, int h, float sigma, float* dst) {
    try {
        copy_out(sift::alg::reduceToNextLevel(wrap(src, w, h), sigma), dst);
        return 0;
    } catch (const std::exception&) {
        return -1;
    }
}
int oracle_increase(const float* src, int w, int h, float sigma, float* dst) {
    try {
        copy_out(sift::alg::increaseToNextLevel(wrap(src, w, h), sigma), dst);
        return 0;
    } catch (const std::exception&) {
        return -1;
    }
}
void oracle_dog(const float* lower, const float* higher, long n, float* out) {
    copy_out(sift::alg::dog(wrap(lower, (int)n, 1), wrap(higher, (int)n, 1)), out);
}

static sift::Matrix<sift::OctaveElem> three_dogs(const float* d0, const float* d1, const float* d2, int w, int h) {
    sift::Matrix<sift::OctaveElem> dogs(1, 3);
    const float* src[3] = {d0, d1, d2};
    for (u16_t i = 0; i < 3; ++i) dogs(0, i).img = wrap(src[i], w, h);
    return dogs;
}

long oracle_extrema(const float* d0, const float* d1, const float* d2, int w, int h, uint16_t* xs, uint16_t* ys, long cap) {
    sift::Sift s(3, 1);
    std::vector<sift::InterestPoint> out;
    s._findScaleSpaceExtrema(three_dogs(d0, d1, d2, w, h), out);
    for (long n = 0; n < (long)out.size() && n < cap; ++n) {
        xs[n] = out[(size_t)n].loc.x;
        ys[n] = out[(size_t)n].loc.y;
    }
    return (long)out.size();
}

void oracle_eliminate(const float* d0, const float* d1, const float* d2, int w, int h, const uint16_t* xs, const uint16_t* ys,
                      long n, uint8_t* filtered) {
    sift::Sift s(3, 1);
    std::vector<sift::InterestPoint> pts;
    for (long i = 0; i < n; ++i) pts.emplace_back(sift::InterestPoint(sift::Point<u16_t, u16_t>(xs[i], ys[i]), 0.0f, 0, 1));
    s._eliminateEdgeResponses(pts, three_dogs(d0, d1, d2, w, h));
    for (long i = 0; i < n; ++i) filtered[i] = pts[(size_t)i].filtered ? 1 : 0;
}

float oracle_vertex_parabola(int lx, float ly, int px, float py, int rx, float ry) {
    return sift::alg::vertexParabola(sift::Point<u16_t, f32_t>((u16_t)lx, ly), sift::Point<u16_t, f32_t>((u16_t)px, py),
                                     sift::Point<u16_t, f32_t>((u16_t)rx, ry));
}

int oracle_find_peaks(const float* histo, float* out36) {
    std::array<f32_t, 36> h;
    std::copy(histo, histo + 36, h.begin());
    sift::Sift s;
    const std::set<f32_t> p = s._findPeaks(h);
    int n = 0;
    for (f32_t v : p) out36[n++] = v;
    return n;
}

// std::sort(cmpByFilter) over real sift::InterestPoint objects (sift.cpp:37); ids ride in loc/octave.
void oracle_sort_order(const uint8_t* flags, long n, uint32_t* order) {
    std::vector<sift::InterestPoint> v((size_t)n);
    for (long i = 0; i < n; ++i) {
        v[(size_t)i].filtered = flags[i] != 0;
        v[(size_t)i].loc.x = (u16_t)(i & 0xffff);
        v[(size_t)i].loc.y = (u16_t)((i >> 16) & 0xffff);
    }
    std::sort(v.begin(), v.end(), sift::InterestPoint::cmpByFilter);
    for (long i = 0; i < n; ++i) order[i] = (uint32_t)v[(size_t)i].loc.x | ((uint32_t)v[(size_t)i].loc.y << 16);
}

void oracle_gradient(const float* img, int w, int h, float* mag, float* ori) {
    Img im = wrap(img, w, h);
    std::memset(mag, 0, sizeof(float) * (size_t)w * (size_t)h);
    std::memset(ori, 0, sizeof(float) * (size_t)w * (size_t)h);
    for (u16_t x = 1; x < w - 1; ++x)
        for (u16_t y = 1; y < h - 1; ++y) {
            mag[(size_t)y * w + x] = sift::alg::gradientMagnitude(im, sift::Point<u16_t, u16_t>(x, y));
            ori[(size_t)y * w + x] = sift::alg::gradientOrientation(im, sift::Point<u16_t, u16_t>(x, y));
        }
}

void oracle_normalize(float* v, int n) {
    std::vector<f32_t> vec(v, v + n);
    sift::alg::normalizeVector(vec);
    std::copy(vec.begin(), vec.end(), v);
}

double oracle_time_calculate(void* vh, const float* img, int w, int h, int* n_out) {
    auto t0 = std::chrono::steady_clock::now();
    int n = oracle_calculate(vh, img, w, h, nullptr, nullptr);
    auto t1 = std::chrono::steady_clock::now();
    if (n_out) *n_out = n;
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
