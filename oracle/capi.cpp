// ORACLE — TEST INFRASTRUCTURE ONLY.  C entry points so tests/, smoke() and bench.py's
// cpu_baseline / --impl reference legs can drive the CPU restatement through ctypes.
#include <chrono>
#include <cstring>
#include <memory>

#include "sift_oracle.hpp"

using namespace oracle;

namespace {
struct Handle {
    Params prm;
    std::unique_ptr<Sift> sift;
    std::vector<KeyPoint> result;
    std::string error;
    std::string text;
    int out_w = 0, out_h = 0;
};
Image wrap(const float* p, int w, int h) {
    Image im(w, h);
    std::memcpy(im.px.data(), p, sizeof(float) * (size_t)w * (size_t)h);
    return im;
}
}  // namespace

extern "C" {

void* oracle_create(int dogs_per_epoch, int octaves, float sigma, float k, int subpixel, int literal, int strict) {
    Handle* h = new Handle();
    h->prm.dogs_per_epoch = (uint16_t)dogs_per_epoch;
    h->prm.octaves = (uint16_t)octaves;
    h->prm.sigma = sigma;
    h->prm.k = k;
    h->prm.subpixel = subpixel != 0;
    h->prm.literal = literal != 0;
    h->prm.strict = strict != 0;
    return h;
}

void oracle_destroy(void* vh) { delete (Handle*)vh; }

const char* oracle_last_error(void* vh) { return ((Handle*)vh)->error.c_str(); }

// Returns the number of returned keypoints, -1 on a precondition violation (what the reference
// surfaces as a std::exception), -2 on a failed assert-equivalent.
int oracle_calculate(void* vh, const float* img, int w, int h, int* out_w, int* out_h) {
    Handle* H = (Handle*)vh;
    H->error.clear();
    try {
        H->sift.reset(new Sift(H->prm));
        Image im = wrap(img, w, h);
        H->result = H->sift->calculate(im);
        H->out_w = im.w;
        H->out_h = im.h;
        if (out_w) *out_w = im.w;
        if (out_h) *out_h = im.h;
        return (int)H->result.size();
    } catch (const Precondition& e) {
        H->error = e.what();
        return -1;
    } catch (const std::exception& e) {
        H->error = e.what();
        return -2;
    }
}

int oracle_level_dims(void* vh, int o, int* w, int* h) {
    Handle* H = (Handle*)vh;
    const Image& im = H->sift->gauss(o, 0).img;
    *w = im.w;
    *h = im.h;
    return 0;
}

float oracle_get_gauss(void* vh, int o, int i, float* out) {
    const Level& l = ((Handle*)vh)->sift->gauss(o, i);
    if (out) std::memcpy(out, l.img.px.data(), sizeof(float) * l.img.px.size());
    return l.scale;
}

float oracle_get_dog(void* vh, int o, int i, float* out) {
    const Level& l = ((Handle*)vh)->sift->dogl(o, i);
    if (out) std::memcpy(out, l.img.px.data(), sizeof(float) * l.img.px.size());
    return l.scale;
}

void oracle_nearest_gaussian(void* vh, float scale, int* o, int* i) { ((Handle*)vh)->sift->nearest_gaussian(scale, o, i); }

int oracle_n_candidates(void* vh) { return (int)((Handle*)vh)->sift->candidates().size(); }

void oracle_get_candidates(void* vh, uint16_t* x, uint16_t* y, uint16_t* octave, uint16_t* index, float* scale,
                           uint8_t* filtered) {
    const auto& c = ((Handle*)vh)->sift->candidates();
    for (size_t n = 0; n < c.size(); ++n) {
        x[n] = c[n].x; y[n] = c[n].y; octave[n] = c[n].octave; index[n] = c[n].index;
        scale[n] = c[n].scale; filtered[n] = c[n].filtered ? 1 : 0;
    }
}

static void export_points(const std::vector<KeyPoint>& v, uint16_t* x, uint16_t* y, uint16_t* octave, uint16_t* index,
                          float* scale, float* orientation, uint8_t* filtered, float* desc, int* desc_len) {
    for (size_t n = 0; n < v.size(); ++n) {
        x[n] = v[n].x; y[n] = v[n].y; octave[n] = v[n].octave; index[n] = v[n].index;
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

int oracle_n_survivors(void* vh) { return (int)((Handle*)vh)->sift->survivors().size(); }
void oracle_get_survivors(void* vh, uint16_t* x, uint16_t* y, uint16_t* octave, uint16_t* index, float* scale,
                          uint8_t* filtered) {
    export_points(((Handle*)vh)->sift->survivors(), x, y, octave, index, scale, nullptr, filtered, nullptr, nullptr);
}

int oracle_n_keypoints(void* vh) { return (int)((Handle*)vh)->result.size(); }
void oracle_get_keypoints(void* vh, uint16_t* x, uint16_t* y, uint16_t* octave, uint16_t* index, float* scale,
                          float* orientation, uint8_t* filtered, float* desc, int* desc_len) {
    export_points(((Handle*)vh)->result, x, y, octave, index, scale, orientation, filtered, desc, desc_len);
}

// Text of main.cpp:78-89 for the last result; returns the length needed (excluding NUL).
long oracle_format_results(void* vh, char* buf, long cap) {
    Handle* H = (Handle*)vh;
    H->text = format_results(H->result);
    if (buf && cap > 0) {
        long n = std::min<long>(cap - 1, (long)H->text.size());
        std::memcpy(buf, H->text.data(), (size_t)n);
        buf[n] = 0;
    }
    return (long)H->text.size();
}

// ---- unit entry points -----------------------------------------------------------------
int oracle_gaussian_taps(float sigma, float* taps, int cap) {
    int r = 0;
    std::vector<float> t = gaussian_taps(sigma, &r);
    for (int i = 0; i < (int)t.size() && i < cap; ++i) taps[i] = t[(size_t)i];
    return r;
}

int oracle_convolve(const float* src, int w, int h, float sigma, float* dst) {
    try {
        Image r = convolve_with_gauss(wrap(src, w, h), sigma);
        std::memcpy(dst, r.px.data(), sizeof(float) * r.px.size());
        return 0;
    } catch (const Precondition&) {
        return -1;
    }
}

void oracle_resize_map(int n_old, int n_new, int* map) {
    std::vector<int> m = resize_index_map(n_old, n_new);
    std::memcpy(map, m.data(), sizeof(int) * m.size());
}

int oracle_resize(const float* src, int w, int h, float* dst, int nw, int nh) {
    try {
        Image r = resize_no_interpolation(wrap(src, w, h), nw, nh);
        std::memcpy(dst, r.px.data(), sizeof(float) * r.px.size());
        return 0;
    } catch (const Precondition&) {
        return -1;
    }
}

int oracle_reduce(const float* src, int w, int h, float sigma, float* dst) {
    try {
        Image r = reduce_to_next_level(wrap(src, w, h), sigma);
        std::memcpy(dst, r.px.data(), sizeof(float) * r.px.size());
        return 0;
    } catch (const Precondition&) {
        return -1;
    }
}

int oracle_increase(const float* src, int w, int h, float sigma, float* dst) {
    try {
        Image r = increase_to_next_level(wrap(src, w, h), sigma);
        std::memcpy(dst, r.px.data(), sizeof(float) * r.px.size());
        return 0;
    } catch (const Precondition&) {
        return -1;
    }
}

void oracle_dog(const float* lower, const float* higher, long n, float* out) {
    for (long i = 0; i < n; ++i) {
        const float dif = higher[i] - lower[i];
        out[i] = 128 + dif;
    }
}

// Extrema of the middle layer d1 given (d0, d1, d2); canonical x-outer / y-inner order.
long oracle_extrema(const float* d0, const float* d1, const float* d2, int w, int h, uint16_t* xs, uint16_t* ys,
                    long cap) {
    std::vector<Level> dogs(3);
    const float* src[3] = {d0, d1, d2};
    for (int i = 0; i < 3; ++i) dogs[(size_t)i].img = wrap(src[i], w, h);
    std::vector<KeyPoint> out;
    Sift::find_scale_space_extrema(dogs, 1, 3, out);
    for (long n = 0; n < (long)out.size() && n < cap; ++n) {
        xs[n] = out[(size_t)n].x;
        ys[n] = out[(size_t)n].y;
    }
    return (long)out.size();
}

// Elimination flags for candidates of the middle layer d1.
void oracle_eliminate(const float* d0, const float* d1, const float* d2, int w, int h, const uint16_t* xs,
                      const uint16_t* ys, long n, uint8_t* filtered) {
    Params prm;
    prm.octaves = 1;
    prm.dogs_per_epoch = 3;
    Sift s(prm);
    std::vector<KeyPoint> pts((size_t)n);
    for (long i = 0; i < n; ++i) {
        pts[(size_t)i].x = xs[i]; pts[(size_t)i].y = ys[i];
        pts[(size_t)i].octave = 0; pts[(size_t)i].index = 1;
    }
    s.inject_dogs(wrap(d0, w, h), wrap(d1, w, h), wrap(d2, w, h));
    s.eliminate_edge_responses(pts);
    for (long i = 0; i < n; ++i) filtered[i] = pts[(size_t)i].filtered ? 1 : 0;
}

int oracle_inverse3(const float* a, float* out) { return inverse3(a, out) ? 1 : 0; }
int oracle_linear_solve3(const float* a, const float* b, float* out) { return linear_solve3(a, b, out) ? 1 : 0; }
float oracle_vertex_parabola(int lx, float ly, int px, float py, int rx, float ry) {
    return vertex_parabola((uint16_t)lx, ly, (uint16_t)px, py, (uint16_t)rx, ry);
}
int oracle_find_peaks(const float* histo, float* out36) {
    std::vector<float> p = find_peaks(histo);
    for (size_t i = 0; i < p.size(); ++i) out36[i] = p[i];
    return (int)p.size();
}
void oracle_sort_order(const uint8_t* flags, long n, uint32_t* order) { sort_by_filter_order(flags, (size_t)n, order); }

void oracle_gradient(const float* img, int w, int h, float* mag, float* ori) {
    Image im = wrap(img, w, h);
    std::memset(mag, 0, sizeof(float) * (size_t)w * (size_t)h);
    std::memset(ori, 0, sizeof(float) * (size_t)w * (size_t)h);
    for (int x = 1; x < w - 1; ++x)
        for (int y = 1; y < h - 1; ++y) {
            mag[(size_t)y * w + x] = gradient_magnitude(im, x, y);
            ori[(size_t)y * w + x] = gradient_orientation(im, x, y);
        }
}

// Wall-clock seconds of one calculate() on this thread (bench.py cpu_baseline).
double oracle_time_calculate(void* vh, const float* img, int w, int h, int* n_out) {
    auto t0 = std::chrono::steady_clock::now();
    int n = oracle_calculate(vh, img, w, h, nullptr, nullptr);
    auto t1 = std::chrono::steady_clock::now();
    if (n_out) *n_out = n;
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
