// ORACLE — TEST INFRASTRUCTURE ONLY.  Parity pinned against the reference's source build and its shipped executable (see sift_oracle.hpp).
// CPU restatement of the reference's hot path; every function cites the reference lines it
// follows (paths relative to the reference tree) and, where the arithmetic lives in Vigra,
// the SURVEY.md Appendix A item (each confirmed in the reference's shipped binary).
#include "sift_oracle.hpp"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <set>
#include <sstream>

#include "vigra_linalg.hpp"

namespace oracle {

// ---------------------------------------------------------------------------------------
// Vigra Kernel1D<float>::initGaussian(std_dev) as called from algorithms.cpp:13-14
// (binary @0x41a900).  radius=(int)(3*sd+0.5) (>=1); tap = n*expf(s2*x*x) in fp32
// with s2=(float)(-0.5/sf/sf), n=(float)(0.3989422804014327/sf); fp32 sequential sum,
// scale=1/sum, tap*=scale.
std::vector<float> gaussian_taps(float sigma, int* radius_out) { return gaussian_taps_d((double)sigma, radius_out); }

std::vector<float> gaussian_taps_d(double std_dev, int* radius_out) {
    if (!(std_dev >= 0.0))
        throw Precondition("Kernel1D::initGaussian(): Standard deviation must be >= 0.");
    std::vector<float> taps;
    int radius = 0;
    if (std_dev > 0.0) {
        const float sf = (float)std_dev;
        const float s2 = (float)(-0.5 / (double)sf / (double)sf);
        const float nrm = (float)(0.3989422804014327 / (double)sf);
        radius = (int)(3.0 * std_dev + 0.5);
        if (radius == 0) radius = 1;
        taps.reserve((size_t)(2 * radius + 1));
        for (float x = -(float)radius; x <= (float)radius; ++x) {
            float x2 = x * x;
            taps.push_back(nrm * std::exp(x2 * s2));  // std::exp(float) == expf
        }
    } else {
        taps.push_back(1.0f);
    }
    float sum = 0.0f;
    for (float t : taps) sum += t;
    if (sum == 0.0f) throw Precondition("Kernel1D<ARITHTYPE>::normalize(): Cannot normalize a kernel with sum = 0");
    sum = 1.0f / sum;
    for (float& t : taps) t = t * sum;
    if (radius_out) *radius_out = radius;
    return taps;
}

// One line of Vigra's internalConvolveLineReflect (binary @0x41ba90): for each
// output the taps are applied to virtual source indices x-r .. x+r in ascending order, kernel
// walked from +r down; out-of-range indices reflect about the edge pixel without repeating it;
// separate fp32 multiply and add.  `src`/`dst` are strided so the same routine serves X and Y.
static void convolve_line_reflect(const float* src, long sstride, float* dst, long dstride, int n,
                                  const float* taps, int r) {
    for (int x = 0; x < n; ++x) {
        float sum = 0.0f;
        for (int j = -r; j <= r; ++j) {
            int s = x + j;
            if (s < 0) s = -s;
            if (s >= n) s = 2 * (n - 1) - s;
            sum += taps[r - j] * src[(long)s * sstride];
        }
        dst[(long)x * dstride] = sum;
    }
}

// All lines of one pass over a strided 2-D array (axis 0: lines along x, axis 1: along y).  When x is
// contiguous the y pass runs row-wise: acc[x] walks the same taps in the same order with the same separate
// fp32 multiply and add as the per-line loop above, so every pixel sees identical arithmetic.
void convolve_lines_reflect(const float* src, long ssx, long ssy, float* dst, long dsx, long dsy, int w, int h,
                            const float* taps, int r, int axis) {
    if (axis == 0) {
        for (int y = 0; y < h; ++y) convolve_line_reflect(src + y * ssy, ssx, dst + y * dsy, dsx, w, taps, r);
        return;
    }
    if (ssx != 1 || dsx != 1) {
        for (int x = 0; x < w; ++x) convolve_line_reflect(src + x * ssx, ssy, dst + x * dsx, dsy, h, taps, r);
        return;
    }
    std::vector<float> acc((size_t)w);
    for (int y = 0; y < h; ++y) {
        std::fill(acc.begin(), acc.end(), 0.0f);
        for (int j = -r; j <= r; ++j) {
            int s = y + j;
            if (s < 0) s = -s;
            if (s >= h) s = 2 * (h - 1) - s;
            const float t = taps[r - j];
            const float* row = src + (long)s * ssy;
            float* a = acc.data();
            for (int x = 0; x < w; ++x) a[x] += t * row[x];
        }
        std::memcpy(dst + (long)y * dsy, acc.data(), sizeof(float) * (size_t)w);
    }
}

// algorithms.cpp:10-22: Kernel1D.initGaussian(sigma); separableConvolveX -> tmp; separableConvolveY.
// Precondition: kernel longer than line when w <= r or h <= r.
Image convolve_with_gauss(const Image& img, float sigma) {
    int r = 0;
    std::vector<float> taps = gaussian_taps(sigma, &r);
    if (img.w < r + 1) throw Precondition("separableConvolveX(): kernel longer than line");
    Image tmp(img.w, img.h), res(img.w, img.h);
    convolve_lines_reflect(img.px.data(), 1, img.w, tmp.px.data(), 1, img.w, img.w, img.h, taps.data(), r, 0);
    if (img.h < r + 1) throw Precondition("separableConvolveY(): kernel longer than line");
    convolve_lines_reflect(tmp.px.data(), 1, img.w, res.px.data(), 1, img.w, img.w, img.h, taps.data(), r, 1);
    return res;
}

// Vigra resizeImageNoInterpolation line walk (binary @0x419ac0): x=0.5 accumulated in
// double with dx=(n_old-1)/(n_new-1), truncation.
std::vector<int> resize_index_map(int n_old, int n_new) {
    std::vector<int> m((size_t)n_new);
    if (n_new == 1) {
        m[0] = 0;
        return m;
    }
    double dx = (double)(n_old - 1) / (double)(n_new - 1);
    double x = 0.5;
    for (int i = 0; i < n_new; ++i, x += dx) m[(size_t)i] = (int)x;
    return m;
}

Image resize_no_interpolation(const Image& src, int nw, int nh) {
    if (!(src.w > 1 && src.h > 1)) throw Precondition("resizeImageNoInterpolation(): Source image too small.");
    if (!(nw > 1 && nh > 1)) throw Precondition("resizeImageNoInterpolation(): Destination image too small.");
    std::vector<int> mx = resize_index_map(src.w, nw), my = resize_index_map(src.h, nh);
    Image out(nw, nh);
    for (int y = 0; y < nh; ++y)
        for (int x = 0; x < nw; ++x) out(x, y) = src(mx[(size_t)x], my[(size_t)y]);
    return out;
}

// algorithms.cpp:24-36: blur with sigma, then NN resize to ((w+1)/2, (h+1)/2).
Image reduce_to_next_level(const Image& img, float sigma) {
    return resize_no_interpolation(convolve_with_gauss(img, sigma), (img.w + 1) / 2, (img.h + 1) / 2);
}

// algorithms.cpp:38-49: blur with sigma, then NN resize to (2w, 2h).
Image increase_to_next_level(const Image& img, float sigma) {
    return resize_no_interpolation(convolve_with_gauss(img, sigma), img.w * 2, img.h * 2);
}

// algorithms.cpp:52-64: dif = higher - lower; 128 + dif (two fp32 roundings).
Image dog(const Image& lower, const Image& higher) {
    Image res(lower.w, lower.h);
    for (size_t i = 0; i < res.px.size(); ++i) {
        const float dif = higher.px[i] - lower.px[i];
        res.px[i] = 128 + dif;
    }
    return res;
}

// algorithms.cpp:66-77 (sign-flipped first differences, /2).
void fo_derivative(const Image* const d[3], int x, int y, float out[3]) {
    out[0] = ((*d[1])(x - 1, y) - (*d[1])(x + 1, y)) / 2;
    out[1] = ((*d[1])(x, y - 1) - (*d[1])(x, y + 1)) / 2;
    out[2] = ((*d[0])(x, y) - (*d[2])(x, y)) / 2;
}

// algorithms.cpp:79-106 (mixed terms /2; dys first two terms cancel, :91).
void so_derivative(const Image* const d[3], int x, int y, float h[3][3]) {
    const Image &d0 = *d[0], &d1 = *d[1], &d2 = *d[2];
    const float dxx = d1(x + 1, y) + d1(x - 1, y) - 2 * d1(x, y);
    const float dyy = d1(x, y + 1) + d1(x, y - 1) - 2 * d1(x, y);
    const float dss = d2(x, y) + d0(x, y) - 2 * d1(x, y);
    const float dxy = (d1(x + 1, y + 1) - d1(x - 1, y + 1) - d1(x + 1, y - 1) + d1(x - 1, y - 1)) / 2;
    const float dxs = (d2(x + 1, y) - d2(x - 1, y) - d0(x + 1, y) + d0(x - 1, y)) / 2;
    const float dys = (d2(x, y + 1) - d2(x, y + 1) - d0(x, y + 1) + d0(x, y - 1)) / 2;
    h[0][0] = dxx; h[1][0] = dxy; h[2][0] = dxs;
    h[0][1] = dxy; h[1][1] = dyy; h[2][1] = dys;
    h[0][2] = dxs; h[1][2] = dys; h[2][2] = dss;
}

// algorithms.cpp:108-111: f32 differences, pow/sqrt in double, result narrowed to f32.
float gradient_magnitude(const Image& img, int x, int y) {
    const float dx = img(x + 1, y) - img(x - 1, y);
    const float dy = img(x, y + 1) - img(x, y - 1);
    return (float)std::sqrt(std::pow((double)dx, 2.0) + std::pow((double)dy, 2.0));
}

// algorithms.cpp:113-116: atan2f in radians, +360 in f32, fmod in double.
float gradient_orientation(const Image& img, int x, int y) {
    const float result = std::atan2(img(x, y + 1) - img(x, y - 1), img(x + 1, y) - img(x - 1, y));
    return (float)std::fmod((double)(result + 360), 360.0);
}

// algorithms.cpp:153-178: A=[[x^2, x, 0]] rows=points, linearSolve (return ignored), -r1/(2 r0).
float vertex_parabola(uint16_t lx, float ly, uint16_t px, float py, uint16_t rx, float ry) {
    la::Mat a(3, 3), b(3, 1), res(3, 1);
    a(0, 0) = (float)std::pow((double)lx, 2.0);
    a(1, 0) = (float)std::pow((double)px, 2.0);
    a(2, 0) = (float)std::pow((double)rx, 2.0);
    a(0, 1) = lx; a(1, 1) = px; a(2, 1) = rx;
    a(0, 2) = 0;  a(1, 2) = 0;  a(2, 2) = 0;
    b(0, 0) = ly; b(1, 0) = py; b(2, 0) = ry;
    la::linear_solve(a.view(), b.view(), res.view());
    return -res(1, 0) / (2 * res(0, 0));
}

// algorithms.cpp:210-223: divide by the plain sum (L1), skip when the sum is 0.
void normalize_vector(std::vector<float>& v) {
    float length = 0;
    for (float n : v) length += n;
    if (length == 0) return;
    for (float& n : v) n /= length;
}

// sift.cpp:220-286.
std::vector<float> find_peaks(const float histo[36]) {
    std::set<float> result;
    std::array<float, 36> peaks_only;
    std::copy(histo, histo + 36, peaks_only.begin());
    const uint16_t max_index = (uint16_t)std::distance(peaks_only.begin(),
                                                       std::max_element(peaks_only.begin(), peaks_only.end()));
    const float range = histo[max_index] * 0.8;  // double product narrowed to f32 (sift.cpp:228)
    for (float& e : peaks_only)
        if (e < range) e = -1;
    for (uint16_t i = 1; i < 35; i++)
        if (peaks_only[i] < peaks_only[i - 1] || peaks_only[i] < peaks_only[i + 1]) peaks_only[i] = -1;

    auto vertex_for = [&](uint16_t i) {
        uint16_t lx, rx;
        float ly, ry;
        if (i == 0) { lx = 35 * 10 + 5; ly = histo[35]; } else { lx = (uint16_t)((i - 1) * 10 + 5); ly = histo[i - 1]; }
        if (i == 35) { rx = 5; ry = histo[0]; } else { rx = (uint16_t)((i + 1) * 10 + 5); ry = histo[i + 1]; }
        return vertex_parabola(lx, ly, (uint16_t)(i * 10 + 5), histo[i], rx, ry);
    };
    result.emplace(vertex_for(max_index));
    for (uint16_t i = 0; i < 36; i++)
        if (peaks_only[i] > -1 && i != max_index) result.emplace(vertex_for(i));
    return std::vector<float>(result.begin(), result.end());
}

bool inverse3(const float a[9], float out[9]) {
    la::Mat m(3, 3), r(3, 3);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) m(i, j) = a[i * 3 + j];
    bool ok = la::inverse(m.view(), r.view());
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) out[i * 3 + j] = r(i, j);
    return ok;
}

bool linear_solve3(const float a[9], const float b[3], float out[3]) {
    la::Mat m(3, 3), bb(3, 1), r(3, 1);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) m(i, j) = a[i * 3 + j];
        bb(i, 0) = b[i];
    }
    bool ok = la::linear_solve(m.view(), bb.view(), r.view());
    for (int i = 0; i < 3; ++i) out[i] = r(i, 0);
    return ok;
}

// interestpoint.hpp:57-62.
static bool cmp_by_filter(const KeyPoint& a, const KeyPoint& b) { return !a.filtered && b.filtered; }

void sort_by_filter_order(const uint8_t* filtered, size_t n, uint32_t* order) {
    struct E {
        uint32_t idx;
        uint8_t f;
    };
    std::vector<E> v(n);
    for (size_t i = 0; i < n; ++i) v[i] = E{(uint32_t)i, filtered[i]};
    std::sort(v.begin(), v.end(), [](const E& a, const E& b) { return !a.f && b.f; });
    for (size_t i = 0; i < n; ++i) order[i] = v[i].idx;
}

// sift.cpp:37-42 / :49-54: unstable std::sort, count of leading unfiltered truncated to u16.
void Sift::cleanup(std::vector<KeyPoint>& pts) {
    std::sort(pts.begin(), pts.end(), cmp_by_filter);
    auto it = std::find_if(pts.begin(), pts.end(), [](const KeyPoint& p) { return p.filtered; });
    uint16_t size = (uint16_t)std::distance(pts.begin(), it);
    pts.resize(size);
}

// sift.cpp:381-417.
void Sift::create_dogs(const Image& img) {
    if (!(prm.octaves > 0)) throw std::logic_error("assert(_octaves > 0)");
    if (!(prm.dogs_per_epoch >= 3)) throw std::logic_error("assert(_dogsPerEpoch >= 3)");
    const int O = prm.octaves, D = prm.dogs_per_epoch, G = D + 1;
    gaussians.assign((size_t)(O * G), Level());
    dogs.assign((size_t)(O * D), Level());
    auto g = [&](int o, int i) -> Level& { return gaussians[(size_t)(o * G + i)]; };
    auto d = [&](int o, int i) -> Level& { return dogs[(size_t)(o * D + i)]; };

    g(0, 0).scale = prm.sigma;
    g(0, 0).img = convolve_with_gauss(img, prm.sigma);
    uint16_t exp = 0;
    for (int i = 0; i < O; i++) {
        for (int j = 1; j < D + 1; j++) {
            float scale = (float)(std::pow((double)prm.k, (double)exp) * (double)prm.sigma);
            g(i, j).scale = scale;
            g(i, j).img = convolve_with_gauss(g(i, j - 1).img, scale);
            d(i, j - 1).scale = g(i, j).scale - g(i, j - 1).scale;
            d(i, j - 1).img = dog(g(i, j - 1).img, g(i, j).img);
            exp++;
        }
        if (i < O - 1) {
            g(i + 1, 0).img = reduce_to_next_level(g(i, D - 1).img, g(i, D - 1).scale);
            g(i + 1, 0).scale = g(i, D - 1).scale;
            exp -= 2;
        }
    }
}

// sift.cpp:348-379: half-open 2x2 neighbourhood {x-1,x}x{y-1,y} in three layers;
// candidate iff none strictly greater OR none strictly less; order e, i, x (outer), y (inner).
void Sift::find_scale_space_extrema(const std::vector<Level>& dogs, int octaves, int n_dogs,
                                    std::vector<KeyPoint>& out) {
    for (int e = 0; e < octaves; e++)
        for (int i = 1; i < n_dogs - 1; i++) {
            const Image& cur = dogs[(size_t)(e * n_dogs + i)].img;
            const Image& und = dogs[(size_t)(e * n_dogs + i - 1)].img;
            const Image& abv = dogs[(size_t)(e * n_dogs + i + 1)].img;
            for (int x = 1; x < cur.w - 1; x++)
                for (int y = 1; y < cur.h - 1; y++) {
                    const float v = cur(x, y);
                    bool any_gt = false, any_lt = false;
                    const Image* L[3] = {&cur, &und, &abv};
                    for (const Image* im : L)
                        for (int yy = y - 1; yy <= y; ++yy)
                            for (int xx = x - 1; xx <= x; ++xx) {
                                const float n = (*im)(xx, yy);
                                any_gt |= (n > v);
                                any_lt |= (n < v);
                            }
                    if (!any_gt || !any_lt) {
                        KeyPoint p;
                        p.x = (uint16_t)x; p.y = (uint16_t)y;
                        p.scale = dogs[(size_t)(e * n_dogs + i)].scale;
                        p.octave = (uint16_t)e; p.index = (uint16_t)i;
                        out.push_back(p);
                    }
                }
        }
}

// sift.cpp:288-346.
void Sift::eliminate_edge_responses(std::vector<KeyPoint>& pts) const {
    const int D = prm.dogs_per_epoch;
    const float t = (float)(std::pow(10.0 + 1.0, 2.0) / 10);  // sift.cpp:294
    la::Mat extremum(3, 1), inverse_matrix(3, 3);
    for (KeyPoint& p : pts) {
        const Level& d = dogs[(size_t)(p.octave * D + p.index)];
        // sift.cpp:297-298 deep-copies the three DoG images per candidate; only the literal flavour keeps that.
        std::array<Image, 3> copies;
        const Image* param[3];
        for (int s = 0; s < 3; ++s) {
            const Image& src = dogs[(size_t)(p.octave * D + p.index - 1 + s)].img;
            if (prm.literal) { copies[(size_t)s] = src; param[s] = &copies[(size_t)s]; } else { param[s] = &src; }
        }
        float dv[3], h[3][3];
        fo_derivative(param, p.x, p.y, dv);
        so_derivative(param, p.x, p.y, h);
        la::Mat deriv(3, 1), neg(3, 3);
        for (int i = 0; i < 3; ++i) {
            deriv(i, 0) = dv[i];
            for (int j = 0; j < 3; ++j) neg(i, j) = h[i][j] * -1.0f;
        }
        if (!la::inverse(neg.view(), inverse_matrix.view())) { p.filtered = true; continue; }
        if (!la::linear_solve(inverse_matrix.view(), deriv.view(), extremum.view())) { p.filtered = true; continue; }
        if (extremum(0, 0) > 127.5 || extremum(1, 0) > 127.5 || extremum(2, 0) > 127.5) { p.filtered = true; continue; }
        float func_val_extremum = la::dot_vec(deriv.view().T(), extremum.view());
        func_val_extremum = (float)((double)func_val_extremum * (0.5 + (double)d.img(p.x, p.y)));
        if ((double)func_val_extremum < 7.65) { p.filtered = true; continue; }
        const float dxx = h[0][0], dyy = h[1][1];
        const float hessian_tr = dxx + dyy;
        const float hessian_det = (float)((double)(dxx * dyy) - std::pow((double)h[0][1], 2.0));
        if (hessian_det < 0) { p.filtered = true; continue; }
        if (std::pow((double)hessian_tr, 2.0) / (double)hessian_det > (double)t) p.filtered = true;
    }
}

// sift.cpp:205-218.
void Sift::nearest_gaussian(float scale, int* o_out, int* i_out) const {
    float lowest_diff = 100;
    int bo = 0, bi = 0;
    for (int o = 0; o < prm.octaves; o++)
        for (int i = 0; i < n_gauss(); i++) {
            const float cur = std::abs(gauss(o, i).scale - scale);
            if (cur < lowest_diff) { lowest_diff = cur; bo = o; bi = i; }
        }
    *o_out = bo; *i_out = bi;
}

// sift.cpp:130-160: interior pixels only, border stays 0.  The reference fills all levels; the
// hoisted flavour fills only the levels _findNearestGaussian can return for this run (the only
// ones ever read), which leaves results unchanged.
void Sift::create_gradient_pyramids() {
    const int O = prm.octaves, G = n_gauss();
    magnitudes.assign((size_t)(O * G), Image());
    orientations.assign((size_t)(O * G), Image());
    std::vector<char> need((size_t)(O * G), prm.literal ? 1 : 0);
    if (!prm.literal) {
        for (int e = 0; e < O; ++e)
            for (int i = 1; i < n_dogs() - 1; ++i) {
                int o, gi;
                nearest_gaussian(dogl(e, i).scale, &o, &gi);
                need[(size_t)(o * G + gi)] = 1;
            }
    }
    for (int o = 0; o < O; o++)
        for (int i = 0; i < G; i++) {
            if (!need[(size_t)(o * G + i)]) continue;
            const Image& cg = gauss(o, i).img;
            Image mag(cg.w, cg.h), ori(cg.w, cg.h);
            for (int x = 1; x < cg.w - 1; x++)
                for (int y = 1; y < cg.h - 1; y++) {
                    mag(x, y) = gradient_magnitude(cg, x, y);
                    ori(x, y) = gradient_orientation(cg, x, y);
                }
            magnitudes[(size_t)(o * G + i)] = std::move(mag);
            orientations[(size_t)(o * G + i)] = std::move(ori);
        }
}

static Image window_copy(const Image& src, int x0, int y0, int n) {
    Image out(n, n);
    for (int y = 0; y < n; ++y)
        for (int x = 0; x < n; ++x) out(x, y) = src(x0 + x, y0 + y);
    return out;
}

// sift.cpp:163-203 + algorithms.cpp:118-133 (hist36: bins[(u16)floorf(o/10) % 35] += mag*gauss, x outer).
void Sift::orientation_assignment(std::vector<KeyPoint>& pts) {
    const int region = 8;
    std::vector<KeyPoint> additional;
    for (KeyPoint& p : pts) {
        int co, ci;
        nearest_gaussian(p.scale, &co, &ci);
        const Image& closest = gauss(co, ci).img;
        if ((p.x < region || p.x >= closest.w - region) || (p.y < region || p.y >= closest.h - region)) {
            p.filtered = true;
            continue;
        }
        const int x0 = p.x - region, y0 = p.y - region;
        const Image gauss_region = window_copy(closest, x0, y0, 2 * region);

        // sift.cpp:184: result unused; only its precondition exception is observable.
        {
            const float dead_sigma = (float)(1.5 * (double)p.scale);
            if (prm.literal) {
                try { (void)convolve_with_gauss(gauss_region, dead_sigma); }
                catch (const Precondition&) { if (prm.strict) throw; }
            } else if (prm.strict) {
                int r = 0;
                (void)gaussian_taps(dead_sigma, &r);
                if (2 * region < r + 1) throw Precondition("separableConvolveX(): kernel longer than line");
            }
        }
        const Image orientation = window_copy(orientations[(size_t)(co * n_gauss() + ci)], x0, y0, 2 * region);
        const Image magnitude = window_copy(magnitudes[(size_t)(co * n_gauss() + ci)], x0, y0, 2 * region);

        float bins[36] = {0};
        for (int x = 0; x < orientation.w; x++)
            for (int y = 0; y < orientation.h; y++) {
                const float sum = magnitude(x, y) * gauss_region(x, y);
                uint16_t i = (uint16_t)(int)std::floor(orientation(x, y) / 10);
                i = i % 35;
                bins[i] += sum;
            }
        const std::vector<float> peaks = find_peaks(bins);
        p.orientation = peaks.front();
        if (peaks.size() > 1)
            for (float v : peaks) {  // `peaks.begin()++` yields begin(): the first peak is duplicated too
                KeyPoint temp = p;
                temp.orientation = v;
                additional.push_back(temp);
            }
    }
    pts.insert(pts.end(), additional.begin(), additional.end());
}

// sift.cpp:60-110: views into the orientation/magnitude pyramids are mutated in place.
// algorithms.cpp:135-150 (hist8: bins[(u16)floorf(o/45) % 7]); sift.cpp:113-128 net effect = L1 normalise.
void Sift::create_descriptors(std::vector<KeyPoint>& pts) {
    const int region = 8;
    const int G = n_gauss();
    std::vector<Image> hoisted_weighting((size_t)(prm.octaves * G));
    for (KeyPoint& p : pts) {
        int co, ci;
        nearest_gaussian(p.scale, &co, &ci);
        const Image& current = gauss(co, ci).img;
        if (p.x < region || p.x > current.w - region || p.y < region || p.y > current.h - region) {
            p.filtered = true;
            continue;
        }
        const int x0 = p.x - region, y0 = p.y - region;
        Image& ori = orientations[(size_t)(co * G + ci)];
        Image& mag = magnitudes[(size_t)(co * G + ci)];

        for (int x = 0; x < 2 * region; x++)
            for (int y = 0; y < 2 * region; y++) ori(x0 + x, y0 + y) += p.orientation;

        // sift.cpp:87: full-image blur per keypoint, of which only the top-left 16x16 is read (:88-92).
        Image literal_weighting;
        const Image* weighting;
        if (prm.literal) {
            literal_weighting = convolve_with_gauss(current, 1.6f);
            weighting = &literal_weighting;
        } else {
            Image& hw = hoisted_weighting[(size_t)(co * G + ci)];
            if (hw.w == 0) hw = convolve_with_gauss(current, 1.6f);
            weighting = &hw;
        }
        for (int x = 0; x < 2 * region; x++)
            for (int y = 0; y < 2 * region; y++) mag(x0 + x, y0 + y) += (*weighting)(x, y);

        std::vector<float> descriptors;
        for (int cx = 0; cx < 2 * region; cx += 4)
            for (int cy = 0; cy < 2 * region; cy += 4) {
                std::vector<float> bins(8, 0);
                for (int x = 0; x < 4; x++)
                    for (int y = 0; y < 4; y++) {
                        const int gx = x0 + cx + x, gy = y0 + cy + y;
                        const float sum = mag(gx, gy) * current(gx, gy);
                        uint16_t i = (uint16_t)(int)std::floor(ori(gx, gy) / 45);
                        i = i % 7;
                        bins[i] += sum;
                    }
                normalize_vector(bins);  // the clipped/renormalised copy of sift.cpp:115-127 is discarded at :103
                descriptors.insert(descriptors.end(), bins.begin(), bins.end());
            }
        p.descriptors = descriptors;
    }
}

// sift.cpp:19-57.
std::vector<KeyPoint> Sift::calculate(Image& img) {
    if (prm.subpixel) img = increase_to_next_level(img, 1.0f);
    create_dogs(img);

    std::vector<KeyPoint> pts;
    find_scale_space_extrema(dogs, prm.octaves, n_dogs(), pts);
    eliminate_edge_responses(pts);

    cands.clear();
    cands.reserve(pts.size());
    for (const KeyPoint& p : pts) cands.push_back(Candidate{p.x, p.y, p.octave, p.index, p.scale, p.filtered});

    cleanup(pts);
    after_first_trim = pts;

    create_gradient_pyramids();
    orientation_assignment(pts);
    cleanup(pts);
    create_descriptors(pts);
    return pts;
}

// main.cpp:78-89 (default ostream float formatting, trailing ", " inside the brackets).
std::string format_results(const std::vector<KeyPoint>& pts) {
    std::ostringstream out;
    out << "Location\tscale\torientation\tdescriptors\n";
    for (const KeyPoint& p : pts) {
        out << "[" << p.x << ", " << p.y << "]\t" << p.scale << "\t" << p.orientation << "\t" << "[";
        for (float d : p.descriptors) out << d << ", ";
        out << "]\n";
    }
    return out.str();
}

}  // namespace oracle
