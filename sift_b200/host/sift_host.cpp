// sift::Sift over the C ABI: the host-side drop-in for the reference's Sift::calculate (sift.cpp:19-57).
#include <cassert>
#include <cstring>
#include <fstream>
#include <sstream>

#include "../../include/sift/sift.hpp"
#include "../../include/sift_gpu.h"

namespace sift {

Sift::Sift(u16_t dogsPerEpoch, u16_t octaves, f32_t sigma, f32_t k, bool subpixel_)
    : subpixel(subpixel_), _sigma(sigma), _k(k), _dogsPerEpoch(dogsPerEpoch), _octaves(octaves) {}

Sift::~Sift() { sift_gpu_destroy(ctx_); }

void Sift::ensureContext(std::ptrdiff_t w, std::ptrdiff_t h, int batch) {
    if (ctx_ && w <= ctx_w_ && h <= ctx_h_ && batch <= ctx_batch_) return;
    sift_gpu_destroy(ctx_);
    ctx_ = nullptr;
    sift_gpu_params p;
    std::memset(&p, 0, sizeof p);
    p.sigma = _sigma; p.k = _k; p.octaves = _octaves; p.dogs_per_epoch = _dogsPerEpoch;
    p.subpixel = subpixel ? 1 : 0;
    p.device = device_;
    p.max_width = (int32_t)std::max(w, ctx_w_);
    p.max_height = (int32_t)std::max(h, ctx_h_);
    p.max_batch = std::max(batch, std::max(ctx_batch_, max_batch_));
    p.flags = flags_ | (subpixel ? SIFT_GPU_FLAG_KEEP_UPSAMPLED : 0u);
    const int rc = sift_gpu_create(&p, &ctx_);
    // the reference's preconditions (sift.cpp:382-383) are asserts
    assert(!(rc == SIFT_GPU_E_ASSERT) && "_octaves > 0 && _dogsPerEpoch >= 3");
    if (rc != SIFT_GPU_OK) throw std::runtime_error(std::string("sift_gpu_create: ") + sift_gpu_last_error(nullptr));
    ctx_w_ = p.max_width; ctx_h_ = p.max_height; ctx_batch_ = p.max_batch;
}

std::vector<std::vector<InterestPoint>> Sift::calculateBatch(std::vector<Image>& imgs) {
    std::vector<std::vector<InterestPoint>> out(imgs.size());
    if (imgs.empty()) return out;
    std::ptrdiff_t mw = 0, mh = 0;
    for (const Image& im : imgs) { mw = std::max(mw, im.width()); mh = std::max(mh, im.height()); }
    ensureContext(mw, mh, (int)std::min<size_t>(imgs.size(), (size_t)std::max(1, max_batch_)));

    std::vector<sift_gpu_image> descs(imgs.size());
    std::vector<Image> doubled(subpixel ? imgs.size() : 0);
    for (size_t i = 0; i < imgs.size(); ++i) {
        std::memset(&descs[i], 0, sizeof(sift_gpu_image));
        descs[i].data = imgs[i].data();
        descs[i].width = (int32_t)imgs[i].width();
        descs[i].height = (int32_t)imgs[i].height();
        descs[i].dtype = SIFT_GPU_DTYPE_F32;
        descs[i].memory = SIFT_GPU_MEM_HOST;
        if (subpixel) {
            doubled[i] = Image(imgs[i].width() * 2, imgs[i].height() * 2);
            descs[i].upsampled_out = doubled[i].data();
        }
    }
    std::vector<sift_gpu_result> res(imgs.size());
    const int rc = sift_gpu_run(ctx_, descs.data(), (int)descs.size(), res.data());
    if (rc == SIFT_GPU_E_PRECONDITION) throw PreconditionViolation(std::string("Precondition violation!\n") + sift_gpu_last_error(ctx_));
    if (rc != SIFT_GPU_OK) throw std::runtime_error(std::string("sift_gpu_run: ") + sift_gpu_last_error(ctx_));
    for (size_t i = 0; i < imgs.size(); ++i) {
        std::vector<InterestPoint>& pts = out[i];
        pts.resize(res[i].n);
        for (uint32_t n = 0; n < res[i].n; ++n) {
            const sift_gpu_keypoint& k = res[i].kps[n];
            InterestPoint& p = pts[n];
            p.scale = k.scale; p.octave = k.octave; p.index = k.index; p.filtered = k.filtered != 0;
            p.loc = Point<u16_t, u16_t>(k.x, k.y);
            p.orientation = k.orientation;
            if (k.desc_len) p.descriptors.assign(res[i].desc + (size_t)n * 128, res[i].desc + (size_t)n * 128 + k.desc_len);
        }
        if (subpixel) imgs[i] = std::move(doubled[i]);  // sift.cpp:21 overwrites the caller's image
    }
    return out;
}

std::vector<InterestPoint> Sift::calculate(Image& img) {
    std::vector<Image> one(1);
    one[0] = std::move(img);
    std::vector<std::vector<InterestPoint>> r;
    try {
        r = calculateBatch(one);
    } catch (...) {
        img = std::move(one[0]);
        throw;
    }
    img = std::move(one[0]);
    return std::move(r[0]);
}

std::string formatResults(const std::vector<InterestPoint>& points) {
    std::ostringstream out;
    out << "Location\tscale\torientation\tdescriptors\n";
    for (const InterestPoint& p : points) {
        out << "[" << p.loc.x << ", " << p.loc.y << "]\t" << p.scale << "\t" << p.orientation << "\t" << "[";
        for (f32_t d : p.descriptors) out << d << ", ";
        out << "]\n";
    }
    return out.str();
}

void writeResults(const std::string& path, const std::vector<InterestPoint>& points) {
    std::ofstream out(path.c_str());
    out << formatResults(points);
    out.close();
}

}  // namespace sift
