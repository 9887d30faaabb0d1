// Host side of libsift_gpu.so: context, pyramid schedule, device-pass orchestration, the host
// replay of the reference's std::sort ordering, and the C ABI of include/sift_gpu.h.
//
// Path replaced: Sift::calculate (reference sift.cpp:19-57) and everything it calls.  There is no
// CPU fallback: every stage below runs as a CUDA kernel; the only host arithmetic is the Gaussian
// tap table (Vigra Kernel1D::initGaussian), the resize index maps, the
// nearest-Gaussian lookup (sift.cpp:205-218) and the std::sort permutation (sift.cpp:37,49).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <tuple>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sift_gpu.h"
#include "common.cuh"
#include "order_replay.h"
#include "pool.h"
#include "tma.cuh"

namespace siftgpu {

static thread_local std::string g_last_error;

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    g_last_error = buf;
    return SIFT_GPU_E_CUDA;
}

// ---- Vigra Kernel1D<float>::initGaussian (reference algorithms.cpp:13-14) -------------
static std::vector<float> gaussian_taps(float sigma, int* radius_out) {
    const double std_dev = (double)sigma;
    std::vector<float> taps;
    int radius = 0;
    if (std_dev > 0.0) {
        const float sf = (float)std_dev;
        const float s2 = (float)(-0.5 / (double)sf / (double)sf);
        const float nrm = (float)(0.3989422804014327 / (double)sf);
        radius = (int)(3.0 * std_dev + 0.5);
        if (radius == 0) radius = 1;
        for (float x = -(float)radius; x <= (float)radius; ++x) {
            const float x2 = x * x;
            taps.push_back(nrm * expf(x2 * s2));
        }
    } else {
        taps.push_back(1.0f);
    }
    float sum = 0.0f;
    for (float t : taps) sum += t;
    sum = 1.0f / sum;
    for (float& t : taps) t = t * sum;
    *radius_out = radius;
    return taps;
}

// ---- Vigra resizeImageNoInterpolation index walk --------------------------------------
static std::vector<int> resize_index_map(int n_old, int n_new) {
    std::vector<int> m((size_t)n_new);
    if (n_new == 1) { m[0] = 0; return m; }
    const double dx = (double)(n_old - 1) / (double)(n_new - 1);
    double x = 0.5;
    for (int i = 0; i < n_new; ++i, x += dx) m[(size_t)i] = (int)x;
    return m;
}

struct BlurSpec {
    float sigma = 0;
    int r = 0;
    size_t tap_off = 0;  // offset into the device tap pool
};

static int pitch_of(int w) { return (w + kPitchAlign - 1) / kPitchAlign * kPitchAlign; }

constexpr int kSlots = 6;            // most device passes in flight (default 5, SIFT_GPU_SLOTS): stage A of later passes (upload + pyramid
                                     // .. elimination) overlaps the host replay of pass k and stage B / download of earlier ones
constexpr uint32_t kSurvFirst = 8192; // survivors per image copied back speculatively with the counters (the rest on demand)

// The part of a plan that holds device addresses of one slot's buffers.
struct PlanSlot {
    CUtensorMap map_up[2], map_base[2], map_chain[kMaxOctaves][kMaxGauss][2], map_reduce[kMaxOctaves][2];  // [0]: 8-row box, [1]: 1-row box
    bool has_up = false, has_base = false, has_chain[kMaxOctaves][kMaxGauss] = {{false}}, has_reduce[kMaxOctaves] = {false};
    std::vector<ScanLayer> layers_host;
    ScanLayer* layers_dev = nullptr;
    std::vector<LevelRef> targets_host;
    LevelRef* targets_dev = nullptr;
};

struct Plan {
    int in_w = 0, in_h = 0, in_pitch = 0;
    int ow[kMaxOctaves] = {0}, oh[kMaxOctaves] = {0}, pitch[kMaxOctaves] = {0};
    int status = SIFT_GPU_OK;
    std::string why;
    int* d_maps = nullptr;  // all index maps, one allocation
    std::vector<int> h_maps; // host copy (launch geometry of the decimating kernel)
    size_t up_mx = 0, up_my = 0;
    size_t sel_x[kMaxOctaves] = {0}, sel_y[kMaxOctaves] = {0};  // decimation: inverse index maps (offsets into d_maps)
    int total_cols = 0;
    uint32_t mask_words = 0;
    std::vector<int> class_target;  // (octave*dpe + index) -> target slot
    std::vector<int> target_w, target_h;
    PlanSlot ps[kSlots];
};

}  // namespace siftgpu

using namespace siftgpu;

struct HostImageOut {
    std::vector<sift_gpu_keypoint> kps;
};

struct ReplayOut {
    int status = SIFT_GPU_OK;
    uint32_t n_survivors = 0;
    std::vector<sift_gpu_keypoint> kps;  // final vector order (orientation/descriptor filled later)
    std::vector<KeyIn> keys;             // the subset that goes to the device, same order
    std::vector<uint32_t> key_of;        // kp -> index in keys or ~0u
    // kept for the rare pass that has to be redone because a keypoint came back with several orientations
    std::vector<Surv> l1;                // the points after the first cleanup, vector order
    std::vector<uint32_t> inside;        // positions in l1 that pass the orientation-stage bounds test
    std::vector<uint32_t> kp_l1;         // kp -> position in l1
    bool truncated = false;              // the u16 size of the second cleanup dropped points
};

struct ChunkImage {
    int result_index;
    const sift_gpu_image* img;
};

// One device pass in flight.
struct Slot {
    cudaStream_t stream = nullptr;
    cudaStream_t aux[3]{};      // the pyramid of a pass runs as up to four image groups on stream + aux[] (tails overlap); 2 groups measured best
    cudaEvent_t ev_fork = nullptr, ev_join[3]{};
    cudaEvent_t ev[12]{};
    // device buffers
    uint8_t* d_in_u8 = nullptr;
    float* d_in = nullptr;
    float* d_up_tmp = nullptr;  // blur(img, 1.0) at input resolution
    float* d_up = nullptr;      // 2x image
    float* d_gauss[kMaxOctaves][kMaxGauss]{};
    float* d_dog[kMaxOctaves][kMaxGauss]{};
    uint32_t* d_mask = nullptr;
    uint32_t *d_col_count = nullptr, *d_col_off = nullptr;
    Cand* d_cands = nullptr;
    Surv* d_surv = nullptr;
    uint32_t *d_n_cand = nullptr, *d_n_surv = nullptr;
    uint32_t* d_slice = nullptr;   // survivor compaction scratch: B x kCompactSlices
    uint32_t *h_n_cand = nullptr, *h_n_surv = nullptr;  // pinned
    Surv* h_surv = nullptr;                              // pinned: B x kSurvFirst (speculative copy)
    std::vector<std::vector<Surv>> surv_overflow;        // per image, only when n_surv > kSurvFirst
    // host f32 frames of the next pass packed to bytes (pack_host.cpp) while the current passes run
    uint8_t* h_pack = nullptr;      // pinned, B x max_in_px, laid out like d_in_u8
    bool pack_job = false;          // a packing job for this slot is on the pack pool
    std::atomic<int> pack_bad{0};   // some pixel of the pass is not an exact 8-bit value: the pass travels as f32
    int pack_blocks = 1;            // work items per image of the job
    bool packed = false;            // stage A uploads the first n_packed frames from h_pack as bytes, the rest as they are
    int n_packed = 0;
    // keypoint stage buffers (grow on demand)
    size_t key_cap = 0;
    KeyIn* d_keys = nullptr; KeyIn* h_keys = nullptr;
    uint32_t* d_key_img = nullptr; uint32_t* h_key_img = nullptr;
    uint32_t* d_key_first = nullptr; uint32_t* h_key_first = nullptr;
    float* d_orient = nullptr; float* h_orient = nullptr;
    uint32_t* d_npeaks = nullptr; uint32_t* h_npeaks = nullptr;
    float* d_peaks = nullptr;
    float* d_desc = nullptr;
    float2* d_grad = nullptr;     // per key: the 256 (magnitude, orientation) pairs of its window (orientation kernel -> descriptor kernel)
    float* d_tables = nullptr;
    size_t tables_cap = 0;
    KeyGrid grid{};               // descriptor kernel's spatial index of the pass's keys (grown with the tables / the key capacity)
    size_t grid_words = 0, grid_keys = 0;
    // state of the pass in flight
    Plan* plan = nullptr;
    std::vector<ChunkImage> imgs;
    std::vector<ReplayOut> rep;
    size_t n_keys = 0;
    float* h_desc = nullptr;
    double t_replay0 = 0.0;
    uint64_t launches = 0;
    // stage A between upload and the survivor copy-back as an instantiated CUDA graph per (plan, images in the pass, input dtype)
    struct StageGraph { cudaGraphExec_t exec = nullptr; uint64_t launches = 0; int seen = 0; };
    std::map<std::tuple<const void*, int, int>, StageGraph> graphs;
    bool busy = false;
};

struct sift_gpu_ctx {
    sift_gpu_params prm{};
    int O = 0, D = 0, G = 0;
    bool fma = false;
    std::string error;

    // schedule (size independent)
    float g_scale[kMaxOctaves][kMaxGauss]{};
    float d_scale[kMaxOctaves][kMaxGauss]{};
    BlurSpec base_blur, up_blur, w16_blur;
    BlurSpec chain_blur[kMaxOctaves][kMaxGauss];
    BlurSpec reduce_blur[kMaxOctaves];
    float* d_taps = nullptr;
    std::vector<float> h_taps;
    int dead_blur_r[kMaxOctaves][kMaxGauss]{};
    bool top_needed[kMaxOctaves]{};  // g(o, D) is somebody's nearest Gaussian (sift.cpp:205-218); otherwise nothing ever reads it and it is not stored

    // sizing
    int B = 1;
    int n_slots = 1;
    int max_in_w = 0, max_in_h = 0, max_in_pitch = 0;
    size_t max_in_px = 0;       // per-image stride of the input staging buffers (pitched)
    size_t maxP[kMaxOctaves]{};
    size_t cand_cap = 0, mask_cap = 0, col_cap = 0;

    Slot slots[kSlots];
    int last_slot = -1;  // slot of the most recent pass (stage-level getters read it)

    // per-run result storage (pinned descriptor blocks + host vectors)
    std::vector<float*> desc_blocks;
    std::vector<size_t> desc_block_cap;
    size_t desc_blocks_used = 0;
    std::vector<HostImageOut> outs;

    std::map<std::pair<int, int>, Plan*> plans;
    Pool* pool = nullptr;
    Pool* pack_pool = nullptr;   // second pool: packing the next pass's frames overlaps the order replay of an earlier one
    int host_threads = 0;

    sift_gpu_timings tm{};
    cudaEvent_t ev_first = nullptr, ev_last = nullptr;
};

static int set_error(sift_gpu_ctx* c, int code, const std::string& msg) {
    if (c) c->error = msg;
    g_last_error = msg;
    return code;
}

#define CTX_TRY(expr)                                       \
    do {                                                    \
        int _rc = (expr);                                   \
        if (_rc != 0) { if (c) c->error = g_last_error; return _rc; } \
    } while (0)
#define CTX_CUDA(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) { int _rc = cuda_fail(_e, #expr, __FILE__, __LINE__); if (c) c->error = g_last_error; return _rc; } \
    } while (0)

static size_t level_px(int w, int h) { return (size_t)w * (size_t)h; }

// Schedule of Sift::_createDOGs (sift.cpp:381-417): scale labels, radii, taps.  Host arithmetic only (no device is touched:
// sift_gpu_debug_host_replay runs it on machines without one).
static void build_schedule_host(sift_gpu_ctx* c) {
    const int O = c->O, D = c->D;
    std::vector<float> pool;
    auto add = [&](float sigma) {
        BlurSpec b;
        b.sigma = sigma;
        std::vector<float> t = gaussian_taps(sigma, &b.r);
        while (pool.size() % 4) pool.push_back(0.0f);
        b.tap_off = pool.size();
        pool.insert(pool.end(), t.begin(), t.end());
        return b;
    };
    c->base_blur = add(c->prm.sigma);
    c->up_blur = add(1.0f);     // sift.cpp:21
    c->w16_blur = add(1.6f);    // sift.cpp:87
    c->g_scale[0][0] = c->prm.sigma;
    uint16_t exp = 0;
    for (int i = 0; i < O; i++) {
        for (int j = 1; j < D + 1; j++) {
            const float scale = (float)(std::pow((double)c->prm.k, (double)exp) * (double)c->prm.sigma);
            c->g_scale[i][j] = scale;
            c->chain_blur[i][j] = add(scale);
            c->d_scale[i][j - 1] = c->g_scale[i][j] - c->g_scale[i][j - 1];
            exp++;
        }
        if (i < O - 1) {
            c->g_scale[i + 1][0] = c->g_scale[i][D - 1];
            c->reduce_blur[i] = add(c->g_scale[i][D - 1]);
            exp -= 2;
        }
    }
    for (int e = 0; e < O; ++e)
        for (int i = 0; i < D; ++i) {
            int r = 0;
            (void)gaussian_taps((float)(1.5 * (double)c->d_scale[e][i]), &r);  // sift.cpp:184
            c->dead_blur_r[e][i] = r;
        }
    c->h_taps = pool;
}

static int build_schedule(sift_gpu_ctx* c) {
    build_schedule_host(c);
    CTX_CUDA(cudaMalloc(&c->d_taps, sizeof(float) * c->h_taps.size()));
    CTX_CUDA(cudaMemcpy(c->d_taps, c->h_taps.data(), sizeof(float) * c->h_taps.size(), cudaMemcpyHostToDevice));
    return 0;
}

static void octave_dims(const sift_gpu_ctx* c, int in_w, int in_h, int* ow, int* oh) {
    int w = c->prm.subpixel ? in_w * 2 : in_w, h = c->prm.subpixel ? in_h * 2 : in_h;
    for (int o = 0; o < c->O; ++o) {
        ow[o] = w; oh[o] = h;
        w = (w + 1) / 2; h = (h + 1) / 2;
    }
}

// Sift::_findNearestGaussian (sift.cpp:205-218).
static void nearest_gaussian(const sift_gpu_ctx* c, float scale, int* o_out, int* i_out) {
    float lowest = 100;
    int bo = 0, bi = 0;
    for (int o = 0; o < c->O; o++)
        for (int i = 0; i < c->G; i++) {
            const float cur = std::abs(c->g_scale[o][i] - scale);
            if (cur < lowest) { lowest = cur; bo = o; bi = i; }
        }
    *o_out = bo; *i_out = bi;
}

// Nearest-Gaussian targets per keypoint class (identical for every slot; only the base pointers differ).  Host arithmetic only.
static void fill_class_targets(sift_gpu_ctx* c, Plan* p, std::vector<std::pair<int, int>>* target_level_out) {
    const int O = c->O, D = c->D;
    p->class_target.assign((size_t)(O * D), -1);
    std::vector<std::pair<int, int>>& target_level = *target_level_out;
    for (int e = 0; e < O; ++e)
        for (int i = 1; i < D - 1; ++i) {
            int to, ti;
            nearest_gaussian(c, c->d_scale[e][i], &to, &ti);
            int slot = -1;
            for (size_t s = 0; s < target_level.size(); ++s)
                if (target_level[s] == std::make_pair(to, ti)) slot = (int)s;
            if (slot < 0) {
                slot = (int)target_level.size();
                target_level.push_back(std::make_pair(to, ti));
                p->target_w.push_back(p->ow[to]);
                p->target_h.push_back(p->oh[to]);
            }
            p->class_target[(size_t)(e * D + i)] = slot;
            if (ti == D) c->top_needed[to] = true;
        }
}

static Plan* get_plan(sift_gpu_ctx* c, int in_w, int in_h) {
    auto key = std::make_pair(in_w, in_h);
    auto it = c->plans.find(key);
    if (it != c->plans.end()) return it->second;
    Plan* p = new Plan();
    c->plans[key] = p;
    p->in_w = in_w; p->in_h = in_h; p->in_pitch = pitch_of(in_w);
    octave_dims(c, in_w, in_h, p->ow, p->oh);
    const int O = c->O, D = c->D;
    for (int o = 0; o < O; ++o) p->pitch[o] = pitch_of(p->ow[o]);

    auto fail = [&](int code, const char* why) { if (p->status == SIFT_GPU_OK) { p->status = code; p->why = why; } };
    auto check_blur = [&](const BlurSpec& b, int w, int h) {
        if (w < b.r + 1 || h < b.r + 1) fail(SIFT_GPU_E_PRECONDITION, "separableConvolveX/Y(): kernel longer than line");
        if (b.r > max_generic_radius() && !stream_box_width(b.r, false)) fail(SIFT_GPU_E_UNSUPPORTED, "blur radius exceeds the tile kernel's shared memory");
    };
    if (in_w < 1 || in_h < 1) fail(SIFT_GPU_E_INVALID, "empty image");
    if (c->prm.subpixel) {
        check_blur(c->up_blur, in_w, in_h);
        if (!(in_w > 1 && in_h > 1)) fail(SIFT_GPU_E_PRECONDITION, "resizeImageNoInterpolation(): Source image too small.");
    }
    check_blur(c->base_blur, p->ow[0], p->oh[0]);
    for (int o = 0; o < O; ++o) {
        for (int j = 1; j <= D; ++j) check_blur(c->chain_blur[o][j], p->ow[o], p->oh[o]);
        if (o < O - 1) {
            check_blur(c->reduce_blur[o], p->ow[o], p->oh[o]);
            if (!(p->ow[o] > 1 && p->oh[o] > 1)) fail(SIFT_GPU_E_PRECONDITION, "resizeImageNoInterpolation(): Source image too small.");
            if (!(p->ow[o + 1] > 1 && p->oh[o + 1] > 1)) fail(SIFT_GPU_E_PRECONDITION, "resizeImageNoInterpolation(): Destination image too small.");
        }
    }
    if (p->status != SIFT_GPU_OK) return p;

    // index maps
    std::vector<int> maps;
    auto push_map = [&](int n_old, int n_new) {
        size_t off = maps.size();
        std::vector<int> m = resize_index_map(n_old, n_new);
        maps.insert(maps.end(), m.begin(), m.end());
        return off;
    };
    auto push_inverse = [&](int n_old, int n_new) {  // source index -> destination index, or -1
        size_t off = maps.size();
        std::vector<int> m = resize_index_map(n_old, n_new), inv((size_t)n_old, -1);
        for (int i = 0; i < n_new; ++i) inv[(size_t)m[(size_t)i]] = i;
        maps.insert(maps.end(), inv.begin(), inv.end());
        return off;
    };
    if (c->prm.subpixel) { p->up_mx = push_map(in_w, in_w * 2); p->up_my = push_map(in_h, in_h * 2); }
    for (int o = 0; o + 1 < O; ++o) { p->sel_x[o] = push_inverse(p->ow[o], p->ow[o + 1]); p->sel_y[o] = push_inverse(p->oh[o], p->oh[o + 1]); }
    p->h_maps = maps;
    if (!maps.empty()) {
        if (cudaMalloc(&p->d_maps, sizeof(int) * maps.size()) != cudaSuccess ||
            cudaMemcpy(p->d_maps, maps.data(), sizeof(int) * maps.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
            fail(SIFT_GPU_E_CUDA, "index map upload failed");
            return p;
        }
    }
    std::vector<std::pair<int, int>> target_level;
    fill_class_targets(c, p, &target_level);
    for (int si = 0; si < c->n_slots; ++si) {
        Slot& S = c->slots[si];
        PlanSlot& ps = p->ps[si];
        // extrema scan layers, ordered (octave, index)
        uint32_t mask_off = 0, col_base = 0;
        for (int e = 0; e < O; ++e)
            for (int i = 1; i < D - 1; ++i) {
                ScanLayer L{};
                L.d0 = S.d_dog[e][i - 1]; L.d1 = S.d_dog[e][i]; L.d2 = S.d_dog[e][i + 1];
                L.stride = c->maxP[e];
                L.pitch = p->pitch[e];
                L.w = p->ow[e]; L.h = p->oh[e];
                L.n_yw = (L.h + 31) / 32;
                L.mask_off = mask_off; L.col_base = col_base;
                L.octave = (uint8_t)e; L.index = (uint8_t)i;
                mask_off += (uint32_t)L.n_yw * (uint32_t)L.w;
                col_base += (uint32_t)L.w;
                ps.layers_host.push_back(L);
            }
        p->mask_words = mask_off;
        p->total_cols = (int)col_base;
        set_scan_tiles(ps.layers_host.data(), (int)ps.layers_host.size());
        for (auto& tl : target_level)
            ps.targets_host.push_back(LevelRef{S.d_gauss[tl.first][tl.second], c->maxP[tl.first], p->pitch[tl.first], p->ow[tl.first], p->oh[tl.first]});
        if (cudaMalloc(&ps.layers_dev, sizeof(ScanLayer) * ps.layers_host.size()) != cudaSuccess ||
            cudaMemcpy(ps.layers_dev, ps.layers_host.data(), sizeof(ScanLayer) * ps.layers_host.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMalloc(&ps.targets_dev, sizeof(LevelRef) * ps.targets_host.size()) != cudaSuccess ||
            cudaMemcpy(ps.targets_dev, ps.targets_host.data(), sizeof(LevelRef) * ps.targets_host.size(), cudaMemcpyHostToDevice) != cudaSuccess)
            fail(SIFT_GPU_E_CUDA, "plan upload failed");
        // TMA descriptors (box width depends on the blur radius); a missing descriptor only means the generic kernel runs
        auto mk = [&](CUtensorMap* m, const float* base, int w, int h, int pitch, size_t stride, int r, bool decimate = false) {
            const int bw = stream_box_width(r, decimate);
            return bw > 0 && tma::make_image_map(&m[0], base, w, h, c->B, (size_t)pitch, stride, bw, stream_box_rows()) &&
                   tma::make_image_map(&m[1], base, w, h, c->B, (size_t)pitch, stride, bw, 1);
        };
        if (c->prm.subpixel) {
            ps.has_up = mk(ps.map_up, S.d_in, in_w, in_h, p->in_pitch, c->max_in_px, c->up_blur.r);
            ps.has_base = mk(ps.map_base, S.d_up, p->ow[0], p->oh[0], p->pitch[0], c->maxP[0], c->base_blur.r);
        } else {
            ps.has_base = mk(ps.map_base, S.d_in, in_w, in_h, p->in_pitch, c->max_in_px, c->base_blur.r);
        }
        for (int o = 0; o < O; ++o) {
            for (int j = 1; j <= D; ++j)
                ps.has_chain[o][j] = mk(ps.map_chain[o][j], S.d_gauss[o][j - 1], p->ow[o], p->oh[o], p->pitch[o], c->maxP[o], c->chain_blur[o][j].r);
            if (o < O - 1)
                ps.has_reduce[o] = mk(ps.map_reduce[o], S.d_gauss[o][D - 1], p->ow[o], p->oh[o], p->pitch[o], c->maxP[o], c->reduce_blur[o].r, true);
        }
    }
    return p;
}

static int alloc_buffers(sift_gpu_ctx* c) {
    const int O = c->O, D = c->D, B = c->B;
    int ow[kMaxOctaves], oh[kMaxOctaves];
    octave_dims(c, c->max_in_w, c->max_in_h, ow, oh);
    c->max_in_pitch = pitch_of(c->max_in_w);
    c->max_in_px = level_px(c->max_in_pitch, c->max_in_h);
    size_t cand_cap = 0, mask_cap = 0, col_cap = 0;
    for (int o = 0; o < O; ++o) {
        c->maxP[o] = level_px(pitch_of(ow[o]), oh[o]);
        cand_cap += level_px(ow[o], oh[o]) * (size_t)(D - 2);
        mask_cap += (size_t)((oh[o] + 31) / 32 + 1) * (size_t)ow[o] * (size_t)(D - 2);
        col_cap += (size_t)ow[o] * (size_t)(D - 2);
    }
    c->cand_cap = cand_cap; c->mask_cap = mask_cap; c->col_cap = col_cap;
    for (int si = 0; si < c->n_slots; ++si) {
        Slot& S = c->slots[si];
        CTX_CUDA(cudaStreamCreateWithFlags(&S.stream, cudaStreamNonBlocking));
        for (auto& st : S.aux) CTX_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        CTX_CUDA(cudaEventCreateWithFlags(&S.ev_fork, cudaEventDisableTiming));
        for (auto& e : S.ev_join) CTX_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        // the host waits on these: block instead of spinning, a spinning waiter per GPU starves the replay workers of the
        // other ranks on a host with few cores per GPU
        for (auto& e : S.ev) CTX_CUDA(cudaEventCreateWithFlags(&e, cudaEventBlockingSync));
        CTX_CUDA(cudaMalloc(&S.d_in_u8, c->max_in_px * (size_t)B));
        CTX_CUDA(cudaMalloc(&S.d_in, sizeof(float) * c->max_in_px * (size_t)B));
        if (c->prm.subpixel) {
            CTX_CUDA(cudaMalloc(&S.d_up_tmp, sizeof(float) * c->max_in_px * (size_t)B));
            CTX_CUDA(cudaMalloc(&S.d_up, sizeof(float) * c->maxP[0] * (size_t)B));
        }
        for (int o = 0; o < O; ++o) {
            for (int i = 0; i <= D; ++i) CTX_CUDA(cudaMalloc(&S.d_gauss[o][i], sizeof(float) * c->maxP[o] * (size_t)B));
            for (int i = 0; i < D; ++i) CTX_CUDA(cudaMalloc(&S.d_dog[o][i], sizeof(float) * c->maxP[o] * (size_t)B));
        }
        CTX_CUDA(cudaMalloc(&S.d_mask, sizeof(uint32_t) * mask_cap * (size_t)B * 2));  // candidate plane + cheap-test plane
        CTX_CUDA(cudaMalloc(&S.d_col_count, sizeof(uint32_t) * col_cap * (size_t)B));
        CTX_CUDA(cudaMalloc(&S.d_col_off, sizeof(uint32_t) * col_cap * (size_t)B));
        CTX_CUDA(cudaMalloc(&S.d_cands, sizeof(Cand) * cand_cap * (size_t)B));
        CTX_CUDA(cudaMalloc(&S.d_surv, sizeof(Surv) * cand_cap * (size_t)B));
        CTX_CUDA(cudaMalloc(&S.d_n_cand, sizeof(uint32_t) * (size_t)B));
        CTX_CUDA(cudaMalloc(&S.d_n_surv, sizeof(uint32_t) * (size_t)B));
        CTX_CUDA(cudaMalloc(&S.d_slice, sizeof(uint32_t) * (size_t)B * kCompactSlices));
        CTX_CUDA(cudaHostAlloc(&S.h_n_cand, sizeof(uint32_t) * (size_t)B, cudaHostAllocDefault));
        CTX_CUDA(cudaHostAlloc(&S.h_n_surv, sizeof(uint32_t) * (size_t)B, cudaHostAllocDefault));
        CTX_CUDA(cudaHostAlloc(&S.h_surv, sizeof(Surv) * (size_t)kSurvFirst * (size_t)B, cudaHostAllocDefault));
        CTX_CUDA(cudaMalloc(&S.d_key_first, sizeof(uint32_t) * (size_t)(B + 1)));
        CTX_CUDA(cudaHostAlloc(&S.h_key_first, sizeof(uint32_t) * (size_t)(B + 1), cudaHostAllocDefault));
        S.surv_overflow.resize((size_t)B);
    }
    CTX_CUDA(cudaEventCreate(&c->ev_first));
    CTX_CUDA(cudaEventCreateWithFlags(&c->ev_last, cudaEventBlockingSync));
    return 0;
}

static int ensure_key_capacity(sift_gpu_ctx* c, Slot& S, size_t n) {
    if (n <= S.key_cap) return 0;
    size_t cap = std::max<size_t>(n * 3 / 2, 4096);
    cudaFree(S.d_keys); cudaFree(S.d_key_img); cudaFree(S.d_orient); cudaFree(S.d_npeaks); cudaFree(S.d_peaks); cudaFree(S.d_desc); cudaFree(S.d_grad);
    cudaFreeHost(S.h_keys); cudaFreeHost(S.h_key_img); cudaFreeHost(S.h_orient); cudaFreeHost(S.h_npeaks);
    S.key_cap = 0;
    CTX_CUDA(cudaMalloc(&S.d_keys, sizeof(KeyIn) * cap));
    CTX_CUDA(cudaMalloc(&S.d_key_img, sizeof(uint32_t) * cap));
    CTX_CUDA(cudaMalloc(&S.d_orient, sizeof(float) * cap));
    CTX_CUDA(cudaMalloc(&S.d_npeaks, sizeof(uint32_t) * cap));
    CTX_CUDA(cudaMalloc(&S.d_peaks, sizeof(float) * 36 * cap));
    CTX_CUDA(cudaMalloc(&S.d_desc, sizeof(float) * kDescLen * cap));
    CTX_CUDA(cudaMalloc(&S.d_grad, sizeof(float2) * 256 * cap));
    CTX_CUDA(cudaHostAlloc(&S.h_keys, sizeof(KeyIn) * cap, cudaHostAllocDefault));
    CTX_CUDA(cudaHostAlloc(&S.h_key_img, sizeof(uint32_t) * cap, cudaHostAllocDefault));
    CTX_CUDA(cudaHostAlloc(&S.h_orient, sizeof(float) * cap, cudaHostAllocDefault));
    CTX_CUDA(cudaHostAlloc(&S.h_npeaks, sizeof(uint32_t) * cap, cudaHostAllocDefault));
    S.key_cap = cap;
    return 0;
}

static float* take_desc_block(sift_gpu_ctx* c, size_t floats) {
    if (floats == 0) floats = 1;
    if (c->desc_blocks_used < c->desc_blocks.size()) {
        size_t i = c->desc_blocks_used;
        if (c->desc_block_cap[i] < floats) {
            cudaFreeHost(c->desc_blocks[i]);
            c->desc_blocks[i] = nullptr;
            size_t cap = floats * 3 / 2;
            if (cudaHostAlloc(&c->desc_blocks[i], sizeof(float) * cap, cudaHostAllocDefault) != cudaSuccess) return nullptr;
            c->desc_block_cap[i] = cap;
        }
        ++c->desc_blocks_used;
        return c->desc_blocks[i];
    }
    float* p = nullptr;
    size_t cap = floats * 3 / 2;
    if (cudaHostAlloc(&p, sizeof(float) * cap, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    c->desc_blocks.push_back(p);
    c->desc_block_cap.push_back(cap);
    ++c->desc_blocks_used;
    return p;
}

// ---- device passes ------------------------------------------------------------------------------
__global__ void fetch_keys_kernel(const uint2* __restrict__ h_keys, uint2* __restrict__ d_keys, const uint32_t* __restrict__ h_img,
                                  uint32_t* __restrict__ d_img, const uint32_t* __restrict__ h_first, uint32_t* __restrict__ d_first,
                                  uint32_t n_keys, uint32_t n_first) {
    static_assert(sizeof(KeyIn) == sizeof(uint2), "KeyIn is copied as uint2");
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_keys; i += stride) {
        d_keys[i] = h_keys[i];
        d_img[i] = h_img[i];
    }
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_first; i += stride) d_first[i] = h_first[i];
}

static BlurArgs blur_args(const sift_gpu_ctx* c, const BlurSpec& b, const float* src, size_t sstride, int spitch, float* dst, size_t dstride,
                          int dpitch, float* dog, size_t gstride, int gpitch, int w, int h, const CUtensorMap* map) {
    BlurArgs a{};
    a.src = src; a.dst = dst; a.dog = dog;
    a.src_stride = sstride; a.dst_stride = dstride; a.dog_stride = gstride;
    a.src_pitch = spitch; a.dst_pitch = dpitch; a.dog_pitch = gpitch;
    a.w = w; a.h = h; a.taps = c->d_taps + b.tap_off; a.taps_host = c->h_taps.data() + b.tap_off; a.r = b.r;
    a.map = map;
    a.map_box = 0;   // set by the caller that knows which box the descriptors were encoded with
    return a;
}

static bool keep_top_levels() {
    static const bool v = getenv("SIFT_GPU_KEEP_TOP_LEVELS") != nullptr;
    return v;
}

// Sift::calculate's upsample (sift.cpp:20-21) + _createDOGs (sift.cpp:381-417) for the images [z0, z0 + cnt) of the pass.
static int run_pyramid_group(sift_gpu_ctx* c, Slot& S, const Plan* p, const PlanSlot& ps, int z0, int cnt, int share, cudaStream_t s) {
    const int O = c->O, D = c->D;
    uint64_t* L = &S.launches;
    // SIFT_GPU_TRACE_PYR: per-launch device times of the pyramid in situ (warm L2, unlike ncu's replays)
    static const bool trace = getenv("SIFT_GPU_TRACE_PYR") != nullptr;
    std::vector<cudaEvent_t> tev;
    std::vector<std::string> tname;
    auto mark = [&](const char* what, int r, int w, int h) {
        if (!trace) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, s);
        tev.push_back(e);
        char buf[64];
        snprintf(buf, sizeof buf, "%s r=%d %dx%d", what, r, w, h);
        tname.push_back(buf);
    };
    auto launch = [&](BlurArgs a) {
        a.z0 = z0;
        a.share = share;
        a.map_box = a.map ? stream_box_width(a.r, a.sel_x != nullptr) : 0;
        mark(a.sel_x ? "reduce" : (a.dog ? "blur+dog" : "blur"), a.r, a.w, a.h);
        return launch_blur(a, cnt, c->fma, s, L);
    };
    const float* base_src = S.d_in;
    size_t base_stride = c->max_in_px;
    int base_pitch = p->in_pitch;
    if (c->prm.subpixel) {
        CTX_TRY(launch(blur_args(c, c->up_blur, S.d_in, c->max_in_px, p->in_pitch, S.d_up_tmp, c->max_in_px, p->in_pitch, nullptr, 0, 0,
                                 p->in_w, p->in_h, ps.has_up ? ps.map_up : nullptr)));
        mark("resize x2", 0, p->ow[0], p->oh[0]);
        CTX_TRY(launch_resize_nn(S.d_up_tmp, c->max_in_px, p->in_pitch, S.d_up, c->maxP[0], p->pitch[0], p->ow[0], p->oh[0],
                                 p->d_maps + p->up_mx, p->d_maps + p->up_my, z0, cnt, s, L));
        base_src = S.d_up;
        base_stride = c->maxP[0];
        base_pitch = p->pitch[0];
    }
    CTX_TRY(launch(blur_args(c, c->base_blur, base_src, base_stride, base_pitch, S.d_gauss[0][0], c->maxP[0], p->pitch[0], nullptr, 0, 0,
                             p->ow[0], p->oh[0], ps.has_base ? ps.map_base : nullptr)));
    for (int o = 0; o < O; ++o) {
        for (int j = 1; j <= D; ++j) {
            // the top Gaussian of an octave is only stored when a keypoint class takes it as its nearest Gaussian (octaves >= 5 with the
            // default schedule); its DoG is all the pipeline needs otherwise (sift_gpu_debug_get_level recomputes it on demand)
            float* g_out = (j == D && !c->top_needed[o] && !keep_top_levels()) ? nullptr : S.d_gauss[o][j];
            CTX_TRY(launch(blur_args(c, c->chain_blur[o][j], S.d_gauss[o][j - 1], c->maxP[o], p->pitch[o], g_out, c->maxP[o],
                                     p->pitch[o], S.d_dog[o][j - 1], c->maxP[o], p->pitch[o], p->ow[o], p->oh[o],
                                     ps.has_chain[o][j] ? ps.map_chain[o][j] : nullptr)));
        }
        if (o < O - 1) {
            // alg::reduceToNextLevel: blur with the level's own label sigma, keep only the pixels the resize picks
            BlurArgs a = blur_args(c, c->reduce_blur[o], S.d_gauss[o][D - 1], c->maxP[o], p->pitch[o], S.d_gauss[o + 1][0], c->maxP[o + 1],
                                   p->pitch[o + 1], nullptr, 0, 0, p->ow[o], p->oh[o], ps.has_reduce[o] ? ps.map_reduce[o] : nullptr);
            a.sel_x = p->d_maps + p->sel_x[o];
            a.sel_y = p->d_maps + p->sel_y[o];
            a.sel_x_host = p->h_maps.data() + p->sel_x[o];
            a.sel_y_host = p->h_maps.data() + p->sel_y[o];
            CTX_TRY(launch(a));
        }
    }
    if (trace) {
        mark("end", 0, 0, 0);
        cudaEventSynchronize(tev.back());
        for (size_t i = 0; i + 1 < tev.size(); ++i) {
            float ms = 0.0f;
            cudaEventElapsedTime(&ms, tev[i], tev[i + 1]);
            fprintf(stderr, "[pyr z0=%d n=%d] %-24s %8.1f us\n", z0, cnt, tname[i].c_str(), ms * 1e3f);
        }
        for (cudaEvent_t e : tev) cudaEventDestroy(e);
    }
    return 0;
}

// The images of a pass are independent, so their pyramids run as up to four groups on parallel streams: while one group's
// kernel drains its last CTAs the other groups' kernels keep the SMs busy.
static int run_pyramid(sift_gpu_ctx* c, Slot& S, const Plan* p, const PlanSlot& ps, int nb) {
    static const int max_groups = [] { const char* e = getenv("SIFT_GPU_PYR_GROUPS"); int g = e ? atoi(e) : 2; return g < 1 ? 1 : (g > 4 ? 4 : g); }();
    const int groups = std::min(max_groups, nb >= 8 ? 4 : (nb >= 2 ? 2 : 1));
    if (groups == 1) return run_pyramid_group(c, S, p, ps, 0, nb, 1, S.stream);
    CTX_CUDA(cudaEventRecord(S.ev_fork, S.stream));
    int z0 = 0;
    for (int g = 0; g < groups; ++g) {
        const int cnt = nb / groups + (g < nb % groups ? 1 : 0);
        cudaStream_t s = g == 0 ? S.stream : S.aux[g - 1];
        if (g > 0) CTX_CUDA(cudaStreamWaitEvent(s, S.ev_fork, 0));
        CTX_TRY(run_pyramid_group(c, S, p, ps, z0, cnt, groups, s));
        if (g > 0) {
            CTX_CUDA(cudaEventRecord(S.ev_join[g - 1], s));
            CTX_CUDA(cudaStreamWaitEvent(S.stream, S.ev_join[g - 1], 0));
        }
        z0 += cnt;
    }
    return 0;
}

// The reference's cleanup (sift.cpp:37-42) of an n-element vector whose unfiltered elements sit at the ascending positions
// zero_pos[0, m): returns the kept elements, as indices into zero_pos, in their post-sort order (valid until the sorter's next
// run).  The std::sort permutation is replayed by SparseFilterSort (order_replay.h); the canonical mode keeps the original
// order (what a stable partition would do).  *kept_n = (uint16_t)m: the u16 size = distance(begin, first filtered) (sift.cpp:41).
static const uint32_t* cleanup_order(SparseFilterSort& sorter, std::vector<uint32_t>& identity, uint32_t n, const uint32_t* zero_pos, size_t m,
                                     bool canonical, size_t* kept_n) {
    *kept_n = (size_t)(uint16_t)m;
    if (canonical) {
        const size_t have = identity.size();
        if (have < m) {
            identity.resize(m);
            for (size_t i = have; i < m; ++i) identity[i] = (uint32_t)i;
        }
        return identity.data();
    }
    return sorter.run(n, zero_pos, m).data();
}

static std::atomic<long> g_rep_ns[6];  // SIFT_GPU_TRACE: CPU time inside replay_image by phase (summed over worker threads)
static const bool g_rep_trace = getenv("SIFT_GPU_TRACE") != nullptr || getenv("SIFT_GPU_REPLAY_REPS") != nullptr;

// Fills the result record of a point that reached _createDecriptors; returns whether it gets a descriptor.
static inline bool make_keypoint(const sift_gpu_ctx* c, const Plan* p, const Surv& s, sift_gpu_keypoint* k, KeyIn* ki) {
    const int slot = p->class_target[(size_t)(s.octave * c->D + s.index)];
    const int tw = p->target_w[(size_t)slot], th = p->target_h[(size_t)slot];
    k->x = s.x; k->y = s.y; k->octave = s.octave; k->index = s.index;
    k->scale = c->d_scale[s.octave][s.index];
    k->orientation = 0.0f;
    k->reserved = 0;
    // _createDecriptors bounds test (sift.cpp:65-70)
    const bool reject = s.x < kRegion || s.x > tw - kRegion || s.y < kRegion || s.y > th - kRegion;
    k->filtered = reject ? 1 : 0;
    k->desc_len = reject ? 0 : kDescLen;
    ki->x = s.x; ki->y = s.y; ki->octave = s.octave; ki->index = s.index; ki->tgt = (uint8_t)slot; ki->pad = 0;
    return !reject;
}

// Host half of Sift::calculate between _eliminateEdgeResponses and _createDecriptors (sift.cpp:37-55).  Runs once per image on
// the worker pool; `out` is reused from pass to pass (every field is rewritten here, the vectors keep their capacity) and the
// scratch arrays are per thread, so the steady state allocates nothing.
static void replay_image(const sift_gpu_ctx* c, const Plan* p, uint32_t n_cand, const Surv* S, uint32_t n_surv, ReplayOut* out) {
    const bool canonical = (c->prm.flags & SIFT_GPU_FLAG_ORDER_CANONICAL) != 0;
    const bool strict = (c->prm.flags & SIFT_GPU_FLAG_STRICT) != 0;
    const int D = c->D;
    static thread_local SparseFilterSort sorter;
    static thread_local std::vector<uint32_t> zero_pos, identity;
    std::chrono::steady_clock::time_point t_ph;
    if (g_rep_trace) t_ph = std::chrono::steady_clock::now();
    auto lap = [&](int i) {
        if (!g_rep_trace) return;
        const auto n = std::chrono::steady_clock::now();
        g_rep_ns[i] += std::chrono::duration_cast<std::chrono::nanoseconds>(n - t_ph).count();
        t_ph = n;
    };
    out->status = SIFT_GPU_OK;
    out->truncated = false;
    // first cleanup over all candidates: the device compacts survivors in canonical order (survivors_count / offsets / write
    // kernels, eliminate.cu), so `canon` ascends
    zero_pos.resize(n_surv);
    for (uint32_t s = 0; s < n_surv; ++s) zero_pos[s] = S[s].canon;
    size_t n1 = 0;
    const uint32_t* L1 = cleanup_order(sorter, identity, n_cand, zero_pos.data(), n_surv, canonical, &n1);  // survivor slots in vector order
    out->n_survivors = (uint32_t)n1;
    out->l1.resize(n1);
    Surv* l1 = out->l1.data();
    for (size_t i = 0; i < n1; ++i) l1[i] = S[L1[i]];
    lap(1);
    // _orientationAssignment bounds test (sift.cpp:173-178) and the dead blur's precondition (sift.cpp:184); both depend on the
    // point's class (octave, index) only through two limits and one flag
    int lim_x[kMaxOctaves * kMaxGauss], lim_y[kMaxOctaves * kMaxGauss];
    bool throws[kMaxOctaves * kMaxGauss];
    for (int cls = 0; cls < c->O * D; ++cls) {
        const int slot = p->class_target[(size_t)cls];
        lim_x[cls] = slot < 0 ? 0 : p->target_w[(size_t)slot] - kRegion;
        lim_y[cls] = slot < 0 ? 0 : p->target_h[(size_t)slot] - kRegion;
        throws[cls] = strict && 2 * kRegion < c->dead_blur_r[cls / D][cls % D] + 1;
    }
    out->inside.resize(n1 + 1);  // positions in l1 that pass the bounds test (compacted without branching: one slot of slack)
    uint32_t* inside = out->inside.data();
    size_t n_in = 0;
    bool thrown = false;
    for (size_t i = 0; i < n1; ++i) {
        const Surv& s = l1[i];
        const int cls = s.octave * D + s.index;
        const bool in = !((s.x < kRegion || s.x >= lim_x[cls]) || (s.y < kRegion || s.y >= lim_y[cls]));
        inside[n_in] = (uint32_t)i;
        n_in += in ? 1 : 0;
        thrown |= in && throws[cls];
    }
    out->inside.resize(n_in);
    if (thrown) {
        out->status = SIFT_GPU_E_PRECONDITION;
        return;
    }
    lap(2);
    size_t n2 = 0;
    const uint32_t* L2 = cleanup_order(sorter, identity, (uint32_t)n1, inside, n_in, canonical, &n2);  // indices into `inside`
    lap(3);
    out->truncated = n2 != n_in;
    out->kps.resize(n2);
    out->key_of.resize(n2);
    out->kp_l1.resize(n2);
    out->keys.resize(n2);
    sift_gpu_keypoint* kps = out->kps.data();
    KeyIn* keys = out->keys.data();
    size_t n_keys = 0;
    for (size_t i = 0; i < n2; ++i) {
        const uint32_t q = inside[L2[i]];
        out->kp_l1[i] = q;
        const bool has = make_keypoint(c, p, l1[q], &kps[i], &keys[n_keys]);
        out->key_of[i] = has ? (uint32_t)n_keys : ~0u;
        n_keys += has ? 1 : 0;
    }
    out->keys.resize(n_keys);
    lap(4);
}

static double g_trace[8];  // SIFT_GPU_TRACE: host wall time per phase of the pass loop (diagnostics only)
static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Stage A: upload, pyramid, extrema, elimination, counters + first survivors back.  Asynchronous.
static int enqueue_stage_a(sift_gpu_ctx* c, Slot& S, int slot_index) {
    Plan* p = S.plan;
    const PlanSlot& ps = p->ps[slot_index];
    const int nb = (int)S.imgs.size();
    cudaStream_t s = S.stream;
    uint64_t* L = &S.launches;
    S.launches = 0;
    CTX_CUDA(cudaEventRecord(S.ev[0], s));
    // upload (main.cpp:52-54 leaves band 0 as float 0..255; u8 input is widened on the device)
    const int n_packed = S.packed ? S.n_packed : 0;
    if (n_packed) {
        // the first n_packed f32 frames of the pass were packed to bytes on the host (pack_begin): they go up as bytes (same
        // layout as d_in_u8, one transfer when dense) and are widened right here, so that the rest of stage A is that of any
        // f32 pass; the frames the host threads did not get to follow as they are
        const size_t img_bytes = (size_t)p->in_pitch * (size_t)p->in_h;
        if (img_bytes == c->max_in_px) {
            CTX_CUDA(cudaMemcpyAsync(S.d_in_u8, S.h_pack, img_bytes * (size_t)n_packed, cudaMemcpyHostToDevice, s));
        } else {
            for (int b = 0; b < n_packed; ++b)
                CTX_CUDA(cudaMemcpyAsync(S.d_in_u8 + (size_t)b * c->max_in_px, S.h_pack + (size_t)b * c->max_in_px, img_bytes, cudaMemcpyHostToDevice, s));
        }
        CTX_TRY(launch_u8_to_f32(S.d_in_u8, c->max_in_px, p->in_pitch, S.d_in, c->max_in_px, p->in_pitch, p->in_w, p->in_h, n_packed, s, L));
        c->tm.h2d_bytes += img_bytes * (size_t)n_packed;
        c->tm.packed_images += (uint32_t)n_packed;
    }
    for (int b = n_packed; b < nb;) {
        const sift_gpu_image& im = *S.imgs[(size_t)b].img;
        const size_t esz = im.dtype == SIFT_GPU_DTYPE_U8 ? 1 : 4;
        const size_t pitch = im.row_stride_bytes ? (size_t)im.row_stride_bytes : (size_t)im.width * esz;
        void* dst = im.dtype == SIFT_GPU_DTYPE_U8 ? (void*)(S.d_in_u8 + (size_t)b * c->max_in_px) : (void*)(S.d_in + (size_t)b * c->max_in_px);
        const cudaMemcpyKind kind = im.memory == SIFT_GPU_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
        const size_t bytes = pitch * (size_t)im.height;
        if (pitch == (size_t)im.width * esz && (size_t)p->in_pitch == (size_t)im.width) {
            // dense on both sides: one linear copy, extended over the following frames while they are adjacent in the
            // caller's memory and in the staging buffer (a frame stack uploads as a single transfer)
            int n = 1;
            if ((size_t)p->in_pitch * (size_t)p->in_h == c->max_in_px)
                while (b + n < nb && S.imgs[(size_t)(b + n)].img->memory == im.memory && (!S.imgs[(size_t)(b + n)].img->row_stride_bytes || (size_t)S.imgs[(size_t)(b + n)].img->row_stride_bytes == pitch) &&
                       (const char*)S.imgs[(size_t)(b + n)].img->data == (const char*)im.data + (size_t)n * bytes)
                    ++n;
            CTX_CUDA(cudaMemcpyAsync(dst, im.data, bytes * (size_t)n, kind, s));
            if (im.memory != SIFT_GPU_MEM_DEVICE) c->tm.h2d_bytes += bytes * (size_t)n;
            b += n;
        } else {
            CTX_CUDA(cudaMemcpy2DAsync(dst, (size_t)p->in_pitch * esz, im.data, pitch, (size_t)im.width * esz, (size_t)im.height, kind, s));
            if (im.memory != SIFT_GPU_MEM_DEVICE) c->tm.h2d_bytes += (size_t)im.width * esz * (size_t)im.height;
            ++b;
        }
    }
    // Everything from here to the survivor copy-back has the same launches, the same device addresses and the same kernel
    // parameters every time this slot sees a pass of this shape: the second time it is captured into a CUDA graph (both
    // pyramid streams, the stage-boundary events as external event-record nodes) and from then on one graph launch
    // replaces ~50 launch calls — less host time per pass and shorter gaps between the short kernels of the small octaves.
    static const bool graphs_on = [] { const char* e = getenv("SIFT_GPU_GRAPHS"); return !(e && atoi(e) == 0) && !getenv("SIFT_GPU_TRACE_PYR"); }();
    const bool u8 = S.imgs[0].img->dtype == SIFT_GPU_DTYPE_U8;  // (packed frames belong to f32 passes and are widened above)
    Slot::StageGraph* G = graphs_on ? &S.graphs[std::make_tuple((const void*)p, nb, u8 ? 1 : 0)] : nullptr;
    if (G && G->exec) {
        CTX_CUDA(cudaGraphLaunch(G->exec, s));
        S.launches += G->launches;
    } else {
        const bool capture = G && G->seen >= 1;  // the first pass of a shape runs eagerly (attribute set-up, occupancy queries)
        if (G) ++G->seen;
        const uint64_t l0 = S.launches;
        if (capture) CTX_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        auto mark = [&](int i) { return capture ? cudaEventRecordWithFlags(S.ev[i], s, cudaEventRecordExternal) : cudaEventRecord(S.ev[i], s); };
        int rc = 0;
        auto body = [&]() -> int {
            if (u8) CTX_TRY(launch_u8_to_f32(S.d_in_u8, c->max_in_px, p->in_pitch, S.d_in, c->max_in_px, p->in_pitch, p->in_w, p->in_h, nb, s, L));
            CTX_CUDA(mark(1));
            CTX_TRY(run_pyramid(c, S, p, ps, nb));
            CTX_CUDA(mark(2));
            CTX_TRY(launch_extrema(ps.layers_dev, ps.layers_host.data(), (int)ps.layers_host.size(), p->total_cols, p->mask_words, S.d_mask,
                                   S.d_mask + c->mask_cap * (size_t)c->B, S.d_col_count, S.d_col_off, S.d_cands, c->cand_cap, S.d_n_cand, nb, s, L));
            CTX_CUDA(mark(3));
            CTX_TRY(launch_eliminate(ps.layers_dev, (int)ps.layers_host.size(), S.d_cands, c->cand_cap, S.d_n_cand, S.d_surv, c->cand_cap,
                                     S.d_n_surv, S.d_slice, c->D, nb, s, L));
            CTX_CUDA(mark(4));
            CTX_CUDA(cudaMemcpyAsync(S.h_n_cand, S.d_n_cand, sizeof(uint32_t) * (size_t)nb, cudaMemcpyDeviceToHost, s));
            CTX_CUDA(cudaMemcpyAsync(S.h_n_surv, S.d_n_surv, sizeof(uint32_t) * (size_t)nb, cudaMemcpyDeviceToHost, s));
            // speculative: the first kSurvFirst survivors of every image travel with the counters (one strided copy)
            const size_t first = std::min<size_t>(kSurvFirst, c->cand_cap);
            CTX_CUDA(cudaMemcpy2DAsync(S.h_surv, sizeof(Surv) * kSurvFirst, S.d_surv, sizeof(Surv) * c->cand_cap, sizeof(Surv) * first, (size_t)nb,
                                       cudaMemcpyDeviceToHost, s));
            return 0;
        };
        rc = body();
        if (capture) {
            cudaGraph_t graph = nullptr;
            const cudaError_t e = cudaStreamEndCapture(s, &graph);
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            CTX_CUDA(e);
            const cudaError_t ei = cudaGraphInstantiate(&G->exec, graph, 0);
            cudaGraphDestroy(graph);
            CTX_CUDA(ei);
            G->launches = S.launches - l0;
            CTX_CUDA(cudaGraphLaunch(G->exec, s));
        } else if (rc) {
            return rc;
        }
    }
    if (c->prm.subpixel && (c->prm.flags & SIFT_GPU_FLAG_KEEP_UPSAMPLED))
        for (int b = 0; b < nb; ++b)
            if (S.imgs[(size_t)b].img->upsampled_out)
                CTX_CUDA(cudaMemcpy2DAsync(S.imgs[(size_t)b].img->upsampled_out, sizeof(float) * (size_t)p->ow[0], S.d_up + (size_t)b * c->maxP[0],
                                           sizeof(float) * (size_t)p->pitch[0], sizeof(float) * (size_t)p->ow[0], (size_t)p->oh[0], cudaMemcpyDeviceToHost, s));
    CTX_CUDA(cudaEventRecord(S.ev[5], s));
    return 0;
}

// Host order replay of one pass, first half: blocks on stage A, then starts the per-image replay on the worker pool and
// returns (the caller collects an older pass and enqueues the next stage A meanwhile).
static int begin_replay(sift_gpu_ctx* c, Slot& S) {
    Plan* p = S.plan;
    const int nb = (int)S.imgs.size();
    cudaStream_t s = S.stream;
    const double t_w0 = now_ms();
    CTX_CUDA(cudaEventSynchronize(S.ev[5]));
    g_trace[1] += now_ms() - t_w0;
    // images with more survivors than the speculative copy holds fetch the remainder now (rare)
    bool overflow = false;
    for (int b = 0; b < nb; ++b) {
        S.surv_overflow[(size_t)b].clear();
        const uint32_t n = S.h_n_surv[b];
        if (n > c->cand_cap) return set_error(c, SIFT_GPU_E_CAPACITY, "survivor list overflow");
        if (n > kSurvFirst) {
            S.surv_overflow[(size_t)b].resize(n);
            CTX_CUDA(cudaMemcpyAsync(S.surv_overflow[(size_t)b].data(), S.d_surv + (size_t)b * c->cand_cap, sizeof(Surv) * n, cudaMemcpyDeviceToHost, s));
            overflow = true;
        }
    }
    if (overflow) CTX_CUDA(cudaStreamSynchronize(s));

    S.t_replay0 = now_ms();
    S.rep.resize((size_t)nb);  // the records (and their vectors' capacity) are reused from pass to pass; replay_image rewrites every field
    Slot* Sp = &S;
    c->pool->begin(nb, [c, p, Sp](int b) {
        const uint32_t n = Sp->h_n_surv[b];
        const Surv* sv = n > kSurvFirst ? Sp->surv_overflow[(size_t)b].data() : Sp->h_surv + (size_t)b * kSurvFirst;
        replay_image(c, p, Sp->h_n_cand[b], sv, n, &Sp->rep[(size_t)b]);
    });
    return 0;
}

// Second half: waits for the replay, then stage B: keypoints up, orientation, descriptors, results back.
static int end_replay_and_enqueue_stage_b(sift_gpu_ctx* c, Slot& S, int slot_index) {
    Plan* p = S.plan;
    const PlanSlot& ps = p->ps[slot_index];
    const int nb = (int)S.imgs.size();
    cudaStream_t s = S.stream;
    uint64_t* L = &S.launches;
    c->pool->end();
    const double t_host0 = S.t_replay0;
    size_t n_keys = 0;
    for (int b = 0; b < nb; ++b) {
        S.h_key_first[b] = (uint32_t)n_keys;
        if (S.rep[(size_t)b].status == SIFT_GPU_OK) n_keys += S.rep[(size_t)b].keys.size();
    }
    S.h_key_first[nb] = (uint32_t)n_keys;
    S.n_keys = n_keys;
    CTX_TRY(ensure_key_capacity(c, S, n_keys));
    for (int b = 0; b < nb; ++b) {
        if (S.rep[(size_t)b].status != SIFT_GPU_OK) continue;
        const size_t off = S.h_key_first[b];
        std::copy(S.rep[(size_t)b].keys.begin(), S.rep[(size_t)b].keys.end(), S.h_keys + off);
        std::fill(S.h_key_img + off, S.h_key_img + off + S.rep[(size_t)b].keys.size(), (uint32_t)b);
    }
    c->tm.host_order_ms += (float)(now_ms() - t_host0);
    g_trace[2] += now_ms() - t_host0;
    const double t_e0 = now_ms();

    CTX_CUDA(cudaEventRecord(S.ev[6], s));
    S.h_desc = take_desc_block(c, n_keys * kDescLen);
    if (!S.h_desc) return set_error(c, SIFT_GPU_E_CUDA, "pinned descriptor block allocation failed");
    if (n_keys) {
        // the key lists are pulled out of pinned host memory by a kernel: a cudaMemcpyAsync would queue on the
        // host-to-device copy engine behind the frame uploads of the next passes (several milliseconds)
        const int blocks = (int)std::min<size_t>(64, (n_keys + 255) / 256);
        fetch_keys_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const uint2*>(S.h_keys), reinterpret_cast<uint2*>(S.d_keys), S.h_key_img,
                                                  S.d_key_img, S.h_key_first, S.d_key_first, (uint32_t)n_keys, (uint32_t)(nb + 1));
        ++*L;
        CTX_CUDA(cudaGetLastError());
    }
    CTX_CUDA(cudaEventRecord(S.ev[7], s));
    const int n_targets = (int)ps.targets_host.size();
    if (n_keys) {
        const size_t tables_need = (size_t)nb * (size_t)n_targets * 256;
        if (tables_need > S.tables_cap) {
            cudaFree(S.d_tables);
            S.tables_cap = 0;
            CTX_CUDA(cudaMalloc(&S.d_tables, sizeof(float) * tables_need));
            S.tables_cap = tables_need;
        }
        // key grid: 16-pixel cells over the largest target level, per image and target
        {
            int tw = 1, th = 1;
            for (size_t t = 0; t < p->target_w.size(); ++t) { tw = std::max(tw, p->target_w[t]); th = std::max(th, p->target_h[t]); }
            S.grid.cw = (tw + 15) / 16;
            S.grid.ch = (th + 15) / 16;
            S.grid.cells_per_image = (uint32_t)(n_targets * S.grid.cw * S.grid.ch);
            const size_t words = (size_t)nb * S.grid.cells_per_image;
            if (words > S.grid_words) {
                cudaFree(S.grid.count); cudaFree(S.grid.offset); cudaFree(S.grid.cursor);
                S.grid.count = S.grid.offset = S.grid.cursor = nullptr;
                S.grid_words = 0;
                CTX_CUDA(cudaMalloc(&S.grid.count, sizeof(uint32_t) * words));
                CTX_CUDA(cudaMalloc(&S.grid.offset, sizeof(uint32_t) * words));
                CTX_CUDA(cudaMalloc(&S.grid.cursor, sizeof(uint32_t) * words));
                S.grid_words = words;
            }
            if (n_keys > S.grid_keys) {
                cudaFree(S.grid.cell_keys);
                S.grid.cell_keys = nullptr;
                S.grid_keys = 0;
                CTX_CUDA(cudaMalloc(&S.grid.cell_keys, sizeof(uint32_t) * S.key_cap));
                S.grid_keys = S.key_cap;
            }
        }
        CTX_TRY(launch_orientation(ps.targets_dev, n_targets, S.d_keys, S.d_key_img, (uint32_t)n_keys, S.d_orient, S.d_npeaks, S.d_peaks, S.d_grad, s, L));
    }
    CTX_CUDA(cudaEventRecord(S.ev[8], s));
    if (n_keys) {
        CTX_TRY(launch_weight_tables(ps.targets_dev, n_targets, c->d_taps + c->w16_blur.tap_off, c->w16_blur.r, S.d_tables, c->fma, nb, s, L));
        CTX_TRY(launch_descriptors(ps.targets_dev, n_targets, S.d_tables, S.d_keys, S.d_key_img, S.d_key_first, (uint32_t)n_keys, S.d_orient,
                                   S.d_desc, S.d_grad, &S.grid, nb, s, L));
    }
    CTX_CUDA(cudaEventRecord(S.ev[9], s));
    if (n_keys) {
        CTX_CUDA(cudaMemcpyAsync(S.h_orient, S.d_orient, sizeof(float) * n_keys, cudaMemcpyDeviceToHost, s));
        CTX_CUDA(cudaMemcpyAsync(S.h_npeaks, S.d_npeaks, sizeof(uint32_t) * n_keys, cudaMemcpyDeviceToHost, s));
        CTX_CUDA(cudaMemcpyAsync(S.h_desc, S.d_desc, sizeof(float) * kDescLen * n_keys, cudaMemcpyDeviceToHost, s));
    }
    CTX_CUDA(cudaEventRecord(S.ev[10], s));
    g_trace[3] += now_ms() - t_e0;
    return 0;
}

// Rare path (sift.cpp:194-200): some keypoint of the pass came back with more than one orientation peak.  The
// reference then appends one copy of the point per peak (`peaks.begin()++` is begin(): the first peak is copied
// too) behind all the points, which changes what the second cleanup sort sees and therefore the order the
// descriptors are built in.  The pass's host replay is redone from the first cleanup on with the peaks known, and
// the descriptors are recomputed for the new key list with the orientations supplied by the host.  Synchronous.
static int redo_with_extra_orientations(sift_gpu_ctx* c, Slot& S, int slot_index) {
    Plan* p = S.plan;
    const PlanSlot& ps = p->ps[slot_index];
    const int nb = (int)S.imgs.size();
    const bool canonical = (c->prm.flags & SIFT_GPU_FLAG_ORDER_CANONICAL) != 0;
    cudaStream_t s = S.stream;
    std::vector<float> peaks(S.n_keys * 36);
    CTX_CUDA(cudaMemcpy(peaks.data(), S.d_peaks, sizeof(float) * peaks.size(), cudaMemcpyDeviceToHost));
    std::vector<std::vector<float>> new_orient((size_t)nb);
    for (int b = 0; b < nb; ++b) {
        ReplayOut& ro = S.rep[(size_t)b];
        if (ro.status != SIFT_GPU_OK) continue;
        if (ro.truncated) {  // orientations of the dropped points were never computed
            ro.status = SIFT_GPU_E_UNSUPPORTED;
            c->error = "more than 65535 keypoints together with extra orientation peaks: not supported";
            continue;
        }
        const size_t off = S.h_key_first[b];
        std::vector<int64_t> key_of_l1(ro.l1.size(), -1);
        for (size_t i = 0; i < ro.kps.size(); ++i)
            if (ro.key_of[i] != ~0u) key_of_l1[ro.kp_l1[i]] = (int64_t)(off + ro.key_of[i]);
        // the vector as _orientationAssignment leaves it: every point of l1 (filtered unless inside), then the copies
        struct Entry { uint32_t l1; float orientation; };
        std::vector<Entry> V(ro.l1.size());
        std::vector<uint32_t> zero_pos;
        for (size_t q = 0; q < ro.l1.size(); ++q) V[q] = Entry{(uint32_t)q, 0.0f};
        for (uint32_t q : ro.inside) {
            zero_pos.push_back(q);
            V[q].orientation = S.h_orient[key_of_l1[q]];
        }
        for (uint32_t q : ro.inside) {
            const size_t k = (size_t)key_of_l1[q];
            const uint32_t np = S.h_npeaks[k];
            if (np <= 1) continue;
            for (uint32_t j = 0; j < np; ++j) {
                zero_pos.push_back((uint32_t)V.size());
                V.push_back(Entry{q, peaks[k * 36 + j]});
            }
        }
        SparseFilterSort sorter;
        std::vector<uint32_t> identity;
        size_t n_kept = 0;
        const uint32_t* kept = cleanup_order(sorter, identity, (uint32_t)V.size(), zero_pos.data(), zero_pos.size(), canonical, &n_kept);
        ReplayOut nr;
        nr.n_survivors = ro.n_survivors;
        nr.kps.resize(n_kept);
        nr.key_of.assign(n_kept, ~0u);
        for (size_t i = 0; i < n_kept; ++i) {
            const Entry& e = V[zero_pos[kept[i]]];
            KeyIn ki;
            if (make_keypoint(c, p, ro.l1[e.l1], &nr.kps[i], &ki)) {
                nr.key_of[i] = (uint32_t)nr.keys.size();
                nr.keys.push_back(ki);
                new_orient[(size_t)b].push_back(e.orientation);
            }
            nr.kps[i].orientation = e.orientation;
        }
        ro.kps.swap(nr.kps);
        ro.key_of.swap(nr.key_of);
        ro.keys.swap(nr.keys);
    }
    size_t n_keys = 0;
    for (int b = 0; b < nb; ++b) {
        S.h_key_first[b] = (uint32_t)n_keys;
        if (S.rep[(size_t)b].status == SIFT_GPU_OK) n_keys += S.rep[(size_t)b].keys.size();
    }
    S.h_key_first[nb] = (uint32_t)n_keys;
    S.n_keys = n_keys;
    CTX_TRY(ensure_key_capacity(c, S, n_keys));
    for (int b = 0; b < nb; ++b) {
        const ReplayOut& ro = S.rep[(size_t)b];
        if (ro.status != SIFT_GPU_OK) continue;
        const size_t off = S.h_key_first[b];
        std::copy(ro.keys.begin(), ro.keys.end(), S.h_keys + off);
        std::fill(S.h_key_img + off, S.h_key_img + off + ro.keys.size(), (uint32_t)b);
        std::copy(new_orient[(size_t)b].begin(), new_orient[(size_t)b].end(), S.h_orient + off);
        std::fill(S.h_npeaks + off, S.h_npeaks + off + ro.keys.size(), 1u);
    }
    S.h_desc = take_desc_block(c, n_keys * kDescLen);
    if (!S.h_desc) return set_error(c, SIFT_GPU_E_CUDA, "pinned descriptor block allocation failed");
    if (n_keys) {
        CTX_CUDA(cudaMemcpyAsync(S.d_keys, S.h_keys, sizeof(KeyIn) * n_keys, cudaMemcpyHostToDevice, s));
        CTX_CUDA(cudaMemcpyAsync(S.d_key_img, S.h_key_img, sizeof(uint32_t) * n_keys, cudaMemcpyHostToDevice, s));
        CTX_CUDA(cudaMemcpyAsync(S.d_key_first, S.h_key_first, sizeof(uint32_t) * (size_t)(nb + 1), cudaMemcpyHostToDevice, s));
        CTX_CUDA(cudaMemcpyAsync(S.d_orient, S.h_orient, sizeof(float) * n_keys, cudaMemcpyHostToDevice, s));
        CTX_TRY(launch_descriptors(ps.targets_dev, (int)ps.targets_host.size(), S.d_tables, S.d_keys, S.d_key_img, S.d_key_first, (uint32_t)n_keys,
                                   S.d_orient, S.d_desc, nullptr /* another key list: recompute the gradients */,
                                   n_keys <= S.grid_keys ? &S.grid : nullptr, (int)S.imgs.size(), s, &S.launches));
        CTX_CUDA(cudaMemcpyAsync(S.h_desc, S.d_desc, sizeof(float) * kDescLen * n_keys, cudaMemcpyDeviceToHost, s));
    }
    CTX_CUDA(cudaStreamSynchronize(s));
    return 0;
}

// Waits for stage B of the pass in `S` and fills the caller's results.
static int finish_pass(sift_gpu_ctx* c, Slot& S, int slot_index, sift_gpu_result* results) {
    Plan* p = S.plan;
    const int nb = (int)S.imgs.size();
    const double t_w0 = now_ms();
    CTX_CUDA(cudaEventSynchronize(S.ev[10]));
    g_trace[4] += now_ms() - t_w0;
    const double t_f0 = now_ms();
    {
        bool extra = false;
        for (size_t k = 0; k < S.n_keys && !extra; ++k) extra = S.h_npeaks[k] > 1;
        if (extra) CTX_TRY(redo_with_extra_orientations(c, S, slot_index));
    }
    for (int b = 0; b < nb; ++b) {
        const int ri = S.imgs[(size_t)b].result_index;
        sift_gpu_result& R = results[ri];
        HostImageOut& HO = c->outs[(size_t)ri];
        ReplayOut& ro = S.rep[(size_t)b];
        R.n_candidates = S.h_n_cand[b];
        R.n_survivors = ro.n_survivors;
        R.out_width = p->ow[0]; R.out_height = p->oh[0];
        R.status = ro.status;
        R.n = 0; R.kps = nullptr; R.desc = nullptr;
        if (ro.status != SIFT_GPU_OK) continue;
        const size_t off = S.h_key_first[b];
        // descriptors are contiguous per image only if every keypoint went to the device (always so in
        // practice, sift.cpp:65 can never reject what sift.cpp:173 accepted); otherwise rows are spread out.
        bool all = true;
        for (size_t i = 0; i < ro.kps.size(); ++i) {
            const uint32_t ko = ro.key_of[i];
            if (ko == ~0u) { all = false; continue; }
            ro.kps[i].orientation = S.h_orient[off + ko];
        }
        HO.kps.swap(ro.kps);
        R.n = (uint32_t)HO.kps.size();
        R.kps = HO.kps.data();
        if (all) {
            R.desc = S.h_desc + off * kDescLen;
        } else {
            float* blk = take_desc_block(c, HO.kps.size() * kDescLen);
            if (!blk) return set_error(c, SIFT_GPU_E_CUDA, "pinned descriptor block allocation failed");
            for (size_t i = 0; i < HO.kps.size(); ++i) {
                if (ro.key_of[i] == ~0u) std::memset(blk + i * kDescLen, 0, sizeof(float) * kDescLen);
                else std::memcpy(blk + i * kDescLen, S.h_desc + (off + ro.key_of[i]) * kDescLen, sizeof(float) * kDescLen);
            }
            R.desc = blk;
        }
    }
    // stage times of this pass (device time of each stage; passes overlap, so their sum can exceed span_ms)
    float ms = 0.0f;
    auto el = [&](int a, int b2) { cudaEventElapsedTime(&ms, S.ev[a], S.ev[b2]); return ms; };
    c->tm.h2d_ms += el(0, 1);
    c->tm.pyramid_ms += el(1, 2);
    c->tm.extrema_ms += el(2, 3);
    c->tm.eliminate_ms += el(3, 4);
    c->tm.d2h_survivors_ms += el(4, 5);
    c->tm.h2d_keypoints_ms += el(6, 7);
    c->tm.orientation_ms += el(7, 8);
    c->tm.descriptor_ms += el(8, 9);
    c->tm.d2h_results_ms += el(9, 10);
    c->tm.kernel_launches += S.launches;
    if (getenv("SIFT_GPU_TRACE")) {
        float t[11];
        for (int i = 0; i < 11; ++i) cudaEventElapsedTime(&t[i], c->ev_first, S.ev[i]);
        fprintf(stderr, "[sift_gpu pass] h2d %.3f-%.3f pyr -%.3f ext -%.3f elim -%.3f d2h -%.3f | keys %.3f-%.3f ori -%.3f desc -%.3f d2h -%.3f\n", t[0], t[1],
                t[2], t[3], t[4], t[5], t[6], t[7], t[8], t[9], t[10]);
    }
    S.busy = false;
    g_trace[5] += now_ms() - t_f0;
    return 0;
}

// =================================================================================================
extern "C" {

const char* sift_gpu_version(void) { return "sift_b200 0.2 (sm_100a)"; }

const char* sift_gpu_last_error(const sift_gpu_ctx* ctx) { return ctx ? ctx->error.c_str() : g_last_error.c_str(); }

int sift_gpu_create(const sift_gpu_params* params, sift_gpu_ctx** out) {
    if (!params || !out) return set_error(nullptr, SIFT_GPU_E_INVALID, "null argument");
    *out = nullptr;
    // reference asserts (sift.cpp:382-383)
    if (!(params->octaves > 0)) return set_error(nullptr, SIFT_GPU_E_ASSERT, "assert(_octaves > 0)");
    if (!(params->dogs_per_epoch >= 3)) return set_error(nullptr, SIFT_GPU_E_ASSERT, "assert(_dogsPerEpoch >= 3)");
    if (params->octaves > kMaxOctaves || params->dogs_per_epoch + 1 > kMaxGauss)
        return set_error(nullptr, SIFT_GPU_E_UNSUPPORTED, "too many octaves / DoGs per octave");
    if (params->max_width < 1 || params->max_height < 1 || params->max_batch < 1)
        return set_error(nullptr, SIFT_GPU_E_INVALID, "max_width/max_height/max_batch must be positive");
    if (params->max_batch > 4096) return set_error(nullptr, SIFT_GPU_E_UNSUPPORTED, "max_batch above 4096 (the batch is a grid dimension)");
    if ((params->subpixel ? 2 : 1) * (long)params->max_width > 65535 || (params->subpixel ? 2 : 1) * (long)params->max_height > 32767)
        return set_error(nullptr, SIFT_GPU_E_UNSUPPORTED, "image too large for u16 keypoint coordinates");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= params->device || params->device < 0)
        return set_error(nullptr, SIFT_GPU_E_CUDA, "no such CUDA device (this library has no CPU fallback)");
    sift_gpu_ctx* c = new sift_gpu_ctx();
    c->prm = *params;
    c->O = params->octaves; c->D = params->dogs_per_epoch; c->G = c->D + 1;
    c->fma = (params->flags & SIFT_GPU_FLAG_FMA_BLUR) != 0;
    c->B = params->max_batch;
    c->n_slots = (params->flags & SIFT_GPU_FLAG_SERIAL) ? 1 : std::min(5, kSlots);
    if (const char* e = getenv("SIFT_GPU_SLOTS")) c->n_slots = std::max(1, std::min(kSlots, atoi(e)));
    c->max_in_w = params->max_width; c->max_in_h = params->max_height;
    int rc = 0;
    auto body = [&]() -> int {
        CTX_CUDA(cudaSetDevice(params->device));
        CTX_TRY(blur_prepare_device());   // function attributes / occupancy figures are per device
        CTX_TRY(build_schedule(c));
        CTX_TRY(alloc_buffers(c));
        return 0;
    };
    rc = body();
    if (rc != 0) {
        std::string err = c->error;
        sift_gpu_destroy(c);
        g_last_error = err;
        return rc;
    }
    int nthreads = 0;
    if (const char* e = getenv("SIFT_GPU_HOST_THREADS")) nthreads = atoi(e);
    // default: the host's hardware threads shared out over its GPUs (one process per GPU is the deployment), at most 16
    if (nthreads <= 0) nthreads = (int)std::min<unsigned>(16u, std::max(4u, std::thread::hardware_concurrency() / (unsigned)std::max(1, ndev)));
    c->pool = new Pool(nthreads - 1);
    c->pack_pool = new Pool(nthreads - 1);
    c->host_threads = nthreads;
    *out = c;
    return SIFT_GPU_OK;
}

void sift_gpu_destroy(sift_gpu_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->prm.device);
    for (Slot& S : c->slots)
        if (S.stream) cudaStreamSynchronize(S.stream);
    delete c->pool;
    delete c->pack_pool;
    for (auto& kv : c->plans) {
        Plan* p = kv.second;
        cudaFree(p->d_maps);
        for (PlanSlot& ps : p->ps) { cudaFree(ps.layers_dev); cudaFree(ps.targets_dev); }
        delete p;
    }
    cudaFree(c->d_taps);
    for (Slot& S : c->slots) {
        cudaFree(S.d_in_u8); cudaFree(S.d_in); cudaFree(S.d_up_tmp); cudaFree(S.d_up);
        for (int o = 0; o < kMaxOctaves; ++o)
            for (int i = 0; i < kMaxGauss; ++i) { cudaFree(S.d_gauss[o][i]); cudaFree(S.d_dog[o][i]); }
        for (auto& kv : S.graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
        S.graphs.clear();
        cudaFree(S.d_mask); cudaFree(S.d_col_count); cudaFree(S.d_col_off); cudaFree(S.d_cands); cudaFree(S.d_surv);
        cudaFree(S.d_n_cand); cudaFree(S.d_n_surv); cudaFree(S.d_slice); cudaFreeHost(S.h_n_cand); cudaFreeHost(S.h_n_surv); cudaFreeHost(S.h_surv); cudaFreeHost(S.h_pack);
        cudaFree(S.d_keys); cudaFree(S.d_key_img); cudaFree(S.d_key_first); cudaFree(S.d_orient); cudaFree(S.d_npeaks);
        cudaFree(S.d_peaks); cudaFree(S.d_desc); cudaFree(S.d_grad); cudaFree(S.d_tables);
        cudaFree(S.grid.count); cudaFree(S.grid.offset); cudaFree(S.grid.cursor); cudaFree(S.grid.cell_keys);
        cudaFreeHost(S.h_keys); cudaFreeHost(S.h_key_img); cudaFreeHost(S.h_key_first); cudaFreeHost(S.h_orient); cudaFreeHost(S.h_npeaks);
        for (auto& e : S.ev) if (e) cudaEventDestroy(e);
        for (auto& e : S.ev_join) if (e) cudaEventDestroy(e);
        if (S.ev_fork) cudaEventDestroy(S.ev_fork);
        for (auto& st : S.aux) if (st) cudaStreamDestroy(st);
        if (S.stream) cudaStreamDestroy(S.stream);
    }
    for (float* b : c->desc_blocks) cudaFreeHost(b);
    if (c->ev_first) cudaEventDestroy(c->ev_first);
    if (c->ev_last) cudaEventDestroy(c->ev_last);
    delete c;
}

struct PassPlan {
    Plan* plan;
    std::vector<ChunkImage> imgs;
};

// ---- lossless f32 -> u8 packing of host frames ahead of their upload (pack_host.cpp) ------------------------------
extern "C" int sift_gpu_debug_pack_rows_u8(const float* src, size_t src_stride_bytes, int w, int rows, uint8_t* dst, size_t dst_pitch);

constexpr int kPackMinImages = 8;   // below this a pass is latency work: packing would only delay its upload
constexpr int kPackRows = 128;      // rows per work item
constexpr int kPackDefaultMode = 1;  // measured on one B200 with 16 host threads: whole passes 12.3 k images/s, split 9.4 k, off 6.3 k

// Policy (read per run): SIFT_GPU_HOST_PACK=0 off; 1 whole passes (stage A waits until every frame of its pass is packed);
// 2 split (stage A never waits: the frames packed by the time the pass is due travel as bytes, the rest as f32 — the host
// threads and the PCIe link share the upload in whatever proportion they manage; measured slower than 1 where the host has
// the threads, kept for hosts that do not).  Default: whole passes when this context has at least 8 host threads to itself
// (a frame costs about a millisecond of one core: with fewer threads the order replay needs them more), else off.
static int pack_mode(const sift_gpu_ctx* c) {
    if (const char* e = getenv("SIFT_GPU_HOST_PACK")) { const int v = atoi(e); return v < 0 ? 0 : (v > 2 ? 2 : v); }
    return c->host_threads >= 8 ? kPackDefaultMode : 0;
}

// Starts packing the frames of pass `pp` into the slot's pinned staging buffer on the pack pool and returns; pack_end joins.
// The slot's previous pass has been uploaded long ago (its stage A was waited for before its order replay began).
static int pack_begin(sift_gpu_ctx* c, Slot& S, const PassPlan& pp) {
    S.pack_job = false;
    const int nb = (int)pp.imgs.size();
    if (nb < kPackMinImages || !c->pack_pool) return 0;
    for (const ChunkImage& ci : pp.imgs)
        if (ci.img->dtype != SIFT_GPU_DTYPE_F32 || ci.img->memory != SIFT_GPU_MEM_HOST) return 0;
    if (!S.h_pack) {
        if (cudaHostAlloc((void**)&S.h_pack, c->max_in_px * (size_t)c->B, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            S.h_pack = nullptr;
            return 0;  // no staging memory: the pass travels as f32
        }
    }
    CTX_CUDA(cudaEventSynchronize(S.ev[5]));  // belt and braces: the last upload out of h_pack has completed
    const Plan* p = pp.plan;
    const int blocks = (p->in_h + kPackRows - 1) / kPackRows;
    S.pack_bad.store(0);
    S.pack_job = true;
    S.pack_blocks = blocks;
    Slot* Sp = &S;
    const PassPlan* ppp = &pp;
    const size_t img_stride = c->max_in_px;
    c->pack_pool->begin(nb * blocks, [Sp, ppp, p, blocks, img_stride](int item) {
        if (Sp->pack_bad.load(std::memory_order_relaxed)) return;
        const int b = item / blocks, y0 = (item % blocks) * kPackRows;
        const sift_gpu_image& im = *ppp->imgs[(size_t)b].img;
        const size_t stride = im.row_stride_bytes ? (size_t)im.row_stride_bytes : (size_t)im.width * 4;
        const int rows = std::min(kPackRows, p->in_h - y0);
        const float* src = reinterpret_cast<const float*>(reinterpret_cast<const char*>(im.data) + (size_t)y0 * stride);
        uint8_t* dst = Sp->h_pack + (size_t)b * img_stride + (size_t)y0 * (size_t)p->in_pitch;
        if (!sift_gpu_debug_pack_rows_u8(src, stride, p->in_w, rows, dst, (size_t)p->in_pitch)) Sp->pack_bad.store(1, std::memory_order_relaxed);
    });
    return 0;
}

// Joins the slot's packing job.  `split`: items not yet started are dropped, only whole frames count.
static void pack_end(sift_gpu_ctx* c, Slot& S, int nb, bool split) {
    S.packed = false;
    S.n_packed = 0;
    if (!S.pack_job) return;
    int items = nb * S.pack_blocks;
    if (split) items = c->pack_pool->cancel_rest();
    c->pack_pool->end();
    S.pack_job = false;
    if (S.pack_bad.load() != 0) return;
    S.n_packed = std::min(nb, items / S.pack_blocks);
    S.packed = S.n_packed > 0;
}

int sift_gpu_run(sift_gpu_ctx* c, const sift_gpu_image* images, int n_images, sift_gpu_result* results) {
    if (!c || (n_images > 0 && (!images || !results)) || n_images < 0) return set_error(c, SIFT_GPU_E_INVALID, "null argument");
    const double t0 = now_ms();
    CTX_CUDA(cudaSetDevice(c->prm.device));
    c->error.clear();
    c->tm = sift_gpu_timings{};
    c->desc_blocks_used = 0;
    if (c->outs.size() < (size_t)n_images) c->outs.resize((size_t)n_images);  // keypoint vectors keep their capacity from call to call
    int first_error = SIFT_GPU_OK;
    // Whatever way this call is left — a CUDA error in the middle of the pipelined loop included — no replay job may keep
    // running on the worker pool and no slot may keep device work queued on buffers the next call reuses.
    struct Quiesce {
        sift_gpu_ctx* c;
        bool clean = false;
        ~Quiesce() {
            if (clean) return;
            if (c->pool) c->pool->end();
            if (c->pack_pool) c->pack_pool->end();
            for (int si = 0; si < c->n_slots; ++si) {
                Slot& S = c->slots[si];
                S.pack_job = false;
                if (S.stream) cudaStreamSynchronize(S.stream);
                for (cudaStream_t a : S.aux)
                    if (a) cudaStreamSynchronize(a);
                S.busy = false;
            }
            cudaGetLastError();
        }
    } quiesce{c};
    // split the images into device passes: runs of equal shape and dtype, at most max_batch each
    std::vector<PassPlan> passes;
    int i = 0;
    while (i < n_images) {
        const sift_gpu_image& im = images[i];
        sift_gpu_result& R = results[i];
        std::memset(&R, 0, sizeof R);
        if (!im.data || im.width < 1 || im.height < 1 || (im.dtype != SIFT_GPU_DTYPE_F32 && im.dtype != SIFT_GPU_DTYPE_U8)) {
            R.status = SIFT_GPU_E_INVALID;
            if (!first_error) first_error = set_error(c, SIFT_GPU_E_INVALID, "bad image descriptor");
            ++i;
            continue;
        }
        if (im.width > c->max_in_w || im.height > c->max_in_h) {
            R.status = SIFT_GPU_E_CAPACITY;
            if (!first_error) first_error = set_error(c, SIFT_GPU_E_CAPACITY, "image larger than max_width x max_height");
            ++i;
            continue;
        }
        Plan* p = get_plan(c, im.width, im.height);
        if (p->status != SIFT_GPU_OK) {
            R.status = p->status;
            if (!first_error) first_error = set_error(c, p->status, p->why);
            ++i;
            continue;
        }
        PassPlan pp;
        pp.plan = p;
        while (i < n_images && (int)pp.imgs.size() < c->B && images[i].data && images[i].width == im.width &&
               images[i].height == im.height && images[i].dtype == im.dtype) {
            std::memset(&results[i], 0, sizeof(sift_gpu_result));
            pp.imgs.push_back(ChunkImage{i, &images[i]});
            ++i;
        }
        passes.push_back(std::move(pp));
    }
    // software pipeline over the passes: A(k) is enqueued before the host replays pass k-1, whose stage B then
    // runs while A(k+1) is being enqueued and pass k-2 is collected
    const int np = (int)passes.size(), ns = c->n_slots;
    // the host replays pass k - lag_b while stage A of passes up to k is queued on the device (queueing ahead keeps the
    // device fed while the host waits for a stage A to end), and collects ns-1 passes behind the enqueue front
    int lag_b = ns >= 4 ? ns - 2 : (ns >= 2 ? 1 : 0);
    if (const char* e = getenv("SIFT_GPU_LAG")) lag_b = std::max(ns >= 2 ? 1 : 0, std::min(ns - 1, atoi(e)));
    const int lag_f = ns - 1;
    auto collect = [&](int kf) -> int {
        Slot& S = c->slots[kf % ns];
        if (kf == np - 1) CTX_CUDA(cudaEventRecord(c->ev_last, S.stream));
        CTX_TRY(finish_pass(c, S, kf % ns, results));
        c->last_slot = kf % ns;
        for (const ChunkImage& ci : S.imgs)
            if (results[ci.result_index].status != SIFT_GPU_OK && !first_error) {
                first_error = results[ci.result_index].status;
                if (first_error == SIFT_GPU_E_PRECONDITION) c->error = "separableConvolveX(): kernel longer than line";
            }
        return 0;
    };
    const int packing = pack_mode(c);
    const bool pack_early = lag_b >= 1 && ns >= 3;
    // frames of pass k+1 are packed on the second pool while pass k's stage A is enqueued and an earlier pass is replayed;
    // stage_a(k+1) joins that job before it uploads
    auto pack_ahead = [&](int k) -> int {
        if (packing && k >= 0 && k < np) CTX_TRY(pack_begin(c, c->slots[k % ns], passes[(size_t)k]));
        return 0;
    };
    auto stage_a = [&](int k) -> int {
        Slot& S = c->slots[k % ns];
        S.plan = passes[(size_t)k].plan;
        S.imgs = passes[(size_t)k].imgs;
        S.busy = true;
        if (k == 0) CTX_CUDA(cudaEventRecord(c->ev_first, S.stream));
        const double t_p0 = now_ms();
        pack_end(c, S, (int)S.imgs.size(), packing == 2);
        // pipelined loop: the host threads go straight on to the next pass (its slot's previous pass was uploaded iterations ago)
        if (pack_early) CTX_TRY(pack_ahead(k + 1));
        g_trace[6] += now_ms() - t_p0;
        const double t_a0 = now_ms();
        CTX_TRY(enqueue_stage_a(c, S, k % ns));
        g_trace[0] += now_ms() - t_a0;
        return 0;
    };
    CTX_TRY(pack_ahead(0));
    if (lag_b >= 1 && ns >= 3) {
        // Iteration k: the replay of pass k - lag_b runs on the worker pool while this thread collects pass k - ns (which
        // frees the slot) and enqueues stage A of pass k; then it joins the replay and enqueues that pass's stage B.
        for (int k = 0; k < np + ns; ++k) {
            const int kb = k - lag_b;
            const bool rb = kb >= 0 && kb < np;
            if (rb) CTX_TRY(begin_replay(c, c->slots[kb % ns]));
            if (k - ns >= 0 && k - ns < np) CTX_TRY(collect(k - ns));
            if (k < np) CTX_TRY(stage_a(k));
            if (rb) CTX_TRY(end_replay_and_enqueue_stage_b(c, c->slots[kb % ns], kb % ns));
        }
    } else {
        for (int k = 0; k < np + lag_f; ++k) {
            if (k < np) CTX_TRY(stage_a(k));
            const int kb = k - lag_b;
            if (kb >= 0 && kb < np) {
                CTX_TRY(begin_replay(c, c->slots[kb % ns]));
                CTX_TRY(end_replay_and_enqueue_stage_b(c, c->slots[kb % ns], kb % ns));
            }
            if (k - lag_f >= 0 && k - lag_f < np) CTX_TRY(collect(k - lag_f));
            CTX_TRY(pack_ahead(k + 1));
        }
    }
    if (np > 0) {
        float ms = 0.0f;
        CTX_CUDA(cudaEventSynchronize(c->ev_last));
        cudaEventElapsedTime(&ms, c->ev_first, c->ev_last);
        c->tm.span_ms = ms;
    }
    c->tm.device_total_ms = c->tm.h2d_ms + c->tm.pyramid_ms + c->tm.extrema_ms + c->tm.eliminate_ms + c->tm.d2h_survivors_ms +
                            c->tm.h2d_keypoints_ms + c->tm.orientation_ms + c->tm.descriptor_ms + c->tm.d2h_results_ms;
    c->tm.wall_ms = (float)(now_ms() - t0);
    if (getenv("SIFT_GPU_TRACE")) {
        fprintf(stderr, "[sift_gpu trace] %d passes, wall %.2f ms: enqueueA %.2f  waitA %.2f  replay %.2f  enqueueB %.2f  waitC %.2f  collect %.2f  packjoin %.2f (%u frames packed)\n", np,
                c->tm.wall_ms, g_trace[0], g_trace[1], g_trace[2], g_trace[3], g_trace[4], g_trace[5], g_trace[6], c->tm.packed_images);
    }
    if (getenv("SIFT_GPU_TRACE"))
        fprintf(stderr, "[sift_gpu replay cpu] per image: radix %.1f us  sort1 %.1f  bounds %.1f  sort2 %.1f  keypoints %.1f\n", g_rep_ns[0] * 1e-3 / std::max(1, n_images),
                g_rep_ns[1] * 1e-3 / std::max(1, n_images), g_rep_ns[2] * 1e-3 / std::max(1, n_images), g_rep_ns[3] * 1e-3 / std::max(1, n_images),
                g_rep_ns[4] * 1e-3 / std::max(1, n_images));
    for (auto& v : g_rep_ns) v = 0;
    for (double& v : g_trace) v = 0.0;
    quiesce.clean = true;
    return first_error;
}

int sift_gpu_get_timings(const sift_gpu_ctx* c, sift_gpu_timings* out) {
    if (!c || !out) return SIFT_GPU_E_INVALID;
    *out = c->tm;
    return SIFT_GPU_OK;
}

// ---- stage-level entry points -------------------------------------------------------------------
int sift_gpu_debug_get_level(sift_gpu_ctx* c, int image_idx, int octave, int elem, int kind, float* out, int* width,
                             int* height, float* scale) {
    if (!c || c->last_slot < 0) return set_error(c, SIFT_GPU_E_INVALID, "no pass has run yet");
    const Slot& S = c->slots[c->last_slot];
    const Plan* p = S.plan;
    if (image_idx < 0 || image_idx >= (int)S.imgs.size() || octave < 0 || octave >= c->O || elem < 0 ||
        elem >= (kind == SIFT_GPU_KIND_DOG ? c->D : c->G))
        return set_error(c, SIFT_GPU_E_INVALID, "level index out of range");
    CTX_CUDA(cudaSetDevice(c->prm.device));
    const float* src = (kind == SIFT_GPU_KIND_DOG ? S.d_dog[octave][elem] : S.d_gauss[octave][elem]) + (size_t)image_idx * c->maxP[octave];
    if (out && kind == SIFT_GPU_KIND_GAUSS && elem == c->D && !c->top_needed[octave] && !keep_top_levels()) {
        // the pass did not store this level (nothing reads it): blur g(o, D-1) of this one image now
        BlurArgs a = blur_args(c, c->chain_blur[octave][elem], S.d_gauss[octave][elem - 1] + (size_t)image_idx * c->maxP[octave], 0, p->pitch[octave],
                               S.d_gauss[octave][elem] + (size_t)image_idx * c->maxP[octave], 0, p->pitch[octave], nullptr, 0, 0, p->ow[octave],
                               p->oh[octave], nullptr);
        a.z0 = 0;
        a.share = 1;
        CTX_TRY(launch_blur(a, 1, c->fma, S.stream, nullptr));
        CTX_CUDA(cudaStreamSynchronize(S.stream));
    }
    if (width) *width = p->ow[octave];
    if (height) *height = p->oh[octave];
    if (scale) *scale = kind == SIFT_GPU_KIND_DOG ? c->d_scale[octave][elem] : c->g_scale[octave][elem];
    if (out) CTX_CUDA(cudaMemcpy2D(out, sizeof(float) * (size_t)p->ow[octave], src, sizeof(float) * (size_t)p->pitch[octave],
                                   sizeof(float) * (size_t)p->ow[octave], (size_t)p->oh[octave], cudaMemcpyDeviceToHost));
    return SIFT_GPU_OK;
}

static int upload(sift_gpu_ctx* c, const float* h, size_t n, float** d) {
    CTX_CUDA(cudaMalloc(d, sizeof(float) * n));
    CTX_CUDA(cudaMemcpy(*d, h, sizeof(float) * n, cudaMemcpyHostToDevice));
    return 0;
}

static int debug_blur_impl(sift_gpu_ctx* c, const float* src, int w, int h, float sigma, float* dst, int mode) {
    if (!c || !src || !dst || w < 1 || h < 1) return set_error(c, SIFT_GPU_E_INVALID, "bad argument");
    CTX_CUDA(cudaSetDevice(c->prm.device));
    int r = 0;
    std::vector<float> taps = gaussian_taps(sigma, &r);
    if (w < r + 1 || h < r + 1) return set_error(c, SIFT_GPU_E_PRECONDITION, "separableConvolveX/Y(): kernel longer than line");
    if (r > max_generic_radius() && !stream_box_width(r, false)) return set_error(c, SIFT_GPU_E_UNSUPPORTED, "radius too large");
    int dw = w, dh = h;
    if (mode == 1) { dw = (w + 1) / 2; dh = (h + 1) / 2; }
    if (mode == 2) { dw = 2 * w; dh = 2 * h; }
    if (mode != 0 && !(w > 1 && h > 1 && dw > 1 && dh > 1)) return set_error(c, SIFT_GPU_E_PRECONDITION, "resizeImageNoInterpolation(): image too small");
    const int sp = pitch_of(w), dp = pitch_of(dw);
    float *d_src = nullptr, *d_blur = nullptr, *d_taps = nullptr, *d_out = nullptr;
    int* d_map = nullptr;
    int rc = 0;
    auto cu = [&](cudaError_t e) { if (e != cudaSuccess && !rc) rc = cuda_fail(e, "debug blur", __FILE__, __LINE__); };
    cu(cudaMalloc(&d_src, sizeof(float) * level_px(sp, h)));
    cu(cudaMalloc(&d_blur, sizeof(float) * level_px(sp, h)));
    cu(cudaMalloc(&d_out, sizeof(float) * level_px(dp, dh)));
    cu(cudaMalloc(&d_taps, sizeof(float) * taps.size()));
    if (!rc) {
        cu(cudaMemcpy2D(d_src, sizeof(float) * (size_t)sp, src, sizeof(float) * (size_t)w, sizeof(float) * (size_t)w, (size_t)h, cudaMemcpyHostToDevice));
        cu(cudaMemcpy(d_taps, taps.data(), sizeof(float) * taps.size(), cudaMemcpyHostToDevice));
    }
    CUtensorMap map[2];
    const int bw = stream_box_width(r, mode == 1);
    const bool has_map = !rc && bw > 0 && tma::make_image_map(&map[0], d_src, w, h, 1, (size_t)sp, level_px(sp, h), bw, stream_box_rows()) &&
                         tma::make_image_map(&map[1], d_src, w, h, 1, (size_t)sp, level_px(sp, h), bw, 1);
    if (!rc) {
        BlurArgs a{};
        a.src = d_src; a.src_pitch = sp; a.w = w; a.h = h; a.taps = d_taps; a.taps_host = taps.data(); a.r = r;
        a.map = has_map ? map : nullptr;
        a.map_box = has_map ? bw : 0;
        if (mode == 1) {
            // reduceToNextLevel: decimation fused into the blur epilogue
            std::vector<int> mx = resize_index_map(w, dw), my = resize_index_map(h, dh), inv((size_t)(w + h), -1);
            for (int i = 0; i < dw; ++i) inv[(size_t)mx[(size_t)i]] = i;
            for (int i = 0; i < dh; ++i) inv[(size_t)(w + my[(size_t)i])] = i;
            cu(cudaMalloc(&d_map, sizeof(int) * inv.size()));
            if (!rc) cu(cudaMemcpy(d_map, inv.data(), sizeof(int) * inv.size(), cudaMemcpyHostToDevice));
            a.dst = d_out; a.dst_pitch = dp; a.sel_x = d_map; a.sel_y = d_map + w;
            a.sel_x_host = inv.data(); a.sel_y_host = inv.data() + w;
            if (!rc) rc = launch_blur(a, 1, c->fma, c->slots[0].stream, nullptr);
        } else {
            a.dst = mode == 0 ? d_out : d_blur; a.dst_pitch = sp;
            rc = launch_blur(a, 1, c->fma, c->slots[0].stream, nullptr);
            if (!rc && mode == 2) {
                std::vector<int> both = resize_index_map(w, dw), my = resize_index_map(h, dh);
                both.insert(both.end(), my.begin(), my.end());
                cu(cudaMalloc(&d_map, sizeof(int) * both.size()));
                if (!rc) cu(cudaMemcpy(d_map, both.data(), sizeof(int) * both.size(), cudaMemcpyHostToDevice));
                if (!rc) rc = launch_resize_nn(d_blur, 0, sp, d_out, 0, dp, dw, dh, d_map, d_map + dw, 0, 1, c->slots[0].stream, nullptr);
            }
        }
    }
    if (!rc) cu(cudaStreamSynchronize(c->slots[0].stream));
    if (!rc) cu(cudaMemcpy2D(dst, sizeof(float) * (size_t)dw, d_out, sizeof(float) * (size_t)dp, sizeof(float) * (size_t)dw, (size_t)dh, cudaMemcpyDeviceToHost));
    cudaFree(d_src); cudaFree(d_blur); cudaFree(d_taps); cudaFree(d_out); cudaFree(d_map);
    if (rc) c->error = g_last_error;
    return rc;
}

int sift_gpu_debug_blur(sift_gpu_ctx* c, const float* src, int w, int h, float sigma, float* dst) { return debug_blur_impl(c, src, w, h, sigma, dst, 0); }
int sift_gpu_debug_reduce(sift_gpu_ctx* c, const float* src, int w, int h, float sigma, float* dst) { return debug_blur_impl(c, src, w, h, sigma, dst, 1); }
int sift_gpu_debug_increase(sift_gpu_ctx* c, const float* src, int w, int h, float sigma, float* dst) { return debug_blur_impl(c, src, w, h, sigma, dst, 2); }

struct DebugLayers {
    float* d[3] = {nullptr, nullptr, nullptr};
    ScanLayer L{};
    ScanLayer* dev = nullptr;
    uint32_t *mask = nullptr, *cc = nullptr, *co = nullptr, *n_cand = nullptr;
    Cand* cands = nullptr;
    ~DebugLayers() {
        for (float* p : d) cudaFree(p);
        cudaFree(dev); cudaFree(mask); cudaFree(cc); cudaFree(co); cudaFree(n_cand); cudaFree(cands);
    }
};

static int debug_layers_setup(sift_gpu_ctx* c, DebugLayers& S, const float* d0, const float* d1, const float* d2, int w, int h) {
    const float* hs[3] = {d0, d1, d2};
    const size_t n = level_px(w, h);
    const int pitch = pitch_of(w);   // the kernels read rows 16 bytes at a time: same row pitch rule as the pyramid levels
    for (int i = 0; i < 3; ++i) {
        CTX_CUDA(cudaMalloc(&S.d[i], sizeof(float) * level_px(pitch, h)));
        CTX_CUDA(cudaMemset(S.d[i], 0, sizeof(float) * level_px(pitch, h)));
        CTX_CUDA(cudaMemcpy2D(S.d[i], sizeof(float) * (size_t)pitch, hs[i], sizeof(float) * (size_t)w, sizeof(float) * (size_t)w, (size_t)h, cudaMemcpyHostToDevice));
    }
    S.L.d0 = S.d[0]; S.L.d1 = S.d[1]; S.L.d2 = S.d[2];
    S.L.stride = 0; S.L.pitch = pitch; S.L.w = w; S.L.h = h; S.L.n_yw = (h + 31) / 32; S.L.mask_off = 0; S.L.col_base = 0; S.L.octave = 0; S.L.index = 1;
    set_scan_tiles(&S.L, 1);
    CTX_CUDA(cudaMalloc(&S.dev, sizeof(ScanLayer)));
    CTX_CUDA(cudaMemcpy(S.dev, &S.L, sizeof(ScanLayer), cudaMemcpyHostToDevice));
    CTX_CUDA(cudaMalloc(&S.mask, sizeof(uint32_t) * (size_t)S.L.n_yw * (size_t)w));
    CTX_CUDA(cudaMalloc(&S.cc, sizeof(uint32_t) * (size_t)w));
    CTX_CUDA(cudaMalloc(&S.co, sizeof(uint32_t) * (size_t)w));
    CTX_CUDA(cudaMalloc(&S.n_cand, sizeof(uint32_t)));
    CTX_CUDA(cudaMalloc(&S.cands, sizeof(Cand) * n));
    return 0;
}

int sift_gpu_debug_extrema(sift_gpu_ctx* c, const float* d0, const float* d1, const float* d2, int w, int h, uint16_t* xs,
                           uint16_t* ys, uint32_t capacity, uint32_t* n_out) {
    if (!c || !d0 || !d1 || !d2 || w < 1 || h < 1 || !n_out) return set_error(c, SIFT_GPU_E_INVALID, "bad argument");
    CTX_CUDA(cudaSetDevice(c->prm.device));
    DebugLayers S;
    CTX_TRY(debug_layers_setup(c, S, d0, d1, d2, w, h));
    CTX_TRY(launch_extrema(S.dev, &S.L, 1, w, (uint32_t)S.L.n_yw * (uint32_t)w, S.mask, nullptr, S.cc, S.co, S.cands, 0, S.n_cand, 1, c->slots[0].stream, nullptr));
    CTX_CUDA(cudaStreamSynchronize(c->slots[0].stream));
    uint32_t n = 0;
    CTX_CUDA(cudaMemcpy(&n, S.n_cand, sizeof n, cudaMemcpyDeviceToHost));
    *n_out = n;
    std::vector<Cand> hc(n);
    if (n) CTX_CUDA(cudaMemcpy(hc.data(), S.cands, sizeof(Cand) * n, cudaMemcpyDeviceToHost));
    for (uint32_t i = 0; i < n && i < capacity; ++i) { xs[i] = hc[i].x; ys[i] = hc[i].y; }
    return SIFT_GPU_OK;
}

int sift_gpu_debug_eliminate(sift_gpu_ctx* c, const float* d0, const float* d1, const float* d2, int w, int h,
                             const uint16_t* xs, const uint16_t* ys, uint32_t n, uint8_t* filtered) {
    if (!c || !d0 || !d1 || !d2 || w < 3 || h < 3 || (n && (!xs || !ys || !filtered))) return set_error(c, SIFT_GPU_E_INVALID, "bad argument");
    CTX_CUDA(cudaSetDevice(c->prm.device));
    DebugLayers S;
    CTX_TRY(debug_layers_setup(c, S, d0, d1, d2, w, h));
    std::vector<Cand> hc(n);
    for (uint32_t i = 0; i < n; ++i) {
        if (xs[i] < 1 || xs[i] > w - 2 || ys[i] < 1 || ys[i] > h - 2) return set_error(c, SIFT_GPU_E_INVALID, "candidate on the border");
        hc[i] = Cand{xs[i], ys[i], 0, 1, 0, 0};
    }
    Cand* d_c = nullptr; Surv* d_s = nullptr; uint32_t* d_ns = nullptr;
    int rc = 0;
    if (cudaMalloc(&d_c, sizeof(Cand) * std::max<uint32_t>(n, 1)) != cudaSuccess || cudaMalloc(&d_s, sizeof(Surv) * std::max<uint32_t>(n, 1)) != cudaSuccess ||
        cudaMalloc(&d_ns, sizeof(uint32_t)) != cudaSuccess)
        rc = SIFT_GPU_E_CUDA;
    if (!rc && n && cudaMemcpy(d_c, hc.data(), sizeof(Cand) * n, cudaMemcpyHostToDevice) != cudaSuccess) rc = SIFT_GPU_E_CUDA;
    if (!rc && cudaMemcpy(S.n_cand, &n, sizeof n, cudaMemcpyHostToDevice) != cudaSuccess) rc = SIFT_GPU_E_CUDA;
    if (!rc) rc = launch_eliminate(S.dev, 1, d_c, 0, S.n_cand, d_s, std::max<uint32_t>(n, 1), d_ns, c->slots[0].d_slice, 3, 1, c->slots[0].stream, nullptr);
    if (!rc && cudaStreamSynchronize(c->slots[0].stream) != cudaSuccess) rc = SIFT_GPU_E_CUDA;
    if (!rc && n && cudaMemcpy(hc.data(), d_c, sizeof(Cand) * n, cudaMemcpyDeviceToHost) != cudaSuccess) rc = SIFT_GPU_E_CUDA;
    cudaFree(d_c); cudaFree(d_s); cudaFree(d_ns);
    if (rc) return set_error(c, rc, "CUDA failure in debug eliminate");
    for (uint32_t i = 0; i < n; ++i) filtered[i] = hc[i].filtered;
    return SIFT_GPU_OK;
}

int sift_gpu_debug_get_candidates(sift_gpu_ctx* c, int image_idx, uint16_t* xs, uint16_t* ys, uint16_t* octave, uint16_t* index,
                                  uint8_t* filtered, uint32_t capacity, uint32_t* n_out) {
    if (!c || c->last_slot < 0 || image_idx < 0 || image_idx >= (int)c->slots[c->last_slot].imgs.size() || !n_out)
        return set_error(c, SIFT_GPU_E_INVALID, "bad argument");
    const Slot& S = c->slots[c->last_slot];
    CTX_CUDA(cudaSetDevice(c->prm.device));
    uint32_t n = 0;
    CTX_CUDA(cudaMemcpy(&n, S.d_n_cand + image_idx, sizeof n, cudaMemcpyDeviceToHost));
    *n_out = n;
    const uint32_t m = std::min(n, capacity);
    std::vector<Cand> hc(m);
    if (m) CTX_CUDA(cudaMemcpy(hc.data(), S.d_cands + (size_t)image_idx * c->cand_cap, sizeof(Cand) * m, cudaMemcpyDeviceToHost));
    for (uint32_t i = 0; i < m; ++i) {
        if (xs) xs[i] = hc[i].x;
        if (ys) ys[i] = hc[i].y;
        if (octave) octave[i] = hc[i].octave;
        if (index) index[i] = hc[i].index;
        if (filtered) filtered[i] = hc[i].filtered;
    }
    return SIFT_GPU_OK;
}

int sift_gpu_debug_host_replay(const sift_gpu_params* params, int width, int height, uint32_t n_candidates, const uint32_t* canon,
                               const uint16_t* xs, const uint16_t* ys, const uint8_t* octave, const uint8_t* index, uint32_t n_unfiltered,
                               sift_gpu_keypoint* kps, uint32_t capacity, uint32_t* n_kps, uint32_t* n_survivors) {
    if (!params || !n_kps || (n_unfiltered && (!canon || !xs || !ys || !octave || !index)) || (capacity && !kps)) return SIFT_GPU_E_INVALID;
    if (!(params->octaves > 0) || !(params->dogs_per_epoch >= 3)) return SIFT_GPU_E_ASSERT;
    if (params->octaves > kMaxOctaves || params->dogs_per_epoch + 1 > kMaxGauss) return SIFT_GPU_E_UNSUPPORTED;
    if (width < 1 || height < 1) return SIFT_GPU_E_INVALID;
    // the host-side parts of a context and a plan: nothing here touches a device
    std::unique_ptr<sift_gpu_ctx> c(new sift_gpu_ctx());
    c->prm = *params;
    c->O = params->octaves; c->D = params->dogs_per_epoch; c->G = c->D + 1;
    build_schedule_host(c.get());
    std::unique_ptr<Plan> p(new Plan());
    p->in_w = width; p->in_h = height;
    octave_dims(c.get(), width, height, p->ow, p->oh);
    std::vector<std::pair<int, int>> target_level;
    fill_class_targets(c.get(), p.get(), &target_level);
    std::vector<Surv> surv(n_unfiltered);
    for (uint32_t i = 0; i < n_unfiltered; ++i) {
        if (canon[i] >= n_candidates || (i && canon[i] <= canon[i - 1]) || octave[i] >= c->O || index[i] < 1 || index[i] >= c->D - 1) return SIFT_GPU_E_INVALID;
        surv[i].canon = canon[i]; surv[i].x = xs[i]; surv[i].y = ys[i]; surv[i].octave = octave[i]; surv[i].index = index[i];
    }
    ReplayOut out;
    replay_image(c.get(), p.get(), n_candidates, surv.data(), n_unfiltered, &out);
    if (const char* e = getenv("SIFT_GPU_REPLAY_REPS")) {  // development: CPU cost of the replay, per phase, buffers reused as in a slot
        const int reps = atoi(e);
        for (auto& a : g_rep_ns) a = 0;
        const double t0 = now_ms();
        for (int i = 0; i < reps; ++i) replay_image(c.get(), p.get(), n_candidates, surv.data(), n_unfiltered, &out);
        const double us = (now_ms() - t0) * 1e3 / std::max(1, reps);
        fprintf(stderr, "[sift_gpu replay cpu] %.1f us per image: sort1 %.1f  bounds %.1f  sort2 %.1f  keypoints %.1f\n", us,
                g_rep_ns[1] * 1e-3 / reps, g_rep_ns[2] * 1e-3 / reps, g_rep_ns[3] * 1e-3 / reps, g_rep_ns[4] * 1e-3 / reps);
    }
    if (n_survivors) *n_survivors = out.n_survivors;
    *n_kps = (uint32_t)out.kps.size();
    if (out.status != SIFT_GPU_OK) return out.status;
    if (out.kps.size() > capacity) return SIFT_GPU_E_CAPACITY;
    std::copy(out.kps.begin(), out.kps.end(), kps);
    return SIFT_GPU_OK;
}

int sift_gpu_debug_sort_order_fast(const uint8_t* filtered, uint32_t n, uint32_t* unfiltered_order, uint32_t* n_unfiltered) {
    if ((n && !filtered) || !unfiltered_order || !n_unfiltered) return SIFT_GPU_E_INVALID;
    std::vector<uint32_t> zero_pos;
    for (uint32_t i = 0; i < n; ++i)
        if (!filtered[i]) zero_pos.push_back(i);
    SparseFilterSort sorter;
    std::vector<uint32_t> order = sorter.run(n, zero_pos);
    for (size_t i = 0; i < order.size(); ++i) unfiltered_order[i] = zero_pos[order[i]];
    *n_unfiltered = (uint32_t)order.size();
    return SIFT_GPU_OK;
}

int sift_gpu_debug_sort_order(const uint8_t* filtered, uint32_t n, uint32_t* order) {
    if ((n && (!filtered || !order))) return SIFT_GPU_E_INVALID;
    std::vector<uint32_t> v(n);
    for (uint32_t i = 0; i < n; ++i) v[i] = i | (filtered[i] ? 0x80000000u : 0u);
    std::sort(v.begin(), v.end(), [](uint32_t a, uint32_t b) { return !(a >> 31) && (b >> 31); });
    for (uint32_t i = 0; i < n; ++i) order[i] = v[i] & 0x7fffffffu;
    return SIFT_GPU_OK;
}

}  // extern "C"
