// Orientation assignment and descriptors, one warp per keypoint.
//
// Orientation (reference sift.cpp:163-203, :220-286; algorithms.cpp:108-133, :153-178):
//   gradient magnitude (f32 differences, sqrt in double) and orientation (atan2f radians, +360 in
//   f32, fmod 360) of the nearest Gaussian level G*, 16x16 window, 36 bins
//   bins[(u16)floorf(o/10) % 35] += mag * G*, accumulated in the reference's x-outer / y-inner order
//   (one lane per bin walks the samples in that order, so every bin's fp32 sum is sequential);
//   peak logic and the rank-deficient least-squares parabola exactly as written.
// Descriptors (sift.cpp:60-128; algorithms.cpp:135-150, :210-223): the reference adds
//   p.orientation to the orientation pyramid and the top-left 16x16 of blur(G*, 1.6) to the
//   magnitude pyramid IN PLACE for every keypoint in vector order, so a keypoint sees the
//   accumulated contributions of every earlier keypoint whose window overlaps its own.  Each warp
//   replays, in vector order, the earlier keypoints of the same level that overlap its window, then
//   its own contribution, bins with (u16)floorf(o/45) % 7 and L1-normalises each 4x4 cell.
#include <cstdlib>

#include "common.cuh"
#include "vigra_qr.cuh"

namespace siftgpu {

constexpr int kWin = 2 * kRegion;  // 16

// alg::gradientMagnitude / gradientOrientation at interior pixel (x, y); 0 on the 1-px border
// (the pyramids are zero-initialised and only the interior is filled, sift.cpp:137-138).
__device__ __forceinline__ void gradient_at(const float* __restrict__ G, int pitch, int w, int h, int x, int y, float* mag,
                                            float* ori) {
    if (x < 1 || y < 1 || x > w - 2 || y > h - 2) {
        *mag = 0.0f;
        *ori = 0.0f;
        return;
    }
    const float dx = G[(size_t)y * pitch + x + 1] - G[(size_t)y * pitch + x - 1];
    const float dy = G[(size_t)(y + 1) * pitch + x] - G[(size_t)(y - 1) * pitch + x];
    *mag = (float)sqrt((double)dx * (double)dx + (double)dy * (double)dy);
    const float r = atan2f(dy, dx);
    const float s = r + 360.0f;
    // (float)fmod((double)s, 360.0): s lies in [360-pi, 360+pi], where the remainder is s or s - 360, and s - 360 is exact
    // in fp32 (s is a multiple of 2^-15 and the difference is below 4)
    *ori = s >= 360.0f ? s - 360.0f : s;
}

// alg::vertexParabola (algorithms.cpp:153-178).
__device__ float vertex_parabola(int lx, float ly, int px, float py, int rx, float ry) {
    float a[9], b[3], res[3] = {0.0f, 0.0f, 0.0f};
    a[0] = (float)((double)lx * (double)lx); a[1] = (float)lx; a[2] = 0.0f;
    a[3] = (float)((double)px * (double)px); a[4] = (float)px; a[5] = 0.0f;
    a[6] = (float)((double)rx * (double)rx); a[7] = (float)rx; a[8] = 0.0f;
    b[0] = ly; b[1] = py; b[2] = ry;
    qr::solve3_fast(a, b, res);
    return -res[1] / (2 * res[0]);
}

// Sift::_findPeaks (sift.cpp:220-286).  Emulates std::set<float>: sorted, unique; a NaN is only
// ever kept when it is the first value inserted, and then nothing else is.
__device__ int find_peaks(const float* histo, float* p, float* out) {
    int max_index = 0;
    for (int i = 0; i < 36; ++i) p[i] = histo[i];
    for (int i = 1; i < 36; ++i)
        if (p[max_index] < p[i]) max_index = i;  // std::max_element: first of the largest
    const float range = (float)((double)histo[max_index] * 0.8);
    for (int i = 0; i < 36; ++i)
        if (p[i] < range) p[i] = -1.0f;
    for (int i = 1; i < 35; ++i)
        if (p[i] < p[i - 1] || p[i] < p[i + 1]) p[i] = -1.0f;
    int n = 0;
    bool nan_first = false;
    for (int pass = 0; pass < 37; ++pass) {
        int i;
        if (pass == 0) i = max_index;
        else {
            i = pass - 1;
            if (!(p[i] > -1.0f) || i == max_index) continue;
        }
        int lx, rx;
        float ly, ry;
        if (i == 0) { lx = 35 * 10 + 5; ly = histo[35]; } else { lx = (i - 1) * 10 + 5; ly = histo[i - 1]; }
        if (i == 35) { rx = 5; ry = histo[0]; } else { rx = (i + 1) * 10 + 5; ry = histo[i + 1]; }
        const float v = vertex_parabola(lx, ly, i * 10 + 5, histo[i], rx, ry);
        if (pass == 0) {
            out[n++] = v;
            nan_first = (v != v);
            continue;
        }
        if (nan_first || v != v) continue;
        int pos = 0;
        bool dup = false;
        while (pos < n && out[pos] < v) ++pos;
        if (pos < n && out[pos] == v) dup = true;
        if (dup) continue;
        for (int k = n; k > pos; --k) out[k] = out[k - 1];
        out[pos] = v;
        ++n;
    }
    return n;
}

// Two phases per block of 32 keypoints.  Phase 1, warp per keypoint (8 keypoints per warp, one after the other): the 256
// window samples and the 36-bin histogram.  Phase 2, THREAD per keypoint: peak logic and the least-squares parabola —
// long scalar code that would otherwise run on one lane of a warp with the other 31 idle.
constexpr int kOriKeys = 32;    // keypoints per CTA pass (phase 2 runs on warp 0)
constexpr int kOriThreads = 128;
constexpr int kHistPitch = 37;  // odd pitch: phase 2's column walk over 32 histograms is conflict-free

__global__ void __launch_bounds__(kOriThreads) orientation_kernel(const LevelRef* __restrict__ targets, const KeyIn* __restrict__ keys,
                                                               const uint32_t* __restrict__ key_img, uint32_t n_keys,
                                                               float* __restrict__ orientation, uint32_t* __restrict__ n_peaks,
                                                               float* __restrict__ peaks, float2* __restrict__ grad_cache) {
    __shared__ __align__(16) float s_val[4][kWin * kWin];
    __shared__ uint16_t s_bin[4][kWin * kWin];
    __shared__ float s_hist[kOriKeys][kHistPitch];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    for (uint32_t base = blockIdx.x * kOriKeys; base < n_keys; base += gridDim.x * kOriKeys) {
        constexpr int per_warp = kOriKeys / (kOriThreads / 32);
        for (int i = 0; i < per_warp; ++i) {
            const int slot = wib * per_warp + i;
            const uint32_t k = base + (uint32_t)slot;
            if (k >= n_keys) break;  // warp-uniform
            float* hist = s_hist[slot];
            const KeyIn key = keys[k];
            const LevelRef T = targets[key.tgt];
            const float* G = T.base + (size_t)key_img[k] * T.stride;
            const int x0 = key.x - kRegion, y0 = key.y - kRegion;
            // sample s = wx*16 + wy is the reference's (x outer, y inner) visiting order
            bool same_bin = true;  // do all 256 samples fall into the bin of sample 0?
            uint16_t bin0 = 0;
#pragma unroll
            for (int j = 0; j < (kWin * kWin) / 32; ++j) {
                const int s = lane + 32 * j;
                const int wx = s >> 4, wy = s & 15;
                float mag, ori;
                gradient_at(G, T.pitch, T.w, T.h, x0 + wx, y0 + wy, &mag, &ori);
                // the descriptor kernel needs the same 256 (magnitude, orientation) pairs of this window: the double sqrt and the
                // atan2f are half of its instructions, 2 KB per keypoint through L2 / HBM is cheaper
                if (grad_cache) grad_cache[(size_t)k * (kWin * kWin) + s] = make_float2(mag, ori);
                const float g = G[(size_t)(y0 + wy) * T.pitch + x0 + wx];
                s_val[wib][s] = mag * g;
                uint16_t bi = (uint16_t)(int)floorf(ori / 10);
                bi = bi % 35;
                s_bin[wib][s] = bi;
                if (j == 0) bin0 = (uint16_t)__shfl_sync(0xffffffffu, (int)bi, 0);
                same_bin = same_bin && bi == bin0;
            }
            same_bin = __all_sync(0xffffffffu, same_bin);
            __syncwarp();
            if (same_bin) {
                // Every sample lands in one bin (with the reference's radians-as-degrees binning that is always bin 0):
                // its sum is the plain sequential sum of the 256 products in visiting order; the other 35 bins are 0.
                if (lane < 4) hist[32 + lane] = 0.0f;
                hist[lane] = 0.0f;
                __syncwarp();
                if (lane == 0) {
                    float acc = 0.0f;
                    const float4* v4 = reinterpret_cast<const float4*>(s_val[wib]);
#pragma unroll 4
                    for (int q = 0; q < (kWin * kWin) / 4; ++q) {
                        const float4 v = v4[q];
                        acc = acc + v.x; acc = acc + v.y; acc = acc + v.z; acc = acc + v.w;
                    }
                    hist[bin0] = acc;
                }
            } else {
                // general case: lane b owns bin b (lanes 0..3 also bins 32..35) and walks the samples in visiting order
                float acc0 = 0.0f, acc1 = 0.0f;
                const int binb = 32 + lane;
#pragma unroll 4
                for (int s = 0; s < kWin * kWin; ++s) {
                    const int bi = s_bin[wib][s];
                    const float v = s_val[wib][s];
                    if (bi == lane) acc0 = acc0 + v;
                    if (bi == binb) acc1 = acc1 + v;
                }
                hist[lane] = acc0;
                if (lane < 4) hist[binb] = acc1;
            }
            __syncwarp();
        }
        __syncthreads();
        {
            const uint32_t k = base + threadIdx.x;
            if (threadIdx.x < kOriKeys && k < n_keys) {
                float histo[36], work[36], out[36];
#pragma unroll
                for (int i = 0; i < 36; ++i) histo[i] = s_hist[threadIdx.x][i];
                const int n = find_peaks(histo, work, out);
                orientation[k] = out[0];
                n_peaks[k] = (uint32_t)n;
                if (n > 1)
                    for (int i = 0; i < n; ++i) peaks[(size_t)k * 36 + i] = out[i];
            }
        }
        __syncthreads();
    }
}

// blur(level, 1.6f) evaluated on [0,16)^2 only (the part sift.cpp:88-92 reads).
template <bool FMA>
__global__ void __launch_bounds__(256) weight_table_kernel(const LevelRef* __restrict__ targets, const float* __restrict__ taps,
                                                           int r, float* __restrict__ tables, int n_targets) {
    extern __shared__ float s_tmp[];  // rows 0..(15+r) x 16
    const int t = blockIdx.x, b = blockIdx.y;
    const LevelRef T = targets[t];
    const float* G = T.base + (size_t)b * T.stride;
    const int rows = (kWin + r) < T.h ? (kWin + r) : T.h;
    auto refl = [](int v, int n) { if (v < 0) v = -v; if (v >= n) v = 2 * (n - 1) - v; return v; };
    for (int i = threadIdx.x; i < rows * kWin; i += blockDim.x) {
        const int y = i / kWin, x = i - y * kWin;
        float sum = 0.0f;
        if (x < T.w)
            for (int j = 0; j <= 2 * r; ++j) {
                const float v = G[(size_t)y * T.pitch + refl(x + j - r, T.w)];
                sum = FMA ? fmaf(taps[2 * r - j], v, sum) : __fadd_rn(sum, __fmul_rn(taps[2 * r - j], v));
            }
        s_tmp[i] = sum;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kWin * kWin; i += blockDim.x) {
        const int y = i / kWin, x = i - y * kWin;
        float sum = 0.0f;
        if (x < T.w && y < T.h)
            for (int j = 0; j <= 2 * r; ++j) {
                const float v = s_tmp[refl(y + j - r, T.h) * kWin + x];
                sum = FMA ? fmaf(taps[2 * r - j], v, sum) : __fadd_rn(sum, __fmul_rn(taps[2 * r - j], v));
            }
        tables[((size_t)b * n_targets + t) * (kWin * kWin) + y * kWin + x] = sum;
    }
}

constexpr int kMaxGridHits = 256;   // overlapping predecessors a warp sorts in shared memory; more than that: full scan

// ---- key grid (KeyGrid, common.cuh): count, per-image exclusive scan, fill -------------------------------------------------
__device__ __forceinline__ uint32_t key_cell(const KeyIn& k, const KeyGrid& g) {
    const int cx = min((int)k.x >> 4, g.cw - 1), cy = min((int)k.y >> 4, g.ch - 1);
    return (uint32_t)k.tgt * (uint32_t)(g.cw * g.ch) + (uint32_t)(cy * g.cw + cx);
}
__global__ void key_grid_count_kernel(const KeyIn* __restrict__ keys, const uint32_t* __restrict__ key_img, uint32_t n_keys, KeyGrid g) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_keys) return;
    atomicAdd(g.count + (size_t)key_img[k] * g.cells_per_image + key_cell(keys[k], g), 1u);
}
// one CTA per image: exclusive scan of its cell counts (offsets are relative to the image's first key)
__global__ void __launch_bounds__(1024) key_grid_scan_kernel(KeyGrid g) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    const uint32_t* in = g.count + (size_t)blockIdx.x * g.cells_per_image;
    uint32_t* out = g.offset + (size_t)blockIdx.x * g.cells_per_image;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint32_t base = 0; base < g.cells_per_image; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < g.cells_per_image ? in[i] : 0u;
        uint32_t s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        if (lane == 31) warp_sums[wid] = s;
        __syncthreads();
        if (wid == 0) {
            uint32_t ws = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, ws, o);
                if (lane >= o) ws += t;
            }
            warp_sums[lane] = ws;
        }
        __syncthreads();
        const uint32_t before = carry + (wid ? warp_sums[wid - 1] : 0u);
        if (i < g.cells_per_image) out[i] = before + s - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + s;
        __syncthreads();
    }
}
__global__ void key_grid_fill_kernel(const KeyIn* __restrict__ keys, const uint32_t* __restrict__ key_img, const uint32_t* __restrict__ key_first,
                                     uint32_t n_keys, KeyGrid g) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_keys) return;
    const uint32_t img = key_img[k];
    const size_t cell = (size_t)img * g.cells_per_image + key_cell(keys[k], g);
    const uint32_t pos = atomicAdd(g.cursor + cell, 1u);
    g.cell_keys[key_first[img] + g.offset[cell] + pos] = k;
}

__global__ void __launch_bounds__(128) descriptor_kernel(const LevelRef* __restrict__ targets, int n_targets,
                                                         const float* __restrict__ tables, const KeyIn* __restrict__ keys,
                                                         const uint32_t* __restrict__ key_img,
                                                         const uint32_t* __restrict__ key_first, uint32_t n_keys,
                                                         const float* __restrict__ orientation, float* __restrict__ desc,
                                                         const float2* __restrict__ grad_cache, const KeyGrid grid) {
    __shared__ float s_val[4][kWin * kWin];
    __shared__ uint32_t s_hit[4][2][kMaxGridHits];   // overlapping predecessors found through the grid: unsorted, sorted
    __shared__ uint16_t s_bin[4][kWin * kWin];
    __shared__ float s_w[4][kWin * kWin];   // this key's weight table: every overlapping predecessor reads it once per covered pixel
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < n_keys; k += warps) {
        const KeyIn key = keys[k];
        const uint32_t img = key_img[k];
        const LevelRef T = targets[key.tgt];
        const float* G = T.base + (size_t)img * T.stride;
        const float* Wg = tables + ((size_t)img * n_targets + key.tgt) * (kWin * kWin);  // W[y*16 + x]
        float* W = s_w[wib];
#pragma unroll
        for (int j = 0; j < 8; ++j) W[lane + 32 * j] = Wg[lane + 32 * j];
        __syncwarp();
        const int x0 = key.x - kRegion, y0 = key.y - kRegion;

        // lane owns window pixels s = lane + 32*j, s = wx*16 + wy
        float O[8], M[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int s = lane + 32 * j, wx = s >> 4, wy = s & 15;
            if (grad_cache) {   // written by the orientation kernel for this very key list
                const float2 mo = grad_cache[(size_t)k * (kWin * kWin) + s];
                M[j] = mo.x; O[j] = mo.y;
            } else {
                gradient_at(G, T.pitch, T.w, T.h, x0 + wx, y0 + wy, &M[j], &O[j]);
            }
        }
        // replay earlier keypoints of this image and level whose window overlaps, in vector order
        auto apply_overlap = [&](int mx, int my, float theta) {
            // This lane's pixels are (wx, wy) = (2j + (lane >> 4), lane & 15): its row inside the earlier window is the same
            // for all eight, its column advances by two.  No branch per pixel: the table index is clamped and the two
            // additions are selected (the sums only change where the windows overlap).
            const int mx0 = mx - kRegion, my0 = my - kRegion;
            const int ly = y0 + (lane & 15) - my0;
            const bool row_ok = (unsigned)ly < (unsigned)kWin;
            const int lx0 = x0 + (lane >> 4) - mx0;
            const float* wrow = W + (row_ok ? ly : 0) * kWin;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int lx = lx0 + 2 * j;
                const bool ok = row_ok && (unsigned)lx < (unsigned)kWin;
                const float wv = wrow[ok ? lx : 0];
                O[j] = ok ? O[j] + theta : O[j];
                M[j] = ok ? M[j] + wv : M[j];
            }
        };
        bool use_scan = grid.count == nullptr;
        if (!use_scan) {
            // candidates: the keys of the 3 x 3 cells around this key's cell (same image, same target level), unordered
            const uint32_t* cnt = grid.count + (size_t)img * grid.cells_per_image + (size_t)key.tgt * grid.cw * grid.ch;
            const uint32_t* off = grid.offset + (size_t)img * grid.cells_per_image + (size_t)key.tgt * grid.cw * grid.ch;
            const int cx = key.x >> 4, cy = key.y >> 4;
            uint32_t* hits = s_hit[wib][0];
            uint32_t* sorted = s_hit[wib][1];
            uint32_t n_hit = 0;
            for (int dy = -1; dy <= 1 && n_hit <= kMaxGridHits; ++dy) {
                const int yy = cy + dy;
                if (yy < 0 || yy >= grid.ch) continue;
                for (int dx = -1; dx <= 1 && n_hit <= kMaxGridHits; ++dx) {
                    const int xx = cx + dx;
                    if (xx < 0 || xx >= grid.cw) continue;
                    const uint32_t cell = (uint32_t)(yy * grid.cw + xx);
                    const uint32_t c_n = cnt[cell], c_off = key_first[img] + off[cell];
                    for (uint32_t base = 0; base < c_n; base += 32) {
                        const uint32_t i = base + lane;
                        uint32_t m = 0xffffffffu;
                        bool hit = false;
                        if (i < c_n) {
                            m = grid.cell_keys[c_off + i];
                            if (m < k) {
                                const KeyIn km = keys[m];
                                const int ddx = (int)km.x - (int)key.x, ddy = (int)km.y - (int)key.y;
                                hit = ddx > -kWin && ddx < kWin && ddy > -kWin && ddy < kWin;
                            }
                        }
                        const unsigned ballot = __ballot_sync(0xffffffffu, hit);
                        const uint32_t slot = n_hit + (uint32_t)__popc(ballot & ((1u << lane) - 1u));
                        if (hit && slot < kMaxGridHits) hits[slot] = m;
                        n_hit += (uint32_t)__popc(ballot);
                    }
                }
            }
            if (n_hit > kMaxGridHits) {
                use_scan = true;   // a crowd (hundreds of windows over one spot): the full scan below handles any number
            } else if (n_hit) {
                __syncwarp();
                // vector order = ascending key index: rank sort (indices are distinct)
                for (uint32_t e = lane; e < n_hit; e += 32) {
                    const uint32_t v = hits[e];
                    uint32_t rank = 0;
                    for (uint32_t i = 0; i < n_hit; ++i) rank += hits[i] < v ? 1u : 0u;
                    sorted[rank] = v;
                }
                __syncwarp();
                for (uint32_t base = 0; base < n_hit; base += 32) {
                    const uint32_t i = base + lane;
                    KeyIn km = key;
                    float th = 0.0f;
                    if (i < n_hit) { km = keys[sorted[i]]; th = orientation[sorted[i]]; }
                    const uint32_t n_here = min(32u, n_hit - base);
                    for (uint32_t src = 0; src < n_here; ++src)
                        apply_overlap(__shfl_sync(0xffffffffu, (int)km.x, src), __shfl_sync(0xffffffffu, (int)km.y, src), __shfl_sync(0xffffffffu, th, src));
                }
                __syncwarp();
            }
        }
        if (use_scan) {
        const uint32_t first = key_first[img];
        constexpr int kAhead = 4;  // key chunks loaded per step: their global-load latencies overlap
        for (uint32_t base0 = first; base0 < k; base0 += 32 * kAhead) {
            KeyIn kmv[kAhead];
            float thv[kAhead];   // the earlier keys' orientations travel with them (no dependent load per overlap)
#pragma unroll
            for (int c = 0; c < kAhead; ++c) {
                const uint32_t m = base0 + 32 * c + lane;
                kmv[c] = keys[m < k ? m : k];
                thv[c] = orientation[m < k ? m : k];
            }
#pragma unroll
            for (int c = 0; c < kAhead; ++c) {
                const uint32_t base = base0 + 32 * c;
                if (base >= k) break;
                const uint32_t m = base + lane;
                const KeyIn km = kmv[c];
                bool hit = false;
                if (m < k) {
                    const int ddx = (int)km.x - (int)key.x, ddy = (int)km.y - (int)key.y;
                    hit = km.tgt == key.tgt && ddx > -kWin && ddx < kWin && ddy > -kWin && ddy < kWin;
                }
                unsigned ballot = __ballot_sync(0xffffffffu, hit);
                while (ballot) {
                    const int src = __ffs(ballot) - 1;
                    ballot &= ballot - 1;
                    apply_overlap(__shfl_sync(0xffffffffu, (int)km.x, src), __shfl_sync(0xffffffffu, (int)km.y, src), __shfl_sync(0xffffffffu, thv[c], src));
                }
            }
        }
        }
        const float theta = orientation[k];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int s = lane + 32 * j, wx = s >> 4, wy = s & 15;
            O[j] = O[j] + theta;
            M[j] = M[j] + W[wy * kWin + wx];
            const float g = G[(size_t)(y0 + wy) * T.pitch + x0 + wx];
            s_val[wib][s] = M[j] * g;
            uint16_t bi = (uint16_t)(int)floorf(O[j] / 45);
            s_bin[wib][s] = bi % 7;
        }
        __syncwarp();
        // 16 cells (cx outer, cy inner); lanes 0..15 take one cell each
        if (lane < 16) {
            const int cx = (lane >> 2) * 4, cy = (lane & 3) * 4;
            float bins[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) bins[q] = 0.0f;
            for (int x = 0; x < 4; ++x)
                for (int y = 0; y < 4; ++y) {
                    const int s = (cx + x) * kWin + cy + y;
                    const int bi = s_bin[wib][s];
                    const float v = s_val[wib][s];
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        if (q == bi) bins[q] = bins[q] + v;
                }
            float length = 0.0f;
#pragma unroll
            for (int q = 0; q < 8; ++q) length = length + bins[q];
            float* out = desc + (size_t)k * kDescLen + lane * 8;
#pragma unroll
            for (int q = 0; q < 8; ++q) out[q] = (length == 0.0f) ? bins[q] : bins[q] / length;
        }
        __syncwarp();
    }
}

int launch_weight_tables(const LevelRef* targets_dev, int n_targets, const float* taps16, int r16, float* tables,
                         bool fma, int batch, cudaStream_t s, uint64_t* launches) {
    dim3 grid(n_targets, batch);
    const size_t smem = sizeof(float) * (size_t)(kWin + r16) * kWin;
    if (fma) weight_table_kernel<true><<<grid, 256, smem, s>>>(targets_dev, taps16, r16, tables, n_targets);
    else weight_table_kernel<false><<<grid, 256, smem, s>>>(targets_dev, taps16, r16, tables, n_targets);
    if (launches) ++*launches;
    SIFT_CUDA_TRY(cudaGetLastError());
    return 0;
}

int launch_orientation(const LevelRef* targets_dev, int n_targets, const KeyIn* keys, const uint32_t* key_img,
                       uint32_t n_keys, float* orientation, uint32_t* n_peaks, float* peaks, float2* grad_cache, cudaStream_t s,
                       uint64_t* launches) {
    (void)n_targets;
    if (n_keys == 0) return 0;
    const unsigned blocks = (unsigned)((n_keys + kOriKeys - 1) / kOriKeys);
    orientation_kernel<<<blocks < 148u * 8u ? blocks : 148u * 8u, kOriThreads, 0, s>>>(targets_dev, keys, key_img, n_keys, orientation,
                                                                                n_peaks, peaks, grad_cache);
    if (launches) ++*launches;
    SIFT_CUDA_TRY(cudaGetLastError());
    return 0;
}

int launch_descriptors(const LevelRef* targets_dev, int n_targets, const float* tables, const KeyIn* keys,
                       const uint32_t* key_img, const uint32_t* key_first, uint32_t n_keys, const float* orientation,
                       float* desc, const float2* grad_cache, const KeyGrid* grid, int batch, cudaStream_t s, uint64_t* launches) {
    if (n_keys == 0) return 0;
    KeyGrid g{};
    static const bool grid_off = getenv("SIFT_GPU_NO_KEY_GRID") != nullptr;
    // (one CTA per image scans the cell counts: worth it for video-sized levels, not for the 400 k cells of a 7680x4320 one)
    if (grid && grid->count && !grid_off && grid->cells_per_image <= 32768u) {
        g = *grid;
        const size_t words = (size_t)batch * g.cells_per_image;
        SIFT_CUDA_TRY(cudaMemsetAsync(g.count, 0, sizeof(uint32_t) * words, s));
        SIFT_CUDA_TRY(cudaMemsetAsync(g.cursor, 0, sizeof(uint32_t) * words, s));
        const unsigned kb = (n_keys + 255) / 256;
        key_grid_count_kernel<<<kb, 256, 0, s>>>(keys, key_img, n_keys, g);
        key_grid_scan_kernel<<<batch, 1024, 0, s>>>(g);
        key_grid_fill_kernel<<<kb, 256, 0, s>>>(keys, key_img, key_first, n_keys, g);
        if (launches) *launches += 3;
    }
    const unsigned blocks = (unsigned)((n_keys + 3) / 4);
    descriptor_kernel<<<blocks < 148u * 8u ? blocks : 148u * 8u, 128, 0, s>>>(targets_dev, n_targets, tables, keys, key_img,
                                                                               key_first, n_keys, orientation, desc, grad_cache, g);
    if (launches) ++*launches;
    SIFT_CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace siftgpu
