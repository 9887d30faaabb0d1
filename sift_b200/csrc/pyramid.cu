// Gaussian pyramid kernels: alg::convolveWithGauss (reference algorithms.cpp:10-22) with the DoG
// subtraction (algorithms.cpp:52-64) and the decimation of alg::reduceToNextLevel (algorithms.cpp:24-36)
// fused into the epilogue, nearest-neighbour upsampling with the Vigra index walk (algorithms.cpp:46;
// SURVEY A.3) and the u8 -> f32 widening of importImage (main.cpp:52-54).
// Borders are BORDER_TREATMENT_REFLECT (index -j -> j, w-1+j -> w-1-j); taps are applied to source
// indices x-r..x+r in ascending order; the row pass is rounded to fp32 before the column pass
// (SURVEY A.2).  Compiled with -fmad=false: the exact path is mul-then-add like the reference's
// mulss/addss, the FMA path asks for fmaf explicitly.
//
// Two implementations:
//  * blur_stream_kernel<R>: the hot one.  A CTA owns a 256-column strip of a row segment and marches
//    down it in chunks of 8 rows.  Input rows are staged in shared memory by TMA (3-D tensor map
//    x, y, image; out-of-bounds zero fill, then a reflect patch on edge strips), one warp per row:
//    the warp's leader issues the row's boxes, the warp row-filters it with a register-blocked
//    window (4 outputs per float4 group) into a ring of row-filtered lines, and after one
//    __syncthreads per chunk each thread column-filters 8 rows of its own column from the ring.
//    The main loop is unrolled over the ring period so every shared-memory address is a constant.
//  * blur_tile_kernel: any radius / tiny images; plain shared-memory tile.
#include "common.cuh"
#include "tma.cuh"

namespace siftgpu {

__host__ __device__ __forceinline__ int reflect101(int v, int n) {
    if (v < 0) v = -v;
    if (v >= n) v = 2 * (n - 1) - v;
    v = v < 0 ? 0 : v;            // only reachable for padding that no valid output reads
    return v >= n ? n - 1 : v;
}

template <bool FMA>
__device__ __forceinline__ float tap_acc(float acc, float t, float v) {
    if (FMA) return fmaf(t, v, acc);
    return __fadd_rn(acc, __fmul_rn(t, v));
}

__device__ __forceinline__ void store_out(const BlurArgs& a, int b, int gx, int gy, float sum, float lower) {
    if (a.sel_x) {
        const int sx = a.sel_x[gx], sy = a.sel_y[gy];
        if (sx >= 0 && sy >= 0) a.dst[(size_t)b * a.dst_stride + (size_t)sy * a.dst_pitch + sx] = sum;
        return;
    }
    if (a.dst) a.dst[(size_t)b * a.dst_stride + (size_t)gy * a.dst_pitch + gx] = sum;
    if (a.dog) {
        const float dif = __fsub_rn(sum, lower);  // higher - lower, then 128 + dif (algorithms.cpp:58-60)
        a.dog[(size_t)b * a.dog_stride + (size_t)gy * a.dog_pitch + gx] = __fadd_rn(128.0f, dif);
    }
}

// ---------------------------------------------------------------------------------------------
// Generic tile kernel.
constexpr int kTW = 64, kTH = 32, kBlurThreads = 256;

template <bool FMA>
__global__ void __launch_bounds__(kBlurThreads) blur_tile_kernel(BlurArgs a) {
    extern __shared__ float smem[];
    const int r = a.r, w = a.w, h = a.h;
    const int iw = kTW + 2 * r, ih = kTH + 2 * r;
    float* s_taps = smem;                  // 2r+1 (padded to a multiple of 4)
    float* s_in = smem + ((2 * r + 1 + 3) & ~3);
    float* s_tmp = s_in + iw * ih;         // ih rows x kTW

    const int b = blockIdx.z;
    const float* src = a.src + (size_t)b * a.src_stride;
    const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH;

    for (int i = threadIdx.x; i < 2 * r + 1; i += kBlurThreads) s_taps[i] = a.taps[i];
    for (int i = threadIdx.x; i < iw * ih; i += kBlurThreads) {
        const int ty = i / iw, tx = i - ty * iw;
        const int gx = reflect101(x0 - r + tx, w), gy = reflect101(y0 - r + ty, h);
        s_in[i] = src[(size_t)gy * a.src_pitch + gx];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kTW * ih; i += kBlurThreads) {
        const int ty = i / kTW, tx = i - ty * kTW;
        const float* p = s_in + ty * iw + tx;
        float sum = 0.0f;
        for (int j = 0; j <= 2 * r; ++j) sum = tap_acc<FMA>(sum, s_taps[2 * r - j], p[j]);
        s_tmp[i] = sum;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kTW * kTH; i += kBlurThreads) {
        const int ty = i / kTW, tx = i - ty * kTW;
        const int gx = x0 + tx, gy = y0 + ty;
        if (gx >= w || gy >= h) continue;
        const float* p = s_tmp + ty * kTW + tx;
        float sum = 0.0f;
        for (int j = 0; j <= 2 * r; ++j) sum = tap_acc<FMA>(sum, s_taps[2 * r - j], p[j * kTW]);
        store_out(a, b, gx, gy, sum, s_in[(ty + r) * iw + tx + r]);
    }
}

static size_t tile_smem_bytes(int r) {
    return sizeof(float) * (size_t)(((2 * r + 1 + 3) & ~3) + (kTW + 2 * r) * (kTH + 2 * r) + kTW * (kTH + 2 * r));
}

int max_generic_radius() {
    int r = 1;
    while (tile_smem_bytes(r + 1) <= 227 * 1024) ++r;
    return r;
}

// ---------------------------------------------------------------------------------------------
// Streaming kernel.
template <int R>
struct SC {
    static constexpr int CH = 8;        // rows per chunk = warps per CTA
    static constexpr int TW = 256;      // strip width = threads per CTA
    static constexpr int NS = 4;        // staging stages
    static constexpr int RPAD = (R + 3) & ~3;
    static constexpr int SW = (TW + 2 * RPAD <= 288) ? 288 : 320;  // staged floats per row (multiple of 32: 128-B aligned rows)
    static constexpr int NBOX = SW == 288 ? 3 : 2;
    static constexpr int BOXW = SW / NBOX;                          // 96 or 160 floats: 128-B multiples
    static constexpr int RC = ((R + CH - 1) / CH) * CH;             // rows loaded above the first output row
    static constexpr int LAG = (R + RC + CH - 1) / CH;              // chunks between a row entering and its output leaving
    static constexpr int NEED = (LAG + 2) * CH - RC + R;
    static constexpr int RING = NEED <= 32 ? 32 : (NEED <= 64 ? 64 : 128);
    static constexpr int PERIOD = RING / CH;
    static constexpr int NW = 4 + 2 * RPAD;                         // row-pass window floats per group
    static constexpr int OFF = RPAD - R;
    static constexpr size_t SMEM = sizeof(float) * (size_t)(NS * CH * SW + RING * TW) + NS * sizeof(uint64_t);
    static_assert(TW + 2 * RPAD <= SW, "radius too large for the streaming kernel");
    static_assert(PERIOD % NS == 0, "stage index must be a function of the phase");
};

template <int R>
struct TapsP {
    float t[2 * R + 1];
};

struct StreamArgs {
    BlurArgs a;
    int seg;  // output rows per CTA
};

template <int R, bool FMA, bool DECIMATE>
__global__ void __launch_bounds__(256) blur_stream_kernel(const __grid_constant__ CUtensorMap map, const StreamArgs sa,
                                                          const TapsP<R> taps) {
    using C = SC<R>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];     // TMA destinations must be 128-B aligned
    float* stage = reinterpret_cast<float*>(smem_raw);            // [NS][CH][SW]
    float* ring = stage + C::NS * C::CH * C::SW;                  // [RING][TW]
    uint64_t* full = (uint64_t*)(ring + C::RING * C::TW);         // [NS]

    const BlurArgs& a = sa.a;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.z;
    const int w = a.w, h = a.h;
    const int x0 = blockIdx.x * C::TW;
    const int y0 = blockIdx.y * sa.seg;
    const int y1 = min(y0 + sa.seg, h);
    const int n_out_chunks = (y1 - y0 + C::CH - 1) / C::CH;
    const int n_in = n_out_chunks + C::LAG;
    const bool edge = (x0 - C::RPAD < 0) || (x0 - C::RPAD + C::SW > w);
    const float* src = a.src + (size_t)b * a.src_stride;

    if (tid == 0) {
        tma::prefetch_map(&map);
        for (int s = 0; s < C::NS; ++s) tma::mbar_init(&full[s], C::CH);
        tma::fence_barrier_init();
    }
    __syncthreads();

    // warp `warp` owns staging row `warp` of every chunk: its leader issues the row's TMA boxes
    auto issue_row = [&](int i, int s) {
        const int v = y0 - C::RC + i * C::CH + warp;
        const int sy = reflect101(v, h);
        float* dst = stage + (s * C::CH + warp) * C::SW;
        tma::mbar_arrive_expect_tx(&full[s], C::SW * (int)sizeof(float));
#pragma unroll
        for (int q = 0; q < C::NBOX; ++q) tma::load_3d(dst + q * C::BOXW, &map, &full[s], x0 - C::RPAD + q * C::BOXW, sy, b);
    };
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < C::NS; ++i)
            if (i < n_in) issue_row(i, i);
    }

    const int x = x0 + tid;              // this thread's column in the column pass
    const bool active = x < w;
    int sel_col = -1;
    if (DECIMATE && active) sel_col = a.sel_x[x];
    // column bases; per row only a 32-bit y*pitch offset is added
    float* const dst_col = a.dst ? a.dst + (size_t)b * a.dst_stride + (DECIMATE ? (sel_col >= 0 ? sel_col : 0) : x) : nullptr;
    float* const dog_col = a.dog ? a.dog + (size_t)b * a.dog_stride + x : nullptr;
    const float* const src_col = src + x;

    for (int i0 = 0; i0 < n_in; i0 += C::PERIOD) {
#pragma unroll
        for (int p = 0; p < C::PERIOD; ++p) {
            const int i = i0 + p;
            if (i < n_in) {
                const int s = p % C::NS;  // compile-time after unrolling (PERIOD % NS == 0)
                tma::mbar_wait(&full[s], (uint32_t)((i / C::NS) & 1));
                float* st = stage + (s * C::CH + warp) * C::SW;
                if (edge) {
                    // reflect patch of this warp's row: staged column c holds x = x0 - RPAD + c
                    const int v = y0 - C::RC + i * C::CH + warp;
                    const float* srow = src + (size_t)reflect101(v, h) * a.src_pitch;
                    for (int c = lane; c < C::SW; c += 32) {
                        const int gx = x0 - C::RPAD + c;
                        if ((gx < 0 && gx >= -R) || (gx >= w && gx <= w - 1 + R)) st[c] = srow[reflect101(gx, w)];
                    }
                    tma::fence_proxy_async();  // generic-proxy writes above vs. the next TMA write into this row
                    __syncwarp();
                }
                // ---- row pass: two float4 groups per lane, columns 4*lane and 128 + 4*lane ----
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    const int c0 = 4 * lane + 128 * g;
                    float wv[C::NW];
#pragma unroll
                    for (int k = 0; k < C::NW / 4; ++k) {
                        const float4 q = *reinterpret_cast<const float4*>(st + c0 + 4 * k);
                        wv[4 * k] = q.x; wv[4 * k + 1] = q.y; wv[4 * k + 2] = q.z; wv[4 * k + 3] = q.w;
                    }
                    float acc[4];
#pragma unroll
                    for (int j = 0; j <= 2 * R; ++j) {
                        const float t = taps.t[2 * R - j];
#pragma unroll
                        for (int o = 0; o < 4; ++o) acc[o] = j == 0 ? __fmul_rn(t, wv[o + C::OFF]) : tap_acc<FMA>(acc[o], t, wv[o + C::OFF + j]);
                    }
                    const int slot = (p * C::CH) % C::RING;  // + warp (< CH) never wraps: RING is a multiple of CH
                    *reinterpret_cast<float4*>(ring + (slot + warp) * C::TW + c0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
                }
                __syncthreads();  // ring rows of chunk i visible; every lane of this warp is done with staging row (s, warp)
                if (lane == 0 && i + C::NS < n_in) issue_row(i + C::NS, s);
                // ---- column pass for output chunk j = i - LAG ----
                const int j = i - C::LAG;
                if (j >= 0 && active) {
                    const int ub = ((((p - C::LAG) * C::CH + C::RC - R) % C::RING) + C::RING) % C::RING;  // compile-time after unrolling
                    const int yb = y0 + j * C::CH;
                    if (!DECIMATE) {
                        float win[C::CH + 2 * R];
#pragma unroll
                        for (int k = 0; k < C::CH + 2 * R; ++k) win[k] = ring[((ub + k) % C::RING) * C::TW + tid];
                        float acc[C::CH];
#pragma unroll
                        for (int jj = 0; jj <= 2 * R; ++jj) {
                            const float t = taps.t[2 * R - jj];
#pragma unroll
                            for (int o = 0; o < C::CH; ++o) acc[o] = jj == 0 ? __fmul_rn(t, win[o]) : tap_acc<FMA>(acc[o], t, win[o + jj]);
                        }
                        const int nrows = y1 - yb;  // >= 1; a full chunk unless this is the segment's last one
                        if (dst_col) {
#pragma unroll
                            for (int o = 0; o < C::CH; ++o)
                                if (o < nrows) dst_col[(yb + o) * a.dst_pitch] = acc[o];
                        }
                        if (dog_col) {
                            float lower[C::CH];
#pragma unroll
                            for (int o = 0; o < C::CH; ++o) lower[o] = o < nrows ? src_col[(yb + o) * a.src_pitch] : 0.0f;
#pragma unroll
                            for (int o = 0; o < C::CH; ++o)
                                if (o < nrows) dog_col[(yb + o) * a.dog_pitch] = __fadd_rn(128.0f, __fsub_rn(acc[o], lower[o]));
                        }
                    } else {
#pragma unroll
                        for (int o = 0; o < C::CH; ++o) {
                            const int y = yb + o;
                            const int sy = y < y1 ? a.sel_y[y] : -1;  // uniform across the CTA
                            if (sy >= 0) {
                                float acc = 0.0f;
#pragma unroll
                                for (int jj = 0; jj <= 2 * R; ++jj) {
                                    const float v = ring[((ub + o + jj) % C::RING) * C::TW + tid];
                                    acc = jj == 0 ? __fmul_rn(taps.t[2 * R], v) : tap_acc<FMA>(acc, taps.t[2 * R - jj], v);
                                }
                                if (sel_col >= 0) dst_col[sy * a.dst_pitch] = acc;
                            }
                        }
                    }
                }
            }
        }
    }
}

template <int R>
static int launch_stream_r(const BlurArgs& a, int batch, bool fma, cudaStream_t s) {
    using C = SC<R>;
    StreamArgs sa;
    sa.a = a;
    // rows per CTA: enough CTAs for ~2 waves when the batch allows, never less than 64 rows (halo amortisation)
    const int strips = (a.w + C::TW - 1) / C::TW;
    int segs = (2 * 148 * 2 + strips * batch - 1) / (strips * batch);
    int seg = (a.h + segs - 1) / segs;
    if (seg < 64) seg = 64;
    seg = (seg + C::CH - 1) / C::CH * C::CH;
    if (seg > a.h) seg = (a.h + C::CH - 1) / C::CH * C::CH;
    sa.seg = seg;
    dim3 grid(strips, (a.h + seg - 1) / seg, batch);
    TapsP<R> tp;
    for (int i = 0; i < 2 * R + 1; ++i) tp.t[i] = a.taps_host[i];
    const bool dec = a.sel_x != nullptr;
#define LAUNCH(F, D)                                                                                                        \
    do {                                                                                                                    \
        static bool attr = false;                                                                                           \
        if (!attr) {                                                                                                        \
            SIFT_CUDA_TRY(cudaFuncSetAttribute(blur_stream_kernel<R, F, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM)); \
            attr = true;                                                                                                    \
        }                                                                                                                   \
        blur_stream_kernel<R, F, D><<<grid, 256, C::SMEM, s>>>(*a.map, sa, tp);                                             \
    } while (0)
    if (fma) { if (dec) LAUNCH(true, true); else LAUNCH(true, false); }
    else { if (dec) LAUNCH(false, true); else LAUNCH(false, false); }
#undef LAUNCH
    SIFT_CUDA_TRY(cudaGetLastError());
    return 0;
}

int stream_box_width(int r) {
    switch (r) {
        case 3: return SC<3>::BOXW;
        case 5: return SC<5>::BOXW;
        case 7: return SC<7>::BOXW;
        case 10: return SC<10>::BOXW;
        case 14: return SC<14>::BOXW;
        case 19: return SC<19>::BOXW;
        default: return 0;
    }
}

int launch_blur(const BlurArgs& a, int batch, bool fma, cudaStream_t s, uint64_t* launches) {
    if (launches) ++*launches;
    if (a.map && a.taps_host && a.w >= 32 && a.h >= 16) {
        switch (a.r) {
            case 3: return launch_stream_r<3>(a, batch, fma, s);
            case 5: return launch_stream_r<5>(a, batch, fma, s);
            case 7: return launch_stream_r<7>(a, batch, fma, s);
            case 10: return launch_stream_r<10>(a, batch, fma, s);
            case 14: return launch_stream_r<14>(a, batch, fma, s);
            case 19: return launch_stream_r<19>(a, batch, fma, s);
            default: break;
        }
    }
    const int r = a.r;
    const size_t smem = tile_smem_bytes(r);
    dim3 grid((a.w + kTW - 1) / kTW, (a.h + kTH - 1) / kTH, batch);
    static bool attr_set[2] = {false, false};
    if (!attr_set[fma ? 1 : 0]) {
        if (fma) SIFT_CUDA_TRY(cudaFuncSetAttribute(blur_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        else SIFT_CUDA_TRY(cudaFuncSetAttribute(blur_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set[fma ? 1 : 0] = true;
    }
    if (fma) blur_tile_kernel<true><<<grid, kBlurThreads, smem, s>>>(a);
    else blur_tile_kernel<false><<<grid, kBlurThreads, smem, s>>>(a);
    SIFT_CUDA_TRY(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// resizeImageNoInterpolation: dst(x, y) = src(map_x[x], map_y[y]); the maps are produced on the host
// by the literal accumulated-double walk.
__global__ void resize_nn_kernel(const float* __restrict__ src, size_t src_stride, int src_pitch, float* __restrict__ dst,
                                 size_t dst_stride, int dst_pitch, int dw, int dh, const int* __restrict__ map_x,
                                 const int* __restrict__ map_y) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= dw || y >= dh) return;
    const int b = blockIdx.z;
    dst[(size_t)b * dst_stride + (size_t)y * dst_pitch + x] = src[(size_t)b * src_stride + (size_t)map_y[y] * src_pitch + map_x[x]];
}

int launch_resize_nn(const float* src, size_t src_stride, int src_pitch, float* dst, size_t dst_stride, int dst_pitch, int dw,
                     int dh, const int* map_x, const int* map_y, int batch, cudaStream_t s, uint64_t* launches) {
    dim3 grid((dw + 255) / 256, dh, batch);
    resize_nn_kernel<<<grid, 256, 0, s>>>(src, src_stride, src_pitch, dst, dst_stride, dst_pitch, dw, dh, map_x, map_y);
    if (launches) ++*launches;
    SIFT_CUDA_TRY(cudaGetLastError());
    return 0;
}

__global__ void u8_to_f32_kernel(const uint8_t* __restrict__ src, size_t src_stride, int src_pitch, float* __restrict__ dst,
                                 size_t dst_stride, int dst_pitch, int w, int h) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.z;
    if (x >= w) return;
    for (int y = blockIdx.y; y < h; y += gridDim.y)
        dst[(size_t)b * dst_stride + (size_t)y * dst_pitch + x] = (float)src[(size_t)b * src_stride + (size_t)y * src_pitch + x];
}

int launch_u8_to_f32(const uint8_t* src, size_t src_stride, int src_pitch, float* dst, size_t dst_stride, int dst_pitch, int w,
                     int h, int batch, cudaStream_t s, uint64_t* launches) {
    dim3 grid((w + 255) / 256, h < 256 ? h : 256, batch);
    u8_to_f32_kernel<<<grid, 256, 0, s>>>(src, src_stride, src_pitch, dst, dst_stride, dst_pitch, w, h);
    if (launches) ++*launches;
    SIFT_CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace siftgpu
