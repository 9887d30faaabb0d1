// Gaussian pyramid kernels: alg::convolveWithGauss (reference algorithms.cpp:10-22) with the DoG
// subtraction (algorithms.cpp:52-64) and the decimation of alg::reduceToNextLevel (algorithms.cpp:24-36)
// fused into the epilogue, nearest-neighbour upsampling with the Vigra index walk (algorithms.cpp:46;
// SURVEY A.3) and the u8 -> f32 widening of importImage (main.cpp:52-54).
// Borders are BORDER_TREATMENT_REFLECT (index -j -> j, w-1+j -> w-1-j); taps are applied to source
// indices x-r..x+r in ascending order; the row pass is rounded to fp32 before the column pass
// (SURVEY A.2).  Compiled with -fmad=false: the exact path is mul-then-add like the reference's
// mulss/addss, the FMA path asks for fmaf explicitly.
//
// Three implementations (launch_blur picks):
//  * blur_stream_kernel<R>: the hot one (radii 3, 5, 7, 10, 14, 19 on levels taller than 300 rows).
//    Every warp is an independent pipeline over a 64-column strip of a row segment: TMA stages 8-row
//    boxes (3-D tensor map x, y, image; out-of-bounds zero fill, then a reflect patch on edge strips),
//    the warp row-filters them into its private ring of row-filtered lines and column-filters from the
//    ring; DoG and decimation are fused into the epilogue.  No CTA-wide barrier.
//  * blur_strip_kernel: small levels (height <= 300), any radius: one CTA per full-height column strip.
//  * blur_tile_kernel: any radius on large levels; plain shared-memory tile.
#include <cmath>
#include <cstdlib>

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

    const int b = blockIdx.z + a.z0;
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
// Column-strip kernel for the small levels of the pyramid (any radius).  On a 240 x 135 or 120 x 68 level the
// streaming kernel's segments are mostly halo and a launch is pure latency, so here one CTA takes a strip of SW
// columns over the FULL height: no vertical halo is recomputed and both passes are spread over all the CTA's
// threads.  Each thread produces 4 adjacent outputs (row pass: 4 columns of a row, column pass: 4 rows of a
// column) and walks the taps in groups of 4 with a sliding register window, so one shared-memory value feeds
// 4 multiply-adds (1 B of shared-memory traffic per tap instead of 4).  The radius is a run-time value: the tap
// list is padded with zeros to a multiple of 4 at both ends, which leaves every partial sum unchanged
// (acc + 0*v == acc for the finite grey values of an image; the padded window only touches real pixels).
constexpr int kStripMaxThreads = 512;

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

template <bool FMA>
__device__ __forceinline__ void window4(float& o0, float& o1, float& o2, float& o3, const float4 t, const float a0, const float a1,
                                        const float a2, const float a3, const float b0, const float b1, const float b2) {
    o0 = tap_acc<FMA>(o0, t.x, a0); o0 = tap_acc<FMA>(o0, t.y, a1); o0 = tap_acc<FMA>(o0, t.z, a2); o0 = tap_acc<FMA>(o0, t.w, a3);
    o1 = tap_acc<FMA>(o1, t.x, a1); o1 = tap_acc<FMA>(o1, t.y, a2); o1 = tap_acc<FMA>(o1, t.z, a3); o1 = tap_acc<FMA>(o1, t.w, b0);
    o2 = tap_acc<FMA>(o2, t.x, a2); o2 = tap_acc<FMA>(o2, t.y, a3); o2 = tap_acc<FMA>(o2, t.z, b0); o2 = tap_acc<FMA>(o2, t.w, b1);
    o3 = tap_acc<FMA>(o3, t.x, a3); o3 = tap_acc<FMA>(o3, t.y, b0); o3 = tap_acc<FMA>(o3, t.z, b1); o3 = tap_acc<FMA>(o3, t.w, b2);
}

// Shared-memory layout (rp = r rounded up to 4, ntp = padded tap count):
//   s_taps[ntp]            padded taps; s_taps[rp - r + j] multiplies source index x - r + j
//   s_in  [h][iw]          column c holds source column x0 - rp + c (reflected), c < SW + ntp
//   s_mid [h4 + ntp][SW]   row m holds the row-filtered source row m - rp (reflected above/below, zero further out)
template <int SW, bool FMA>
__global__ void __launch_bounds__(kStripMaxThreads) blur_strip_kernel(BlurArgs a, int iw, int ntp) {
    extern __shared__ __align__(16) float smem[];
    const int r = a.r, w = a.w, h = a.h;
    const int rp = (r + 3) & ~3, off = rp - r, h4 = (h + 3) & ~3;
    float* s_taps = smem;
    float* s_in = s_taps + ntp;
    float* s_mid = s_in + h * iw;

    const int T = blockDim.x, tid = threadIdx.x;
    const int b = blockIdx.z + a.z0;
    const int x0 = blockIdx.x * SW;
    const float* src = a.src + (size_t)b * a.src_stride;

    // every element is fetched with its own 4-byte cp.async (the source index is reflected per element), all of
    // them in flight at once: the whole staging phase costs one round trip to L2 instead of one per loop trip
    {
        const int warp = tid >> 5, lane = tid & 31, nwarp = T >> 5, span = SW + ntp;
        for (int tx = lane; tx < span; tx += 32) {
            const float* colp = src + reflect101(x0 - rp + tx, w);
            for (int y = warp; y < h; y += nwarp) cp_async4(&s_in[y * iw + tx], colp + (size_t)y * a.src_pitch);
        }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    for (int i = tid; i < ntp; i += T) s_taps[i] = (i >= off && i <= off + 2 * r) ? a.taps[2 * r - (i - off)] : 0.0f;
    // rows of s_mid the row pass does not write: reflected rows are copied after the row pass, the rest is zero
    for (int i = tid; i < (h4 + ntp) * SW; i += T) {
        const int m = i / SW;
        if (m < rp || m >= rp + h) s_mid[i] = 0.0f;
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncthreads();

    constexpr int CGS = SW / 4;
    for (int id = tid; id < h * CGS; id += T) {
        const int row = id / CGS, cg = id % CGS;
        const float* p = s_in + row * iw + 4 * cg;
        float4 A = *reinterpret_cast<const float4*>(p);
        float o0 = 0.0f, o1 = 0.0f, o2 = 0.0f, o3 = 0.0f;
        for (int g = 0; g < ntp; g += 4) {
            const float4 B = *reinterpret_cast<const float4*>(p + g + 4);
            const float4 t = *reinterpret_cast<const float4*>(s_taps + g);
            window4<FMA>(o0, o1, o2, o3, t, A.x, A.y, A.z, A.w, B.x, B.y, B.z);
            A = B;
        }
        *reinterpret_cast<float4*>(s_mid + (row + rp) * SW + 4 * cg) = make_float4(o0, o1, o2, o3);
    }
    __syncthreads();
    for (int i = tid; i < 2 * r * SW; i += T) {
        const int q = i / SW, c = i % SW;
        const int yy = q < r ? q - r : h + (q - r);  // source rows -r .. -1 and h .. h + r - 1
        s_mid[(yy + rp) * SW + c] = s_mid[(reflect101(yy, h) + rp) * SW + c];
    }
    __syncthreads();

    for (int id = tid; id < (h4 / 4) * SW; id += T) {
        const int rg = id / SW, col = id % SW;
        const float* p = s_mid + (4 * rg) * SW + col;  // output row y reads s_mid rows y .. y + ntp - 1
        float a0 = p[0], a1 = p[SW], a2 = p[2 * SW], a3 = p[3 * SW];
        float o0 = 0.0f, o1 = 0.0f, o2 = 0.0f, o3 = 0.0f;
        for (int g = 0; g < ntp; g += 4) {
            const float* q = p + (g + 4) * SW;
            const float b0 = q[0], b1 = q[SW], b2 = q[2 * SW], b3 = q[3 * SW];
            const float4 t = *reinterpret_cast<const float4*>(s_taps + g);
            window4<FMA>(o0, o1, o2, o3, t, a0, a1, a2, a3, b0, b1, b2);
            a0 = b0; a1 = b1; a2 = b2; a3 = b3;
        }
        const int gx = x0 + col, y = 4 * rg;
        if (gx < w) {
            const float* centre = s_in + y * iw + rp + col;
            store_out(a, b, gx, y, o0, centre[0]);
            if (y + 1 < h) store_out(a, b, gx, y + 1, o1, centre[iw]);
            if (y + 2 < h) store_out(a, b, gx, y + 2, o2, centre[2 * iw]);
            if (y + 3 < h) store_out(a, b, gx, y + 3, o3, centre[3 * iw]);
        }
    }
}

static int strip_ntp(int r) { return (((r + 3) & ~3) - r + 2 * r + 1 + 3) & ~3; }

static int strip_iw(int sw, int r) {
    int iw = sw + strip_ntp(r);                 // multiple of 4: rows stay 16-B aligned
    if (sw == 16) while (iw % 32 != 16) iw += 4;  // two rows per quarter-warp: keep them on disjoint banks
    return iw;
}

static size_t strip_smem_bytes(int sw, int r, int h) {
    const int h4 = (h + 3) & ~3, ntp = strip_ntp(r);
    return sizeof(float) * (size_t)(ntp + h * strip_iw(sw, r) + (h4 + ntp) * sw);
}

static int strip_max_height() {
    static const int v = [] { const char* e = getenv("SIFT_GPU_STRIP_MAX_H"); return e ? atoi(e) : 300; }();
    return v;
}
// Levels between this height and strip_max_height() go to the streaming kernel when it has the radius (measured faster
// from 270 rows up) and to the strip kernel otherwise.
static int strip_preferred_height() {
    static const int v = [] { const char* e = getenv("SIFT_GPU_STRIP_PREF_H"); return e ? atoi(e) : 200; }();
    return v;
}

template <int SW>
static int launch_strip_sw(const BlurArgs& a, int batch, bool fma, size_t smem, cudaStream_t s) {
    dim3 grid((a.w + SW - 1) / SW, 1, batch);
    const int iw = strip_iw(SW, a.r), ntp = strip_ntp(a.r);
    int threads = (a.h * (SW / 4) + 31) / 32 * 32;
    threads = threads < 64 ? 64 : (threads > kStripMaxThreads ? kStripMaxThreads : threads);
    if (fma) blur_strip_kernel<SW, true><<<grid, threads, smem, s>>>(a, iw, ntp);
    else blur_strip_kernel<SW, false><<<grid, threads, smem, s>>>(a, iw, ntp);
    SIFT_CUDA_TRY(cudaGetLastError());
    return 0;
}

// returns -1 when the level does not qualify
static int launch_strip(const BlurArgs& a, int batch, bool fma, cudaStream_t s, int max_h) {
    if (a.h > max_h || a.r >= a.h || a.r >= a.w) return -1;
    const size_t cap = 200 * 1024;
    int sw = 32;
    if (strip_smem_bytes(32, a.r, a.h) > cap || ((a.w + 31) / 32) * batch < 2 * 148) sw = 16;
    const size_t smem = strip_smem_bytes(sw, a.r, a.h);
    if (smem > cap) return -1;
    return sw == 32 ? launch_strip_sw<32>(a, batch, fma, smem, s) : launch_strip_sw<16>(a, batch, fma, smem, s);
}

// ---------------------------------------------------------------------------------------------
// Streaming kernel.  Every warp is an independent pipeline over its own 64-column strip: no CTA-wide
// barrier anywhere.  Per 8-row chunk the warp (1) waits for its TMA box (one 8-row box for interior
// chunks, eight 1-row boxes where rows are reflected at the image top/bottom), (2) patches reflected
// columns on edge strips, (3) row-filters the 8 x 64 block into its private ring of row-filtered
// lines (lane = 4 columns x 4 rows, float4 windows in registers), (4) column-filters 8 rows of its
// two columns (lane = columns 2*lane, 2*lane+1) from the ring and writes dst / DoG with 8-byte stores.
// The chunk loop is unrolled over the ring period, so every shared-memory address is a constant.
//
// Arithmetic: exact mode multiplies with packed FMUL2 (two taps of one output in the row pass, two
// columns in the column pass) and adds with scalar FADD in the reference's order, so results are
// bit-identical to separate mulss/addss; FMA mode uses FFMA2 (row pass: even/odd tap partial sums).
template <int R>
struct SC {
    static constexpr int CH = 8;        // rows per chunk
    static constexpr int WC = 64;       // columns per warp
    static constexpr int NWARP = 2;     // warps per CTA (adjacent strips); small CTAs pack shared memory better
    static constexpr int SR = 8;        // rows per staging stage (SR = 4 with NS = 3 was measured slower: more per-stage overhead)
    static constexpr int NS = 2;        // staging stages per warp (TMA runs one stage ahead)
    static constexpr int RPAD = (R + 3) & ~3;
    static constexpr int SWW = (WC + 2 * RPAD <= 96) ? 96 : 128;   // staged floats per row (128-B multiple)
    static constexpr int RC = ((R + CH - 1) / CH) * CH;             // rows loaded above the first output row
    static constexpr int LAG = (R + RC + CH - 1) / CH;              // chunks between a row entering and its output leaving
    static constexpr int RING = (LAG + 1) * CH - RC + R <= (LAG + 1) * CH ? (LAG + 1) * CH : (LAG + 2) * CH;  // >= (LAG+1)*CH - RC + R, multiple of CH
    static constexpr int WL = CH + 2 * R;                           // column-pass window rows
    static constexpr int NW = 4 + 2 * RPAD;                         // row-pass window floats
    static constexpr int OFF = RPAD - R;
    static constexpr int WARP_FLOATS = NS * SR * SWW + RING * WC;
    static constexpr size_t SMEM = sizeof(float) * (size_t)(NWARP * WARP_FLOATS) + NWARP * NS * sizeof(uint64_t);
    static_assert(WC + 2 * RPAD <= SWW, "radius too large for the streaming kernel");
    static_assert(RING >= (LAG + 1) * CH - RC + R && RING % CH == 0 && RING >= WL, "ring too small");
    static_assert((SR * SWW * 4) % 128 == 0 && (RING * WC * 4) % 128 == 0 && CH % SR == 0, "TMA destinations must stay 128-B aligned");
};

template <int R>
struct TapsP {           // tk[j] = kernel tap applied to source index x - R + j
    float2 dup[2 * R + 1];  // (tk[j], tk[j])
    float2 pe[R];           // (tk[2m], tk[2m+1])
    float2 po[R];           // (tk[2m+1], tk[2m+2])
    float2 one;             // (1, 1), deliberately a run-time value: see tap_acc2
};

struct StreamArgs {
    BlurArgs a;
    int seg;  // output rows per CTA
};

// one output of the row pass: sum over j of tk[j] * wv[base + j], ascending j
template <int R, bool FMA, int BASE, int NWV>
__device__ __forceinline__ float row_output(const float (&wv)[NWV], const TapsP<R>& tp) {
    if (FMA) {
        float2 acc2 = make_float2(0.0f, 0.0f);
        float lead = 0.0f;
        if (BASE % 2 == 0) {
#pragma unroll
            for (int m = 0; m < R; ++m) acc2 = __ffma2_rn(tp.pe[m], make_float2(wv[BASE + 2 * m], wv[BASE + 2 * m + 1]), acc2);
            lead = tp.dup[2 * R].x * wv[BASE + 2 * R];
        } else {
            lead = tp.dup[0].x * wv[BASE];
#pragma unroll
            for (int m = 0; m < R; ++m) acc2 = __ffma2_rn(tp.po[m], make_float2(wv[BASE + 2 * m + 1], wv[BASE + 2 * m + 2]), acc2);
        }
        return (acc2.x + acc2.y) + lead;
    } else {
        float acc;
        if (BASE % 2 == 0) {
            float2 m0 = __fmul2_rn(tp.pe[0], make_float2(wv[BASE], wv[BASE + 1]));
            acc = __fadd_rn(m0.x, m0.y);
#pragma unroll
            for (int m = 1; m < R; ++m) {
                const float2 pm = __fmul2_rn(tp.pe[m], make_float2(wv[BASE + 2 * m], wv[BASE + 2 * m + 1]));
                acc = __fadd_rn(__fadd_rn(acc, pm.x), pm.y);
            }
            acc = __fadd_rn(acc, __fmul_rn(tp.dup[2 * R].x, wv[BASE + 2 * R]));
        } else {
            acc = __fmul_rn(tp.dup[0].x, wv[BASE]);
#pragma unroll
            for (int m = 0; m < R; ++m) {
                const float2 pm = __fmul2_rn(tp.po[m], make_float2(wv[BASE + 2 * m + 1], wv[BASE + 2 * m + 2]));
                acc = __fadd_rn(__fadd_rn(acc, pm.x), pm.y);
            }
        }
        return acc;
    }
}

template <bool FMA>
__device__ __forceinline__ float2 tap_acc2(float2 acc, float2 t, float2 v, float2 one) {
    if (FMA) return __ffma2_rn(t, v, acc);
    // exact: the product is rounded by the packed multiply; the packed add is written as fma(m, 1, acc), which rounds
    // m + acc exactly like an add.  A plain packed add after a packed multiply would be contracted into one FFMA2 by
    // ptxas (single rounding) even with -fmad=false; `one` comes from the kernel parameters so nothing can fold it away.
    const float2 m = __fmul2_rn(t, v);
    return __ffma2_rn(m, one, acc);
}

// Reflect patch of one staged chunk (edge strips only): staged column c holds x = xs - rpad + c.  TMA zero-filled
// what lies outside the image; columns -1..-r and w..w+r-1 get their mirrored pixels, which are staged in the same row
// (x = -k mirrors to k <= r <= rpad + strip; x = w-1+k mirrors to w-1-k >= xs - rpad because rpad >= r and xs < w).
__device__ __noinline__ void patch_reflected_columns(float* st, int rows, int w, int xs, int rpad, int r, int sww, int lane) {
    const int n = 2 * r;  // per row: r columns left of 0, r columns right of w-1
    for (int e = lane; e < rows * n; e += 32) {
        const int rr = e / n, k = e - rr * n;
        const int gx = k < r ? -1 - k : w + (k - r);
        const int c = gx - (xs - rpad);
        const int cs = reflect101(gx, w) - (xs - rpad);
        if (c >= 0 && c < sww && cs >= 0 && cs < sww) st[rr * sww + c] = st[rr * sww + cs];
    }
    tma::fence_proxy_async();  // generic-proxy writes above vs. the next TMA write into this stage
    __syncwarp();
}

// lane 0: fetch the rows v0..v0+rows-1 of columns [x, x + box) of image b into dst (one box when no row is reflected)
__device__ __noinline__ void issue_chunk_tma(float* dst, uint64_t* bar, const CUtensorMap* mapn, const CUtensorMap* map1, int x, int v0,
                                             int rows, int h, int b, int sww) {
    tma::mbar_arrive_expect_tx(bar, rows * sww * (int)sizeof(float));
    if (v0 >= 0 && v0 + rows <= h) {
        tma::load_3d(dst, mapn, bar, x, v0, b);
    } else {
#pragma unroll 1
        for (int rr = 0; rr < rows; ++rr) tma::load_3d(dst + rr * sww, map1, bar, x, reflect101(v0 + rr, h), b);
    }
}

// Column-pass window: rows ub .. ub+WL-1 (mod RING) of this lane's two columns.  ub only takes RING/CH values
// (it advances by CH per chunk), so each value gets its own fully static load sequence.
template <int R, int M>
__device__ __forceinline__ void load_window_phase(const float* ring_col, float2 (&win)[SC<R>::WL]) {
    using C = SC<R>;
    constexpr int UB = (C::RC - R + C::CH * M) % C::RING;
#pragma unroll
    for (int k = 0; k < C::WL; ++k) win[k] = *reinterpret_cast<const float2*>(ring_col + ((UB + k) % C::RING) * C::WC);
}
template <int R>
__device__ __forceinline__ void load_window(int phase, const float* ring_col, float2 (&win)[SC<R>::WL]) {
    constexpr int NP = SC<R>::RING / SC<R>::CH;
    static_assert(NP <= 8, "add cases");
    switch (phase) {
        case 0: load_window_phase<R, 0>(ring_col, win); break;
        case 1: if (NP > 1) load_window_phase<R, 1 % NP>(ring_col, win); break;
        case 2: if (NP > 2) load_window_phase<R, 2 % NP>(ring_col, win); break;
        case 3: if (NP > 3) load_window_phase<R, 3 % NP>(ring_col, win); break;
        case 4: if (NP > 4) load_window_phase<R, 4 % NP>(ring_col, win); break;
        case 5: if (NP > 5) load_window_phase<R, 5 % NP>(ring_col, win); break;
        case 6: if (NP > 6) load_window_phase<R, 6 % NP>(ring_col, win); break;
        default: if (NP > 7) load_window_phase<R, 7 % NP>(ring_col, win); break;
    }
}

template <int R, bool FMA, bool DECIMATE>
__global__ void __launch_bounds__(64) blur_stream_kernel(const __grid_constant__ CUtensorMap map8, const __grid_constant__ CUtensorMap map1,
                                                         const StreamArgs sa, const TapsP<R> taps) {
    using C = SC<R>;
    static_assert(C::CH == 8, "helpers assume 8-row chunks");
    extern __shared__ __align__(1024) unsigned char smem_raw[];     // TMA destinations must be 128-B aligned
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* stage = reinterpret_cast<float*>(smem_raw) + warp * C::WARP_FLOATS;   // [NS][CH][SWW]
    float* ring = stage + C::NS * C::SR * C::SWW;                                // [RING][WC]
    uint64_t* full = reinterpret_cast<uint64_t*>(reinterpret_cast<float*>(smem_raw) + C::NWARP * C::WARP_FLOATS) + warp * C::NS;

    const BlurArgs& a = sa.a;
    const int b = blockIdx.z + a.z0;
    const int w = a.w, h = a.h;
    const int xs = (blockIdx.x * C::NWARP + warp) * C::WC;   // first column of this warp's strip
    if (xs >= w) return;                                     // warps are independent: no barrier below
    const int y0 = blockIdx.y * sa.seg;
    const int y1 = min(y0 + sa.seg, h);
    const int n_out_chunks = (y1 - y0 + C::CH - 1) / C::CH;
    const int n_in = n_out_chunks + C::LAG;
    const bool edge = (xs - C::RPAD < 0) || (xs - C::RPAD + C::SWW > w);
    const float* src = a.src + (size_t)b * a.src_stride;

    if (lane == 0) {
        tma::prefetch_map(&map8);
        for (int s = 0; s < C::NS; ++s) tma::mbar_init(&full[s], 1);
        tma::fence_barrier_init();
    }
    __syncwarp();
    constexpr int HPC = C::CH / C::SR;       // staging stages (half-chunks) per chunk
    const int n_half = n_in * HPC;
    if (lane == 0) {
        for (int i = 0; i < C::NS && i < n_half; ++i)
            issue_chunk_tma(stage + i * C::SR * C::SWW, &full[i], &map8, &map1, xs - C::RPAD, y0 - C::RC + i * C::SR, C::SR, h, b, C::SWW);
    }

    // row pass: lane = (row parity, column group): columns 4*cg..4*cg+3 of rows rsub, rsub+2, rsub+4, rsub+6
    const int cg = lane & 15, rsub = lane >> 4;
    // column pass: columns xs + 2*lane, +1 (for an odd width the second column lies in the row's padding: pitch % 32 == 0)
    const int x = xs + 2 * lane;
    const bool active = x < w;
    int sel0 = -1, sel1 = -1;
    if (DECIMATE && active) {
        sel0 = a.sel_x[x];
        if (x + 1 < w) sel1 = a.sel_x[x + 1];
    }
    float* const dst_col = a.dst ? a.dst + (size_t)b * a.dst_stride + (DECIMATE ? 0 : x) : nullptr;
    float* const dog_col = a.dog ? a.dog + (size_t)b * a.dog_stride + x : nullptr;
    const float* const src_col = src + (active ? x : 0);
    const float* const ring_col = ring + 2 * lane;

    // running row pointers of the output chunk (advance by 8 rows per chunk; rows inside a chunk add a multiple of the pitch)
    float* dst_row = dst_col ? dst_col + (size_t)y0 * a.dst_pitch : nullptr;
    float* dog_row = dog_col ? dog_col + (size_t)y0 * a.dog_pitch : nullptr;
    const float* low_row = src_col + (size_t)y0 * a.src_pitch;   // centre values of the chunk whose `lower` is fetched next
    const size_t dstp = (size_t)a.dst_pitch, dogp = (size_t)a.dog_pitch, srcp = (size_t)a.src_pitch;
    // DoG centre values of the next output chunk, fetched one chunk ahead so their latency never shows
    float2 lower[C::CH];
    auto fetch_lower = [&](int rows_left) {
        if (rows_left >= C::CH) {
#pragma unroll
            for (int o = 0; o < C::CH; ++o) lower[o] = *reinterpret_cast<const float2*>(low_row + o * srcp);
        } else {
#pragma unroll
            for (int o = 0; o < C::CH; ++o) lower[o] = *reinterpret_cast<const float2*>(low_row + (size_t)min(o, rows_left - 1) * srcp);
        }
        low_row += C::CH * srcp;
    };
    if (!DECIMATE && dog_col && active) fetch_lower(y1 - y0);

    int s = 0;                              // staging stage of chunk i and its mbarrier phase parity
    uint32_t parity = 0;
    int wslot = 0;                          // ring slot of the chunk being row-filtered
    int phase = 0;                          // output chunk index mod RING/CH: selects the static window-load sequence
#pragma unroll 1
    for (int i = 0; i < n_in; ++i) {
#pragma unroll 1
        for (int hh = 0; hh < HPC; ++hh) {
            float* const st = stage + s * C::SR * C::SWW;
            tma::mbar_wait(&full[s], parity);
            if (edge) patch_reflected_columns(st, C::SR, w, xs, C::RPAD, R, C::SWW, lane);
            // ---- row pass of SR staged rows: lane = 4 columns x rows rsub, rsub+2 ----
#pragma unroll
            for (int q = 0; q < C::SR / 2; ++q) {
                const int rr = rsub + 2 * q;
                const float* srow = st + rr * C::SWW + 4 * cg;
                float wv[C::NW];
#pragma unroll
                for (int k = 0; k < C::NW / 4; ++k) {
                    const float4 f = *reinterpret_cast<const float4*>(srow + 4 * k);
                    wv[4 * k] = f.x; wv[4 * k + 1] = f.y; wv[4 * k + 2] = f.z; wv[4 * k + 3] = f.w;
                }
                float4 o;
                o.x = row_output<R, FMA, C::OFF + 0>(wv, taps);
                o.y = row_output<R, FMA, C::OFF + 1>(wv, taps);
                o.z = row_output<R, FMA, C::OFF + 2>(wv, taps);
                o.w = row_output<R, FMA, C::OFF + 3>(wv, taps);
                *reinterpret_cast<float4*>(ring + (wslot + hh * C::SR + rr) * C::WC + 4 * cg) = o;
            }
            __syncwarp();  // ring rows visible to the warp; every lane is done with stage s
            const int hnext = i * HPC + hh + C::NS;
            if (lane == 0 && hnext < n_half) {
                const int v0 = y0 - C::RC + hnext * C::SR;
                if (v0 >= 0 && v0 + C::SR <= h) {
                    tma::mbar_arrive_expect_tx(&full[s], C::SR * C::SWW * (int)sizeof(float));
                    tma::load_3d(st, &map8, &full[s], xs - C::RPAD, v0, b);
                } else {
                    issue_chunk_tma(st, &full[s], &map8, &map1, xs - C::RPAD, v0, C::SR, h, b, C::SWW);
                }
            }
            if (++s == C::NS) { s = 0; parity ^= 1u; }
        }
        wslot = wslot + C::CH == C::RING ? 0 : wslot + C::CH;
        // ---- column pass for output chunk j = i - LAG ----
        const int j = i - C::LAG;
        if (j >= 0) {
            if (active) {
                const int yb = y0 + j * C::CH;
                const int nrows = y1 - yb;  // >= 1; a full chunk unless this is the segment's last one
                float2 win[C::WL];
                load_window<R>(phase, ring_col, win);
                if (!DECIMATE) {
                    float2 acc[C::CH];
#pragma unroll
                    for (int jj = 0; jj <= 2 * R; ++jj) {
#pragma unroll
                        for (int o = 0; o < C::CH; ++o) acc[o] = jj == 0 ? __fmul2_rn(taps.dup[0], win[o]) : tap_acc2<FMA>(acc[o], taps.dup[jj], win[o + jj], taps.one);
                    }
                    if (nrows >= C::CH) {  // full chunk: no per-row predicates
                        if (dst_col) {
#pragma unroll
                            for (int o = 0; o < C::CH; ++o) *reinterpret_cast<float2*>(dst_row + o * dstp) = acc[o];
                        }
                        if (dog_col) {
#pragma unroll
                            for (int o = 0; o < C::CH; ++o)
                                *reinterpret_cast<float2*>(dog_row + o * dogp) =
                                    make_float2(__fadd_rn(128.0f, __fsub_rn(acc[o].x, lower[o].x)), __fadd_rn(128.0f, __fsub_rn(acc[o].y, lower[o].y)));
                        }
                    } else {
#pragma unroll
                        for (int o = 0; o < C::CH; ++o)
                            if (o < nrows) {
                                if (dst_col) *reinterpret_cast<float2*>(dst_row + o * dstp) = acc[o];
                                if (dog_col)
                                    *reinterpret_cast<float2*>(dog_row + o * dogp) =
                                        make_float2(__fadd_rn(128.0f, __fsub_rn(acc[o].x, lower[o].x)), __fadd_rn(128.0f, __fsub_rn(acc[o].y, lower[o].y)));
                            }
                    }
                    if (dst_col) dst_row += C::CH * dstp;
                    if (dog_col) {
                        dog_row += C::CH * dogp;
                        if (j + 1 < n_out_chunks) fetch_lower(nrows - C::CH);
                    }
                } else {
#pragma unroll
                    for (int o = 0; o < C::CH; ++o) {
                        const int sy = o < nrows ? a.sel_y[yb + o] : -1;  // uniform across the warp
                        if (sy >= 0) {
                            float2 acc = __fmul2_rn(taps.dup[0], win[o]);
#pragma unroll
                            for (int jj = 1; jj <= 2 * R; ++jj) acc = tap_acc2<FMA>(acc, taps.dup[jj], win[o + jj], taps.one);
                            if (sel0 >= 0) dst_col[sy * a.dst_pitch + sel0] = acc.x;
                            if (sel1 >= 0) dst_col[sy * a.dst_pitch + sel1] = acc.y;
                        }
                    }
                }
            }
            phase = phase + 1 == C::RING / C::CH ? 0 : phase + 1;
        }
    }
}

// Per-device launch data of the streaming kernels (function attributes and occupancy are per device: a process may hold
// contexts on several GPUs).  Filled by blur_prepare_device(), which sift_gpu_create calls after cudaSetDevice.
struct StreamDevInfo { bool ready = false; int n_sm = 148; int cps[6][2][2] = {}; };   // [radius index][fma][decimate]
static StreamDevInfo g_stream_dev[64];
static int stream_radius_index(int r) { return r == 3 ? 0 : r == 5 ? 1 : r == 7 ? 2 : r == 10 ? 3 : r == 14 ? 4 : 5; }
static const StreamDevInfo& stream_dev_info() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!g_stream_dev[dev].ready) blur_prepare_device();
    return g_stream_dev[dev];
}

template <int R, bool FMA, bool DEC>
static int stream_prepare_one(int* cps) {
    using C = SC<R>;
    SIFT_CUDA_TRY(cudaFuncSetAttribute(blur_stream_kernel<R, FMA, DEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    int n = 0;
    SIFT_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, blur_stream_kernel<R, FMA, DEC>, C::NWARP * 32, C::SMEM));
    *cps = n > 0 ? n : 1;
    return 0;
}
template <int R>
static int stream_prepare_r(StreamDevInfo& d) {
    const int i = stream_radius_index(R);
    int rc;
    if ((rc = stream_prepare_one<R, false, false>(&d.cps[i][0][0])) || (rc = stream_prepare_one<R, false, true>(&d.cps[i][0][1])) ||
        (rc = stream_prepare_one<R, true, false>(&d.cps[i][1][0])) || (rc = stream_prepare_one<R, true, true>(&d.cps[i][1][1])))
        return rc;
    return 0;
}

int blur_prepare_device() {
    int dev = 0;
    SIFT_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) dev = 0;
    StreamDevInfo& d = g_stream_dev[dev];
    if (d.ready) return 0;
    if (cudaDeviceGetAttribute(&d.n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || d.n_sm <= 0) d.n_sm = 148;
    int rc;
    if ((rc = stream_prepare_r<3>(d)) || (rc = stream_prepare_r<5>(d)) || (rc = stream_prepare_r<7>(d)) || (rc = stream_prepare_r<10>(d)) ||
        (rc = stream_prepare_r<14>(d)) || (rc = stream_prepare_r<19>(d)))
        return rc;
    SIFT_CUDA_TRY(cudaFuncSetAttribute(blur_strip_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SIFT_CUDA_TRY(cudaFuncSetAttribute(blur_strip_kernel<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SIFT_CUDA_TRY(cudaFuncSetAttribute(blur_strip_kernel<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SIFT_CUDA_TRY(cudaFuncSetAttribute(blur_strip_kernel<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SIFT_CUDA_TRY(cudaFuncSetAttribute(blur_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SIFT_CUDA_TRY(cudaFuncSetAttribute(blur_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if ((rc = slide_prepare_device())) return rc;
    d.ready = true;
    return 0;
}

template <int R>
static int launch_stream_r(const BlurArgs& a, int batch, bool fma, cudaStream_t s) {
    using C = SC<R>;
    StreamArgs sa;
    sa.a = a;
    // Rows per CTA.  More segments = more parallelism but RC + R extra rows to stage and row-filter per segment; fewer
    // segments = less overhead but a ragged last wave.  Pick the count with the best (last-wave fill) / (halo overhead).
    const int strips = (a.w + C::NWARP * C::WC - 1) / (C::NWARP * C::WC);
    const StreamDevInfo& dv = stream_dev_info();
    const int cps = dv.cps[stream_radius_index(R)][fma ? 1 : 0][a.sel_x ? 1 : 0];
    const int n_sm = dv.n_sm;
    const double slots = (double)n_sm * cps / (a.share > 1 ? a.share : 1);
    int seg = a.h;
    double best = -1.0;
    for (int nseg = 1; nseg <= (a.h + 15) / 16; ++nseg) {
        int sg = ((a.h + nseg - 1) / nseg + C::CH - 1) / C::CH * C::CH;
        const int real_segs = (a.h + sg - 1) / sg;
        const double tiles = (double)strips * real_segs * batch;
        const double waves = std::ceil(tiles / slots);
        const double fill = tiles / (waves * slots);
        const double overhead = (double)(sg + C::RC + R + C::LAG * C::CH * 0.25) / sg;
        const double score = fill / overhead;
        if (score > best + 1e-9) { best = score; seg = sg; }
    }
    sa.seg = seg;
    dim3 grid(strips, (a.h + seg - 1) / seg, batch);
    TapsP<R> tp;
    auto tk = [&](int j) { return a.taps_host[2 * R - j]; };  // kernel walked from +r down while the source index ascends
    tp.one = make_float2(1.0f, 1.0f);
    for (int j = 0; j < 2 * R + 1; ++j) tp.dup[j] = make_float2(tk(j), tk(j));
    for (int m = 0; m < R; ++m) {
        tp.pe[m] = make_float2(tk(2 * m), tk(2 * m + 1));
        tp.po[m] = make_float2(tk(2 * m + 1), tk(2 * m + 2));
    }
    const bool dec = a.sel_x != nullptr;
#define LAUNCH(F, D) blur_stream_kernel<R, F, D><<<grid, C::NWARP * 32, C::SMEM, s>>>(a.map[0], a.map[1], sa, tp)
    if (fma) { if (dec) LAUNCH(true, true); else LAUNCH(true, false); }
    else { if (dec) LAUNCH(false, true); else LAUNCH(false, false); }
#undef LAUNCH
    SIFT_CUDA_TRY(cudaGetLastError());
    return 0;
}

int stream_box_rows() { return SC<5>::SR; }

static int stream_old_box_width(int r) {
    switch (r) {
        case 3: return SC<3>::SWW;
        case 5: return SC<5>::SWW;
        case 7: return SC<7>::SWW;
        case 10: return SC<10>::SWW;
        case 14: return SC<14>::SWW;
        case 19: return SC<19>::SWW;
        default: return 0;
    }
}

int stream_box_width(int r, bool decimate) {
    if (!decimate && slide_box_width(r) > 0) return slide_box_width(r);
    if (decimate && slide_dec_box_width(r) > 0) return slide_dec_box_width(r);
    switch (r) {
        case 3: return SC<3>::SWW;
        case 5: return SC<5>::SWW;
        case 7: return SC<7>::SWW;
        case 10: return SC<10>::SWW;
        case 14: return SC<14>::SWW;
        case 19: return SC<19>::SWW;
        default: return 0;
    }
}

int launch_blur(const BlurArgs& a, int batch, bool fma, cudaStream_t s, uint64_t* launches) {
    if (launches) ++*launches;
    {
        const int rc = launch_strip(a, batch, fma, s, strip_preferred_height());
        if (rc >= 0) return rc;
    }
    if (a.map && a.map_box == stream_box_width(a.r, a.sel_x != nullptr)) {
        const int rc = a.sel_x ? launch_slide_dec(a, batch, fma, s) : launch_slide(a, batch, fma, s);
        if (rc >= 0) return rc;
    }
    {
        const int rc = launch_strip(a, batch, fma, s, strip_max_height());
        if (rc >= 0) return rc;
    }
    if (a.map && a.taps_host && a.w >= 32 && a.h >= 16 && a.map_box == stream_old_box_width(a.r)) {
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
    (void)stream_dev_info();   // makes sure this device's function attributes are set
    if (fma) blur_tile_kernel<true><<<grid, kBlurThreads, smem, s>>>(a);
    else blur_tile_kernel<false><<<grid, kBlurThreads, smem, s>>>(a);
    SIFT_CUDA_TRY(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// resizeImageNoInterpolation: dst(x, y) = src(map_x[x], map_y[y]); the maps are produced on the host
// by the literal accumulated-double walk.  A thread writes four adjacent destination pixels as one 16-byte store (row
// pitches are multiples of 32 floats); its four gathered sources are neighbours in one row, i.e. L1 hits.
__global__ void __launch_bounds__(256) resize_nn_kernel(const float* __restrict__ src, size_t src_stride, int src_pitch, float* __restrict__ dst,
                                                        size_t dst_stride, int dst_pitch, int dw, int dh, const int* __restrict__ map_x,
                                                        const int* __restrict__ map_y, int z0) {
    const int x = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (x >= dw) return;
    const int b = blockIdx.z + z0;
    int mx[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) mx[j] = map_x[x + j < dw ? x + j : dw - 1];
    for (int y = blockIdx.y; y < dh; y += gridDim.y) {
        const float* srow = src + (size_t)b * src_stride + (size_t)map_y[y] * src_pitch;
        float4 v;
        v.x = srow[mx[0]]; v.y = srow[mx[1]]; v.z = srow[mx[2]]; v.w = srow[mx[3]];
        float* d = dst + (size_t)b * dst_stride + (size_t)y * dst_pitch + x;
        if (x + 3 < dw) *reinterpret_cast<float4*>(d) = v;
        else {
            d[0] = v.x;
            if (x + 1 < dw) d[1] = v.y;
            if (x + 2 < dw) d[2] = v.z;
        }
    }
}

int launch_resize_nn(const float* src, size_t src_stride, int src_pitch, float* dst, size_t dst_stride, int dst_pitch, int dw,
                     int dh, const int* map_x, const int* map_y, int z0, int batch, cudaStream_t s, uint64_t* launches) {
    dim3 grid((dw + 1023) / 1024, dh < 1024 ? dh : 1024, batch);
    resize_nn_kernel<<<grid, 256, 0, s>>>(src, src_stride, src_pitch, dst, dst_stride, dst_pitch, dw, dh, map_x, map_y, z0);
    if (launches) ++*launches;
    SIFT_CUDA_TRY(cudaGetLastError());
    return 0;
}

// importImage's widening of 8-bit pixels (main.cpp:52-54): four pixels per thread (one 4-byte load, one 16-byte store) where the
// row start is aligned, single pixels otherwise.
__global__ void __launch_bounds__(256) u8_to_f32_kernel(const uint8_t* __restrict__ src, size_t src_stride, int src_pitch, float* __restrict__ dst,
                                                        size_t dst_stride, int dst_pitch, int w, int h) {
    const int x = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
    const int b = blockIdx.z;
    if (x >= w) return;
    const bool vec = (src_pitch & 3) == 0 && (src_stride & 3) == 0 && ((uintptr_t)src & 3) == 0 && x + 3 < w;
    for (int y = blockIdx.y; y < h; y += gridDim.y) {
        const uint8_t* sp = src + (size_t)b * src_stride + (size_t)y * src_pitch + x;
        float* dp = dst + (size_t)b * dst_stride + (size_t)y * dst_pitch + x;
        if (vec) {
            const uchar4 u = *reinterpret_cast<const uchar4*>(sp);
            *reinterpret_cast<float4*>(dp) = make_float4((float)u.x, (float)u.y, (float)u.z, (float)u.w);
        } else {
            for (int j = 0; j < 4 && x + j < w; ++j) dp[j] = (float)sp[j];
        }
    }
}

int launch_u8_to_f32(const uint8_t* src, size_t src_stride, int src_pitch, float* dst, size_t dst_stride, int dst_pitch, int w,
                     int h, int batch, cudaStream_t s, uint64_t* launches) {
    dim3 grid((w + 1023) / 1024, h < 540 ? h : 540, batch);
    u8_to_f32_kernel<<<grid, 256, 0, s>>>(src, src_stride, src_pitch, dst, dst_stride, dst_pitch, w, h);
    if (launches) ++*launches;
    SIFT_CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace siftgpu
