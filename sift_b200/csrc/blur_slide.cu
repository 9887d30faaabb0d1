// Sliding-accumulator Gaussian blur: alg::convolveWithGauss (reference algorithms.cpp:10-22) with the DoG of
// algorithms.cpp:52-64 fused into the epilogue, for the large levels of the pyramid.
//
// Every warp is an independent pipeline over a 64-column strip of a row segment (no CTA-wide barrier):
//   * TMA (cp.async.bulk.tensor.3d, mbarrier expect_tx) stages 8-row boxes of SWW = 96 or 128 columns two chunks
//     ahead into a ring of three stages; rows above/below the image arrive as 1-row boxes at their reflected index,
//     columns left/right of it are zero-filled by TMA and patched by a smem -> smem mirror on edge strips;
//   * row pass: lane = 4 adjacent columns x 4 rows, float4 window loads (the lanes of a quarter-warp read 128 contiguous
//     bytes: conflict-free for any row pitch), results to an 8-row buffer `mid`;
//   * column pass: lane = 2 adjacent columns.  The 2R+1 output rows a source row contributes to live in 2R+1 float2
//     ACCUMULATORS IN REGISTERS: each new row-filtered line h(v) is loaded once (one LDS.64) and scattered into them
//     (acc[yo] += tk[v - yo + R] * h(v)); the accumulator whose last tap this was is complete and leaves through the
//     epilogue (dst, and dog = 128 + (dst - src); the row pass parks the centre pixels src in a small buffer `cen`).  An output row therefore
//     receives its taps in ascending source order exactly like the reference's line loop, there is no ring of
//     row-filtered lines to re-read, no vertical halo is recomputed inside a segment, and the 2R+1 operations of a step
//     are independent of each other.  The sums move down one register per step; the FFMA2 that adds a tap writes its
//     result one place lower than it read it, so the shift is free and every step runs the same code.
//
// Arithmetic: exact mode = packed multiply, then the add as fma(m, 1, acc) with a run-time 1 (ptxas contracts a packed
// multiply feeding a packed add into one FFMA2 even under -fmad=false; see pyramid.cu) -> bit-identical to the
// reference's mulss/addss; FMA mode = one FFMA2 per tap pair.  Radii are instantiated for a few R; a blur of radius
// r runs on the smallest R >= r with zero taps added symmetrically (acc + 0*v == acc for the finite values of an image).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "tma.cuh"

namespace siftgpu {

__host__ __device__ __forceinline__ int sl_reflect(int v, int n) {
    if (v < 0) v = -v;
    if (v >= n) v = 2 * (n - 1) - v;
    v = v < 0 ? 0 : v;            // only reachable for padding rows/columns that no valid output reads
    return v >= n ? n - 1 : v;
}

template <int R, bool DOG>
struct SL {
    static constexpr int WC = 64;                     // output columns per warp
    static constexpr int CH = 8;                      // rows per chunk = rows per TMA stage
    static constexpr int NWARP = 2;                   // warps per CTA (adjacent strips)
    static constexpr int RPAD = (R + 3) & ~3;
    static constexpr int OFF = RPAD - R;
    static constexpr int SWW = (WC + 2 * RPAD <= 96) ? 96 : 128;   // staged floats per row: a multiple of 128 B, so that the 1-row boxes
                                                                   // of reflected rows land on 128-B aligned shared-memory addresses
    static constexpr int NACC = 2 * R;                // partial sums carried from step to step (the 2R+1-th is born and the oldest leaves in a step)
    static constexpr int NB = DOG ? (R + CH - 1) / CH + 1 : 0;   // 8-row blocks of centre pixels kept for the DoG: rows v - R .. end of the chunk
    static constexpr int MAX_AHEAD = 4;               // chunks in flight: a launch parameter (SlideArgs::ahead), stages = ahead + 1
    static constexpr int MW = 64;                     // row pitch of `mid` and `cen`
    static constexpr int STAGE_FLOATS = CH * SWW;
    static constexpr int MID_FLOATS = CH * MW;
    static constexpr int CEN_FLOATS = NB * CH * MW;
    static constexpr int NW = 4 + 2 * RPAD;           // row-pass window of one group of 4 outputs
    static constexpr int warp_floats(int nstg) { return nstg * STAGE_FLOATS + MID_FLOATS + CEN_FLOATS; }
    static constexpr size_t smem(int nstg) { return sizeof(float) * (size_t)(NWARP * warp_floats(nstg)) + NWARP * (MAX_AHEAD + 1) * sizeof(uint64_t); }
    static_assert(WC + 2 * RPAD <= SWW && (SWW * 4) % 128 == 0 && (STAGE_FLOATS * 4) % 128 == 0 && ((MID_FLOATS + CEN_FLOATS) * 4) % 128 == 0,
                  "TMA destinations must stay 128-B aligned");
};

// Resident CTAs per SM the kernels are compiled for (__launch_bounds__): what shared memory allows with two stages, capped where
// the register budget that goes with it (65536 / 64 / n) would spill the 2R accumulators and the row-pass windows.
template <int R, bool FMA, int MODE>
constexpr int slide_min_blocks() {
    const int by_smem = (int)((227 * 1024) / (SL<R, MODE != 0>::smem(2) + 1024));
    const int cap = R <= 10 ? 12 : (R <= 14 ? 7 : (R <= 19 ? (FMA ? 5 : 4) : 3));
    return by_smem < cap ? by_smem : cap;
}

// Resident CTAs per SM each kernel is compiled for (__launch_bounds__ minimum): 1 = leave the register count to ptxas.  The
// values were picked from in-situ launch times (tools/build_variant.sh builds the alternatives).
#ifndef SL_MINB_R5
#define SL_MINB_R5 1
#endif
#ifndef SL_MINB_R7
#define SL_MINB_R7 1
#endif
#ifndef SL_MINB_R10
#define SL_MINB_R10 1
#endif
#ifndef SL_MINB_R14
#define SL_MINB_R14 1
#endif
template <int R>
constexpr int slide_min_blocks() { return R == 5 ? SL_MINB_R5 : R == 7 ? SL_MINB_R7 : R == 10 ? SL_MINB_R10 : R == 14 ? SL_MINB_R14 : 1; }

template <int R>
struct SlideTaps {        // tk[j] = tap applied to source index x - R + j; symmetric (tk[j] == tk[2R - j], checked on the host)
    float tk[R + 1];      // j <= R; the column pass uses them as (tk, tk): ptxas folds that into a scalar-broadcast FFMA2 operand
    float one;            // 1 as a run-time value (see header)
    float2 pe[R];         // (tk[2m], tk[2m+1])
    float2 po[R];         // (tk[2m+1], tk[2m+2])
};

// Work partition of a launch: the rows of all images of the launch form one linear space g = image * h + y, cut into equal
// contiguous shares, one per "row lane"; a row lane is `strips` warps with consecutive global warp numbers, one per
// 64-column strip, which therefore walk down the same rows side by side (whole rows stream from DRAM, and the horizontal
// halo a strip shares with its neighbours is in L2 when the neighbour asks for it).  The grid is what the GPU holds at once
// (SMs x resident CTAs): a grid of independent CTAs a little smaller than the resident slots gets spread unevenly by the
// hardware (4 to 7 CTAs per SM measured).  A share that crosses an image boundary is cut into two pieces; every piece costs
// one vertical halo (2R rows).
// (Measured and dropped: dynamic row ranges with work stealing between the warps of a strip.  The warp schedulers favour the
// warps launched first, so equal shares finish between 150 and 280 us in a 296 us launch, but the launch is bound by what
// the SM sustains with all warps resident, not by that tail: stealing evened the finish times out and left the span where it
// was, at the price of registers.)
struct SlideArgs {
    BlurArgs a;
    int strips;           // 64-column warp strips per image
    int lanes;            // row lanes: warps gw >= lanes * strips have nothing to do
    uint32_t share;       // rows of the linear space per row lane
    uint32_t total;       // images * h
    int ahead;            // chunks in flight per warp (stages = ahead + 1)
    unsigned long long* probe;   // development (SIFT_GPU_SLIDE_PROBE): per warp {start ns, end ns, smid, hardware warp id}, or null
};

// one output of the row pass: sum over j of tk[j] * wv[BASE + j], ascending j
template <int R, bool FMA, int BASE, int NWV>
__device__ __forceinline__ float sl_row_output(const float (&wv)[NWV], const SlideTaps<R>& tp) {
    const float tk0 = tp.tk[0];
    if (FMA) {
        float2 acc2;
        if (BASE % 2 == 0) {
            acc2 = __fmul2_rn(tp.pe[0], make_float2(wv[BASE], wv[BASE + 1]));
#pragma unroll
            for (int m = 1; m < R; ++m) acc2 = __ffma2_rn(tp.pe[m], make_float2(wv[BASE + 2 * m], wv[BASE + 2 * m + 1]), acc2);
            acc2.x = fmaf(tk0, wv[BASE + 2 * R], acc2.x);   // tk[2R] == tk[0]
        } else {
            acc2 = __fmul2_rn(tp.po[0], make_float2(wv[BASE + 1], wv[BASE + 2]));
#pragma unroll
            for (int m = 1; m < R; ++m) acc2 = __ffma2_rn(tp.po[m], make_float2(wv[BASE + 2 * m + 1], wv[BASE + 2 * m + 2]), acc2);
            acc2.x = fmaf(tk0, wv[BASE], acc2.x);
        }
        return acc2.x + acc2.y;
    } else {
        float acc;
        if (BASE % 2 == 0) {
            const float2 m0 = __fmul2_rn(tp.pe[0], make_float2(wv[BASE], wv[BASE + 1]));
            acc = __fadd_rn(m0.x, m0.y);
#pragma unroll
            for (int m = 1; m < R; ++m) {
                const float2 pm = __fmul2_rn(tp.pe[m], make_float2(wv[BASE + 2 * m], wv[BASE + 2 * m + 1]));
                acc = __fadd_rn(__fadd_rn(acc, pm.x), pm.y);
            }
            acc = __fadd_rn(acc, __fmul_rn(tk0, wv[BASE + 2 * R]));
        } else {
            acc = __fmul_rn(tk0, wv[BASE]);
#pragma unroll
            for (int m = 0; m < R; ++m) {
                const float2 pm = __fmul2_rn(tp.po[m], make_float2(wv[BASE + 2 * m + 1], wv[BASE + 2 * m + 2]));
                acc = __fadd_rn(__fadd_rn(acc, pm.x), pm.y);
            }
        }
        return acc;
    }
}

// Reflect patch of a staged chunk (edge strips only): staged column c holds x = x_first + c; TMA zero-filled what lies outside
// the image; columns -1..-r and w..w+r-1 get their mirrored pixels, which are staged in the same row.  A strip's columns are
// fixed, so lane k < r works out once which staged column it patches on the left (x = -1-k) and on the right (x = w+k) and
// where its source sits; per chunk that leaves eight loads and stores per side.  (The first version recomputed the indices
// with a division per element: an edge warp then took 1.6x as long per chunk as an interior one, and the whole launch
// waited for the two edge strips.)
struct SlPatch { uint32_t left, right; };   // staged columns (destination | source << 16) on the left and on the right, or ~0
__device__ __forceinline__ SlPatch sl_patch_setup(int lane, int w, int x_first, int r, int sww) {
    SlPatch p{~0u, ~0u};
    if (lane < r) {
        int gx = -1 - lane, c = gx - x_first, cs = sl_reflect(gx, w) - x_first;
        if (c >= 0 && c < sww && cs >= 0 && cs < sww) p.left = (uint32_t)c | ((uint32_t)cs << 16);
        gx = w + lane; c = gx - x_first; cs = sl_reflect(gx, w) - x_first;
        if (c >= 0 && c < sww && cs >= 0 && cs < sww) p.right = (uint32_t)c | ((uint32_t)cs << 16);
    }
    return p;
}
template <int ROWS>
__device__ __forceinline__ void sl_patch_apply(float* st, const SlPatch& p, int sww) {
    if (p.left != ~0u) {
        float* d = st + (p.left & 0xffffu);
        const float* q = st + (p.left >> 16);
#pragma unroll
        for (int rr = 0; rr < ROWS; ++rr) d[rr * sww] = q[rr * sww];
    }
    if (p.right != ~0u) {
        float* d = st + (p.right & 0xffffu);
        const float* q = st + (p.right >> 16);
#pragma unroll
        for (int rr = 0; rr < ROWS; ++rr) d[rr * sww] = q[rr * sww];
    }
    tma::fence_proxy_async();  // generic-proxy writes above vs. the next TMA write into this stage
    __syncwarp();
}

// shared-memory addresses as 32-bit values (one conversion per kernel instead of one per use)
__device__ __forceinline__ void sl_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sl_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y, int z) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
                 "l"(map), "r"(bar), "r"(x), "r"(y), "r"(z)
                 : "memory");
}
__device__ __forceinline__ void sl_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "SL_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra SL_DONE;\n"
        "bra SL_WAIT;\n"
        "SL_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
// a chunk with rows above/below the image: eight 1-row boxes at the reflected row indices
__device__ __noinline__ void sl_issue_reflected(uint32_t dst, uint32_t bar, const CUtensorMap* map1, int x, int v0, int h, int b, int sww) {
#pragma unroll 1
    for (int rr = 0; rr < 8; ++rr) sl_load_3d(dst + rr * sww * 4, map1, bar, x, sl_reflect(v0 + rr, h), b);
}
__device__ __forceinline__ void sl_issue_chunk(uint32_t dst, uint32_t bar, const CUtensorMap* map8, const CUtensorMap* map1, int x, int v0, int h,
                                               int b, int sww) {
    sl_expect_tx(bar, 8 * sww * (int)sizeof(float));
    if (v0 >= 0 && v0 + 8 <= h) sl_load_3d(dst, map8, bar, x, v0, b);
    else sl_issue_reflected(dst, bar, map1, x, v0, h, b, sww);
}

__device__ __forceinline__ void sl_store2_if(void* p, float2 v, uint32_t ok) {
    asm volatile(
        "{\n"
        ".reg .pred q;\n"
        "setp.ne.u32 q, %3, 0;\n"
        "@q st.global.v2.f32 [%0], {%1, %2};\n"
        "}\n" ::"l"(p), "f"(v.x), "f"(v.y), "r"(ok)
        : "memory");
}

// One step of the column pass = one new row-filtered line h (this lane's two columns).  acc[k] holds the partial sum of
// output row v - R + k, which has received the taps of source rows up to v - 1.  The line contributes tap 2R - k to acc[k];
// acc[0] is then complete, everything else moves down one place (the FFMA2 writes acc[k] from acc[k + 1], so the shift
// costs nothing) and a new sum is born at the top.  Every output receives its taps in ascending source order.
template <int R, bool FMA>
__device__ __forceinline__ float2 sl_col_step(float2 (&acc)[2 * R], const float2 h, const float (&tk)[R + 1], const float one) {
    auto tap = [&](int j) { return tk[j <= R ? j : 2 * R - j]; };
    auto fma2 = [&](float t, float2 a) {
        if (FMA) return __ffma2_rn(make_float2(t, t), h, a);
        return __ffma2_rn(__fmul2_rn(make_float2(t, t), h), make_float2(one, one), a);
    };
    const float2 res = fma2(tap(2 * R), acc[0]);
#pragma unroll
    for (int k = 0; k + 1 < 2 * R; ++k) acc[k] = fma2(tap(2 * R - 1 - k), acc[k + 1]);
    acc[2 * R - 1] = __fmul2_rn(make_float2(tk[0], tk[0]), h);
    return res;
}

// MODE: 0 = dst only, 1 = dst + DoG, 2 = DoG only.  CHECK: steps outside [2R, n_virtual) must not store (first and last chunks).
template <int R, bool FMA, int MODE, bool CHECK>
__device__ __forceinline__ void sl_col_chunk(float2 (&acc)[2 * R], const float* mid_lane, const float* const (&cb)[SL<R, MODE != 0>::NB + 1],
                                             const float (&tk)[R + 1], const float one, char*& row, const size_t pitch_bytes, const ptrdiff_t dog_delta,
                                             const bool active, const int t0, const int n_virtual) {
    using C = SL<R, MODE != 0>;
#pragma unroll
    for (int r = 0; r < C::CH; ++r) {
        const float2 h = *reinterpret_cast<const float2*>(mid_lane + r * C::MW);
        float2 lo = make_float2(0.0f, 0.0f);
        if (MODE != 0) {   // centre pixels of output row v - R: (R - r) rows back -> block q back, row rr of it
            const int back = R - r;
            const int q = back > 0 ? (back + C::CH - 1) / C::CH : 0;
            const int rr = r - R + C::CH * q;
            lo = *reinterpret_cast<const float2*>(cb[q] + rr * C::MW);
        }
        const float2 res = sl_col_step<R, FMA>(acc, h, tk, one);
        // predicated stores, no branch
        const uint32_t ok = (CHECK ? (active && t0 + r >= 2 * R && t0 + r < n_virtual) : active) ? 1u : 0u;
        if (MODE != 2) sl_store2_if(row, res, ok);
        if (MODE != 0) {
            // higher - lower, then 128 + dif (algorithms.cpp:58-60)
            const float2 dif = __fadd2_rn(res, make_float2(-lo.x, -lo.y));
            sl_store2_if(row + (MODE == 1 ? dog_delta : 0), __fadd2_rn(make_float2(128.0f, 128.0f), dif), ok);
        }
        row += pitch_bytes;
    }
}

template <int R, bool FMA, int MODE>
__global__ void __launch_bounds__(64, slide_min_blocks<R>()) blur_slide_kernel(const __grid_constant__ CUtensorMap map8, const __grid_constant__ CUtensorMap map1,
                                                        const SlideArgs sa, const SlideTaps<R> taps) {
    constexpr bool DOG = MODE != 0;
    using C = SL<R, DOG>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nstg = sa.ahead + 1;
    const int warp_floats = nstg * C::STAGE_FLOATS + C::MID_FLOATS + C::CEN_FLOATS;
    float* stage = reinterpret_cast<float*>(smem_raw) + warp * warp_floats;          // [nstg][CH][SWW]
    float* mid = stage + nstg * C::STAGE_FLOATS;                                     // [CH][MW] row-filtered lines of the current chunk
    float* cen = mid + C::MID_FLOATS;                                               // [NB][CH][MW] centre pixels of the last rows (DoG)
    const uint32_t stage_u = tma::smem_u32(stage);
    const uint32_t full_u = tma::smem_u32(reinterpret_cast<uint64_t*>(reinterpret_cast<float*>(smem_raw) + C::NWARP * warp_floats) + warp * (C::MAX_AHEAD + 1));

    const BlurArgs& a = sa.a;
    const int w = a.w, h = a.h;
    if (lane == 0) {
        tma::prefetch_map(&map8);
        for (int s = 0; s < nstg; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full_u + 8 * s) : "memory");
        tma::fence_barrier_init();
    }
    __syncwarp();

    float tk[R + 1];
#pragma unroll
    for (int j = 0; j <= R; ++j) tk[j] = taps.tk[j];
    const float one = taps.one;

    // row pass: lane = (row parity, column group)
    const int cg = lane & 15, rsub = lane >> 4;
    const float* const mid_lane = mid + 2 * lane;
    const int opitch = MODE == 2 ? a.dog_pitch : a.dst_pitch;
    const size_t pitch_bytes = (size_t)opitch * sizeof(float);

    // this warp's strip and its row lane's share of the linear (image, row) space
    const uint32_t gw = blockIdx.x * C::NWARP + warp;
    const uint32_t row_lane = gw / (uint32_t)sa.strips;
    if (row_lane >= (uint32_t)sa.lanes) return;              // warps are independent: no CTA-wide barrier anywhere
    const int xs = (int)(gw - row_lane * (uint32_t)sa.strips) * C::WC;
    const SlPatch patch = sl_patch_setup(lane, w, xs - C::RPAD, R, C::SWW);
    const bool edge = __any_sync(0xffffffffu, (patch.left & patch.right) != ~0u);
    unsigned long long t_start = 0;
    if (sa.probe) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
    uint32_t g = row_lane * sa.share;
    const uint32_t g_end = min(sa.total, g + sa.share);
    int s = 0;                 // stage of the next chunk, its mbarrier phase parity: the stages are used round-robin across pieces
    uint32_t parity = 0;
    int cur = 0;               // block of `cen` the current chunk's centre pixels go to
    // state of the current piece (one loop over all chunks of all pieces: a nested loop costs ptxas ~40 registers)
    int i = 0, n_chunks = 0, n_virtual = 0, v_first = 0, bz = 0;
    // column pass: columns xs + 2*lane, +1 (for an odd width the second column lies in the row's padding: pitch % 32 == 0)
    const int x = xs + 2 * lane;
    const bool active = x < w;
    char* row = nullptr;
    ptrdiff_t dog_delta = 0;
    float2 acc[C::NACC];
#pragma unroll 1
    for (;;) {
        if (i == n_chunks) {   // next piece
            if (g >= g_end) break;
            const uint32_t b = g / (uint32_t)h;
            const int y0 = (int)(g - b * (uint32_t)h);
            const int y1 = min(h, y0 + (int)(g_end - g));
            g += (uint32_t)(y1 - y0);
            bz = (int)b + a.z0;
            n_virtual = (y1 - y0) + 2 * R;                // virtual rows y0 - R .. y1 + R - 1 feed this piece
            n_chunks = (n_virtual + C::CH - 1) / C::CH;
            v_first = y0 - R;
            i = 0;
            if (lane == 0) {
                int st = s;
                for (int q = 0; q < nstg && q < n_chunks; ++q) {
                    sl_issue_chunk(stage_u + st * C::STAGE_FLOATS * 4, full_u + 8 * st, &map8, &map1, xs - C::RPAD, v_first + q * C::CH, h, bz, C::SWW);
                    if (++st == nstg) st = 0;
                }
            }
            // address of output row (t - 2R) of the piece at the lane's columns: it starts 2R rows above the piece and moves down one
            // row per step; stores only happen once t >= 2R.  MODE 0/1: dst (dog = dst + dog_delta), MODE 2: dog.
            const float* obase = MODE == 2 ? a.dog + (size_t)bz * a.dog_stride : a.dst + (size_t)bz * a.dst_stride;
            row = reinterpret_cast<char*>(const_cast<float*>(obase)) + ((ptrdiff_t)(y0 - 2 * R) * opitch + (active ? x : 0)) * (ptrdiff_t)sizeof(float);
            dog_delta = MODE == 1 ? (reinterpret_cast<const char*>(a.dog + (size_t)bz * a.dog_stride) - reinterpret_cast<const char*>(a.dst + (size_t)bz * a.dst_stride)) : 0;
#pragma unroll
            for (int k = 0; k < C::NACC; ++k) acc[k] = make_float2(0.0f, 0.0f);
        }
        {
            float* const st = stage + s * C::STAGE_FLOATS;
            sl_wait(full_u + 8 * s, parity);
            if (edge) sl_patch_apply<C::CH>(st, patch, C::SWW);

            // blocks of `cen`: cb[q] = the block q chunks back (cb[0] = this chunk's)
            const float* cb[C::NB + 1];
#pragma unroll
            for (int q = 0; q < C::NB; ++q) {
                int blk = cur - q;
                blk = blk < 0 ? blk + C::NB : blk;
                cb[q] = cen + blk * (C::CH * C::MW) + 2 * lane;
            }
            cb[C::NB] = nullptr;

            // ---- row pass: lane = 4 adjacent columns x rows rsub, rsub+2, rsub+4, rsub+6 (a quarter-warp reads 128 contiguous bytes).
            // The window of the next row is loaded before the current one is filtered (two register sets), so the shared-memory
            // latency of one row hides behind the arithmetic of the other.
            {
                float wv[2][C::NW];
                auto load_window = [&](float (&dstw)[C::NW], int rw) {
                    const float* srow = st + rw * C::SWW + 4 * cg;
#pragma unroll
                    for (int k = 0; k < C::NW / 4; ++k) {
                        const float4 f = *reinterpret_cast<const float4*>(srow + 4 * k);
                        dstw[4 * k] = f.x; dstw[4 * k + 1] = f.y; dstw[4 * k + 2] = f.z; dstw[4 * k + 3] = f.w;
                    }
                };
                load_window(wv[0], rsub);
#pragma unroll
                for (int q = 0; q < C::CH / 2; ++q) {
                    const int rw = rsub + 2 * q;
                    if (q + 1 < C::CH / 2) load_window(wv[(q + 1) & 1], rw + 2);
                    const float (&wq)[C::NW] = wv[q & 1];
                    if (DOG)   // the four centre pixels are window elements RPAD .. RPAD+3
                        *reinterpret_cast<float4*>(cen + cur * (C::CH * C::MW) + rw * C::MW + 4 * cg) =
                            make_float4(wq[C::RPAD], wq[C::RPAD + 1], wq[C::RPAD + 2], wq[C::RPAD + 3]);
                    float4 o4;
                    o4.x = sl_row_output<R, FMA, C::OFF + 0>(wq, taps);
                    o4.y = sl_row_output<R, FMA, C::OFF + 1>(wq, taps);
                    o4.z = sl_row_output<R, FMA, C::OFF + 2>(wq, taps);
                    o4.w = sl_row_output<R, FMA, C::OFF + 3>(wq, taps);
                    *reinterpret_cast<float4*>(mid + rw * C::MW + 4 * cg) = o4;
                }
            }
            __syncwarp();   // mid / cen rows visible to the whole warp; every lane is done with stage s
            // the row pass was the stage's only reader (the DoG's centre pixels went to `cen`): refill it with chunk i + nstg right away
            if (lane == 0 && i + nstg < n_chunks)
                sl_issue_chunk(stage_u + s * C::STAGE_FLOATS * 4, full_u + 8 * s, &map8, &map1, xs - C::RPAD, v_first + (i + nstg) * C::CH, h, bz, C::SWW);

            // ---- column pass: 8 steps ----
            const int t0 = i * C::CH;
            if (t0 >= 2 * R && t0 + C::CH <= n_virtual)
                sl_col_chunk<R, FMA, MODE, false>(acc, mid_lane, cb, tk, one, row, pitch_bytes, dog_delta, active, t0, n_virtual);
            else
                sl_col_chunk<R, FMA, MODE, true>(acc, mid_lane, cb, tk, one, row, pitch_bytes, dog_delta, active, t0, n_virtual);
            __syncwarp();   // every lane is done with mid
            if (DOG) cur = cur + 1 == C::NB ? 0 : cur + 1;
            if (++s == nstg) { s = 0; parity ^= 1u; }
            ++i;
        }
    }
    if (sa.probe && lane == 0) {
        unsigned long long t_end;
        unsigned smid, wid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
        unsigned long long* o = sa.probe + 4ull * gw;
        o[0] = t_start; o[1] = t_end; o[2] = smid; o[3] = wid;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Decimating variant: alg::reduceToNextLevel (reference algorithms.cpp:24-36) = blur with the level's own sigma, then the
// nearest-neighbour resize, which picks source column sx(k) = 2k + p and source row sy(m) = 2m + q with p, q in {0, 1}
// constant over long runs (Vigra's accumulated index walk: even sizes switch from 2i to 2i+1 once, at the middle).
// Only the picked pixels are computed along x: a warp owns 64 DESTINATION columns of one parity run; its row pass filters
// the 64 picked source columns of every source row (window of 8 + 2*RPAD staged floats per 4 outputs), the column pass runs
// over those 64 columns exactly like the plain kernel, and a step's result is stored only when its source row is picked.
// Half the row-pass and half the column-pass arithmetic of a full blur, a quarter of the stores.
struct DecRegion { int k_begin, k_end, parity, first_strip; };   // destination columns [k_begin, k_end): source x = 2k + parity; first warp strip of the run
struct SlideDecArgs {
    BlurArgs a;
    int strips;           // warp strips (64 destination columns of one parity run) per image
    int lanes;
    uint32_t share, total;   // as SlideArgs
    int ahead, n_regions;
    DecRegion reg[4];
};

template <int R>
struct SLD {
    using B = SL<R, false>;
    static constexpr int SWW = (128 + 2 * B::RPAD + 31) / 32 * 32;   // staged source floats per row (128-B multiple)
    static constexpr int NW = 8 + 2 * B::RPAD;                       // row-pass window of 4 picked outputs
    static constexpr int STAGE_FLOATS = B::CH * SWW;
    static constexpr int warp_floats(int nstg) { return nstg * STAGE_FLOATS + B::MID_FLOATS; }
    static constexpr size_t smem(int nstg) { return sizeof(float) * (size_t)(B::NWARP * warp_floats(nstg)) + B::NWARP * (B::MAX_AHEAD + 1) * sizeof(uint64_t); }
    static_assert(SWW <= 256 && (STAGE_FLOATS * 4) % 128 == 0 && (B::MID_FLOATS * 4) % 128 == 0, "TMA box / alignment");
};

template <int R, bool FMA, int P, int NWV>
__device__ __forceinline__ float4 sl_dec_outputs(const float (&wv)[NWV], const SlideTaps<R>& tp) {
    using B = SL<R, false>;
    float4 o;
    o.x = sl_row_output<R, FMA, B::OFF + P + 0>(wv, tp);
    o.y = sl_row_output<R, FMA, B::OFF + P + 2>(wv, tp);
    o.z = sl_row_output<R, FMA, B::OFF + P + 4>(wv, tp);
    o.w = sl_row_output<R, FMA, B::OFF + P + 6>(wv, tp);
    return o;
}

template <int R, bool FMA>
__global__ void __launch_bounds__(64) blur_slide_dec_kernel(const __grid_constant__ CUtensorMap map8, const __grid_constant__ CUtensorMap map1,
                                                            const SlideDecArgs sa, const SlideTaps<R> taps) {
    using B = SL<R, false>;
    using C = SLD<R>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nstg = sa.ahead + 1;
    const int warp_floats = nstg * C::STAGE_FLOATS + B::MID_FLOATS;
    float* stage = reinterpret_cast<float*>(smem_raw) + warp * warp_floats;
    float* mid = stage + nstg * C::STAGE_FLOATS;
    const uint32_t stage_u = tma::smem_u32(stage);
    const uint32_t full_u = tma::smem_u32(reinterpret_cast<uint64_t*>(reinterpret_cast<float*>(smem_raw) + B::NWARP * warp_floats) + warp * (B::MAX_AHEAD + 1));

    const BlurArgs& a = sa.a;
    const int w = a.w, h = a.h;
    if (lane == 0) {
        tma::prefetch_map(&map8);
        for (int s = 0; s < nstg; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full_u + 8 * s) : "memory");
        tma::fence_barrier_init();
    }
    __syncwarp();
    float tk[R + 1];
#pragma unroll
    for (int j = 0; j <= R; ++j) tk[j] = taps.tk[j];
    const float one = taps.one;
    const int cg = lane & 15, rsub = lane >> 4;
    const float* const mid_lane = mid + 2 * lane;

    const uint32_t gw = blockIdx.x * B::NWARP + warp;
    const uint32_t row_lane = gw / (uint32_t)sa.strips;
    if (row_lane >= (uint32_t)sa.lanes) return;
    const int strip = (int)(gw - row_lane * (uint32_t)sa.strips);
    int ri = 0;
    while (ri + 1 < sa.n_regions && strip >= sa.reg[ri + 1].first_strip) ++ri;
    const DecRegion rg = sa.reg[ri];
    // first destination column of this strip; runs start at an even column (the TMA box must start on a 16-byte boundary:
    // x = 2*k0 - RPAD a multiple of 4), columns below k_begin are computed with the wrong parity and never stored
    const int k0 = (rg.k_begin & ~1) + (strip - rg.first_strip) * B::WC;
    const int xs = 2 * k0;                                  // staged column c holds source x = xs - RPAD + c
    const int par = rg.parity;
    const SlPatch patch = sl_patch_setup(lane, w, xs - B::RPAD, R, C::SWW);
    const bool edge = __any_sync(0xffffffffu, (patch.left & patch.right) != ~0u);
    const int k = k0 + 2 * lane;                            // this lane's destination columns k, k + 1
    const bool ok0 = k >= rg.k_begin && k < rg.k_end, ok1 = k + 1 >= rg.k_begin && k + 1 < rg.k_end;
    uint32_t g = row_lane * sa.share;
    const uint32_t g_end = min(sa.total, g + sa.share);
    int s = 0;
    uint32_t parity = 0;
    // state of the current piece (one flat loop over the chunks of all pieces, as in blur_slide_kernel)
    int i = 0, n_chunks = 0, n_virtual = 0, v_first = 0, bz = 0, y_top = 0;
    float* dst_img = nullptr;
    float2 acc[B::NACC];
#pragma unroll 1
    for (;;) {
        if (i == n_chunks) {   // next piece
            if (g >= g_end) break;
            const uint32_t b = g / (uint32_t)h;
            const int y0 = (int)(g - b * (uint32_t)h);
            const int y1 = min(h, y0 + (int)(g_end - g));
            g += (uint32_t)(y1 - y0);
            bz = (int)b + a.z0;
            n_virtual = (y1 - y0) + 2 * R;
            n_chunks = (n_virtual + B::CH - 1) / B::CH;
            v_first = y0 - R;
            y_top = y0 - 2 * R;
            i = 0;
            if (lane == 0) {
                int st = s;
                for (int q = 0; q < nstg && q < n_chunks; ++q) {
                    sl_issue_chunk(stage_u + st * C::STAGE_FLOATS * 4, full_u + 8 * st, &map8, &map1, xs - B::RPAD, v_first + q * B::CH, h, bz, C::SWW);
                    if (++st == nstg) st = 0;
                }
            }
            dst_img = a.dst + (size_t)bz * a.dst_stride + k;
#pragma unroll
            for (int q = 0; q < B::NACC; ++q) acc[q] = make_float2(0.0f, 0.0f);
        }
        {
            float* const st = stage + s * C::STAGE_FLOATS;
            // destination row of each of the chunk's eight completed source rows (lane r holds step r's), fetched before the row
            // pass so that the load's latency is off the column pass's critical path
            const int t0 = i * B::CH;
            int my_drow = -1;
            {
                const int t = t0 + lane;
                if (lane < B::CH && t >= 2 * R && t < n_virtual) my_drow = __ldg(a.sel_y + (y_top + t));
            }
            sl_wait(full_u + 8 * s, parity);
            if (edge) sl_patch_apply<B::CH>(st, patch, C::SWW);
            {
                float wv[2][C::NW];
                auto load_window = [&](float (&dstw)[C::NW], int rw) {
                    const float* srow = st + rw * C::SWW + 8 * cg;
#pragma unroll
                    for (int q = 0; q < C::NW / 4; ++q) {
                        const float4 f = *reinterpret_cast<const float4*>(srow + 4 * q);
                        dstw[4 * q] = f.x; dstw[4 * q + 1] = f.y; dstw[4 * q + 2] = f.z; dstw[4 * q + 3] = f.w;
                    }
                };
                load_window(wv[0], rsub);
#pragma unroll
                for (int q = 0; q < B::CH / 2; ++q) {
                    const int rw = rsub + 2 * q;
                    if (q + 1 < B::CH / 2) load_window(wv[(q + 1) & 1], rw + 2);
                    const float (&wq)[C::NW] = wv[q & 1];
                    const float4 o4 = par ? sl_dec_outputs<R, FMA, 1>(wq, taps) : sl_dec_outputs<R, FMA, 0>(wq, taps);
                    *reinterpret_cast<float4*>(mid + rw * B::MW + 4 * cg) = o4;
                }
            }
            __syncwarp();
            if (lane == 0 && i + nstg < n_chunks)
                sl_issue_chunk(stage_u + s * C::STAGE_FLOATS * 4, full_u + 8 * s, &map8, &map1, xs - B::RPAD, v_first + (i + nstg) * B::CH, h, bz, C::SWW);

            // column pass: step r consumes virtual row i*8 + r and completes source row yo = y0 - 2R + i*8 + r
#pragma unroll
            for (int r = 0; r < B::CH; ++r) {
                const float2 hv = *reinterpret_cast<const float2*>(mid_lane + r * B::MW);
                const float2 res = sl_col_step<R, FMA>(acc, hv, tk, one);
                const int drow = __shfl_sync(0xffffffffu, my_drow, r);   // -1: not picked, or outside the piece
                if (drow >= 0) {                                         // warp-uniform
                    float* o = dst_img + (size_t)drow * a.dst_pitch;
                    if (ok0) o[0] = res.x;
                    if (ok1) o[1] = res.y;
                }
            }
            __syncwarp();
            if (++s == nstg) { s = 0; parity ^= 1u; }
            ++i;
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------
static constexpr int kSlideRadii[] = {3, 5, 7, 10, 14, 19, 27};

int slide_radius_for(int r) {
    for (int R : kSlideRadii)
        if (r <= R) return R;
    return 0;
}

template <int R>
static int slide_box_width_r(bool dog) { return dog ? SL<R, true>::SWW : SL<R, false>::SWW; }
static_assert(SL<10, true>::SWW == SL<10, false>::SWW, "the staged width must not depend on the mode");

int slide_box_width(int r) {   // same for DOG and plain (SWW does not depend on it)
    switch (slide_radius_for(r)) {
        case 3: return slide_box_width_r<3>(true);
        case 5: return slide_box_width_r<5>(true);
        case 7: return slide_box_width_r<7>(true);
        case 10: return slide_box_width_r<10>(true);
        case 14: return slide_box_width_r<14>(true);
        case 19: return slide_box_width_r<19>(true);
        case 27: return slide_box_width_r<27>(true);
        default: return 0;
    }
}

// Grid and share of a launch (see SlideArgs): as many CTAs as the device holds at once, fewer when the level is so small that a
// share would be mostly vertical halo.  Returns the number of CTAs.
static int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }
static int slide_partition(int strips, int batch, int h, int R, int nwarp, int n_sm, int cps, int* lanes, uint32_t* share, uint32_t* total) {
    static const int cps_cap = env_int("SIFT_GPU_SLIDE_CPS", 0);          // development knobs
    static const int even = env_int("SIFT_GPU_SLIDE_CPS_EVEN", 0);
    static const int min_mult = env_int("SIFT_GPU_SLIDE_MIN_SHARE", 1);   // x (2R + 8) rows: a single image still spreads over the GPU (measured: 4 -> 1 takes a lone 1080p pyramid from 559 to 332 us, batches do not care)
    const uint64_t tot = (uint64_t)batch * (uint64_t)h;
    if (tot >= (1ull << 31) || strips < 1) return -1;
    if (cps_cap > 0 && cps > cps_cap) cps = cps_cap;
    if (even && cps > 1 && (cps * nwarp) % 4 != 0) --cps;
    const uint64_t min_share = (uint64_t)std::max(64, min_mult * (2 * R + 8));
    uint64_t nl = (uint64_t)n_sm * cps * nwarp / strips;           // row lanes the resident warps can form
    const uint64_t by_work = std::max<uint64_t>(1, tot / min_share);
    if (nl > by_work) nl = by_work;
    if (nl < 1) nl = 1;
    *lanes = (int)nl;
    *share = (uint32_t)((tot + nl - 1) / nl);
    *total = (uint32_t)tot;
    return (int)((nl * strips + nwarp - 1) / nwarp);
}

struct SlideDevInfo { bool ready = false; int n_sm = 148; int ahead = 1; int cps[8][2][3] = {}; };   // [radius index][fma][mode]
static int slide_ahead() {
    static const int v = [] { const char* e = getenv("SIFT_GPU_SLIDE_AHEAD"); int a = e ? atoi(e) : 1; return a < 1 ? 1 : (a > 4 ? 4 : a); }();
    return v;
}
static SlideDevInfo g_slide_dev[64];

template <int R, bool FMA, int MODE>
static int slide_prepare_one(int* cps) {
    using C = SL<R, MODE != 0>;
    const size_t smem = C::smem(slide_ahead() + 1);
    SIFT_CUDA_TRY(cudaFuncSetAttribute(blur_slide_kernel<R, FMA, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int n = 0;
    SIFT_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, blur_slide_kernel<R, FMA, MODE>, C::NWARP * 32, smem));
    *cps = n > 0 ? n : 1;
    return 0;
}
template <int R>
static int slide_prepare_r(SlideDevInfo& d, int idx) {
    int rc;
    if ((rc = slide_prepare_one<R, false, 0>(&d.cps[idx][0][0])) || (rc = slide_prepare_one<R, false, 1>(&d.cps[idx][0][1])) ||
        (rc = slide_prepare_one<R, false, 2>(&d.cps[idx][0][2])) || (rc = slide_prepare_one<R, true, 0>(&d.cps[idx][1][0])) ||
        (rc = slide_prepare_one<R, true, 1>(&d.cps[idx][1][1])) || (rc = slide_prepare_one<R, true, 2>(&d.cps[idx][1][2])))
        return rc;
    return 0;
}

static int slide_dec_prepare_device(int dev);
// Per-device set-up (function attributes are per device): called from sift_gpu_create after cudaSetDevice.
int slide_prepare_device() {
    int dev = 0;
    SIFT_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return 0;
    SlideDevInfo& d = g_slide_dev[dev];
    if (d.ready) return 0;
    if (cudaDeviceGetAttribute(&d.n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || d.n_sm <= 0) d.n_sm = 148;
    int rc;
    if ((rc = slide_prepare_r<3>(d, 0)) || (rc = slide_prepare_r<5>(d, 1)) || (rc = slide_prepare_r<7>(d, 2)) || (rc = slide_prepare_r<10>(d, 3)) ||
        (rc = slide_prepare_r<14>(d, 4)) || (rc = slide_prepare_r<19>(d, 5)) || (rc = slide_prepare_r<27>(d, 6)))
        return rc;
    if ((rc = slide_dec_prepare_device(dev))) return rc;
    d.ready = true;
    return 0;
}

template <int R>
static int launch_slide_r(const BlurArgs& a, int batch, bool fma, int idx, cudaStream_t s) {
    using C = SL<R, true>;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !g_slide_dev[dev].ready) {
        const int rc = slide_prepare_device();
        if (rc) return rc;
    }
    const SlideDevInfo& d = g_slide_dev[dev < 64 && dev >= 0 ? dev : 0];
    const int mode = a.dog ? (a.dst ? 1 : 2) : 0;
    if (!a.dog && !a.dst) return -1;
    // SIFT_GPU_FLAG_FMA_BLUR permits fusing, it does not demand it: the r = 7 blur + DoG launch of a large level is bound by
    // its memory access pattern, and there the exact-arithmetic instantiation (120 registers, no spills) is the faster one
    // (324 vs 345 us for 64 1080p frames, measured on every build variant) — and it is bit-exact on top.
    if (R == 7 && mode == 1 && (size_t)a.w * (size_t)a.h >= (1u << 20)) fma = false;
    if (mode == 1 && a.dst_pitch != a.dog_pitch) return -1;   // one row pointer serves both outputs
    // taps for radius R: tk[j], j = 0..2R, zero outside the real radius; the kernel relies on their symmetry
    const int r = a.r, pad = R - r;
    float tk[2 * R + 1];
    for (int j = 0; j <= 2 * R; ++j) tk[j] = (j >= pad && j <= pad + 2 * r) ? a.taps_host[2 * r - (j - pad)] : 0.0f;
    for (int j = 0; j < R; ++j)
        if (tk[j] != tk[2 * R - j]) return -1;   // not a symmetric kernel: another implementation takes it
    SlideTaps<R> tp;
    tp.one = 1.0f;
    for (int j = 0; j <= R; ++j) tp.tk[j] = tk[j];
    for (int m = 0; m < R; ++m) {
        tp.pe[m] = make_float2(tk[2 * m], tk[2 * m + 1]);
        tp.po[m] = make_float2(tk[2 * m + 1], tk[2 * m + 2]);
    }
    SlideArgs sa;
    sa.a = a;
    const int strips = (a.w + C::WC - 1) / C::WC;
    // (Measured and not adopted: capping the resident CTAs of single launches.  In isolation blur + DoG r = 5 is 15 % faster at
    // 4 CTAs/SM than at 9 and the fused r = 10 launch 7 % at 6, but with the image groups of a pass running side by side the
    // caps cost 1-3 % of the stage.)
    const int cps = d.cps[idx][fma ? 1 : 0][mode];
    sa.strips = strips;
    const int ctas = slide_partition(strips, batch, a.h, R, C::NWARP, d.n_sm, cps, &sa.lanes, &sa.share, &sa.total);
    if (ctas <= 0) return -1;
    sa.ahead = slide_ahead();
    dim3 grid(ctas, 1, 1);
    static const bool probe_on = getenv("SIFT_GPU_SLIDE_PROBE") != nullptr;
    sa.probe = nullptr;
    const size_t n_probe = (size_t)ctas * C::NWARP;
    if (probe_on) {
        SIFT_CUDA_TRY(cudaMalloc(&sa.probe, n_probe * 4 * sizeof(unsigned long long)));
        SIFT_CUDA_TRY(cudaMemsetAsync(sa.probe, 0, n_probe * 4 * sizeof(unsigned long long), s));
    }
#define SL_LAUNCH(F, M) blur_slide_kernel<R, F, M><<<grid, C::NWARP * 32, SL<R, M != 0>::smem(sa.ahead + 1), s>>>(a.map[0], a.map[1], sa, tp)
    if (fma) { if (mode == 0) SL_LAUNCH(true, 0); else if (mode == 1) SL_LAUNCH(true, 1); else SL_LAUNCH(true, 2); }
    else { if (mode == 0) SL_LAUNCH(false, 0); else if (mode == 1) SL_LAUNCH(false, 1); else SL_LAUNCH(false, 2); }
#undef SL_LAUNCH
    SIFT_CUDA_TRY(cudaGetLastError());
    if (probe_on) {
        std::vector<unsigned long long> hp(n_probe * 4);
        SIFT_CUDA_TRY(cudaStreamSynchronize(s));
        SIFT_CUDA_TRY(cudaMemcpy(hp.data(), sa.probe, hp.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        cudaFree(sa.probe);
        unsigned long long t0 = ~0ull, t1 = 0;
        for (size_t i = 0; i < n_probe; ++i)
            if (hp[4 * i + 1]) { t0 = std::min(t0, hp[4 * i]); t1 = std::max(t1, hp[4 * i + 1]); }
        const double T = (double)(t1 - t0);
        std::vector<double> sm_end(256, 0.0), sm_start(256, 1e30);
        std::vector<int> sm_warps(256, 0);
        double dur_sum = 0, dur_min = 1e30, dur_max = 0, q_dur[4] = {0, 0, 0, 0};
        int q_n[4] = {0, 0, 0, 0}, n = 0;
        for (size_t i = 0; i < n_probe; ++i) {
            if (!hp[4 * i + 1]) continue;
            const double a0 = (double)(hp[4 * i] - t0), a1 = (double)(hp[4 * i + 1] - t0), dd = a1 - a0;
            const int sm = (int)hp[4 * i + 2] & 255, q = (int)hp[4 * i + 3] & 3;
            sm_end[sm] = std::max(sm_end[sm], a1); sm_start[sm] = std::min(sm_start[sm], a0); ++sm_warps[sm];
            dur_sum += dd; dur_min = std::min(dur_min, dd); dur_max = std::max(dur_max, dd); q_dur[q] += dd; ++q_n[q]; ++n;
        }
        if (const char* dump = getenv("SIFT_GPU_SLIDE_PROBE_DUMP")) {
            if (FILE* f = fopen(dump, "a")) {
                fprintf(f, "# R=%d mode=%d w=%d h=%d strips=%d lanes=%d share=%u\n", R, mode, a.w, a.h, sa.strips, sa.lanes, sa.share);
                for (size_t i = 0; i < n_probe; ++i)
                    if (hp[4 * i + 1]) fprintf(f, "%zu %llu %llu %llu %llu\n", i, hp[4 * i] - t0, hp[4 * i + 1] - t0, hp[4 * i + 2], hp[4 * i + 3]);
                fclose(f);
            }
        }
        {   // who is slow: by strip position, by row lane, and the slowest warps
            const int ns = sa.strips;
            std::vector<double> sd((size_t)ns, 0.0); std::vector<int> sn((size_t)ns, 0);
            std::vector<std::pair<double, size_t>> all;
            for (size_t i = 0; i < n_probe; ++i) {
                if (!hp[4 * i + 1]) continue;
                const double dd = (double)(hp[4 * i + 1] - hp[4 * i]);
                sd[i % ns] += dd; ++sn[i % ns];
                all.push_back(std::make_pair(dd, i));
            }
            std::sort(all.begin(), all.end());
            fprintf(stderr, "   by strip:");
            for (int k = 0; k < ns; ++k) fprintf(stderr, " %.0f", sd[k] / std::max(1, sn[k]) * 1e-3);
            fprintf(stderr, "\n   slowest:");
            for (size_t k = 0; k < 12 && k < all.size(); ++k) {
                const size_t i = all[all.size() - 1 - k].second;
                fprintf(stderr, " [%.0fus strip %d lane %d sm %d w %d start %.0f]", all[all.size() - 1 - k].first * 1e-3, (int)(i % ns), (int)(i / ns), (int)hp[4 * i + 2],
                        (int)hp[4 * i + 3], (double)(hp[4 * i] - t0) * 1e-3);
            }
            fprintf(stderr, "\n   deciles:");
            for (int k = 0; k <= 10; ++k) fprintf(stderr, " %.0f", all[std::min(all.size() - 1, all.size() * k / 10)].first * 1e-3);
            fprintf(stderr, "\n");
        }
        std::vector<double> ends, wcount;
        for (int i = 0; i < 256; ++i) if (sm_warps[i]) { ends.push_back(sm_end[i] / T); wcount.push_back(sm_warps[i]); }
        std::sort(ends.begin(), ends.end());
        fprintf(stderr, "[slide probe R=%d mode=%d %dx%d] span %.1f us, %d warps: duration mean %.1f min %.1f max %.1f us; by warpid%%4: %.1f(%d) %.1f(%d) %.1f(%d) %.1f(%d); SM end/T: min %.2f p25 %.2f med %.2f p75 %.2f max %.2f; warps/SM min %.0f max %.0f\n",
                R, mode, a.w, a.h, T * 1e-3, n, dur_sum / n * 1e-3, dur_min * 1e-3, dur_max * 1e-3, q_dur[0] / std::max(1, q_n[0]) * 1e-3, q_n[0],
                q_dur[1] / std::max(1, q_n[1]) * 1e-3, q_n[1], q_dur[2] / std::max(1, q_n[2]) * 1e-3, q_n[2], q_dur[3] / std::max(1, q_n[3]) * 1e-3, q_n[3],
                ends.front(), ends[ends.size() / 4], ends[ends.size() / 2], ends[3 * ends.size() / 4], ends.back(),
                *std::min_element(wcount.begin(), wcount.end()), *std::max_element(wcount.begin(), wcount.end()));
    }
    return 0;
}

// returns -1 when the launch does not qualify (decimation, no TMA descriptors, small level, radius above 27)
int launch_slide(const BlurArgs& a, int batch, bool fma, cudaStream_t s) {
    static const bool off = getenv("SIFT_GPU_NO_SLIDE") != nullptr;
    if (off || a.sel_x || !a.map || !a.taps_host || a.w < 32 || a.h < 16) return -1;
    switch (slide_radius_for(a.r)) {
        case 3: return launch_slide_r<3>(a, batch, fma, 0, s);
        case 5: return launch_slide_r<5>(a, batch, fma, 1, s);
        case 7: return launch_slide_r<7>(a, batch, fma, 2, s);
        case 10: return launch_slide_r<10>(a, batch, fma, 3, s);
        case 14: return launch_slide_r<14>(a, batch, fma, 4, s);
        case 19: return launch_slide_r<19>(a, batch, fma, 5, s);
        case 27: return launch_slide_r<27>(a, batch, fma, 6, s);
        default: return -1;
    }
}


int slide_dec_box_width(int r) {
    switch (slide_radius_for(r)) {
        case 3: return SLD<3>::SWW;
        case 5: return SLD<5>::SWW;
        case 7: return SLD<7>::SWW;
        case 10: return SLD<10>::SWW;
        case 14: return SLD<14>::SWW;
        case 19: return SLD<19>::SWW;
        default: return 0;   // 27: the window would not fit the register budget; those launches take another kernel
    }
}

static int g_dec_cps[64][8][2];
template <int R>
static int slide_dec_prepare_r(int dev, int idx) {
    const size_t smem = SLD<R>::smem(slide_ahead() + 1);
    SIFT_CUDA_TRY(cudaFuncSetAttribute(blur_slide_dec_kernel<R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SIFT_CUDA_TRY(cudaFuncSetAttribute(blur_slide_dec_kernel<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int n = 0;
    SIFT_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, blur_slide_dec_kernel<R, false>, 64, smem));
    g_dec_cps[dev][idx][0] = n > 0 ? n : 1;
    SIFT_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, blur_slide_dec_kernel<R, true>, 64, smem));
    g_dec_cps[dev][idx][1] = n > 0 ? n : 1;
    return 0;
}
static int slide_dec_prepare_device(int dev) {
    int rc;
    if ((rc = slide_dec_prepare_r<3>(dev, 0)) || (rc = slide_dec_prepare_r<5>(dev, 1)) || (rc = slide_dec_prepare_r<7>(dev, 2)) ||
        (rc = slide_dec_prepare_r<10>(dev, 3)) || (rc = slide_dec_prepare_r<14>(dev, 4)) || (rc = slide_dec_prepare_r<19>(dev, 5)))
        return rc;
    return 0;
}

template <int R>
static int launch_slide_dec_r(const BlurArgs& a, int batch, bool fma, int idx, cudaStream_t s) {
    using B = SL<R, false>;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!g_slide_dev[dev].ready) {
        const int rc = slide_prepare_device();
        if (rc) return rc;
    }
    const SlideDevInfo& d = g_slide_dev[dev];
    // runs of constant parity in the column pick; every picked source column must be 2k or 2k + 1
    SlideDecArgs sa;
    sa.a = a;
    sa.n_regions = 0;
    int k = 0, blocks = 0;   // blocks: warp strips so far
    for (int x = 0; x < a.w; ++x) {
        const int dk = a.sel_x_host[x];
        if (dk < 0) continue;
        if (dk != k || (x - 2 * k != 0 && x - 2 * k != 1)) return -1;
        const int par = x - 2 * k;
        if (sa.n_regions == 0 || sa.reg[sa.n_regions - 1].parity != par) {
            if (sa.n_regions) {
                DecRegion& pr = sa.reg[sa.n_regions - 1];
                pr.k_end = k;
                blocks += (pr.k_end - (pr.k_begin & ~1) + B::WC - 1) / B::WC;
            }
            if (sa.n_regions == 4) return -1;
            sa.reg[sa.n_regions++] = DecRegion{k, k, par, blocks};
        }
        ++k;
    }
    if (sa.n_regions == 0) return -1;
    {
        DecRegion& pr = sa.reg[sa.n_regions - 1];
        pr.k_end = k;
        blocks += (pr.k_end - (pr.k_begin & ~1) + B::WC - 1) / B::WC;
    }
    for (int y = 0; y < a.h; ++y) {   // rows may be picked in any pattern; only sanity-check the range
        const int dy = a.sel_y_host[y];
        if (dy < -1) return -1;
    }
    const int r = a.r, pad = R - r;
    float tkv[2 * R + 1];
    for (int j = 0; j <= 2 * R; ++j) tkv[j] = (j >= pad && j <= pad + 2 * r) ? a.taps_host[2 * r - (j - pad)] : 0.0f;
    for (int j = 0; j < R; ++j)
        if (tkv[j] != tkv[2 * R - j]) return -1;
    SlideTaps<R> tp;
    tp.one = 1.0f;
    for (int j = 0; j <= R; ++j) tp.tk[j] = tkv[j];
    for (int m = 0; m < R; ++m) {
        tp.pe[m] = make_float2(tkv[2 * m], tkv[2 * m + 1]);
        tp.po[m] = make_float2(tkv[2 * m + 1], tkv[2 * m + 2]);
    }
    const int cps = g_dec_cps[dev][idx][fma ? 1 : 0];
    sa.strips = blocks;
    const int ctas = slide_partition(blocks, batch, a.h, R, B::NWARP, d.n_sm, cps, &sa.lanes, &sa.share, &sa.total);
    if (ctas <= 0) return -1;
    sa.ahead = slide_ahead();
    dim3 grid(ctas, 1, 1);
    if (fma) blur_slide_dec_kernel<R, true><<<grid, 64, SLD<R>::smem(sa.ahead + 1), s>>>(a.map[0], a.map[1], sa, tp);
    else blur_slide_dec_kernel<R, false><<<grid, 64, SLD<R>::smem(sa.ahead + 1), s>>>(a.map[0], a.map[1], sa, tp);
    SIFT_CUDA_TRY(cudaGetLastError());
    return 0;
}

// returns -1 when the launch does not qualify
int launch_slide_dec(const BlurArgs& a, int batch, bool fma, cudaStream_t s) {
    static const bool off = getenv("SIFT_GPU_NO_SLIDE_DEC") != nullptr || getenv("SIFT_GPU_NO_SLIDE") != nullptr;
    if (off || !a.sel_x || !a.sel_x_host || !a.sel_y_host || !a.map || !a.taps_host || !a.dst || a.dog || a.w < 64 || a.h < 16) return -1;
    switch (slide_radius_for(a.r)) {
        case 3: return launch_slide_dec_r<3>(a, batch, fma, 0, s);
        case 5: return launch_slide_dec_r<5>(a, batch, fma, 1, s);
        case 7: return launch_slide_dec_r<7>(a, batch, fma, 2, s);
        case 10: return launch_slide_dec_r<10>(a, batch, fma, 3, s);
        case 14: return launch_slide_dec_r<14>(a, batch, fma, 4, s);
        case 19: return launch_slide_dec_r<19>(a, batch, fma, 5, s);
        default: return -1;
    }
}

}  // namespace siftgpu
