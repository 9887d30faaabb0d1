// Candidate rejection, reference Sift::_eliminateEdgeResponses (sift.cpp:288-346) with
// alg::foDerivative / alg::soDerivative (algorithms.cpp:66-106), op for op in fp32:
//   inverse(-H) fails -> reject; linearSolve(inv, grad) fails -> reject; any component > 127.5 ->
//   reject; f = dot(grad, ext) * (0.5 + D(x,y)) [double], f < 7.65 -> reject; det < 0 -> reject;
//   tr^2/det > 12.1f -> reject.  loc/scale are never moved.
// One thread per candidate writes the verdict into the candidate list; a second kernel (one CTA per image, block-wide
// prefix sums) then compacts the unfiltered candidates IN CANONICAL ORDER into the survivor list, each with its
// position in the candidate vector (the host's std::sort replay works on those positions, and no longer has to sort
// an atomically appended list back into order).
#include "common.cuh"
#include "vigra_qr.cuh"

namespace siftgpu {

__global__ void __launch_bounds__(128) eliminate_kernel(const ScanLayer* __restrict__ layers, int n_layers,
                                                        Cand* __restrict__ cands, size_t cand_stride,
                                                        const uint32_t* __restrict__ n_cand, Surv* __restrict__ survivors,
                                                        size_t surv_stride, uint32_t* __restrict__ n_surv,
                                                        int dogs_per_epoch) {
    const int b = blockIdx.y;
    const uint32_t n = n_cand[b];
    Cand* cl = cands + (size_t)b * cand_stride;
    const float t = (float)(121.0 / 10);  // sift.cpp:294
    const int mid_layers = dogs_per_epoch - 2;
    const uint32_t span = (uint32_t)gridDim.x * blockDim.x;
    // uniform trip count so the warp-level append below is convergent
    for (uint32_t base = (uint32_t)blockIdx.x * blockDim.x; base < n; base += span) {
        const uint32_t i = base + threadIdx.x;
        bool keep = false;
        Cand c;
        if (i < n) c = cl[i];
        if (i < n && !c.filtered) {  // the extrema kernels already applied the det / edge-ratio tests (extrema.cu, PREFILTER)
            const ScanLayer L = layers[c.octave * mid_layers + (c.index - 1)];
            const size_t off = (size_t)b * L.stride;
            const float* D0 = L.d0 + off;
            const float* D1 = L.d1 + off;
            const float* D2 = L.d2 + off;
            const int w = L.pitch, x = c.x, y = c.y;
#define AT(D, xx, yy) D[(size_t)(yy) * w + (xx)]
            // algorithms.cpp:69-71
            const float dx = (AT(D1, x - 1, y) - AT(D1, x + 1, y)) / 2;
            const float dy = (AT(D1, x, y - 1) - AT(D1, x, y + 1)) / 2;
            const float ds = (AT(D0, x, y) - AT(D2, x, y)) / 2;
            // algorithms.cpp:82-92
            const float c11 = AT(D1, x, y);
            const float dxx = AT(D1, x + 1, y) + AT(D1, x - 1, y) - 2 * c11;
            const float dyy = AT(D1, x, y + 1) + AT(D1, x, y - 1) - 2 * c11;
            const float dss = AT(D2, x, y) + AT(D0, x, y) - 2 * c11;
            const float dxy = (AT(D1, x + 1, y + 1) - AT(D1, x - 1, y + 1) - AT(D1, x + 1, y - 1) + AT(D1, x - 1, y - 1)) / 2;
            const float dxs = (AT(D2, x + 1, y) - AT(D2, x - 1, y) - AT(D0, x + 1, y) + AT(D0, x - 1, y)) / 2;
            const float dys = (AT(D2, x, y + 1) - AT(D2, x, y + 1) - AT(D0, x, y + 1) + AT(D0, x, y - 1)) / 2;
#undef AT
            // The reference tests in the order inverse / solve / ext > 127.5 / contrast / det < 0 / edge ratio and stops at the
            // first hit; the verdict is the OR of all of them, so the two tests that need no linear algebra run first.
            bool filtered = false;
            {
                const float tr = dxx + dyy;
                const float det = (float)((double)(dxx * dyy) - (double)dxy * (double)dxy);
                if (det < 0) filtered = true;
                else if (((double)tr * (double)tr) / (double)det > (double)t) filtered = true;
            }
            if (!filtered) {
                const float negH[9] = {dxx * -1.0f, dxy * -1.0f, dxs * -1.0f, dxy * -1.0f, dyy * -1.0f, dys * -1.0f,
                                       dxs * -1.0f, dys * -1.0f, dss * -1.0f};
                const float grad[3] = {dx, dy, ds};
                float inv[9], ext[3];
                if (!qr::inverse3_fast(negH, inv)) filtered = true;
                if (!filtered && !qr::solve3_fullrank(inv, grad, ext)) filtered = true;
                if (!filtered && ((double)ext[0] > 127.5 || (double)ext[1] > 127.5 || (double)ext[2] > 127.5)) filtered = true;
                if (!filtered) {
                    float f = 0.0f;
                    f = f + grad[0] * ext[0];
                    f = f + grad[1] * ext[1];
                    f = f + grad[2] * ext[2];
                    f = (float)((double)f * (0.5 + (double)c11));
                    if ((double)f < 7.65) filtered = true;
                }
            }
            keep = !filtered;
            cl[i].filtered = filtered ? 1 : 0;
        }
        (void)keep;
    }
}

// Ordered compaction of the unfiltered candidates of image blockIdx.x.
__global__ void __launch_bounds__(1024) compact_survivors_kernel(const Cand* __restrict__ cands, size_t cand_stride, const uint32_t* __restrict__ n_cand,
                                                                 Surv* __restrict__ survivors, size_t surv_stride, uint32_t* __restrict__ n_surv) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    const int b = blockIdx.x;
    const uint32_t n = n_cand[b];
    const Cand* cl = cands + (size_t)b * cand_stride;
    Surv* out = survivors + (size_t)b * surv_stride;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    constexpr int PER = 4;   // consecutive candidates per thread and round
    for (uint32_t base = 0; base < n; base += 1024 * PER) {
        const uint32_t i0 = base + (uint32_t)threadIdx.x * PER;
        Cand c[PER];
        uint32_t cnt = 0;
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            if (i0 + q < n) {
                c[q] = cl[i0 + q];
                cnt += c[q].filtered ? 0u : 1u;
            } else {
                c[q].filtered = 1;
            }
        }
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            uint32_t ws = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, ws, o);
                if (lane >= o) ws += t;
            }
            warp_sums[lane] = ws;  // inclusive
        }
        __syncthreads();
        uint32_t slot = carry + (wid ? warp_sums[wid - 1] : 0u) + incl - cnt;
#pragma unroll
        for (int q = 0; q < PER; ++q)
            if (!c[q].filtered) {
                if (slot < surv_stride) {
                    Surv sv;
                    sv.canon = i0 + q;
                    sv.x = c[q].x; sv.y = c[q].y; sv.octave = c[q].octave; sv.index = c[q].index; sv.pad = 0;
                    out[slot] = sv;
                }
                ++slot;
            }
        __syncthreads();
        if (threadIdx.x == 1023) carry = slot;
        __syncthreads();
    }
    if (threadIdx.x == 0) n_surv[b] = carry;
}

int launch_eliminate(const ScanLayer* layers_dev, int n_layers, Cand* cands, size_t cand_stride,
                     const uint32_t* n_cand, Surv* survivors, size_t surv_stride, uint32_t* n_surv, int dogs_per_epoch,
                     int batch, cudaStream_t s, uint64_t* launches) {
    dim3 grid(148 * 4, batch);
    eliminate_kernel<<<grid, 128, 0, s>>>(layers_dev, n_layers, cands, cand_stride, n_cand, survivors, surv_stride, n_surv,
                                         dogs_per_epoch);
    compact_survivors_kernel<<<batch, 1024, 0, s>>>(cands, cand_stride, n_cand, survivors, surv_stride, n_surv);
    if (launches) *launches += 2;
    SIFT_CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace siftgpu
