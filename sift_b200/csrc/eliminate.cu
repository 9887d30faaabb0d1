// Candidate rejection, reference Sift::_eliminateEdgeResponses (sift.cpp:288-346) with
// alg::foDerivative / alg::soDerivative (algorithms.cpp:66-106), op for op in fp32:
//   inverse(-H) fails -> reject; linearSolve(inv, grad) fails -> reject; any component > 127.5 ->
//   reject; f = dot(grad, ext) * (0.5 + D(x,y)) [double], f < 7.65 -> reject; det < 0 -> reject;
//   tr^2/det > 12.1f -> reject.  loc/scale are never moved.
// One thread per candidate; unfiltered candidates are appended (warp-aggregated atomic) to the
// survivor list together with their canonical position, which the host order replay needs.
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
        const unsigned ballot = __ballot_sync(0xffffffffu, keep);
        if (ballot) {
            const int lane = threadIdx.x & 31;
            uint32_t slot0 = 0;
            if (lane == 0) slot0 = atomicAdd(&n_surv[b], (uint32_t)__popc(ballot));
            slot0 = __shfl_sync(0xffffffffu, slot0, 0);
            if (keep) {
                const uint32_t slot = slot0 + (uint32_t)__popc(ballot & ((1u << lane) - 1u));
                if (slot < surv_stride) {
                    Surv s;
                    s.canon = i;
                    s.x = c.x; s.y = c.y; s.octave = c.octave; s.index = c.index; s.pad = 0;
                    survivors[(size_t)b * surv_stride + slot] = s;
                }
            }
        }
    }
}

int launch_eliminate(const ScanLayer* layers_dev, int n_layers, Cand* cands, size_t cand_stride,
                     const uint32_t* n_cand, Surv* survivors, size_t surv_stride, uint32_t* n_surv, int dogs_per_epoch,
                     int batch, cudaStream_t s, uint64_t* launches) {
    SIFT_CUDA_TRY(cudaMemsetAsync(n_surv, 0, sizeof(uint32_t) * (size_t)batch, s));
    dim3 grid(148 * 4, batch);
    eliminate_kernel<<<grid, 128, 0, s>>>(layers_dev, n_layers, cands, cand_stride, n_cand, survivors, surv_stride, n_surv,
                                         dogs_per_epoch);
    if (launches) ++*launches;
    SIFT_CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace siftgpu
