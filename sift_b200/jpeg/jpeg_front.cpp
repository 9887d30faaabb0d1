// nvJPEG front end (include/sift_gpu_jpeg.h): JPEG bytes -> device RGB planes -> band 0 as a device-resident u8 frame of
// sift_gpu_run.  Replaces the decode half of the reference's vigra::importImage (main.cpp:52-54) for JPEG input.
#include "../../include/sift_gpu_jpeg.h"

#include <cuda_runtime.h>
#include <nvjpeg.h>

#include <string>
#include <vector>

struct sift_gpu_jpeg_ctx {
    sift_gpu_ctx* sift = nullptr;
    int device = 0, max_w = 0, max_h = 0, max_images = 0;
    nvjpegHandle_t handle = nullptr;
    nvjpegJpegState_t state = nullptr;
    cudaStream_t stream = nullptr;
    unsigned char* planes = nullptr;  // max_images band-0 planes, then one scratch plane each for G and B
    size_t plane_bytes = 0;
    std::vector<int> w, h;            // of the last run
    std::vector<sift_gpu_image> imgs;
    std::string error;
};

static std::string g_jpeg_error;

static int fail(sift_gpu_jpeg_ctx* c, int code, const std::string& what) {
    (c ? c->error : g_jpeg_error) = what;
    return code;
}

extern "C" {

const char* sift_gpu_jpeg_last_error(const sift_gpu_jpeg_ctx* c) { return c ? c->error.c_str() : g_jpeg_error.c_str(); }

void sift_gpu_jpeg_destroy(sift_gpu_jpeg_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->state) nvjpegJpegStateDestroy(c->state);
    if (c->handle) nvjpegDestroy(c->handle);
    if (c->planes) cudaFree(c->planes);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int sift_gpu_jpeg_create(sift_gpu_ctx* sift, int device, int max_width, int max_height, int max_images, sift_gpu_jpeg_ctx** out) {
    if (!sift || !out || max_width < 1 || max_height < 1 || max_images < 1) return fail(nullptr, SIFT_GPU_E_INVALID, "bad argument");
    *out = nullptr;
    if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, SIFT_GPU_E_CUDA, "no such CUDA device");
    sift_gpu_jpeg_ctx* c = new sift_gpu_jpeg_ctx();
    c->sift = sift; c->device = device; c->max_w = max_width; c->max_h = max_height; c->max_images = max_images;
    c->plane_bytes = (size_t)max_width * (size_t)max_height;
    const nvjpegStatus_t s1 = nvjpegCreateSimple(&c->handle);
    const nvjpegStatus_t s2 = s1 == NVJPEG_STATUS_SUCCESS ? nvjpegJpegStateCreate(c->handle, &c->state) : s1;
    if (s2 != NVJPEG_STATUS_SUCCESS || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc(&c->planes, c->plane_bytes * (size_t)(max_images + 2)) != cudaSuccess) {
        const std::string why = "nvJPEG / CUDA set-up failed (nvjpeg status " + std::to_string((int)s2) + ")";
        sift_gpu_jpeg_destroy(c);
        return fail(nullptr, SIFT_GPU_E_CUDA, why);
    }
    *out = c;
    return SIFT_GPU_OK;
}

int sift_gpu_jpeg_run(sift_gpu_jpeg_ctx* c, const sift_gpu_jpeg* jpegs, int n, sift_gpu_result* results) {
    if (!c || n < 0 || (n > 0 && (!jpegs || !results))) return fail(c, SIFT_GPU_E_INVALID, "null argument");
    if (n > c->max_images) return fail(c, SIFT_GPU_E_CAPACITY, "more JPEGs than max_images");
    if (cudaSetDevice(c->device) != cudaSuccess) return fail(c, SIFT_GPU_E_CUDA, "cudaSetDevice failed");
    c->error.clear();
    c->w.assign((size_t)n, 0);
    c->h.assign((size_t)n, 0);
    c->imgs.assign((size_t)n, sift_gpu_image{});
    unsigned char* scratch_g = c->planes + c->plane_bytes * (size_t)c->max_images;
    unsigned char* scratch_b = scratch_g + c->plane_bytes;
    for (int i = 0; i < n; ++i) {
        int ncomp = 0, ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
        nvjpegChromaSubsampling_t sub;
        if (!jpegs[i].data || jpegs[i].size < 4 ||
            nvjpegGetImageInfo(c->handle, jpegs[i].data, jpegs[i].size, &ncomp, &sub, ws, hs) != NVJPEG_STATUS_SUCCESS || ws[0] < 1 || hs[0] < 1)
            return fail(c, SIFT_GPU_E_INVALID, "image " + std::to_string(i) + " is not a decodable JPEG");
        if (ws[0] > c->max_w || hs[0] > c->max_h) return fail(c, SIFT_GPU_E_CAPACITY, "JPEG larger than max_width x max_height");
        nvjpegImage_t dst{};
        dst.channel[0] = c->planes + c->plane_bytes * (size_t)i;  // R (or the grey values): the band the reference reads
        dst.channel[1] = scratch_g;                               // G and B are decoded and dropped (decodes run in stream order)
        dst.channel[2] = scratch_b;
        dst.pitch[0] = dst.pitch[1] = dst.pitch[2] = (size_t)ws[0];
        const nvjpegStatus_t st = nvjpegDecode(c->handle, c->state, jpegs[i].data, jpegs[i].size, NVJPEG_OUTPUT_RGB, &dst, c->stream);
        if (st != NVJPEG_STATUS_SUCCESS) return fail(c, SIFT_GPU_E_INVALID, "nvjpegDecode failed on image " + std::to_string(i) + " (status " + std::to_string((int)st) + ")");
        c->w[(size_t)i] = ws[0];
        c->h[(size_t)i] = hs[0];
        sift_gpu_image& im = c->imgs[(size_t)i];
        im.data = dst.channel[0];
        im.width = ws[0];
        im.height = hs[0];
        im.row_stride_bytes = 0;
        im.dtype = SIFT_GPU_DTYPE_U8;
        im.memory = SIFT_GPU_MEM_DEVICE;
        im.upsampled_out = nullptr;
    }
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return fail(c, SIFT_GPU_E_CUDA, "decode stream failed");
    const int rc = sift_gpu_run(c->sift, c->imgs.data(), n, results);
    if (rc != SIFT_GPU_OK) c->error = sift_gpu_last_error(c->sift);
    return rc;
}

int sift_gpu_jpeg_decoded(sift_gpu_jpeg_ctx* c, int i, unsigned char* out, int* width, int* height) {
    if (!c || i < 0 || i >= (int)c->w.size()) return fail(c, SIFT_GPU_E_INVALID, "no such image in the last run");
    if (width) *width = c->w[(size_t)i];
    if (height) *height = c->h[(size_t)i];
    if (!out) return SIFT_GPU_OK;
    if (cudaSetDevice(c->device) != cudaSuccess ||
        cudaMemcpy(out, c->planes + c->plane_bytes * (size_t)i, (size_t)c->w[(size_t)i] * (size_t)c->h[(size_t)i], cudaMemcpyDeviceToHost) != cudaSuccess)
        return fail(c, SIFT_GPU_E_CUDA, "copy of the decoded plane failed");
    return SIFT_GPU_OK;
}

}  // extern "C"
