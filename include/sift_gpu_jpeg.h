/* JPEG front end for the device path ("decode on the device"): the frames arrive as JPEG byte streams, are
 * decoded by nvJPEG straight into device memory, and band 0 of the RGB result — what vigra::importImage leaves in the
 * reference's scalar image (main.cpp:52-54: the R channel of a colour file, the grey values of a grey one) — goes into
 * sift_gpu_run as a device-resident 8-bit frame.  Only the compressed bytes cross PCIe.
 *
 * Separate library (libsift_gpu_jpeg.so, links libnvjpeg + libsift_gpu.so) so that the core ABI has no decoder dependency.
 * JPEG decoders differ by a grey level here and there (IDCT, chroma upsampling), so parity against the CPU oracle is
 * defined on the decoded plane: sift_gpu_jpeg_decoded() returns exactly the pixels the pipeline saw. */
#ifndef SIFT_GPU_JPEG_H
#define SIFT_GPU_JPEG_H

#include <stddef.h>

#include "sift_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sift_gpu_jpeg_ctx sift_gpu_jpeg_ctx;

typedef struct sift_gpu_jpeg {
    const unsigned char* data; /* host memory, a complete JPEG stream */
    size_t size;
} sift_gpu_jpeg;

/* `sift`: the pipeline context the decoded frames are fed to (same device; frames must fit its max_width x max_height).
 * max_images: most JPEGs per sift_gpu_jpeg_run call. */
int sift_gpu_jpeg_create(sift_gpu_ctx* sift, int device, int max_width, int max_height, int max_images, sift_gpu_jpeg_ctx** out);

/* Decodes n JPEGs on the device and runs the SIFT path on band 0 of each; results as sift_gpu_run (valid until the next
 * run on `sift`).  Returns SIFT_GPU_OK or a negative SIFT_GPU_E_* code (SIFT_GPU_E_INVALID: not a decodable JPEG). */
int sift_gpu_jpeg_run(sift_gpu_jpeg_ctx* ctx, const sift_gpu_jpeg* jpegs, int n, sift_gpu_result* results);

/* Band 0 of image i of the last run, tightly packed width*height bytes (out may be NULL to query the size). */
int sift_gpu_jpeg_decoded(sift_gpu_jpeg_ctx* ctx, int i, unsigned char* out, int* width, int* height);

const char* sift_gpu_jpeg_last_error(const sift_gpu_jpeg_ctx* ctx);
void sift_gpu_jpeg_destroy(sift_gpu_jpeg_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* SIFT_GPU_JPEG_H */
