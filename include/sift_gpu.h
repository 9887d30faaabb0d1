/* sift_gpu.h — C ABI of the B200-native SIFT feature pipeline (libsift_gpu.so).
 *
 * The reference (snowiow/SIFT) has no FFI or plugin interface: its only API is the C++ class
 * sift::Sift (reference sift.hpp:17-78), called from main.cpp:56-57.  This header is the
 * drop-in boundary that replaces the body of Sift::calculate (reference sift.cpp:19-57): the host
 * class in include/sift/sift.hpp keeps the reference's surface and calls these three entry
 * points.  Plain pointers and sizes only; nothing here throws; every function returns 0 or a
 * negative SIFT_GPU_E_* code.  There is no CPU fallback: without a CUDA device create() fails.
 *
 * Threading: a ctx is bound to one device and owns its streams, buffers and worker threads; it
 * is not thread-safe.  Use one ctx per GPU (and per calling thread); distinct ctxs are independent.
 */
#ifndef SIFT_GPU_H
#define SIFT_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIFT_GPU_OK 0
#define SIFT_GPU_E_INVALID (-1)      /* bad argument */
#define SIFT_GPU_E_CUDA (-2)         /* CUDA runtime error, see sift_gpu_last_error */
#define SIFT_GPU_E_PRECONDITION (-3) /* what the reference surfaces as vigra::PreconditionViolation
                                        ("kernel longer than line", image < 2 px): SURVEY A.7 */
#define SIFT_GPU_E_CAPACITY (-4)     /* image larger than the ctx was created for */
#define SIFT_GPU_E_ASSERT (-5)       /* reference assert (sift.cpp:382-383): octaves == 0 or dogsPerEpoch < 3 */
#define SIFT_GPU_E_UNSUPPORTED (-6)

/* flags */
#define SIFT_GPU_FLAG_ORDER_CANONICAL 0x1u /* keep survivors in (octave,index,x,y) order on the device instead of
                                              replaying the reference's std::sort permutation (sift.cpp:37,49).
                                              Same keypoint set; descriptors of overlapping windows differ. */
#define SIFT_GPU_FLAG_STRICT 0x2u          /* reproduce the reference's exception from the dead 16x16 blur
                                              (sift.cpp:184) as SIFT_GPU_E_PRECONDITION.  NOT the default of the C ABI:
                                              without it a run with >= 6 octaves returns keypoints where the reference
                                              aborts.  sift::Sift (include/sift/sift.hpp) sets it by default. */
#define SIFT_GPU_FLAG_FMA_BLUR 0x4u        /* fused multiply-add in the Gaussian blur (faster; DoG within 1e-4 relative
                                              of the reference instead of bit-identical to its mulss/addss order) */
#define SIFT_GPU_FLAG_KEEP_UPSAMPLED 0x8u  /* subpixel: copy the 2x image back (reference overwrites img, sift.cpp:21) */
#define SIFT_GPU_FLAG_SERIAL 0x10u         /* one device pass at a time (no overlap between passes): clean per-stage timings */

#define SIFT_GPU_DTYPE_F32 0
#define SIFT_GPU_DTYPE_U8 1
#define SIFT_GPU_MEM_HOST 0
#define SIFT_GPU_MEM_DEVICE 1

/* Constructor arguments of sift::Sift (reference sift.hpp:66-71) plus sizing. */
typedef struct sift_gpu_params {
    float sigma;             /* _sigma, default 1.6f */
    float k;                 /* _k, default (float)sqrt(2) */
    uint16_t octaves;        /* _octaves, ctor default 3 (CLI default 4) */
    uint16_t dogs_per_epoch; /* _dogsPerEpoch, default 3 */
    uint8_t subpixel;        /* start from the 2x upsampled image (sift.cpp:20-21) */
    int32_t device;          /* CUDA device ordinal */
    int32_t max_width;       /* largest input image (before upsampling) */
    int32_t max_height;
    int32_t max_batch;       /* images processed per device pass; run() accepts any count */
    uint32_t flags;
} sift_gpu_params;

/* One loaded image: what vigra::importImage leaves in MultiArray<2,float> (main.cpp:52-54):
 * band 0, values 0..255, (x,y) -> y*row_stride + x.  Caller-owned, read-only, valid until run returns. */
typedef struct sift_gpu_image {
    const void* data;
    int32_t width, height;
    int64_t row_stride_bytes; /* 0 = tightly packed */
    int32_t dtype;            /* SIFT_GPU_DTYPE_* */
    int32_t memory;           /* SIFT_GPU_MEM_* (device memory must be on params.device) */
    void* upsampled_out;      /* optional host float[4*w*h] when SIFT_GPU_FLAG_KEEP_UPSAMPLED */
} sift_gpu_image;

/* sift::InterestPoint (reference interestpoint.hpp:13-63) minus the descriptor vector. */
typedef struct sift_gpu_keypoint {
    uint16_t x, y;       /* loc, in the keypoint's own octave coordinates */
    uint16_t octave;
    uint16_t index;
    float scale;
    float orientation;
    uint8_t filtered;    /* only descriptor-stage rejects stay in the output (sift.cpp:65-70) */
    uint8_t desc_len;    /* 128, or 0 when filtered */
    uint16_t reserved;
} sift_gpu_keypoint;

/* Result of one image.  Buffers are ctx-owned host memory (desc: pinned, filled straight from the device; kps: ordinary
 * heap memory assembled on the host), valid until the next run/destroy. */
typedef struct sift_gpu_result {
    int32_t status;              /* per-image SIFT_GPU_* code */
    uint32_t n;                  /* keypoints returned, in the reference's vector order */
    const sift_gpu_keypoint* kps;
    const float* desc;           /* n x 128 */
    uint32_t n_candidates;       /* extrema emitted (sift.cpp:373) */
    uint32_t n_survivors;        /* after _eliminateEdgeResponses + first cleanup */
    int32_t out_width, out_height; /* image size after the optional upsample */
} sift_gpu_result;

/* Stage times of the last run (ms, CUDA events on the ctx stream; host_* are wall clock). */
typedef struct sift_gpu_timings {
    float h2d_ms, pyramid_ms, extrema_ms, eliminate_ms, d2h_survivors_ms;
    float host_order_ms;
    float h2d_keypoints_ms, orientation_ms, descriptor_ms, d2h_results_ms;
    float device_total_ms; /* sum of device stages */
    float wall_ms;         /* whole run(), host clock */
    float span_ms;         /* CUDA-event time from the first enqueue to the last completion of every device pass
                              (includes the host order replay that sits between the two device halves) */
    uint64_t kernel_launches;
    uint64_t h2d_bytes;      /* frame bytes uploaded from host memory, as they travelled (a packed frame counts 1 byte per pixel) */
    uint32_t packed_images;  /* host f32 frames that were 8-bit valued and went up as bytes (lossless; see sift_gpu_run) */
    uint32_t reserved;
} sift_gpu_timings;

typedef struct sift_gpu_ctx sift_gpu_ctx;

/* Replaces: Sift::Sift(...) (sift.hpp:66-71). */
int sift_gpu_create(const sift_gpu_params* params, sift_gpu_ctx** out);
/* Replaces: Sift::calculate (sift.cpp:19-57) for n_images independent images.
 * Host f32 frames whose pixels are all exact 8-bit values (what vigra::importImage delivers, main.cpp:52-54) are packed to
 * bytes by host threads while earlier passes run, uploaded as bytes and widened on the device — a quarter of the PCIe traffic,
 * bit-identical results; any other frame makes its pass travel as f32.  Applies to passes of at least 8 images on contexts
 * with at least 8 host threads (SIFT_GPU_HOST_THREADS); SIFT_GPU_HOST_PACK=0 / 1 in the environment forces it off / on. */
int sift_gpu_run(sift_gpu_ctx* ctx, const sift_gpu_image* images, int n_images, sift_gpu_result* results);
void sift_gpu_destroy(sift_gpu_ctx* ctx);
const char* sift_gpu_last_error(const sift_gpu_ctx* ctx); /* ctx may be NULL: last create() error */
int sift_gpu_get_timings(const sift_gpu_ctx* ctx, sift_gpu_timings* out);
const char* sift_gpu_version(void);

/* ---- stage-level entry points used by the parity tests -------------------------- */
#define SIFT_GPU_KIND_GAUSS 0
#define SIFT_GPU_KIND_DOG 1
/* Pyramid level of image `image_idx` of the last device pass (sift.cpp:381-417). out: w*h floats. */
int sift_gpu_debug_get_level(sift_gpu_ctx* ctx, int image_idx, int octave, int elem, int kind, float* out,
                             int* width, int* height, float* scale);
/* alg::convolveWithGauss (algorithms.cpp:10-22) on a host image. */
int sift_gpu_debug_blur(sift_gpu_ctx* ctx, const float* src, int width, int height, float sigma, float* dst);
/* alg::reduceToNextLevel / increaseToNextLevel (algorithms.cpp:24-49). */
int sift_gpu_debug_reduce(sift_gpu_ctx* ctx, const float* src, int width, int height, float sigma, float* dst);
int sift_gpu_debug_increase(sift_gpu_ctx* ctx, const float* src, int width, int height, float sigma, float* dst);
/* _findScaleSpaceExtrema (sift.cpp:348-379) on caller-supplied DoG layers (d0,d1,d2 = below, current, above):
 * candidates of the middle layer in the reference's emission order.  Returns the count in *n. */
int sift_gpu_debug_extrema(sift_gpu_ctx* ctx, const float* d0, const float* d1, const float* d2, int width,
                           int height, uint16_t* xs, uint16_t* ys, uint32_t capacity, uint32_t* n);
/* _eliminateEdgeResponses (sift.cpp:288-346) flags for candidates of the middle layer. */
int sift_gpu_debug_eliminate(sift_gpu_ctx* ctx, const float* d0, const float* d1, const float* d2, int width,
                             int height, const uint16_t* xs, const uint16_t* ys, uint32_t n, uint8_t* filtered);
/* Candidates of image `image_idx` of the last pass, canonical order, with their elimination flag. */
int sift_gpu_debug_get_candidates(sift_gpu_ctx* ctx, int image_idx, uint16_t* xs, uint16_t* ys, uint16_t* octave,
                                  uint16_t* index, uint8_t* filtered, uint32_t capacity, uint32_t* n);
/* The host half of Sift::calculate between _eliminateEdgeResponses and _createDecriptors (sift.cpp:37-55: first cleanup sort,
 * u16 size, orientation-stage bounds test, second cleanup sort, descriptor-stage bounds test) on its own, exactly as
 * sift_gpu_run executes it between its two device stages, but without a device — so the ordering logic can be tested on
 * machines that have no GPU.  In: the unfiltered candidates of ONE image of `width` x `height` (input size, before the optional
 * upsample) in canonical order, `canon[i]` = their position in the candidate vector of `n_candidates` entries (ascending).
 * Out: the keypoint records in the reference's final vector order (orientation 0, it comes from the device), and the number
 * of points after the first cleanup.  SIFT_GPU_FLAG_STRICT / _ORDER_CANONICAL in params->flags act as in sift_gpu_run. */
int sift_gpu_debug_host_replay(const sift_gpu_params* params, int width, int height, uint32_t n_candidates, const uint32_t* canon,
                               const uint16_t* xs, const uint16_t* ys, const uint8_t* octave, const uint8_t* index,
                               uint32_t n_unfiltered, sift_gpu_keypoint* kps, uint32_t capacity, uint32_t* n_kps,
                               uint32_t* n_survivors);
/* The lossless f32 -> u8 packing sift_gpu_run applies to host frames ahead of their upload (see sift_gpu_run): packs `rows` rows
 * of `w` floats (row stride in bytes) into bytes (row pitch dst_pitch).  Returns 1 when every pixel is an exact 8-bit value (an
 * integer in [0, 255] whose float -> u8 -> float round trip is bit-identical: no fractions, negatives, -0.0f, NaN, Inf), else 0
 * and dst is unspecified.  Host code, no device needed. */
int sift_gpu_debug_pack_rows_u8(const float* src, size_t src_stride_bytes, int w, int rows, uint8_t* dst, size_t dst_pitch);
/* The std::sort(cmpByFilter) permutation the host replays (sift.cpp:37): order[i] = source index. */
int sift_gpu_debug_sort_order(const uint8_t* filtered, uint32_t n, uint32_t* order);
/* The same permutation restricted to the unfiltered elements (all the pipeline needs), computed by the sparse
 * introsort simulation the product path uses: unfiltered_order[i] = source index of the i-th element after the sort. */
int sift_gpu_debug_sort_order_fast(const uint8_t* filtered, uint32_t n, uint32_t* unfiltered_order, uint32_t* n_unfiltered);

#ifdef __cplusplus
}
#endif
#endif /* SIFT_GPU_H */
