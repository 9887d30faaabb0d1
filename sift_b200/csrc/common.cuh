// Shared declarations of the sm_100a kernels behind include/sift_gpu.h.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace siftgpu {

constexpr int kMaxOctaves = 12;
constexpr int kMaxGauss = 12;   // dogs_per_epoch + 1
constexpr int kDescLen = 128;
constexpr int kRegion = 8;      // reference sift.cpp:61,164
constexpr int kPitchAlign = 32; // row pitch of every device image is a multiple of 32 floats (128 B): TMA needs 16-B strides

// One emitted extremum (reference sift.cpp:373), canonical (octave, index, x, y) order.
struct Cand {
    uint16_t x, y;
    uint8_t octave, index, filtered, pad;
};
static_assert(sizeof(Cand) == 8, "Cand");

// Unfiltered candidate after _eliminateEdgeResponses; `canon` = position in the canonical list.
struct Surv {
    uint32_t canon;
    uint16_t x, y;
    uint8_t octave, index;
    uint16_t pad;
};
static_assert(sizeof(Surv) == 12, "Surv");

// Keypoint handed to the orientation/descriptor kernels, in the reference's vector order.
struct KeyIn {
    uint16_t x, y;
    uint8_t octave, index;
    uint8_t tgt;   // slot of the nearest-Gaussian level (_findNearestGaussian, sift.cpp:205-218)
    uint8_t pad;
};
static_assert(sizeof(KeyIn) == 8, "KeyIn");

// A pyramid level as the kernels see it: pixel (x, y) of image b lives at base[b*stride + y*pitch + x].
struct LevelRef {
    float* base;
    size_t stride; // elements between consecutive images of the batch
    int pitch;     // elements between consecutive rows
    int w, h;
};

// One (octave, middle DoG layer) of the extrema scan.
struct ScanLayer {
    const float* d0; // below, current, above (image 0 of the batch)
    const float* d1;
    const float* d2;
    size_t stride;       // elements between images
    int pitch;
    int w, h;
    int n_yw;            // ceil(h/32) mask words per column
    uint32_t mask_off;   // word offset of this layer inside one image's mask block
    uint32_t col_base;   // first global column id of this layer
    uint32_t tile_base;  // first CTA of this layer in the single mask launch (set_scan_tiles)
    uint32_t tiles_x;    // CTAs per row-block: ceil(w / 128)
    uint8_t octave, index;
};

// One Gaussian blur launch: dst = blur(src); optional dog = 128 + (dst - src); optional decimation
// (alg::reduceToNextLevel fused: only the pixels the nearest-neighbour resize picks are stored).
struct BlurArgs {
    const float* src;
    float* dst;      // may be null when only the DoG is wanted
    float* dog;      // or null
    size_t src_stride, dst_stride, dog_stride; // per-image strides (elements)
    int src_pitch, dst_pitch, dog_pitch;
    int w, h;
    const float* taps;      // 2r+1 taps, device memory
    const float* taps_host; // same taps, host memory (kernel-parameter copy for the streaming kernel)
    int r;
    const int* sel_x; // decimation: destination column of source column x, or -1 (device); null = no decimation
    const int* sel_y;
    const int* sel_x_host; // the same maps in host memory (the decimating streaming kernel derives its launch geometry from them), or null
    const int* sel_y_host;
    const CUtensorMap* map; // host pointer to two TMA descriptors of src (boxes map_box x 8 and x 1), or null
    int map_box;            // box width the descriptors were encoded with (each streaming kernel checks it is its own)
    int z0;      // first image of the batch this launch covers (the launch's blockIdx.z counts from here)
    int share;   // number of sibling launches expected to run concurrently (the grid is sized for 1/share of the GPU)
};

#define SIFT_CUDA_TRY(expr)                                         \
    do {                                                            \
        cudaError_t _e = (expr);                                    \
        if (_e != cudaSuccess) return siftgpu::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

// ---- launchers (each returns 0 or SIFT_GPU_E_CUDA; they count their launches in *launches) ----
int launch_blur(const BlurArgs& a, int batch, bool fma, cudaStream_t s, uint64_t* launches);
int stream_box_width(int r, bool decimate);   // TMA box width the streaming kernel for (radius r, decimating or not) needs, or 0 if none
int blur_prepare_device();     // per-device kernel attributes and occupancy figures; call after cudaSetDevice (sift_gpu_create does)
// blur_slide.cu
int launch_slide(const BlurArgs& a, int batch, bool fma, cudaStream_t s);   // -1: does not qualify
int launch_slide_dec(const BlurArgs& a, int batch, bool fma, cudaStream_t s);   // decimating launches; -1: does not qualify
int slide_dec_box_width(int r);
int slide_box_width(int r);
int slide_radius_for(int r);
int slide_prepare_device();
int stream_box_rows();         // rows of the multi-row TMA box (the other descriptor has 1-row boxes)
int max_generic_radius();      // largest radius the generic tile kernel can hold in shared memory
int launch_resize_nn(const float* src, size_t src_stride, int src_pitch, float* dst, size_t dst_stride, int dst_pitch, int dw,
                     int dh, const int* map_x, const int* map_y, int z0, int batch, cudaStream_t s, uint64_t* launches);
int launch_u8_to_f32(const uint8_t* src, size_t src_stride, int src_pitch, float* dst, size_t dst_stride, int dst_pitch, int w,
                     int h, int batch, cudaStream_t s, uint64_t* launches);

// fills tile_base / tiles_x of a layer list (call before the list is copied to the device)
void set_scan_tiles(ScanLayer* layers, int n_layers);
// pass_mask: second bit plane (candidates that also pass the cheap elimination tests), or null to leave every candidate unfiltered
int launch_extrema(const ScanLayer* layers_dev, const ScanLayer* layers_host, int n_layers, int total_cols,
                   uint32_t mask_words_per_image, uint32_t* mask, uint32_t* pass_mask, uint32_t* col_count, uint32_t* col_off,
                   Cand* cands, size_t cand_stride, uint32_t* n_cand, int batch, cudaStream_t s, uint64_t* launches);

constexpr int kCompactSlices = 1024;   // most slices per image of the survivor compaction (slice_scratch: batch x kCompactSlices words)
int launch_eliminate(const ScanLayer* layers_dev, int n_layers, Cand* cands, size_t cand_stride,
                     const uint32_t* n_cand, Surv* survivors, size_t surv_stride, uint32_t* n_surv, uint32_t* slice_scratch, int dogs_per_epoch,
                     int batch, cudaStream_t s, uint64_t* launches);

// weight tables: blur(level, 1.6f) top-left 16x16 of every (image, target slot)
int launch_weight_tables(const LevelRef* targets_dev, int n_targets, const float* taps16, int r16, float* tables,
                         bool fma, int batch, cudaStream_t s, uint64_t* launches);
// keys of all images concatenated; key_img[i] = image of key i; key_first[b] = first key of image b
int launch_orientation(const LevelRef* targets_dev, int n_targets, const KeyIn* keys, const uint32_t* key_img,
                       uint32_t n_keys, float* orientation, uint32_t* n_peaks, float* peaks, float2* grad_cache, cudaStream_t s,
                       uint64_t* launches);
// Spatial index of a pass's keys for the descriptor kernel: cells of 16 x 16 pixels per (image, target level); a key's
// overlapping predecessors can only sit in its own cell and the eight around it.  Scratch owned by the caller:
// count / offset / cursor: images * cells_per_image words each, cell_keys: one word per key.
struct KeyGrid {
    uint32_t* count;
    uint32_t* offset;
    uint32_t* cursor;
    uint32_t* cell_keys;
    int cw, ch;                 // cells per row / column of one target level
    uint32_t cells_per_image;   // n_targets * cw * ch
};
int launch_descriptors(const LevelRef* targets_dev, int n_targets, const float* tables, const KeyIn* keys,
                       const uint32_t* key_img, const uint32_t* key_first, uint32_t n_keys, const float* orientation,
                       float* desc, const float2* grad_cache, const KeyGrid* grid, int batch, cudaStream_t s, uint64_t* launches);
// grad_cache: n_keys x 256 (magnitude, orientation) pairs, written by launch_orientation and read by launch_descriptors when
// both run over the same key list; null = the descriptor kernel computes the gradients itself

}  // namespace siftgpu
