// Test helper (CPU): the frame packer of the upload path (sift_b200/csrc/pack_host.cpp) on ragged widths and row counts with
// exact-size heap blocks, so that a read or write outside a row trips AddressSanitizer (the test builds it with
// -fsanitize=address,undefined where the toolchain has it).  Exit code 0 = every case agrees.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cstdint>
extern "C" int sift_gpu_debug_pack_rows_u8(const float* src, size_t src_stride_bytes, int w, int rows, uint8_t* dst, size_t dst_pitch);
int main() {
    int cases = 0;
    for (int w : {1, 2, 15, 16, 17, 31, 32, 33, 63, 64, 65, 100, 211, 1920})
        for (int rows : {1, 3, 128}) {
            // exact-size heap blocks: any read or write outside them trips the sanitizer
            float* src = (float*)malloc(sizeof(float) * (size_t)w * rows);
            uint8_t* dst = (uint8_t*)malloc((size_t)w * rows);
            for (int i = 0; i < w * rows; ++i) src[i] = (float)((i * 37) & 255);
            if (!sift_gpu_debug_pack_rows_u8(src, sizeof(float) * w, w, rows, dst, w)) { printf("not packable w=%d\n", w); return 1; }
            for (int i = 0; i < w * rows; ++i) if (dst[i] != (uint8_t)((i * 37) & 255)) { printf("wrong byte w=%d i=%d\n", w, i); return 1; }
            src[w * rows - 1] = 0.5f;
            if (sift_gpu_debug_pack_rows_u8(src, sizeof(float) * w, w, rows, dst, w)) { printf("fraction accepted w=%d\n", w); return 1; }
            free(src); free(dst); ++cases;
        }
    printf("%d pack edge cases ok\n", cases);
}
