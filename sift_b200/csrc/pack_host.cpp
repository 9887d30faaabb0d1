// Host side of the frame upload: lossless f32 -> u8 packing of 8-bit-valued frames.
//
// The reference loads an image with vigra::importImage into MultiArray<2, f32_t> (main.cpp:52-54): every pixel that reaches
// Sift::calculate is an integer in [0, 255] stored as a float.  A host f32 frame costs 4 bytes per pixel on PCIe, which is what
// bounds the end-to-end throughput of one GPU (8.3 MB per 1080p frame).  When every pixel of a frame survives
// float -> u8 -> float bit for bit, the library uploads the bytes instead and widens them on the device (u8_to_f32_kernel, the
// path SIFT_GPU_DTYPE_U8 frames take anyway): a quarter of the traffic, identical results by construction.  A frame with any
// other value (fractions, negatives, > 255, -0.0f, NaN, Inf) makes its whole pass travel as f32, as before.
//
// Plain host C++ (SSE2 is part of the x86-64 baseline; other targets take the scalar loop), compiled by the host compiler.
#include <cstddef>
#include <cstdint>
#include <cstring>

#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define SIFT_PACK_AVX2 1
#endif

namespace {

inline bool pack_scalar(const float* src, uint8_t* dst, int n) {
    uint32_t bad = 0;
    for (int i = 0; i < n; ++i) {
        const float v = src[i];
        const int q = (v >= 0.0f && v <= 255.0f) ? (int)v : 0;   // NaN fails both comparisons
        const float back = (float)q;
        uint32_t a, b;
        std::memcpy(&a, &v, 4);
        std::memcpy(&b, &back, 4);
        bad |= a ^ b;                                             // bit pattern, so -0.0f does not pass as 0
        dst[i] = (uint8_t)q;
    }
    return bad == 0;
}

#if defined(SIFT_PACK_AVX2)
// 32 pixels per step on machines that have AVX2 (checked at run time); same test, same bytes.
__attribute__((target("avx2"))) bool pack_row_avx2(const float* s, uint8_t* d, int w, int* done) {
    __m256i diff = _mm256_setzero_si256(), range = _mm256_setzero_si256();
    const __m256i order = _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7);
    const bool aligned = (reinterpret_cast<uintptr_t>(d) & 31u) == 0;
    int x = 0;
    for (; x + 32 <= w; x += 32) {
        // the loop is bound by how many cache misses one core keeps in flight: asking for the two lines 4 KB ahead lifts a
        // thread from about 7.5 to 10 GB/s of frame data (prefetches never fault, so running past the row is harmless)
        _mm_prefetch(reinterpret_cast<const char*>(s + x + 1024), _MM_HINT_T0);
        _mm_prefetch(reinterpret_cast<const char*>(s + x + 1040), _MM_HINT_T0);
        const __m256 f0 = _mm256_loadu_ps(s + x), f1 = _mm256_loadu_ps(s + x + 8), f2 = _mm256_loadu_ps(s + x + 16), f3 = _mm256_loadu_ps(s + x + 24);
        const __m256i i0 = _mm256_cvttps_epi32(f0), i1 = _mm256_cvttps_epi32(f1), i2 = _mm256_cvttps_epi32(f2), i3 = _mm256_cvttps_epi32(f3);
        diff = _mm256_or_si256(diff, _mm256_xor_si256(_mm256_castps_si256(_mm256_cvtepi32_ps(i0)), _mm256_castps_si256(f0)));
        diff = _mm256_or_si256(diff, _mm256_xor_si256(_mm256_castps_si256(_mm256_cvtepi32_ps(i1)), _mm256_castps_si256(f1)));
        diff = _mm256_or_si256(diff, _mm256_xor_si256(_mm256_castps_si256(_mm256_cvtepi32_ps(i2)), _mm256_castps_si256(f2)));
        diff = _mm256_or_si256(diff, _mm256_xor_si256(_mm256_castps_si256(_mm256_cvtepi32_ps(i3)), _mm256_castps_si256(f3)));
        range = _mm256_or_si256(range, _mm256_or_si256(_mm256_or_si256(i0, i1), _mm256_or_si256(i2, i3)));
        // the packs work per 128-bit lane: lane 0 holds the low halves of i0..i3, lane 1 the high halves; `order` undoes that
        const __m256i b = _mm256_permutevar8x32_epi32(_mm256_packus_epi16(_mm256_packs_epi32(i0, i1), _mm256_packs_epi32(i2, i3)), order);
        if (aligned) _mm256_stream_si256(reinterpret_cast<__m256i*>(d + x), b);
        else _mm256_storeu_si256(reinterpret_cast<__m256i*>(d + x), b);
    }
    *done = x;
    // exact round trip of every pixel (bit pattern: -0.0f fails) and every integer inside [0, 255] (NaN/Inf become INT_MIN)
    const __m256i bad = _mm256_or_si256(diff, _mm256_andnot_si256(_mm256_set1_epi32(255), range));
    return _mm256_testz_si256(bad, bad) != 0;
}
#endif

}  // namespace

extern "C" {

// Packs `rows` rows of `w` floats (row stride src_stride_bytes) into bytes (row pitch dst_pitch).  Returns 1 when every pixel
// was an exact 8-bit value (dst is then the frame), 0 otherwise (dst contents unspecified).
int sift_gpu_debug_pack_rows_u8(const float* src, size_t src_stride_bytes, int w, int rows, uint8_t* dst, size_t dst_pitch) {
    bool ok = true;
#if defined(SIFT_PACK_AVX2)
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
#endif
    for (int y = 0; y < rows && ok; ++y) {
        const float* s = reinterpret_cast<const float*>(reinterpret_cast<const char*>(src) + (size_t)y * src_stride_bytes);
        uint8_t* d = dst + (size_t)y * dst_pitch;
        int x = 0;
#if defined(SIFT_PACK_AVX2)
        if (have_avx2) {
            ok = pack_row_avx2(s, d, w, &x);
            if (ok && x < w) ok = pack_scalar(s + x, d + x, w - x);
            continue;
        }
#endif
#if defined(__SSE2__)
        __m128i diff = _mm_setzero_si128();
        const bool aligned = (reinterpret_cast<uintptr_t>(d) & 15u) == 0;
        for (; x + 16 <= w; x += 16) {
            _mm_prefetch(reinterpret_cast<const char*>(s + x + 1024), _MM_HINT_T0);
            const __m128 f0 = _mm_loadu_ps(s + x), f1 = _mm_loadu_ps(s + x + 4), f2 = _mm_loadu_ps(s + x + 8), f3 = _mm_loadu_ps(s + x + 12);
            const __m128i i0 = _mm_cvttps_epi32(f0), i1 = _mm_cvttps_epi32(f1), i2 = _mm_cvttps_epi32(f2), i3 = _mm_cvttps_epi32(f3);
            // saturating packs: anything outside [0, 255] (NaN/Inf convert to INT_MIN) comes back as a different float below
            const __m128i b = _mm_packus_epi16(_mm_packs_epi32(i0, i1), _mm_packs_epi32(i2, i3));
            const __m128i z = _mm_setzero_si128();
            const __m128i w0 = _mm_unpacklo_epi8(b, z), w1 = _mm_unpackhi_epi8(b, z);
            const __m128 r0 = _mm_cvtepi32_ps(_mm_unpacklo_epi16(w0, z)), r1 = _mm_cvtepi32_ps(_mm_unpackhi_epi16(w0, z));
            const __m128 r2 = _mm_cvtepi32_ps(_mm_unpacklo_epi16(w1, z)), r3 = _mm_cvtepi32_ps(_mm_unpackhi_epi16(w1, z));
            diff = _mm_or_si128(diff, _mm_xor_si128(_mm_castps_si128(r0), _mm_castps_si128(f0)));
            diff = _mm_or_si128(diff, _mm_xor_si128(_mm_castps_si128(r1), _mm_castps_si128(f1)));
            diff = _mm_or_si128(diff, _mm_xor_si128(_mm_castps_si128(r2), _mm_castps_si128(f2)));
            diff = _mm_or_si128(diff, _mm_xor_si128(_mm_castps_si128(r3), _mm_castps_si128(f3)));
            if (aligned) _mm_stream_si128(reinterpret_cast<__m128i*>(d + x), b);   // the staging buffer is read by DMA only
            else _mm_storeu_si128(reinterpret_cast<__m128i*>(d + x), b);
        }
        ok = _mm_movemask_epi8(_mm_cmpeq_epi32(diff, _mm_setzero_si128())) == 0xffff;
#endif
        if (ok && x < w) ok = pack_scalar(s + x, d + x, w - x);
    }
#if defined(__SSE2__)
    _mm_sfence();
#endif
    return ok ? 1 : 0;
}

}  // extern "C"
