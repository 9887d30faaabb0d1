// include/sift/imageio.hpp: PNG (zlib), JPEG (nvJPEG) and binary PNM readers for the command-line shim, PNG writer for the
// overlay.  Host-side code: nothing here is on the measured path.
#include "../../include/sift/imageio.hpp"

#include <cuda_runtime.h>
#include <nvjpeg.h>
#include <zlib.h>

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <vector>

namespace sift {
namespace {

using Bytes = std::vector<unsigned char>;

bool slurp(const std::string& path, Bytes* out) {
    std::ifstream f(path.c_str(), std::ios::binary);
    if (!f) return false;
    out->assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    return true;
}

// band0_16: the file's own 16-bit samples of band 0 (what vigra::importImage copies into the float image unscaled), or empty
void fill(const unsigned char* rgb, const std::vector<uint16_t>& band0_16, int w, int h, Image* band0, ColorImage* color) {
    *band0 = Image(w, h);
    *color = ColorImage(w, h);
    std::memcpy(color->rgb.data(), rgb, (size_t)w * (size_t)h * 3);
    f32_t* px = band0->data();
    const size_t n = (size_t)w * (size_t)h;
    if (band0_16.size() == n)
        for (size_t i = 0; i < n; ++i) px[i] = (f32_t)band0_16[i];
    else
        for (size_t i = 0; i < n; ++i) px[i] = (f32_t)rgb[3 * i];   // band 0
}

// ---- binary PGM / PPM ------------------------------------------------------------------------------------------------
bool decode_pnm(const Bytes& b, Bytes* rgb, int* w, int* h, std::string* err) {
    size_t at = 2;
    auto next_int = [&]() -> long {
        for (;;) {
            if (at >= b.size()) return -1;
            if (b[at] == '#') { while (at < b.size() && b[at] != '\n') ++at; }
            else if (isspace(b[at])) ++at;
            else break;
        }
        long v = 0;
        bool any = false;
        while (at < b.size() && isdigit(b[at])) { v = v * 10 + (b[at++] - '0'); any = true; }
        return any ? v : -1;
    };
    const int ch = b[1] == '6' ? 3 : 1;
    const long ww = next_int(), hh = next_int(), maxv = next_int();
    ++at;   // the single whitespace byte after maxval
    if (ww < 1 || hh < 1 || maxv < 1 || maxv > 255 || at + (size_t)ww * (size_t)hh * ch > b.size()) { *err = "truncated or unsupported PNM (8-bit binary P5/P6 only)"; return false; }
    *w = (int)ww; *h = (int)hh;
    rgb->resize((size_t)ww * (size_t)hh * 3);
    for (size_t i = 0, n = (size_t)ww * (size_t)hh; i < n; ++i)
        for (int c = 0; c < 3; ++c) (*rgb)[3 * i + c] = b[at + i * ch + (ch == 3 ? c : 0)];
    return true;
}

// ---- PNG ------------------------------------------------------------------------------------------------------------
uint32_t be32(const unsigned char* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

bool decode_png(const Bytes& b, Bytes* rgb, std::vector<uint16_t>* band0_16, int* w, int* h, std::string* err) {
    size_t at = 8;
    uint32_t width = 0, height = 0;
    int depth = 0, ctype = -1, interlace = 0;
    Bytes idat, plte;
    for (;;) {
        if (at + 12 > b.size()) { *err = "truncated PNG"; return false; }
        const uint32_t len = be32(&b[at]);
        const char* type = reinterpret_cast<const char*>(&b[at + 4]);
        const unsigned char* data = &b[at + 8];
        if (at + 12 + (size_t)len > b.size()) { *err = "truncated PNG chunk"; return false; }
        if (!std::memcmp(type, "IHDR", 4) && len >= 13) {
            width = be32(data); height = be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12];
        } else if (!std::memcmp(type, "PLTE", 4)) plte.assign(data, data + len);
        else if (!std::memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
        else if (!std::memcmp(type, "IEND", 4)) break;
        at += 12 + (size_t)len;
    }
    const int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!width || !height || width > 65535 || height > 65535 || !channels) { *err = "unsupported PNG header"; return false; }
    if ((depth != 8 && depth != 16) || (depth == 16 && ctype == 3) || interlace) { *err = "unsupported PNG (8 or 16 bits per sample, non-interlaced only)"; return false; }
    if (ctype == 3 && plte.size() < 3) { *err = "palette PNG without PLTE"; return false; }
    const int bps = depth / 8;                       // bytes per sample; 16-bit samples are big-endian
    const size_t stride = (size_t)width * channels * bps;
    Bytes raw((stride + 1) * height);
    uLongf raw_len = (uLongf)raw.size();
    if (uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size()) != Z_OK || raw_len != raw.size()) { *err = "PNG inflate failed"; return false; }
    // undo the scanline filters in place (bpp = bytes per complete pixel)
    const int bpp = channels * bps;
    Bytes prev(stride, 0);
    rgb->resize((size_t)width * height * 3);
    if (bps == 2 && band0_16) band0_16->resize((size_t)width * height);
    for (uint32_t y = 0; y < height; ++y) {
        unsigned char* line = &raw[(stride + 1) * y + 1];
        const int ft = raw[(stride + 1) * y];
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= (size_t)bpp ? line[i - bpp] : 0, up = prev[i], c = i >= (size_t)bpp ? prev[i - bpp] : 0;
            int v = line[i];
            switch (ft) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += up; break;
                case 3: v += (a + up) / 2; break;
                case 4: v += paeth(a, up, c); break;
                default: *err = "bad PNG filter type"; return false;
            }
            line[i] = (unsigned char)v;
        }
        std::memcpy(prev.data(), line, stride);
        unsigned char* out = &(*rgb)[(size_t)y * width * 3];
        for (uint32_t x = 0; x < width; ++x) {
            const unsigned char* p = line + (size_t)x * channels * bps;   // 16-bit: the high byte is the 8-bit view (cv::imread)
            if (bps == 2 && band0_16) (*band0_16)[(size_t)y * width + x] = (uint16_t)((p[0] << 8) | p[1]);
            if (ctype == 0 || ctype == 4) { out[3 * x] = out[3 * x + 1] = out[3 * x + 2] = p[0]; }
            else if (ctype == 3) {
                const size_t e = (size_t)p[0] * 3;
                for (int k = 0; k < 3; ++k) out[3 * x + k] = e + 2 < plte.size() ? plte[e + k] : 0;
            } else { out[3 * x] = p[0]; out[3 * x + 1] = p[bps]; out[3 * x + 2] = p[2 * bps]; }
        }
    }
    *w = (int)width; *h = (int)height;
    return true;
}

// ---- JPEG (nvJPEG: decoded on the device, copied back) -----------------------------------------------------------------
bool decode_jpeg(const Bytes& b, Bytes* rgb, int* w, int* h, std::string* err) {
    nvjpegHandle_t handle = nullptr;
    nvjpegJpegState_t state = nullptr;
    unsigned char* dev = nullptr;
    bool ok = false;
    do {
        if (nvjpegCreateSimple(&handle) != NVJPEG_STATUS_SUCCESS || nvjpegJpegStateCreate(handle, &state) != NVJPEG_STATUS_SUCCESS) { *err = "nvJPEG set-up failed (is a CUDA device present?)"; break; }
        int ncomp = 0, ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
        nvjpegChromaSubsampling_t sub;
        if (nvjpegGetImageInfo(handle, b.data(), b.size(), &ncomp, &sub, ws, hs) != NVJPEG_STATUS_SUCCESS || ws[0] < 1 || hs[0] < 1) { *err = "not a decodable JPEG"; break; }
        const size_t pitch = (size_t)ws[0] * 3, bytes = pitch * (size_t)hs[0];
        if (cudaMalloc(&dev, bytes) != cudaSuccess) { *err = "cudaMalloc failed"; break; }
        nvjpegImage_t dst{};
        dst.channel[0] = dev;
        dst.pitch[0] = pitch;
        if (nvjpegDecode(handle, state, b.data(), b.size(), NVJPEG_OUTPUT_RGBI, &dst, nullptr) != NVJPEG_STATUS_SUCCESS) { *err = "nvjpegDecode failed"; break; }
        rgb->resize(bytes);
        if (cudaMemcpy(rgb->data(), dev, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) { *err = "copy of the decoded image failed"; break; }
        *w = ws[0]; *h = hs[0];
        ok = true;
    } while (false);
    cudaFree(dev);
    if (state) nvjpegJpegStateDestroy(state);
    if (handle) nvjpegDestroy(handle);
    return ok;
}

bool decode_any(const std::string& path, Bytes* rgb, std::vector<uint16_t>* band0_16, int* w, int* h, std::string* err) {
    if (band0_16) band0_16->clear();
    Bytes b;
    if (!slurp(path, &b)) { *err = "cannot open '" + path + "'"; return false; }
    static const unsigned char png_magic[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (b.size() >= 8 && !std::memcmp(b.data(), png_magic, 8)) return decode_png(b, rgb, band0_16, w, h, err);
    if (b.size() >= 3 && b[0] == 0xff && b[1] == 0xd8) return decode_jpeg(b, rgb, w, h, err);
    if (b.size() >= 7 && b[0] == 'P' && (b[1] == '5' || b[1] == '6')) return decode_pnm(b, rgb, w, h, err);
    *err = "'" + path + "': not a PNG, JPEG or binary PGM/PPM file";
    return false;
}

void put_chunk(std::ofstream& f, const char* type, const unsigned char* data, size_t len) {
    unsigned char hdr[8] = {(unsigned char)(len >> 24), (unsigned char)(len >> 16), (unsigned char)(len >> 8), (unsigned char)len,
                            (unsigned char)type[0], (unsigned char)type[1], (unsigned char)type[2], (unsigned char)type[3]};
    f.write(reinterpret_cast<const char*>(hdr), 8);
    if (len) f.write(reinterpret_cast<const char*>(data), (std::streamsize)len);
    uLong crc = crc32(0L, hdr + 4, 4);
    if (len) crc = crc32(crc, data, (uInt)len);
    const unsigned char c[4] = {(unsigned char)(crc >> 24), (unsigned char)(crc >> 16), (unsigned char)(crc >> 8), (unsigned char)crc};
    f.write(reinterpret_cast<const char*>(c), 4);
}

}  // namespace

bool readImage(const std::string& path, Image* band0, ColorImage* color, std::string* error) {
    Bytes rgb;
    std::vector<uint16_t> b16;
    int w = 0, h = 0;
    std::string err;
    if (!decode_any(path, &rgb, &b16, &w, &h, &err)) {
        if (error) *error = err;
        return false;
    }
    fill(rgb.data(), b16, w, h, band0, color);
    return true;
}

bool writePNG(const std::string& path, const ColorImage& image) {
    if (image.width < 1 || image.height < 1) return false;
    const size_t stride = (size_t)image.width * 3;
    Bytes raw((stride + 1) * (size_t)image.height);
    for (int y = 0; y < image.height; ++y) {
        raw[(stride + 1) * y] = 0;   // filter type None
        std::memcpy(&raw[(stride + 1) * y + 1], &image.rgb[stride * y], stride);
    }
    uLongf clen = compressBound((uLong)raw.size());
    Bytes z(clen);
    if (compress2(z.data(), &clen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return false;
    std::ofstream f(path.c_str(), std::ios::binary);
    if (!f) return false;
    static const unsigned char magic[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    f.write(reinterpret_cast<const char*>(magic), 8);
    const uint32_t w = (uint32_t)image.width, h = (uint32_t)image.height;
    const unsigned char ihdr[13] = {(unsigned char)(w >> 24), (unsigned char)(w >> 16), (unsigned char)(w >> 8), (unsigned char)w,
                                    (unsigned char)(h >> 24), (unsigned char)(h >> 16), (unsigned char)(h >> 8), (unsigned char)h, 8, 2, 0, 0, 0};
    put_chunk(f, "IHDR", ihdr, 13);
    put_chunk(f, "IDAT", z.data(), clen);
    put_chunk(f, "IEND", nullptr, 0);
    return (bool)f;
}

}  // namespace sift

extern "C" {
int sift_host_read_image(const char* path, unsigned char* rgb, int* width, int* height, char* err, int err_len) {
    sift::Bytes px;
    int w = 0, h = 0;
    std::string why;
    if (!path || !sift::decode_any(path, &px, nullptr, &w, &h, &why)) {
        if (err && err_len > 0) { std::strncpy(err, why.c_str(), (size_t)err_len - 1); err[err_len - 1] = 0; }
        return -1;
    }
    if (width) *width = w;
    if (height) *height = h;
    if (rgb) std::memcpy(rgb, px.data(), px.size());
    return 0;
}
int sift_host_read_band0(const char* path, float* band0, int* width, int* height, char* err, int err_len) {
    sift::Image img;
    sift::ColorImage color;
    std::string why;
    if (!path || !sift::readImage(path, &img, &color, &why)) {
        if (err && err_len > 0) { std::strncpy(err, why.c_str(), (size_t)err_len - 1); err[err_len - 1] = 0; }
        return -1;
    }
    if (width) *width = (int)img.width();
    if (height) *height = (int)img.height();
    if (band0) std::memcpy(band0, img.data(), sizeof(float) * img.size());
    return 0;
}
int sift_host_write_png(const char* path, const unsigned char* rgb, int width, int height) {
    if (!path || !rgb || width < 1 || height < 1) return -1;
    sift::ColorImage img(width, height);
    std::memcpy(img.rgb.data(), rgb, img.rgb.size());
    return sift::writePNG(path, img) ? 0 : -1;
}
}
