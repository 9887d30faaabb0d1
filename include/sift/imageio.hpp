// Image files for the command-line shim (reference main.cpp:52-54 and :59, :76): what vigra::importImage leaves in the
// reference's scalar float image is band 0 of the file, raw 0..255 — the R channel of a colour image, the grey values of
// a grey one — and cv::imread(..., CV_LOAD_IMAGE_COLOR) gives the three-channel copy the overlay is drawn on.
// Decode stays on the host side of the C ABI (north_star):
//   * PNG: 8- or 16-bit grey / RGB / with alpha, 8-bit palette, non-interlaced — chunk parser and the five scanline filters
//     here, inflate by zlib; 16-bit samples reach the float image unscaled (0..65535), as importImage leaves them;
//   * JPEG: the toolkit's nvJPEG (the image has no libjpeg headers), decoded to interleaved RGB and copied back;
//   * binary PGM / PPM.
// The overlay goes out as PNG (`<image>_orientation.png`, main.cpp:76) or PPM.
#ifndef SIFT_IMAGEIO_HPP
#define SIFT_IMAGEIO_HPP

#include <string>

#include "core.hpp"
#include "draw.hpp"

namespace sift {

// Reads `path` (format by its first bytes, not its name).  On success fills `band0` (w x h floats, 0..255) and `color`.
// On failure returns false and says why in *error.
bool readImage(const std::string& path, Image* band0, ColorImage* color, std::string* error);

// 8-bit RGB PNG, no interlace (what cv::imwrite produces for main.cpp:76, up to the compressor's choices).
bool writePNG(const std::string& path, const ColorImage& image);

}  // namespace sift

extern "C" {
// C entry points for bindings and tests.  rgb = width*height*3 bytes, caller-allocated (call with rgb = NULL to get the size).
// Returns 0, or -1 with a message in err (if err_len > 0).
int sift_host_read_image(const char* path, unsigned char* rgb, int* width, int* height, char* err, int err_len);
// band 0 as readImage() delivers it to sift::Sift::calculate (width*height floats; band0 = NULL to get the size)
int sift_host_read_band0(const char* path, float* band0, int* width, int* height, char* err, int err_len);
int sift_host_write_png(const char* path, const unsigned char* rgb, int width, int height);
}
#endif  // SIFT_IMAGEIO_HPP
