// Keypoint overlay of the reference's command-line tool (main.cpp:59-75): every interest point is drawn as a
// rotated square of side scale*10 (truncated to int, as cv::Size does) at its position in input-image pixels,
// turned by its orientation, outlined in cv::Scalar(255, 0, 0) (blue in OpenCV's BGR order).  Host code; the
// corner formula is cv::RotatedRect::points(), the rasteriser is a plain 8-connected Bresenham line.
#ifndef SIFT_DRAW_HPP
#define SIFT_DRAW_HPP

#include <string>
#include <vector>

#include "interestpoint.hpp"

namespace sift {

struct ColorImage {
    int width = 0, height = 0;
    std::vector<unsigned char> rgb;  // interleaved R, G, B; row-major
    ColorImage() = default;
    ColorImage(int w, int h) : width(w), height(h), rgb((size_t)w * (size_t)h * 3, 0) {}
};

// main.cpp:60-74.  `subpixel` is Sift::subpixel (coordinates of a 2x-upsampled run are halved).
void drawInterestPoints(ColorImage& image, const std::vector<InterestPoint>& points, bool subpixel);

// binary PPM (P6); the reference writes "<image>_orientation.png" through OpenCV (main.cpp:76)
bool writePPM(const std::string& path, const ColorImage& image);

}  // namespace sift

extern "C" {
// C entry point for bindings and tests: pts = n x {x, y, octave, scale, orientation} floats (loc in octave pixels).
void sift_host_draw_points(unsigned char* rgb, int width, int height, const float* pts, int n, int subpixel);
}
#endif  // SIFT_DRAW_HPP
