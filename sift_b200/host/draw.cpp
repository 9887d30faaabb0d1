// Keypoint overlay (reference main.cpp:59-76), see include/sift/draw.hpp.
#include "../../include/sift/draw.hpp"

#include <cmath>
#include <cstdlib>
#include <fstream>

namespace sift {

namespace {

struct Pt { float x, y; };

void put(ColorImage& im, int x, int y) {
    if (x < 0 || y < 0 || x >= im.width || y >= im.height) return;
    unsigned char* p = &im.rgb[((size_t)y * (size_t)im.width + (size_t)x) * 3];
    p[0] = 0; p[1] = 0; p[2] = 255;  // cv::Scalar(255, 0, 0) is B, G, R
}

// cv::line(img, a, b, color): integer end points (cv::Point from cv::Point2f rounds), thickness 1, 8-connected
void line(ColorImage& im, Pt a, Pt b) {
    long x0 = std::lround(a.x), y0 = std::lround(a.y);
    const long x1 = std::lround(b.x), y1 = std::lround(b.y);
    // a segment that misses the image entirely is skipped; long off-image runs are walked but not stored
    const long dx = std::labs(x1 - x0), sx = x0 < x1 ? 1 : -1;
    const long dy = -std::labs(y1 - y0), sy = y0 < y1 ? 1 : -1;
    if (dx > 1000000 || -dy > 1000000) return;
    long err = dx + dy;
    for (;;) {
        put(im, (int)x0, (int)y0);
        if (x0 == x1 && y0 == y1) break;
        const long e2 = 2 * err;
        if (e2 >= dy) { err += dy; x0 += sx; }
        if (e2 <= dx) { err += dx; y0 += sy; }
    }
}

void draw_one(ColorImage& image, float px, float py, float octave, float scale, float orientation, bool subpixel) {
    const unsigned divisor = subpixel ? 2 : 1;
    // u16_t x = (p.loc.x * std::pow(2, p.octave)) / subpixel_divisor;  (main.cpp:62-63)
    const u16_t x = (u16_t)((double)px * std::pow(2.0, (double)octave) / divisor);
    const u16_t y = (u16_t)((double)py * std::pow(2.0, (double)octave) / divisor);
    const int side = (int)(scale * 10);  // cv::Size(p.scale * 10, p.scale * 10) holds ints
    // cv::RotatedRect::points()
    const double ang = (double)orientation * 3.14159265358979323846 / 180.0;
    const float b = (float)std::cos(ang) * 0.5f, a = (float)std::sin(ang) * 0.5f;
    const float w = (float)side, h = (float)side;
    Pt pt[4];
    pt[0].x = x - a * h - b * w;
    pt[0].y = y + b * h - a * w;
    pt[1].x = x + a * h - b * w;
    pt[1].y = y - b * h - a * w;
    pt[2].x = 2 * x - pt[0].x;
    pt[2].y = 2 * y - pt[0].y;
    pt[3].x = 2 * x - pt[1].x;
    pt[3].y = 2 * y - pt[1].y;
    if (!(std::isfinite(pt[0].x) && std::isfinite(pt[0].y) && std::isfinite(pt[1].x) && std::isfinite(pt[1].y))) return;
    line(image, pt[0], pt[1]);  // main.cpp:70-73
    line(image, pt[0], pt[3]);
    line(image, pt[2], pt[3]);
    line(image, pt[1], pt[2]);
}

}  // namespace

void drawInterestPoints(ColorImage& image, const std::vector<InterestPoint>& points, bool subpixel) {
    for (const InterestPoint& p : points)
        draw_one(image, (float)p.loc.x, (float)p.loc.y, (float)p.octave, p.scale, p.orientation, subpixel);
}

bool writePPM(const std::string& path, const ColorImage& image) {
    std::ofstream f(path.c_str(), std::ios::binary);
    if (!f) return false;
    f << "P6\n" << image.width << " " << image.height << "\n255\n";
    f.write(reinterpret_cast<const char*>(image.rgb.data()), (std::streamsize)image.rgb.size());
    return (bool)f;
}

}  // namespace sift

extern "C" void sift_host_draw_points(unsigned char* rgb, int width, int height, const float* pts, int n, int subpixel) {
    sift::ColorImage im;
    im.width = width;
    im.height = height;
    im.rgb.assign(rgb, rgb + (size_t)width * (size_t)height * 3);
    for (int i = 0; i < n; ++i)
        sift::draw_one(im, pts[5 * i], pts[5 * i + 1], pts[5 * i + 2], pts[5 * i + 3], pts[5 * i + 4], subpixel != 0);
    std::copy(im.rgb.begin(), im.rgb.end(), rgb);
}
