// Command-line shim equivalent to the reference's main.cpp:21-95 without its decode/drawing libraries (PNM in, PPM overlay out):
//   sift [-i] image.{pgm,ppm} [-s sigma] [-k k] [-o octaves] [-d dogsPerEpoch] [-p 0|1] [-r 0|1] [--out file]
// Defaults as main.cpp:33-38 (sigma 1.6, k sqrt(2), octaves 4, dogsPerEpoch 3, subpixel 0, result 0).
// Image decode stays on the host: binary PGM/PPM, band 0 only, raw 0..255 (what vigra::importImage
// leaves in a scalar image, main.cpp:52-54 / SURVEY A.8).  Exceptions are printed and the exit code
// stays 0, as in main.cpp:90-94.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>

#include "../../include/sift/draw.hpp"
#include "../../include/sift/sift.hpp"

static bool read_pnm_band0(const std::string& path, sift::Image* out, sift::ColorImage* color) {
    std::ifstream f(path.c_str(), std::ios::binary);
    if (!f) return false;
    std::string magic;
    f >> magic;
    if (magic != "P5" && magic != "P6") return false;
    auto next_int = [&]() {
        int c;
        for (;;) {
            c = f.peek();
            if (c == '#') { std::string line; std::getline(f, line); }
            else if (isspace(c)) f.get();
            else break;
        }
        int v; f >> v; return v;
    };
    const int w = next_int(), h = next_int(), maxv = next_int();
    f.get();
    if (w < 1 || h < 1 || maxv > 255) return false;
    const int ch = magic == "P6" ? 3 : 1;
    std::string buf((size_t)w * h * ch, '\0');
    f.read(&buf[0], (std::streamsize)buf.size());
    if (!f) return false;
    *out = sift::Image(w, h);
    *color = sift::ColorImage(w, h);  // cv::imread(..., CV_LOAD_IMAGE_COLOR) of main.cpp:59: grey input is replicated
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const size_t at = ((size_t)y * w + x) * ch;
            (*out)(x, y) = (f32_t)(unsigned char)buf[at];
            for (int c = 0; c < 3; ++c) color->rgb[((size_t)y * w + x) * 3 + c] = (unsigned char)buf[at + (ch == 3 ? c : 0)];
        }
    return true;
}

int main(int argc, char** argv) {
    std::string img_file, out_file = "interstpoints.txt";  // the reference's file name (main.cpp:79)
    f32_t sigma = 1.6, k = std::sqrt(2);
    u16_t octaves = 4, dogsPerEpoch = 3;
    bool subpixel = false, result = false;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto val = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--help") {
            std::cout << "Options\n  --help\n  -i [ --img ] arg\n  -s [ --sigma ] arg (=1.6)\n  -k [ --k ] arg (=1.41421)\n"
                         "  -o [ --octaves ] arg (=4)\n  -d [ --dogsPerEpoch ] arg (=3)\n  -p [ --subpixel ] arg (=0)\n"
                         "  -r [ --result ] arg (=0)\n  --out arg (=interstpoints.txt)\n";
            return 1;
        } else if (a == "-i" || a == "--img") img_file = val();
        else if (a == "-s" || a == "--sigma") sigma = (f32_t)atof(val());
        else if (a == "-k" || a == "--k") k = (f32_t)atof(val());
        else if (a == "-o" || a == "--octaves") octaves = (u16_t)atoi(val());
        else if (a == "-d" || a == "--dogsPerEpoch") dogsPerEpoch = (u16_t)atoi(val());
        else if (a == "-p" || a == "--subpixel") subpixel = atoi(val()) != 0;
        else if (a == "-r" || a == "--result") result = atoi(val()) != 0;
        else if (a == "--out") out_file = val();
        else img_file = a;
    }
    try {
        sift::Image img;
        sift::ColorImage color;
        if (!read_pnm_band0(img_file, &img, &color)) throw std::runtime_error("cannot read '" + img_file + "' (binary PGM/PPM expected)");
        sift::Sift sift(dogsPerEpoch, octaves, sigma, k, subpixel);
        std::vector<sift::InterestPoint> interestPoints = sift.calculate(img);
        std::cout << interestPoints.size() << " interest points\n";
        // main.cpp:59-76: overlay on the colour image, written next to the input (PPM instead of OpenCV's PNG)
        sift::drawInterestPoints(color, interestPoints, sift.subpixel);
        if (!sift::writePPM(img_file + "_orientation.ppm", color)) std::cerr << "cannot write " << img_file << "_orientation.ppm" << std::endl;
        if (result) sift::writeResults(out_file, interestPoints);
    } catch (std::exception& ex) {
        std::cerr << ex.what() << std::endl;
    }
    return 0;
}
