// Command-line shim equivalent to the reference's main.cpp:21-95 without Boost / Vigra / OpenCV:
//   sift [-i] image.{jpg,png,pgm,ppm} [-s sigma] [-k k] [-o octaves] [-d dogsPerEpoch] [-p 0|1] [-r 0|1] [--out file]
// Defaults as main.cpp:33-38 (sigma 1.6, k sqrt(2), octaves 4, dogsPerEpoch 3, subpixel 0, result 0), so
// `./sift example/parrot.jpg -r 1` works as the README says.  Image decode stays on the host side of the C ABI
// (include/sift/imageio.hpp: PNG, JPEG, binary PNM; band 0, raw 0..255 = what vigra::importImage leaves in a scalar
// image, main.cpp:52-54 / SURVEY A.8); the overlay is written as `<image>_orientation.png` (main.cpp:76).
// Exceptions are printed and the exit code stays 0, as in main.cpp:90-94.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>

#include "../../include/sift/draw.hpp"
#include "../../include/sift/imageio.hpp"
#include "../../include/sift/sift.hpp"

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
        std::string why;
        if (!sift::readImage(img_file, &img, &color, &why)) throw std::runtime_error(why);
        sift::Sift sift(dogsPerEpoch, octaves, sigma, k, subpixel);
        std::vector<sift::InterestPoint> interestPoints = sift.calculate(img);
        std::cout << interestPoints.size() << " interest points\n";
        // main.cpp:59-76: overlay on the colour image, written next to the input
        sift::drawInterestPoints(color, interestPoints, sift.subpixel);
        if (!sift::writePNG(img_file + "_orientation.png", color)) std::cerr << "cannot write " << img_file << "_orientation.png" << std::endl;
        if (result) sift::writeResults(out_file, interestPoints);
    } catch (std::exception& ex) {
        std::cerr << ex.what() << std::endl;
    }
    return 0;
}
