// Test driver for the C++ surface (include/sift/sift.hpp): exercises sift::Sift exactly as a caller of the
// reference would (main.cpp:56-57, :90-92) and dumps what it saw so tests/test_gpu_host_class.py can compare it
// with the oracle.  Not part of the product libraries.
//   host_selftest <out_dir> <dpe> <octaves> <subpixel> <w> <h> <raw_f32_file> [<w> <h> <raw_f32_file> ...]
// Writes, per image i:   calc_<i>.txt (formatResults of calculate()), img_<i>.f32 + img_<i>.dims (the caller's image
// after calculate(): overwritten with the 2x image when subpixel, sift.cpp:21), batch_<i>.txt (calculateBatch()).
// Then provokes the exception path (an image smaller than a blur radius) and prints what it caught.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "../../include/sift/sift.hpp"

static sift::Image load(const char* path, int w, int h) {
    std::vector<float> px((size_t)w * h);
    std::ifstream f(path, std::ios::binary);
    f.read(reinterpret_cast<char*>(px.data()), (std::streamsize)(px.size() * sizeof(float)));
    if (!f) { std::fprintf(stderr, "cannot read %s\n", path); std::exit(2); }
    return sift::Image(w, h, px.data());
}

int main(int argc, char** argv) {
    if (argc < 8 || (argc - 5) % 3 != 0) { std::fprintf(stderr, "usage\n"); return 2; }
    const std::string out = argv[1];
    const u16_t dpe = (u16_t)std::atoi(argv[2]), octaves = (u16_t)std::atoi(argv[3]);
    const bool subpixel = std::atoi(argv[4]) != 0;
    std::vector<sift::Image> imgs;
    for (int a = 5; a + 2 < argc; a += 3) imgs.push_back(load(argv[a + 2], std::atoi(argv[a]), std::atoi(argv[a + 1])));

    try {
        sift::Sift sift(dpe, octaves, 1.6, std::sqrt(2), subpixel);  // main.cpp:56
        if (sift.subpixel != subpixel) return 3;
        for (size_t i = 0; i < imgs.size(); ++i) {
            sift::Image img = imgs[i];
            std::vector<sift::InterestPoint> pts = sift.calculate(img);  // main.cpp:57
            sift::writeResults(out + "/calc_" + std::to_string(i) + ".txt", pts);
            std::ofstream(out + "/img_" + std::to_string(i) + ".dims") << img.width() << " " << img.height() << "\n";
            std::ofstream f(out + "/img_" + std::to_string(i) + ".f32", std::ios::binary);
            f.write(reinterpret_cast<const char*>(img.data()), (std::streamsize)(img.size() * sizeof(float)));
        }
        sift::Sift batch(dpe, octaves, 1.6, std::sqrt(2), subpixel);
        batch.setMaxBatch(2);  // three images -> two device passes
        std::vector<sift::Image> copy = imgs;
        std::vector<std::vector<sift::InterestPoint>> all = batch.calculateBatch(copy);
        for (size_t i = 0; i < all.size(); ++i) sift::writeResults(out + "/batch_" + std::to_string(i) + ".txt", all[i]);
        std::printf("batch_images %zu first_width_after %td\n", all.size(), copy[0].width());
    } catch (std::exception& ex) {
        std::printf("unexpected exception: %s\n", ex.what());
        return 4;
    }

    // main.cpp:43,90-92: a violated Vigra precondition reaches the caller as a std::exception
    try {
        sift::Sift sift(3, 4);
        sift::Image tiny(40, 40);
        sift.calculate(tiny);
        std::printf("no exception\n");
        return 5;
    } catch (sift::PreconditionViolation& ex) {
        std::printf("caught PreconditionViolation: %s\n", ex.what());
    } catch (std::exception& ex) {
        std::printf("caught other std::exception: %s\n", ex.what());
        return 6;
    }
    return 0;
}
