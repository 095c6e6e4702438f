// oracle/cvshim/opencv2/highgui.hpp — cv::imread stand-in for the reference's ImageReader (TEST INFRASTRUCTURE ONLY):
// binary PGM (P5, maxval 255) only — the synthetic EuRoC-shaped test datasets are written in that format, so no image
// codec is needed to run the reference's own directory listing / timestamp / synchronisation code.
#ifndef VSO_CVSHIM_HIGHGUI_HPP
#define VSO_CVSHIM_HIGHGUI_HPP
#include <cstdio>
#include <string>
#include "core.hpp"
#define CV_LOAD_IMAGE_GRAYSCALE 0
namespace cv {
inline Mat imread(const std::string& name, int /*flags*/) {
    Mat m;
    FILE* f = fopen(name.c_str(), "rb");
    if (!f) return m;
    int w = 0, h = 0, maxv = 0;
    char magic[3] = {0, 0, 0};
    if (fscanf(f, "%2s %d %d %d", magic, &w, &h, &maxv) == 4 && magic[0] == 'P' && magic[1] == '5' && maxv == 255) {
        fgetc(f);
        m.create(h, w, CV_8U);
        if (fread(m.data(), 1, (size_t)w * h, f) != (size_t)w * h) m.release();
    }
    fclose(f);
    return m;
}
}  // namespace cv
#endif
