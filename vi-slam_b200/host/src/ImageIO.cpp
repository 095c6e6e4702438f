// ImageIO.cpp — what cv::imread(file, CV_LOAD_IMAGE_GRAYSCALE) does for the dataset formats of this path
// (ImageReader.cpp:80-82): EuRoC / KITTI frames are 8-bit grayscale PNG, TUM frames 8-bit RGB PNG; the synthetic test
// datasets use binary PGM.  PNG: non-interlaced, bit depth 8, colour types 0 (gray), 2 (RGB), 4 (gray+alpha),
// 6 (RGBA); the IDAT stream is inflated with zlib and un-filtered here.  Colour is reduced with libpng's
// png_set_rgb_to_gray(0.299, 0.587) fixed-point weights, which is what OpenCV 3.2's PNG decoder requests:
// (9797 R + 19234 G + 3737 B + 16384) >> 15.
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "vislam/DataReader.hpp"

namespace {

bool read_file(const std::string& file, std::vector<unsigned char>& out) {
    FILE* f = std::fopen(file.c_str(), "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? (size_t)n : 0);
    const bool ok = n >= 0 && std::fread(out.data(), 1, out.size(), f) == out.size();
    std::fclose(f);
    return ok;
}

cv::Mat decode_pgm(const std::vector<unsigned char>& buf) {
    // "P5" <ws> width <ws> height <ws> maxval <single ws> raster; '#' comments run to the end of the line
    size_t p = 2;
    int vals[3] = {0, 0, 0};
    for (int k = 0; k < 3; k++) {
        for (;;) {
            while (p < buf.size() && (buf[p] == ' ' || buf[p] == '\t' || buf[p] == '\n' || buf[p] == '\r')) p++;
            if (p < buf.size() && buf[p] == '#') { while (p < buf.size() && buf[p] != '\n') p++; continue; }
            break;
        }
        if (p >= buf.size() || buf[p] < '0' || buf[p] > '9') return cv::Mat();
        while (p < buf.size() && buf[p] >= '0' && buf[p] <= '9') vals[k] = vals[k] * 10 + (buf[p++] - '0');
    }
    p++;   // the single whitespace byte after maxval
    const int w = vals[0], h = vals[1];
    if (w <= 0 || h <= 0 || vals[2] != 255 || p + (size_t)w * h > buf.size()) return cv::Mat();
    cv::Mat m(h, w, CV_8U);
    std::memcpy(m.data, buf.data() + p, (size_t)w * h);
    return m;
}

inline unsigned be32(const unsigned char* p) { return ((unsigned)p[0] << 24) | ((unsigned)p[1] << 16) | ((unsigned)p[2] << 8) | p[3]; }

inline int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

cv::Mat decode_png(const std::vector<unsigned char>& buf) {
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (buf.size() < 8 + 25 || std::memcmp(buf.data(), sig, 8) != 0) return cv::Mat();
    int w = 0, h = 0, depth = 0, ctype = -1, interlace = 0;
    std::vector<unsigned char> idat;
    for (size_t p = 8; p + 12 <= buf.size();) {
        const unsigned len = be32(&buf[p]);
        const unsigned char* type = &buf[p + 4];
        const unsigned char* body = &buf[p + 8];
        if (p + 12 + (size_t)len > buf.size()) return cv::Mat();
        if (!std::memcmp(type, "IHDR", 4) && len >= 13) {
            w = (int)be32(body); h = (int)be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12];
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), body, body + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            break;
        }
        p += 12 + (size_t)len;
    }
    int ch = 0;
    switch (ctype) { case 0: ch = 1; break; case 2: ch = 3; break; case 4: ch = 2; break; case 6: ch = 4; break; default: break; }
    if (w <= 0 || h <= 0 || depth != 8 || ch == 0 || interlace != 0) return cv::Mat();
    const size_t stride = (size_t)w * ch;
    std::vector<unsigned char> raw((stride + 1) * (size_t)h);
    uLongf raw_len = (uLongf)raw.size();
    if (uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size()) != Z_OK || raw_len != raw.size()) return cv::Mat();
    // un-filter in place (PNG spec 9.2): each scanline is <filter type> <stride bytes>
    std::vector<unsigned char> zero(stride, 0);
    for (int y = 0; y < h; y++) {
        unsigned char* cur = &raw[(stride + 1) * (size_t)y + 1];
        const unsigned char* up = y ? &raw[(stride + 1) * (size_t)(y - 1) + 1] : zero.data();
        const int ft = cur[-1];
        for (size_t x = 0; x < stride; x++) {
            const int a = x >= (size_t)ch ? cur[x - ch] : 0, b = up[x], c = x >= (size_t)ch ? up[x - ch] : 0;
            int pred = 0;
            switch (ft) { case 1: pred = a; break; case 2: pred = b; break; case 3: pred = (a + b) >> 1; break; case 4: pred = paeth(a, b, c); break; default: break; }
            cur[x] = (unsigned char)(cur[x] + pred);
        }
    }
    cv::Mat m(h, w, CV_8U);
    for (int y = 0; y < h; y++) {
        const unsigned char* src = &raw[(stride + 1) * (size_t)y + 1];
        unsigned char* dst = m.ptr<unsigned char>(y);
        if (ch <= 2) {
            for (int x = 0; x < w; x++) dst[x] = src[(size_t)x * ch];
        } else {
            for (int x = 0; x < w; x++) {
                const unsigned r = src[(size_t)x * ch], g = src[(size_t)x * ch + 1], b = src[(size_t)x * ch + 2];
                dst[x] = (r == g && g == b) ? (unsigned char)r : (unsigned char)((9797u * r + 19234u * g + 3737u * b + 16384u) >> 15);
            }
        }
    }
    return m;
}

}  // namespace

cv::Mat vi::imread_gray(const std::string& file) {
    std::vector<unsigned char> buf;
    if (!read_file(file, buf) || buf.size() < 4) return cv::Mat();
    if (buf[0] == 'P' && buf[1] == '5') return decode_pgm(buf);
    if (buf[0] == 0x89 && buf[1] == 'P') return decode_png(buf);
    return cv::Mat();
}
