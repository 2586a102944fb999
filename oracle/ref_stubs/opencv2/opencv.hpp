// Stand-in for <opencv2/opencv.hpp> — TEST INFRASTRUCTURE (oracle/_ref build only; see ref_stubs/Eigen/Dense).
// cv::Mat with at<T>() / clone / zeros, Vec3b, and the two 8-bit colour conversions getOccupancyResult uses
// (semantic_dsp_map.h:1335-1350).  The conversions restate OpenCV's published 8-bit RGB<->HSV algorithm (H in
// [0,180), fixed-point division tables with a 12-bit shift; HSV->RGB through float sectors with the two "1 - s*h" terms
// fused and the final scaling truncated) and are checked on all 2^24 inputs against the real cv2 4.13 module in
// tests/test_colour_cpu.py.  GUI / resize / colormap entry points
// exist only so the reference headers compile; they abort if executed.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <queue>
#include <string>
#include <vector>

typedef unsigned char uchar;
typedef unsigned short ushort;

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)

namespace cv {

[[noreturn]] inline void stub_unreachable(const char* what) {
    std::fprintf(stderr, "ref_stubs/opencv2: %s is outside the hot path and not implemented\n", what);
    std::abort();
}

template <typename T, int N>
struct Vec {
    T val[N];
    Vec() { for (int i = 0; i < N; ++i) val[i] = T(0); }
    Vec(T a, T b, T c) { static_assert(N == 3, "3 channels"); val[0] = a; val[1] = b; val[2] = c; }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
};
typedef Vec<uchar, 3> Vec3b;

struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : val{a, b, c, d} {}
    double operator[](int i) const { return val[i]; }
};

struct Size {
    int width = 0, height = 0;
    Size() = default;
    Size(int w, int h) : width(w), height(h) {}
};

inline int elem_size_of(int type) {
    static const int depth_bytes[7] = {1, 1, 2, 2, 4, 4, 8};
    return depth_bytes[type & 7] * ((type >> CV_CN_SHIFT) + 1);
}

inline uchar saturate_u8(double v) {
    const long r = std::lrint(v);
    return static_cast<uchar>(r < 0 ? 0 : (r > 255 ? 255 : r));
}

class Mat {
public:
    int rows = 0, cols = 0;
    size_t step = 0;      // bytes per row
    uchar* data = nullptr;

    Mat() = default;
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, const Scalar& s) {
        create(r, c, type);
        const int cn = (type >> CV_CN_SHIFT) + 1;
        if ((type & 7) != CV_8U) stub_unreachable("Mat(rows, cols, type, Scalar) for non-8U types");
        for (int i = 0; i < r; ++i)
            for (int j = 0; j < c; ++j)
                for (int k = 0; k < cn; ++k) data[i * step + static_cast<size_t>(j) * cn + k] = saturate_u8(s[k]);
    }
    // wraps caller memory without copying (the harness hands the depth image and masks in this way)
    Mat(int r, int c, int type, void* ext, size_t step_bytes = 0)
        : rows(r), cols(c), step(step_bytes ? step_bytes : static_cast<size_t>(c) * elem_size_of(type)), data(static_cast<uchar*>(ext)), type_(type) {}

    void create(int r, int c, int type) {
        rows = r; cols = c; type_ = type;
        step = static_cast<size_t>(c) * elem_size_of(type);
        own_ = std::shared_ptr<uchar>(new uchar[std::max<size_t>(step * r, 1)](), std::default_delete<uchar[]>());
        data = own_.get();
    }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }  // create() value-initialises
    int type() const { return type_; }
    int channels() const { return (type_ >> CV_CN_SHIFT) + 1; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    size_t total() const { return static_cast<size_t>(rows) * cols; }
    Size size() const { return Size(cols, rows); }
    Mat clone() const {
        Mat m;
        if (empty()) return m;
        m.create(rows, cols, type_);
        const size_t row_bytes = static_cast<size_t>(cols) * elem_size_of(type_);
        for (int i = 0; i < rows; ++i) std::memcpy(m.data + i * m.step, data + i * step, row_bytes);
        return m;
    }
    template <typename T>
    T& at(int i, int j) { return *reinterpret_cast<T*>(data + i * step + static_cast<size_t>(j) * sizeof(T)); }
    template <typename T>
    const T& at(int i, int j) const { return *reinterpret_cast<const T*>(data + i * step + static_cast<size_t>(j) * sizeof(T)); }
    template <typename T>
    T* ptr(int i = 0) { return reinterpret_cast<T*>(data + i * step); }
    template <typename T>
    const T* ptr(int i = 0) const { return reinterpret_cast<const T*>(data + i * step); }

private:
    int type_ = 0;
    std::shared_ptr<uchar> own_;
};

enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum { COLORMAP_JET = 2 };
enum { COLOR_BGR2RGB = 4, COLOR_RGB2HSV = 41, COLOR_HSV2RGB = 55 };

inline void resize(const Mat&, Mat&, Size, double = 0, double = 0, int = INTER_LINEAR) { stub_unreachable("cv::resize"); }
inline void applyColorMap(const Mat&, Mat&, int) {}   // display only (the VERBOSE_MODE debug windows, semantic_dsp_map.h:1562-1569), like imshow
inline void imshow(const std::string&, const Mat&) {}
inline int waitKey(int = 0) { return -1; }

// ---- 8-bit RGB <-> HSV (H in [0,180)) ------------------------------------------------------------------------
namespace hsv_detail {
constexpr int kShift = 12;
inline const int* sdiv_table() {
    static int t[256];
    static bool init = false;
    if (!init) {
        t[0] = 0;
        for (int i = 1; i < 256; ++i) t[i] = static_cast<int>(std::lrint((255 << kShift) / (1. * i)));
        init = true;
    }
    return t;
}
inline const int* hdiv_table180() {
    static int t[256];
    static bool init = false;
    if (!init) {
        t[0] = 0;
        for (int i = 1; i < 256; ++i) t[i] = static_cast<int>(std::lrint((180 << kShift) / (6. * i)));
        init = true;
    }
    return t;
}
inline void rgb2hsv_u8(uchar r, uchar g, uchar b, uchar out[3]) {
    const int R = r, G = g, B = b;
    const int v = std::max(R, std::max(G, B)), vmin = std::min(R, std::min(G, B));
    const int diff = v - vmin;
    const int vr = v == R ? -1 : 0, vg = v == G ? -1 : 0;
    const int s = (diff * sdiv_table()[v] + (1 << (kShift - 1))) >> kShift;
    int h = (vr & (G - B)) + (~vr & ((vg & (B - R + 2 * diff)) + ((~vg) & (R - G + 4 * diff))));
    h = (h * hdiv_table180()[diff] + (1 << (kShift - 1))) >> kShift;
    h += h < 0 ? 180 : 0;
    out[0] = static_cast<uchar>(h < 0 ? 0 : (h > 255 ? 255 : h));
    out[1] = static_cast<uchar>(s);
    out[2] = static_cast<uchar>(v);
}
inline void hsv2rgb_u8(uchar hh, uchar ss, uchar vv, uchar out[3]) {
    float h = hh, s = ss * (1.f / 255.f), v = vv * (1.f / 255.f);
    float b, g, r;
    if (s == 0) {
        b = g = r = v;
    } else {
        static const int sector_data[6][3] = {{1, 3, 0}, {1, 0, 2}, {3, 0, 1}, {0, 2, 1}, {0, 1, 3}, {2, 1, 0}};
        float tab[4];
        h *= 6.f / 180.f;
        if (h < 0) do h += 6; while (h < 0);
        else if (h >= 6) do h -= 6; while (h >= 6);
        int sector = static_cast<int>(std::floor(h));
        h -= sector;
        if (static_cast<unsigned>(sector) >= 6u) { sector = 0; h = 0.f; }
        tab[0] = v;
        tab[1] = v * (1.f - s);
        tab[2] = v * std::fmaf(-s, h, 1.f);
        tab[3] = v * std::fmaf(-s, 1.f - h, 1.f);
        b = tab[sector_data[sector][0]];
        g = tab[sector_data[sector][1]];
        r = tab[sector_data[sector][2]];
    }
    auto sat = [](float x) { const float y = std::floor(x * 255.f); return static_cast<uchar>(y < 0.f ? 0.f : (y > 255.f ? 255.f : y)); };
    out[0] = sat(r);
    out[1] = sat(g);
    out[2] = sat(b);
}
}  // namespace hsv_detail

inline void cvtColor(const Mat& src, Mat& dst, int code) {
    if (src.type() != CV_8UC3) stub_unreachable("cvtColor on a non-8UC3 image");
    Mat out(src.rows, src.cols, CV_8UC3);
    for (int i = 0; i < src.rows; ++i)
        for (int j = 0; j < src.cols; ++j) {
            const Vec3b& p = src.at<Vec3b>(i, j);
            uchar o[3];
            if (code == COLOR_RGB2HSV) hsv_detail::rgb2hsv_u8(p[0], p[1], p[2], o);
            else if (code == COLOR_HSV2RGB) hsv_detail::hsv2rgb_u8(p[0], p[1], p[2], o);
            else stub_unreachable("cvtColor code");
            out.at<Vec3b>(i, j) = Vec3b(o[0], o[1], o[2]);
        }
    dst = out;
}

}  // namespace cv
