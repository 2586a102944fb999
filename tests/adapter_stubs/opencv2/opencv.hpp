// Minimal stand-in for cv::Mat (TEST ONLY).
#pragma once
#include <cstddef>
#include <vector>
namespace cv {
struct Mat {
    int rows = 0, cols = 0;
    size_t step = 0;
    unsigned char* data = nullptr;
    std::vector<unsigned char> store;
    Mat() {}
    Mat(int r, int c, size_t elem) : rows(r), cols(c), step(c * elem), store((size_t)r * c * elem) { data = store.data(); }
    Mat(const Mat& o) : rows(o.rows), cols(o.cols), step(o.step), store(o.store) { data = store.data(); }
    Mat& operator=(const Mat& o) { rows = o.rows; cols = o.cols; step = o.step; store = o.store; data = store.data(); return *this; }
};
}  // namespace cv
