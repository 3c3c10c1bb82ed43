// cv_compat.h — the handful of OpenCV types the ORB front-end's class surface mentions, for builds WITHOUT OpenCV
// (this image has no OpenCV C++ headers).  With OpenCV present the real headers are used and this file is empty:
// the shims in ORBextractor.cc / ORBmatcher.cc / StereoMatcher.cc compile against either.
#pragma once
#if defined(ORB_B200_USE_OPENCV) || __has_include(<opencv2/core/core.hpp>)
#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>
#else
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0

namespace cv {

struct Point2f { float x = 0, y = 0; Point2f() {} Point2f(float x_, float y_) : x(x_), y(y_) {} };

// same field order and size (28 bytes) as cv::KeyPoint
struct KeyPoint {
    Point2f pt;
    float size = 0, angle = -1, response = 0;
    int octave = 0, class_id = -1;
};

// 8-bit single-channel matrix with shared ownership and ROI support (what mvImagePyramid needs: step != cols)
class Mat {
public:
    int rows = 0, cols = 0;
    size_t step = 0;
    uint8_t* data = nullptr;
    Mat() {}
    Mat(int r, int c, int /*type*/) { create(r, c, CV_8U); }
    Mat(int r, int c, int /*type*/, void* ext, size_t st = 0) : rows(r), cols(c), step(st ? st : (size_t)c), data((uint8_t*)ext) {}
    void create(int r, int c, int /*type*/) {
        if (r == rows && c == cols && buf_ && step == (size_t)c) return;
        buf_.reset(new std::vector<uint8_t>((size_t)r * c));
        rows = r; cols = c; step = (size_t)c; data = buf_->data();
    }
    void release() { buf_.reset(); rows = cols = 0; step = 0; data = nullptr; }
    bool empty() const { return rows == 0 || cols == 0 || !data; }
    int type() const { return CV_8UC1; }
    bool isContinuous() const { return step == (size_t)cols; }
    uint8_t* ptr(int y = 0) { return data + (size_t)y * step; }
    const uint8_t* ptr(int y = 0) const { return data + (size_t)y * step; }
    Mat row(int y) const { Mat m; m.buf_ = buf_; m.rows = 1; m.cols = cols; m.step = step; m.data = data + (size_t)y * step; return m; }
    Mat roi(int x, int y, int w, int h) const {
        Mat m; m.buf_ = buf_; m.rows = h; m.cols = w; m.step = step; m.data = data + (size_t)y * step + x; return m;
    }
    Mat getMat() const { return *this; }
private:
    std::shared_ptr<std::vector<uint8_t>> buf_;
};

typedef const Mat& InputArray;
typedef Mat& OutputArray;

}  // namespace cv
#endif
