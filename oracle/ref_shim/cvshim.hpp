/*
 * cvshim.hpp — a minimal stand-in for the OpenCV C++ API, TEST INFRASTRUCTURE ONLY.
 *
 * Purpose: this image has no OpenCV C++ headers or libraries (only the python cv2 4.13.0 wheel), so the reference's
 * own translation units (orb_slam2/src/ORBextractor.cc, ORBmatcher.cc, Frame.cc and the DBoW2 sources) cannot be
 * compiled against the real thing.  This header declares exactly the OpenCV surface those files use, so that they
 * compile UNMODIFIED from where they lie under /root/reference into oracle/_ref/liborb_ref.so (oracle/Makefile).
 * The control flow that runs in oracle/_ref is therefore the reference's own code; only the OpenCV primitives
 * below it are ours:
 *
 *   resize(INTER_LINEAR, 8u), copyMakeBorder(REFLECT_101), FAST(9/16 + NMS), GaussianBlur(7x7, sigma 2, 8u),
 *   fastAtan2, cvRound  -> the recipes of oracle/orb_oracle_extract.cpp (orc_* exports), each of which is checked
 *                          against the cv2 4.13.0 wheel by tests/test_oracle_vs_cv2.py
 *   float matrix algebra (operator*, +, -, t(), dot, norm)  -> eager evaluation; products follow cv::gemm's
 *                          arithmetic for the shapes the reference uses (checked against cv2.gemm by
 *                          tests/test_ref_pin.py): 2..4-term inner products of small matrices are accumulated in
 *                          float, left to right; everything else in double.  Real OpenCV evaluates A*B+C lazily as
 *                          ONE gemm; for the float-accumulated shapes the result is the same (the final add is a
 *                          single correctly rounded operation either way).
 *
 * The same header lets the repo's drop-in host sources (orb_slam_2_ros_b200/host/) compile here; with OpenCV
 * installed they compile against the real headers instead.
 */
#ifndef ORB_REF_CVSHIM_HPP
#define ORB_REF_CVSHIM_HPP

#include <algorithm>
#include <cassert>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <map>
#include <set>
#include <sstream>
#include <memory>
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
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> CV_CN_SHIFT) & 511) + 1)
#define CV_MAKETYPE(depth, cn) (CV_MAT_DEPTH(depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_PI 3.1415926535897932384626433832795
#define CV_VERSION "shim-4.13.0"
#define CV_MAJOR_VERSION 4

namespace cv {

using std::string;
typedef std::string String;

[[noreturn]] inline void shim_fail(const char* what) {
    std::fprintf(stderr, "cvshim: %s is not part of the surface the ORB front-end uses\n", what);
    std::abort();
}

/* cvRound: round half to even (SSE cvtss2si / lrint) */
inline int cvRound(double v) { return (int)std::lrint(v); }
inline int cvRound(float v) { return (int)std::lrintf(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
inline int cvFloor(float v) { int i = (int)v; return i - (i > v); }
inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }
inline int cvCeil(float v) { int i = (int)v; return i + (i < v); }
float fastAtan2(float y, float x); /* degrees in [0,360) */

template <class T> inline T saturate_cast(double v) { return (T)v; }
template <> inline uchar saturate_cast<uchar>(double v) { int i = cvRound(v); return (uchar)(i < 0 ? 0 : i > 255 ? 255 : i); }

template <class T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <class U> Point_(const Point_<U>& p) : x((T)p.x), y((T)p.y) {}
    Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }   /* saturate_cast<float> is the identity */
    Point_& operator*=(double s) { x = (T)(x * s); y = (T)(y * s); return *this; }
    Point_& operator+=(const Point_& o) { x += o.x; y += o.y; return *this; }
    Point_& operator-=(const Point_& o) { x -= o.x; y -= o.y; return *this; }
};
template <class T> inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <class T> inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x - b.x, a.y - b.y); }
template <class T> inline bool operator==(const Point_<T>& a, const Point_<T>& b) { return a.x == b.x && a.y == b.y; }
/* Point2i(float, float): the reference builds integer corners from floats (ORBextractor.cc:579-580); C++ converts by
 * truncation at the call site exactly as with OpenCV's Point_<int>(int, int) constructor. */
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

template <class T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
};
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;

template <class T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    T area() const { return width * height; }
};
typedef Size_<int> Size;

template <class T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
};
typedef Rect_<int> Rect;

template <class T, int N> struct Vec { T val[N]; T& operator[](int i) { return val[i]; } const T& operator[](int i) const { return val[i]; } };
typedef Vec<float, 2> Vec2f;

struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    double operator[](int i) const { return val[i]; }
};

struct Range { int start, end; Range(int s, int e) : start(s), end(e) {} };

/* same field order and size (28 bytes) as cv::KeyPoint */
struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(Point2f p, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(p), size(s), angle(a), response(r), octave(o), class_id(c) {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

struct KeyPointsFilter {
    /* only ComputeKeyPointsOld (dead code in the reference, ORBextractor.cc:1045,1063) calls this */
    static void retainBest(std::vector<KeyPoint>& kps, int n);
};

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4, BORDER_REFLECT101 = 4,
       BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3 };
enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4, NORM_HAMMING = 6 };
enum { GEMM_1_T = 1, GEMM_2_T = 2, GEMM_3_T = 4 };

class Mat;
template <class T> class Mat_;
template <class T> class MatCommaInitializer_;

class Mat {
public:
    int flags;  /* the type (depth + channels) */
    int dims;
    int rows, cols;
    uchar* data;
    size_t step; /* bytes per row */

    Mat() : flags(0), dims(2), rows(0), cols(0), data(nullptr), step(0) {}
    Mat(int r, int c, int type) : Mat() { create(r, c, type); }
    Mat(Size sz, int type) : Mat() { create(sz.height, sz.width, type); }
    Mat(int r, int c, int type, const Scalar& s) : Mat() { create(r, c, type); setTo(s); }
    Mat(int r, int c, int type, void* ext, size_t st = 0) : flags(type), dims(2), rows(r), cols(c), data((uchar*)ext), step(st) {
        if (!step) step = (size_t)c * elemSize();
    }
    Mat(const Mat& m, const Rect& r) : Mat(m) { data += (size_t)r.y * step + (size_t)r.x * elemSize(); rows = r.height; cols = r.width; }

    void create(int r, int c, int type) {
        if (data && r == rows && c == cols && type == flags) return;   /* same size and type: keep the buffer (also an ROI's) */
        flags = type; rows = r; cols = c; dims = 2;
        step = (size_t)c * elemSize();
        size_t bytes = (size_t)r * step;
        buf_.reset(new uchar[bytes ? bytes : 1], std::default_delete<uchar[]>());
        data = buf_.get();
    }
    void create(Size sz, int type) { create(sz.height, sz.width, type); }
    void release() { buf_.reset(); data = nullptr; rows = cols = 0; step = 0; }

    int type() const { return flags; }
    int depth() const { return CV_MAT_DEPTH(flags); }
    int channels() const { return CV_MAT_CN(flags); }
    size_t elemSize1() const { static const int s[8] = {1, 1, 2, 2, 4, 4, 8, 2}; return s[depth()]; }
    size_t elemSize() const { return elemSize1() * channels(); }
    size_t step1() const { return step / elemSize1(); }
    size_t total() const { return (size_t)rows * cols; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    bool isContinuous() const { return rows <= 1 || step == (size_t)cols * elemSize(); }
    Size size() const { return Size(cols, rows); }

    uchar* ptr(int y = 0) { return data + (size_t)y * step; }
    const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
    template <class T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
    template <class T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }
    template <class T> T& at(int i, int j) { return ((T*)(data + (size_t)i * step))[j]; }
    template <class T> const T& at(int i, int j) const { return ((const T*)(data + (size_t)i * step))[j]; }
    template <class T> T& at(int i) {   /* cv::Mat::at(int i0): continuous or single row -> flat, single column -> by rows */
        if (isContinuous() || rows == 1) return ((T*)data)[i];
        if (cols == 1) return *(T*)(data + (size_t)i * step);
        return ((T*)(data + (size_t)(i / cols) * step))[i % cols];
    }
    template <class T> const T& at(int i) const { return const_cast<Mat*>(this)->at<T>(i); }

    Mat row(int y) const { return Mat(*this, Rect(0, y, cols, 1)); }
    Mat col(int x) const { return Mat(*this, Rect(x, 0, 1, rows)); }
    Mat rowRange(int a, int b) const { return Mat(*this, Rect(0, a, cols, b - a)); }
    Mat colRange(int a, int b) const { return Mat(*this, Rect(a, 0, b - a, rows)); }
    Mat rowRange(const Range& r) const { return rowRange(r.start, r.end); }
    Mat colRange(const Range& r) const { return colRange(r.start, r.end); }
    Mat operator()(const Rect& r) const { return Mat(*this, r); }

    Mat clone() const { Mat m; copyTo(m); return m; }
    void copyTo(Mat& dst) const {
        if (empty()) { dst.release(); return; }
        if (dst.data == data && dst.rows == rows && dst.cols == cols) return;
        if (!(dst.data && dst.rows == rows && dst.cols == cols && dst.flags == flags)) dst.create(rows, cols, flags);
        const size_t rb = (size_t)cols * elemSize();
        for (int y = 0; y < rows; ++y) std::memcpy(dst.ptr(y), ptr(y), rb);
    }
    void convertTo(Mat& dst, int rtype) const;
    Mat reshape(int cn, int new_rows = 0) const;
    Mat t() const;
    Mat inv() const;
    double dot(const Mat& o) const;
    Mat& setTo(const Scalar& s);
    Mat& operator=(const Scalar& s) { return setTo(s); }

    /* Mat::zeros / ones / eye return a MatExpr in OpenCV; ASSIGNING one to an existing Mat of the same size and type fills
     * that Mat's buffer in place (MatOp_Initializer::assign = create() + setTo) instead of rebinding the header.  The
     * reference relies on it: computeDescriptors zeroes the caller's descriptor rows through an ROI (ORBextractor.cc:1077). */
    struct Init {
        int rows, cols, type, kind;   /* kind: 0 zeros, 1 ones, 2 eye */
        operator Mat() const { Mat m; m.assignInit(*this); return m; }
    };
    static Init zeros(int r, int c, int type) { return Init{r, c, type, 0}; }
    static Init zeros(Size sz, int type) { return Init{sz.height, sz.width, type, 0}; }
    static Init ones(int r, int c, int type) { return Init{r, c, type, 1}; }
    static Init eye(int r, int c, int type) { return Init{r, c, type, 2}; }
    Mat(const Init& e) : Mat() { assignInit(e); }
    Mat& operator=(const Init& e) { assignInit(e); return *this; }
    void assignInit(const Init& e) {
        create(e.rows, e.cols, e.type);
        setTo(Scalar(e.kind == 1 ? 1.0 : 0.0));
        if (e.kind == 2) for (int i = 0; i < std::min(rows, cols); ++i) set1(i, i, 1.0);
    }

    /* element access by depth, as double (float / double / int / uchar) */
    double get1(int i, int j) const {
        switch (depth()) {
            case CV_8U: return at<uchar>(i, j);
            case CV_32S: return at<int>(i, j);
            case CV_32F: return at<float>(i, j);
            case CV_64F: return at<double>(i, j);
        }
        shim_fail("Mat depth");
    }
    void set1(int i, int j, double v) {
        switch (depth()) {
            case CV_8U: at<uchar>(i, j) = saturate_cast<uchar>(v); return;
            case CV_32S: at<int>(i, j) = cvRound(v); return;
            case CV_32F: at<float>(i, j) = (float)v; return;
            case CV_64F: at<double>(i, j) = v; return;
        }
        shim_fail("Mat depth");
    }

protected:
    std::shared_ptr<uchar> buf_;
};

template <class T> struct DepthOf;
template <> struct DepthOf<uchar> { enum { value = CV_8U }; };
template <> struct DepthOf<int> { enum { value = CV_32S }; };
template <> struct DepthOf<float> { enum { value = CV_32F }; };
template <> struct DepthOf<double> { enum { value = CV_64F }; };

template <class T> class MatCommaInitializer_ {
public:
    explicit MatCommaInitializer_(Mat_<T>* m) : m_(m), i_(0) {}
    template <class U> MatCommaInitializer_& operator,(U v) { put((T)v); return *this; }
    void put(T v) { Mat* m = (Mat*)m_; m->template at<T>(i_ / m->cols, i_ % m->cols) = v; ++i_; }
    operator Mat_<T>() const { return *m_; }
    operator Mat() const { return *(Mat*)m_; }
private:
    Mat_<T>* m_;
    int i_;
};

template <class T> class Mat_ : public Mat {
public:
    Mat_() : Mat() { flags = DepthOf<T>::value; }
    Mat_(int r, int c) : Mat(r, c, DepthOf<T>::value) {}
    Mat_(const Mat& m) : Mat(m) {}
    T& operator()(int i, int j) { return this->template at<T>(i, j); }
    const T& operator()(int i, int j) const { return this->template at<T>(i, j); }
    T& operator()(int i) { return this->template at<T>(i); }
};
/* cv::Mat_<float>(3,1) << x, y, z;  (Frame.cc:687) — the temporary lives until the end of the full expression */
template <class T, class U> inline MatCommaInitializer_<T> operator<<(const Mat_<T>& m, U v) {
    MatCommaInitializer_<T> ci(const_cast<Mat_<T>*>(&m));
    ci.put((T)v);
    return ci;
}

/* eager matrix algebra (see the header comment) */
Mat operator*(const Mat& a, const Mat& b);
Mat operator*(const Mat& a, double s);
Mat operator*(double s, const Mat& a);
Mat operator/(const Mat& a, double s);
Mat operator+(const Mat& a, const Mat& b);
Mat operator-(const Mat& a, const Mat& b);
Mat operator-(const Mat& a);
template <class T> inline Mat operator*(const MatCommaInitializer_<T>& a, const Mat& b) { return (Mat)a * b; }
inline Mat operator*(double s, const Mat::Init& e) { return s * (Mat)e; }
inline Mat operator*(const Mat::Init& e, double s) { return (Mat)e * s; }
double norm(const Mat& a, int normType = NORM_L2);
double norm(const Mat& a, const Mat& b, int normType = NORM_L2);
void gemm(const Mat& a, const Mat& b, double alpha, const Mat& c, double beta, Mat& dst, int flags = 0);

class _InputArray {
public:
    _InputArray() : m_(nullptr) {}
    _InputArray(const Mat& m) : m_(const_cast<Mat*>(&m)) {}
    Mat getMat() const { return m_ ? *m_ : Mat(); }
    bool empty() const { return !m_ || m_->empty(); }
    int type() const { return m_ ? m_->type() : 0; }
    Size size() const { return m_ ? m_->size() : Size(); }
protected:
    Mat* m_;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray() {}
    _OutputArray(Mat& m) : _InputArray(m) {}
    void create(int r, int c, int type) const { m_->create(r, c, type); }
    void create(Size sz, int type) const { m_->create(sz.height, sz.width, type); }
    void release() const { if (m_) m_->release(); }
    Mat& getMatRef() const { return *m_; }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
typedef const _OutputArray& InputOutputArray;
inline InputArray noArray() { static _InputArray none; return none; }

/* ---- imgproc / features2d primitives: OpenCV 4.13.0 semantics (oracle/orb_oracle_extract.cpp recipes) ---- */
void resize(InputArray src, OutputArray dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int borderType, const Scalar& value = Scalar());
void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT);
void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);
void undistortPoints(InputArray src, OutputArray dst, InputArray cameraMatrix, InputArray distCoeffs, InputArray R = noArray(),
                     InputArray P = noArray());

/* ---- persistence: DBoW2's TemplatedVocabulary declares yml save / load virtuals (TemplatedVocabulary.h:1444-1640); they
 *      must compile, nothing on the hot path calls them ---- */
class FileNode {
public:
    FileNode operator[](const char*) const { shim_fail("cv::FileNode"); }
    FileNode operator[](const std::string&) const { shim_fail("cv::FileNode"); }
    FileNode operator[](int) const { shim_fail("cv::FileNode"); }
    size_t size() const { return 0; }
    operator int() const { shim_fail("cv::FileNode"); }
    operator float() const { shim_fail("cv::FileNode"); }
    operator double() const { shim_fail("cv::FileNode"); }
    operator std::string() const { shim_fail("cv::FileNode"); }
    bool empty() const { return true; }
};
class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const std::string&, int) { shim_fail("cv::FileStorage"); }
    bool isOpened() const { return false; }
    void release() {}
    FileNode operator[](const char*) const { shim_fail("cv::FileStorage"); }
    FileNode operator[](const std::string&) const { shim_fail("cv::FileStorage"); }
};
template <class T> inline FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }
template <class T> inline void operator>>(const FileNode&, T&) { shim_fail("cv::FileNode"); }

}  // namespace cv

using cv::cvRound;
using cv::cvFloor;
using cv::cvCeil;

#endif /* ORB_REF_CVSHIM_HPP */
