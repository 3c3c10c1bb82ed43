/*
 * cvshim.cpp — the OpenCV stand-in's out-of-line parts (TEST INFRASTRUCTURE, see cvshim.hpp).
 * Mat, the eager float matrix algebra and the non-image helpers; no dependency on the oracle library (the drop-in arm
 * of the harness links this file too).  The image primitives live in cvshim_imgproc.cpp.
 */
#include "cvshim.hpp"


namespace cv {

void KeyPointsFilter::retainBest(std::vector<KeyPoint>& kps, int n) {
    /* features2d/keypoint.cpp: keep the n best responses and everything tied with the n-th */
    if (n >= 0 && (int)kps.size() > n) {
        if (n == 0) { kps.clear(); return; }
        std::nth_element(kps.begin(), kps.begin() + n - 1, kps.end(), [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
        const float ambiguous = kps[n - 1].response;
        auto it = std::partition(kps.begin() + n, kps.end(), [=](const KeyPoint& k) { return k.response >= ambiguous; });
        kps.resize(it - kps.begin());
    }
}

/* ---------------------------------------------------------------- Mat ---------------------------------------------- */
Mat& Mat::setTo(const Scalar& s) {
    for (int i = 0; i < rows; ++i)
        for (int j = 0; j < cols * channels(); ++j) {
            switch (depth()) {
                case CV_8U: ptr<uchar>(i)[j] = saturate_cast<uchar>(s[j % channels()]); break;
                case CV_32S: ptr<int>(i)[j] = cvRound(s[j % channels()]); break;
                case CV_32F: ptr<float>(i)[j] = (float)s[j % channels()]; break;
                case CV_64F: ptr<double>(i)[j] = s[j % channels()]; break;
                default: shim_fail("setTo depth");
            }
        }
    return *this;
}

void Mat::convertTo(Mat& dst, int rtype) const {
    const int ddepth = CV_MAT_DEPTH(rtype);
    Mat out(rows, cols, CV_MAKETYPE(ddepth, channels()));
    const int n = cols * channels();
    for (int i = 0; i < rows; ++i)
        for (int j = 0; j < n; ++j) {
            double v;
            switch (depth()) {
                case CV_8U: v = ptr<uchar>(i)[j]; break;
                case CV_32S: v = ptr<int>(i)[j]; break;
                case CV_32F: v = ptr<float>(i)[j]; break;
                case CV_64F: v = ptr<double>(i)[j]; break;
                default: shim_fail("convertTo depth");
            }
            switch (ddepth) {
                case CV_8U: out.ptr<uchar>(i)[j] = saturate_cast<uchar>(v); break;
                case CV_32S: out.ptr<int>(i)[j] = cvRound(v); break;
                case CV_32F: out.ptr<float>(i)[j] = (float)v; break;
                case CV_64F: out.ptr<double>(i)[j] = v; break;
                default: shim_fail("convertTo depth");
            }
        }
    dst = out;   /* dst may be *this (Frame.cc:599): the source ROI is left untouched, like OpenCV's reallocation */
}

Mat Mat::reshape(int cn, int new_rows) const {
    if (new_rows != 0) shim_fail("reshape(rows)");
    if (!isContinuous() && rows > 1) shim_fail("reshape of a non-continuous Mat");
    Mat m(*this);
    const int total_ch = cols * channels();
    if (cn == 0) cn = channels();
    if (total_ch % cn) shim_fail("reshape channel count");
    m.flags = CV_MAKETYPE(depth(), cn);
    m.cols = total_ch / cn;
    return m;
}

Mat Mat::t() const {
    Mat m(cols, rows, flags);
    const size_t es = elemSize();
    for (int i = 0; i < rows; ++i)
        for (int j = 0; j < cols; ++j) std::memcpy(m.data + (size_t)j * m.step + (size_t)i * es, data + (size_t)i * step + (size_t)j * es, es);
    return m;
}

Mat Mat::inv() const {   /* Gauss-Jordan with partial pivoting in double; not on the hot path */
    if (rows != cols) shim_fail("inv of a non-square Mat");
    const int n = rows;
    std::vector<double> a((size_t)n * 2 * n, 0.0);
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) a[(size_t)i * 2 * n + j] = get1(i, j);
        a[(size_t)i * 2 * n + n + i] = 1.0;
    }
    for (int c = 0; c < n; ++c) {
        int p = c;
        for (int r = c + 1; r < n; ++r) if (std::fabs(a[(size_t)r * 2 * n + c]) > std::fabs(a[(size_t)p * 2 * n + c])) p = r;
        if (a[(size_t)p * 2 * n + c] == 0.0) return Mat::zeros(n, n, flags);
        if (p != c) for (int j = 0; j < 2 * n; ++j) std::swap(a[(size_t)p * 2 * n + j], a[(size_t)c * 2 * n + j]);
        const double d = a[(size_t)c * 2 * n + c];
        for (int j = 0; j < 2 * n; ++j) a[(size_t)c * 2 * n + j] /= d;
        for (int r = 0; r < n; ++r) if (r != c) {
            const double f = a[(size_t)r * 2 * n + c];
            if (f != 0.0) for (int j = 0; j < 2 * n; ++j) a[(size_t)r * 2 * n + j] -= f * a[(size_t)c * 2 * n + j];
        }
    }
    Mat m(n, n, flags);
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) m.set1(i, j, a[(size_t)i * 2 * n + n + j]);
    return m;
}

double Mat::dot(const Mat& o) const {   /* cv::Mat::dot on CV_32F accumulates in double */
    if (total() != o.total() || flags != o.flags) shim_fail("dot of mismatching Mats");
    double s = 0;
    const int n = (int)total();
    for (int k = 0; k < n; ++k) {
        const int i1 = k / cols, j1 = k % cols, i2 = k / o.cols, j2 = k % o.cols;
        s += get1(i1, j1) * o.get1(i2, j2);
    }
    return s;
}

/* cv::gemm arithmetic for D = alpha * A * B + beta * C, no transposes.  For CV_32F with 2 <= len <= 4 and
 * (len == D.cols or len == D.rows) OpenCV uses an unrolled small-matrix path: the inner product is summed in float, left to
 * right, then d = (float)(s * alpha + c * beta) in double; everything else accumulates in double (matmul.simd.hpp). */
void gemm(const Mat& a, const Mat& b, double alpha, const Mat& c, double beta, Mat& dst, int flags) {
    Mat A = (flags & GEMM_1_T) ? a.t() : a, B = (flags & GEMM_2_T) ? b.t() : b, C = (flags & GEMM_3_T) ? c.t() : c;
    if (A.cols != B.rows || A.flags != B.flags) shim_fail("gemm of mismatching Mats");
    const int len = A.cols;
    Mat D(A.rows, B.cols, A.flags);
    const bool has_c = !C.empty() && beta != 0;
    const bool small = flags == 0 && A.depth() == CV_32F && len >= 2 && len <= 4 && (len == D.cols || len == D.rows);
    for (int i = 0; i < D.rows; ++i)
        for (int j = 0; j < D.cols; ++j) {
            double r;
            if (small) {
                float s = A.at<float>(i, 0) * B.at<float>(0, j);
                for (int k = 1; k < len; ++k) s = s + A.at<float>(i, k) * B.at<float>(k, j);
                r = (double)s * alpha;
            } else {
                double s = 0;
                for (int k = 0; k < len; ++k) s += A.get1(i, k) * B.get1(k, j);
                r = s * alpha;
            }
            if (has_c) r += C.get1(i, j) * beta;
            D.set1(i, j, r);
        }
    dst = D;
}

Mat operator*(const Mat& a, const Mat& b) { Mat d; gemm(a, b, 1.0, Mat(), 0.0, d, 0); return d; }

static Mat scaled(const Mat& a, double s, bool divide) {
    Mat m(a.rows, a.cols, a.flags);
    for (int i = 0; i < a.rows; ++i)
        for (int j = 0; j < a.cols; ++j) m.set1(i, j, divide ? a.get1(i, j) / s : a.get1(i, j) * s);   /* convertTo-style: double, one rounding */
    return m;
}
Mat operator*(const Mat& a, double s) { return scaled(a, s, false); }
Mat operator*(double s, const Mat& a) { return scaled(a, s, false); }
Mat operator/(const Mat& a, double s) { return scaled(a, 1.0 / s, false); }   /* MatExpr a / s = a * (1/s) */

static Mat addsub(const Mat& a, const Mat& b, double sign) {
    if (a.rows != b.rows || a.cols != b.cols || a.flags != b.flags) shim_fail("add/sub of mismatching Mats");
    Mat m(a.rows, a.cols, a.flags);
    for (int i = 0; i < a.rows; ++i)
        for (int j = 0; j < a.cols; ++j) m.set1(i, j, a.get1(i, j) + sign * b.get1(i, j));   /* exact in double, one rounding */
    return m;
}
Mat operator+(const Mat& a, const Mat& b) { return addsub(a, b, 1.0); }
Mat operator-(const Mat& a, const Mat& b) { return addsub(a, b, -1.0); }
Mat operator-(const Mat& a) { return scaled(a, -1.0, false); }

double norm(const Mat& a, int normType) {   /* double accumulation (ST = double for CV_32F) */
    double s = 0;
    for (int i = 0; i < a.rows; ++i)
        for (int j = 0; j < a.cols; ++j) {
            const double v = a.get1(i, j);
            if (normType == NORM_L2) s += v * v;
            else if (normType == NORM_L1) s += std::fabs(v);
            else if (normType == NORM_INF) s = std::max(s, std::fabs(v));
            else shim_fail("norm type");
        }
    return normType == NORM_L2 ? std::sqrt(s) : s;
}
double norm(const Mat& a, const Mat& b, int normType) {
    if (a.rows != b.rows || a.cols != b.cols || a.flags != b.flags) shim_fail("norm of mismatching Mats");
    double s = 0;
    for (int i = 0; i < a.rows; ++i)
        for (int j = 0; j < a.cols; ++j) {
            const double v = a.get1(i, j) - b.get1(i, j);   /* float inputs: the difference is taken in float by OpenCV; exact here (integers) */
            if (normType == NORM_L2) s += v * v;
            else if (normType == NORM_L1) s += std::fabs(a.depth() == CV_32F ? (double)(float)v : v);
            else if (normType == NORM_INF) s = std::max(s, std::fabs(v));
            else shim_fail("norm type");
        }
    return normType == NORM_L2 ? std::sqrt(s) : s;
}

void undistortPoints(InputArray, OutputArray, InputArray, InputArray, InputArray, InputArray) {
    shim_fail("cv::undistortPoints (the harness uses rectified / undistorted cameras: mDistCoef[0] == 0, Frame.cc:443-447)");
}

}  // namespace cv
