/*
 * cvshim_imgproc.cpp — the OpenCV stand-in's image primitives (TEST INFRASTRUCTURE, see cvshim.hpp): resize, copyMakeBorder,
 * GaussianBlur, FAST, fastAtan2 forward to the cv2-4.13.0-verified recipes exported by oracle/_build/liborb_oracle.so.
 * Only the reference arm (oracle/_ref) links this file.
 */
#include "cvshim.hpp"

#include "../orb_oracle.h"

namespace cv {

float fastAtan2(float y, float x) { return orc_fast_atan2(y, x); }

/* ------------------------------------------------------------ imgproc ---------------------------------------------- */
static void need_u8(const Mat& m, const char* what) { if (m.type() != CV_8UC1) shim_fail(what); }

void resize(InputArray _src, OutputArray _dst, Size dsize, double, double, int interpolation) {
    Mat src = _src.getMat();
    need_u8(src, "resize of a non-8UC1 Mat");
    if (interpolation != INTER_LINEAR) shim_fail("resize interpolation");
    _dst.create(dsize, src.type());          /* a correctly sized ROI is written in place (ORBextractor.cc:1171) */
    Mat dst = _dst.getMat();
    orc_resize_linear_u8(src.data, src.cols, src.rows, (int)src.step, dst.data, dst.cols, dst.rows, (int)dst.step);
}

void copyMakeBorder(InputArray _src, OutputArray _dst, int top, int bottom, int left, int right, int borderType, const Scalar&) {
    Mat src = _src.getMat();
    need_u8(src, "copyMakeBorder of a non-8UC1 Mat");
    if ((borderType & ~BORDER_ISOLATED) != BORDER_REFLECT_101 || top != bottom || top != left || top != right) shim_fail("copyMakeBorder mode");
    /* without BORDER_ISOLATED OpenCV reads the pixels around an ROI; the harness only passes whole images */
    _dst.create(src.rows + top + bottom, src.cols + left + right, src.type());
    Mat dst = _dst.getMat();
    orc_border_reflect101(src.data, src.cols, src.rows, (int)src.step, dst.data, top, (int)dst.step);
}

void GaussianBlur(InputArray _src, OutputArray _dst, Size ksize, double sigmaX, double sigmaY, int borderType) {
    Mat src = _src.getMat();
    need_u8(src, "GaussianBlur of a non-8UC1 Mat");
    if (ksize.width != 7 || ksize.height != 7 || sigmaX != 2 || sigmaY != 2 || borderType != BORDER_REFLECT_101) shim_fail("GaussianBlur parameters");
    Mat tmp(src.rows, src.cols, src.type());
    orc_gaussian7x7_s2(src.data, src.cols, src.rows, (int)src.step, tmp.data, (int)tmp.step);
    _dst.create(src.rows, src.cols, src.type());
    Mat dst = _dst.getMat();
    tmp.copyTo(dst);
}

void FAST(InputArray _image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression) {
    Mat img = _image.getMat();
    need_u8(img, "FAST of a non-8UC1 Mat");
    static_assert(sizeof(KeyPoint) == sizeof(orc_kp), "KeyPoint layout");
    int cap = 256;
    for (;;) {   /* orc_fast9_16 returns the full count and writes at most cap entries */
        keypoints.resize(cap);
        const int n = orc_fast9_16(img.data, img.cols, img.rows, (int)img.step, threshold, nonmaxSuppression ? 1 : 0, reinterpret_cast<orc_kp*>(keypoints.data()), cap);
        if (n <= cap) { keypoints.resize(n < 0 ? 0 : n); break; }
        cap = n;
    }
}

}  // namespace cv
