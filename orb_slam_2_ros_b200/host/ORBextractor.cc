// ORBextractor.cc — replacement of orb_slam2/src/ORBextractor.cc: a thin host shim over liborb_b200's C ABI.
#include "ORBextractor.h"

#include <stdexcept>
#include <string>

#include <cstdlib>
#include <cstring>

#include "../../include/orb_b200.h"

namespace ORB_SLAM2 {

static void check(int rc, const char* what) {
    if (rc != ORB_OK) throw std::runtime_error(std::string(what) + ": " + orb_last_error());   // OpenCV would throw cv::Exception
}

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST) {
    // the reference constructor has no notion of a device: ORB_B200_DEVICE selects it (default 0)
    const char* dev = getenv("ORB_B200_DEVICE");
    check(orb_create(&ctx_, nfeatures, _scaleFactor, nlevels, iniThFAST, minThFAST, dev ? atoi(dev) : 0, /*max_batch*/ 1), "orb_create");
    mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels);
    mvInvLevelSigma2.resize(nlevels); mnFeaturesPerLevel.resize(nlevels);
    check(orb_get_tables(ctx_, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(), mvInvLevelSigma2.data(),
                         mnFeaturesPerLevel.data()), "orb_get_tables");
    mvImagePyramid.resize(nlevels);
    bordered_.resize(nlevels);
}

ORBextractor::~ORBextractor() { orb_destroy(ctx_); }

void ORBextractor::operator()(cv::InputArray _image, cv::InputArray /*_mask*/, std::vector<cv::KeyPoint>& _keypoints,
                              cv::OutputArray _descriptors) {
    if (_image.empty()) return;                                  // ORBextractor.cc:1086-1087
    cv::Mat image = _image.getMat();
    // the reference only asserts CV_8UC1 (ORBextractor.cc:1090, compiled out in Release) and would then read garbage; a
    // colour or 16-bit Mat here is a caller bug, so fail loudly instead of extracting from misinterpreted bytes
    if (image.type() != CV_8UC1) throw std::runtime_error("ORBextractor: image must be CV_8UC1 (use orb_extract_batch_pix for colour input)");
    const int cap = orb_max_keypoints(ctx_);
    _keypoints.resize(cap);
    static_assert(sizeof(cv::KeyPoint) == sizeof(orb_kp), "cv::KeyPoint must be 28 bytes");
    cv::Mat desc(cap, 32, CV_8U);
    int n = 0;
    check(orb_extract(ctx_, image.data, image.cols, image.rows, (size_t)image.step, reinterpret_cast<orb_kp*>(_keypoints.data()),
                      desc.data, cap, &n), "orb_extract");
    _keypoints.resize(n);
    if (n == 0) {
        _descriptors.release();                                   // ORBextractor.cc:1109
    } else {
        _descriptors.create(n, 32, CV_8U);                        // ORBextractor.cc:1112
        cv::Mat out = _descriptors.getMat();
        memcpy(out.data, desc.data, (size_t)n * 32);
    }
    if (download_pyramid_) {
        // public mvImagePyramid (ORBextractor.h:85): all levels in one pass — 8 asynchronous 2-D copies into one host buffer and
        // ONE synchronisation; the Mats are headers over that buffer: the ROI (step = w + 38) of the bordered buffer, as the
        // reference builds it (ORBextractor.cc:1161-1165)
        size_t need = 0;
        check(orb_pyramid_levels(ctx_, 0, nullptr, 0, nullptr, nullptr, &need), "orb_pyramid_levels");
        if (pyramid_host_.size() < need) pyramid_host_.resize(need);
        std::vector<size_t> off(nlevels), pitch(nlevels);
        check(orb_pyramid_levels(ctx_, 0, pyramid_host_.data(), pyramid_host_.size(), off.data(), pitch.data(), nullptr), "orb_pyramid_levels");
        for (int l = 0; l < nlevels; ++l) {
            int w = 0, h = 0;
            check(orb_level_dims(ctx_, l, &w, &h), "orb_level_dims");
            bordered_[l] = cv::Mat(h + 38, w + 38, CV_8U, pyramid_host_.data() + off[l], pitch[l]);
#if defined(ORB_B200_USE_OPENCV) || __has_include(<opencv2/core/core.hpp>)
            mvImagePyramid[l] = bordered_[l](cv::Rect(19, 19, w, h));
#else
            mvImagePyramid[l] = bordered_[l].roi(19, 19, w, h);
#endif
        }
    }
}

}  // namespace ORB_SLAM2
