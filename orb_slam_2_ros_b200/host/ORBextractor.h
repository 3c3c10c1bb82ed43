// ORBextractor.h — drop-in replacement of orb_slam2/include/ORBextractor.h (reference :46-107): the same class name,
// constructor, operator(), getters and public mvImagePyramid, implemented over the C ABI of liborb_b200.so
// (include/orb_b200.h).  ExtractorNode and the protected helpers of the reference are gone: the quadtree, FAST,
// orientation and descriptor code all run on the GPU.
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <vector>

#include "cv_compat.h"

struct orb_ctx;

namespace ORB_SLAM2 {

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
    ~ORBextractor();
    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;

    // Compute the ORB features and descriptors on an image.  Mask is ignored (like the reference, ORBextractor.h:58).
    void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint>& keypoints, cv::OutputArray descriptors);

    int inline GetLevels() { return nlevels; }
    float inline GetScaleFactor() { return scaleFactor; }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    // interior ROIs (step = cols + 38) of the bordered level buffers, refreshed by every operator() call
    std::vector<cv::Mat> mvImagePyramid;

    // the device context, for Frame::ComputeStereoMatches (StereoMatcher.h) which reads both pyramids on the GPU
    orb_ctx* context() const { return ctx_; }
    // false: skip the device->host copy of the pyramid in operator() (callers that never touch mvImagePyramid)
    void SetPyramidDownload(bool on) { download_pyramid_ = on; }

protected:
    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;
    std::vector<int> mnFeaturesPerLevel;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;

private:
    orb_ctx* ctx_ = nullptr;
    bool download_pyramid_ = true;
    std::vector<cv::Mat> bordered_;   // headers of the bordered levels inside pyramid_host_
    std::vector<unsigned char> pyramid_host_;   // all levels of the last frame in the device layout (one copy pass per call)
};

}  // namespace ORB_SLAM2
#endif
