// Frame_ComputeStereoMatches.cc — drop-in replacement of ONE method of orb_slam2/src/Frame.cc:
// void Frame::ComputeStereoMatches()  (Frame.cc:502-676).  Frame.h and the rest of Frame.cc stay the reference's own; a
// maintainer deletes the body at Frame.cc:502-676 and adds this file (tests/cxx/Makefile does the equivalent at link time:
// it compiles the unmodified Frame.cc, weakens that one symbol and lets this definition win).
//
// Both extractors (host/ORBextractor.h) still hold the pyramids of the frame pair on the device, so nothing is uploaded
// but the two keypoint / descriptor sets; row-band candidate search, Hamming, 11x11 SAD over 11 shifts, parabola fit,
// disparity gate, depth and the 1.5 * 1.4 * median SAD cut run in liborb_b200.so (orb_stereo_match).
#include <stdexcept>
#include <string>

#include "ORBextractor.h"   // this repo's (host/ORBextractor.h), ahead of the reference's on the include path
#include "Frame.h"          // the reference's own
#include "orb_b200.h"

namespace ORB_SLAM2 {

void Frame::ComputeStereoMatches() {
    mvuRight = std::vector<float>(N, -1.0f);   // Frame.cc:504-505
    mvDepth = std::vector<float>(N, -1.0f);
    const int nR = (int)mvKeysRight.size();
    if (N == 0 || nR == 0) return;
    static_assert(sizeof(cv::KeyPoint) == sizeof(orb_kp), "cv::KeyPoint must be 28 bytes");
    int nmatches = 0;
    // mb is what the reference reads as minZ (Frame.cc:533): the constructor assigns it only AFTER this call
    // (Frame.cc:124), so it is the previous frame's value left in the same storage — passed through as it is.
    const int rc = orb_stereo_match(mpORBextractorLeft->context(), mpORBextractorRight->context(), reinterpret_cast<const orb_kp*>(mvKeys.data()),
                                    mDescriptors.data, N, reinterpret_cast<const orb_kp*>(mvKeysRight.data()), mDescriptorsRight.data, nR, mbf, mb,
                                    mvuRight.data(), mvDepth.data(), &nmatches);
    if (rc != ORB_OK) throw std::runtime_error(std::string("orb_stereo_match: ") + orb_last_error());
}

}  // namespace ORB_SLAM2
