// ORBmatcher.cc — replacement of the search loops of orb_slam2/src/ORBmatcher.cc and of Frame::ComputeStereoMatches
// (orb_slam2/src/Frame.cc:502-676): host shims over liborb_b200's C ABI.
#include "ORBmatcher.h"

#include <stdexcept>
#include <string>

#include "../../include/orb_b200.h"
#include "ORBextractor.h"

namespace ORB_SLAM2 {

static void check(int rc, const char* what) {
    if (rc != ORB_OK) throw std::runtime_error(std::string(what) + ": " + orb_last_error());
}

int ORBmatcher::DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {
    orb_top2 r;
    check(orb_hamming_top2(0, a.data, 1, b.data, 1, &r), "orb_hamming_top2");
    return r.best_dist;
}

int ORBmatcher::Search(int mode, const TargetFrame& F, std::vector<uint8_t>& taken, const Queries& q, int thDist,
                       std::vector<int32_t>& matchOfQuery, std::vector<int32_t>& ownerOfTarget) const {
    orb_search_params prm;
    prm.mode = mode; prm.th_dist = thDist; prm.nn_ratio = mfNNratio; prm.check_orientation = mbCheckOrientation ? 1 : 0;
    prm.min_x = F.minX; prm.min_y = F.minY; prm.max_x = F.maxX; prm.max_y = F.maxY;
    taken.resize(F.N, 0);
    matchOfQuery.assign(q.n, -1);
    ownerOfTarget.assign(F.N, -1);
    int nmatches = 0;
    static_assert(sizeof(cv::KeyPoint) == sizeof(orb_kp), "cv::KeyPoint must be 28 bytes");
    check(orb_search_by_projection(device_, &prm, reinterpret_cast<const orb_kp*>(F.keysUn), F.descriptors, F.uRight, F.N, taken.data(),
                                   q.n, q.u, q.v, q.radius, q.minLevel, q.maxLevel, q.descriptors, q.uR, q.erMax, q.angle, q.valid,
                                   q.hasObservations, matchOfQuery.data(), ownerOfTarget.data(), &nmatches),
          "orb_search_by_projection");
    return nmatches;
}

int ORBmatcher::SearchByProjectionLastFrame(const TargetFrame& F, std::vector<uint8_t>& taken, const Queries& q, int thDist,
                                            std::vector<int32_t>& matchOfQuery, std::vector<int32_t>& ownerOfTarget) const {
    return Search(ORB_MODE_TRACK_LAST, F, taken, q, thDist, matchOfQuery, ownerOfTarget);
}

int ORBmatcher::SearchByProjectionLocalPoints(const TargetFrame& F, std::vector<uint8_t>& taken, const Queries& q,
                                              std::vector<int32_t>& matchOfQuery, std::vector<int32_t>& ownerOfTarget) const {
    return Search(ORB_MODE_LOCAL_POINTS, F, taken, q, TH_HIGH, matchOfQuery, ownerOfTarget);
}

int ORBmatcher::MatchNode(const uint8_t* desc1, const float* angle1, int n1, const uint8_t* desc2, const float* angle2, int n2,
                          int thDist, std::vector<int32_t>& match12) const {
    match12.assign(n1, -1);
    int nmatches = 0;
    check(orb_match_bruteforce(device_, desc1, angle1, n1, desc2, angle2, n2, thDist, mfNNratio, mbCheckOrientation ? 1 : 0,
                               match12.data(), &nmatches), "orb_match_bruteforce");
    return nmatches;
}

int ComputeStereoMatches(ORBextractor& left, ORBextractor& right, const std::vector<cv::KeyPoint>& keysL, const cv::Mat& descL,
                         const std::vector<cv::KeyPoint>& keysR, const cv::Mat& descR, float bf, float b,
                         std::vector<float>& mvuRight, std::vector<float>& mvDepth) {
    const int N = (int)keysL.size();
    mvuRight.assign(N, -1.0f);                                    // Frame.cc:504-505
    mvDepth.assign(N, -1.0f);
    int nmatches = 0;
    check(orb_stereo_match(left.context(), right.context(), reinterpret_cast<const orb_kp*>(keysL.data()), descL.data, N,
                           reinterpret_cast<const orb_kp*>(keysR.data()), descR.data, (int)keysR.size(), bf, b, mvuRight.data(),
                           mvDepth.data(), &nmatches), "orb_stereo_match");
    return nmatches;
}

}  // namespace ORB_SLAM2
