// ORBmatcherArrays.cc — array-level form of the search loops of orb_slam2/src/ORBmatcher.cc and of Frame::ComputeStereoMatches
// (orb_slam2/src/Frame.cc:502-676): host shims over liborb_b200's C ABI.
#include "ORBmatcherArrays.h"

#include <cstring>
#include <stdexcept>
#include <string>

#include "../../include/orb_b200.h"
#include "ORBextractor.h"

namespace ORB_SLAM2 {

static void check(int rc, const char* what) {
    if (rc != ORB_OK) throw std::runtime_error(std::string(what) + ": " + orb_last_error());
}

int ORBmatcherArrays::DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {
    // one 256-bit popcount: stays on the host (MapPoint::ComputeDistinctiveDescriptors calls this in an O(N^2) loop,
    // MapPoint.cc:332); the GPU is for the batched entry points
    int dist = 0;
    for (int i = 0; i < 32; i += 8) {
        uint64_t x, y;
        memcpy(&x, a.data + i, 8); memcpy(&y, b.data + i, 8);
        dist += __builtin_popcountll(x ^ y);
    }
    return dist;
}

int ORBmatcherArrays::Search(int mode, const TargetFrame& F, std::vector<uint8_t>& taken, const Queries& q, int thDist,
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

int ORBmatcherArrays::SearchByProjectionLastFrame(const TargetFrame& F, std::vector<uint8_t>& taken, const Queries& q, int thDist,
                                            std::vector<int32_t>& matchOfQuery, std::vector<int32_t>& ownerOfTarget) const {
    return Search(ORB_MODE_TRACK_LAST, F, taken, q, thDist, matchOfQuery, ownerOfTarget);
}

int ORBmatcherArrays::SearchByProjectionLocalPoints(const TargetFrame& F, std::vector<uint8_t>& taken, const Queries& q,
                                              std::vector<int32_t>& matchOfQuery, std::vector<int32_t>& ownerOfTarget) const {
    return Search(ORB_MODE_LOCAL_POINTS, F, taken, q, TH_HIGH, matchOfQuery, ownerOfTarget);
}

int ORBmatcherArrays::SearchForInitialization(const cv::KeyPoint* keys1Un, const uint8_t* desc1, int n1, const TargetFrame& F2,
                                        std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize) const {
    std::vector<float> u(n1), v(n1), radius(n1, (float)windowSize), angle(n1);
    std::vector<int32_t> lvl(n1, 0);
    std::vector<uint8_t> valid(n1), taken;
    for (int i = 0; i < n1; ++i) {
        u[i] = vbPrevMatched[i].x; v[i] = vbPrevMatched[i].y; angle[i] = keys1Un[i].angle;
        valid[i] = keys1Un[i].octave <= 0;                        // level1 > 0: continue (ORBmatcherArrays.cc:423-425)
    }
    Queries q{n1, u.data(), v.data(), radius.data(), lvl.data(), lvl.data(), desc1, nullptr, nullptr, angle.data(), valid.data(), nullptr};
    std::vector<int32_t> m12, m21;
    const int nmatches = Search(ORB_MODE_INITIALIZATION, F2, taken, q, TH_LOW, m12, m21);
    vnMatches12.assign(m12.begin(), m12.end());
    for (int i = 0; i < n1; ++i)
        if (vnMatches12[i] >= 0) vbPrevMatched[i] = F2.keysUn[vnMatches12[i]].pt;   // ORBmatcherArrays.cc:515-518
    return nmatches;
}

void ORBmatcherArrays::BestTwoOverCandidates(const uint8_t* desc1, int n1, const uint8_t* desc2, int n2, const std::vector<int32_t>& candOff,
                                       const std::vector<int32_t>& candIdx, std::vector<Best2>& out) const {
    static_assert(sizeof(Best2) == sizeof(orb_top2), "Best2 mirrors orb_top2");
    out.resize(n1);
    check(orb_hamming_top2_csr(device_, desc1, n1, desc2, n2, candOff.data(), candIdx.data(), reinterpret_cast<orb_top2*>(out.data())),
          "orb_hamming_top2_csr");
}

int ORBmatcherArrays::MatchNode(const uint8_t* desc1, const float* angle1, int n1, const uint8_t* desc2, const float* angle2, int n2,
                          int thDist, std::vector<int32_t>& match12) const {
    match12.assign(n1, -1);
    int nmatches = 0;
    check(orb_match_bruteforce(device_, desc1, angle1, n1, desc2, angle2, n2, thDist, mfNNratio, mbCheckOrientation ? 1 : 0,
                               match12.data(), &nmatches), "orb_match_bruteforce");
    return nmatches;
}

static void flatten(const DBoW2::FeatureVector& fv, std::vector<int32_t>& node, std::vector<int32_t>& start, std::vector<int32_t>& feat) {
    node.clear(); start.assign(1, 0); feat.clear();
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it) {
        node.push_back((int32_t)it->first);
        feat.insert(feat.end(), it->second.begin(), it->second.end());
        start.push_back((int32_t)feat.size());
    }
}

int ORBmatcherArrays::SearchByBoW(const uint8_t* desc1, const float* angle1, const uint8_t* valid1, int n1, const DBoW2::FeatureVector& fv1,
                            const uint8_t* desc2, const float* angle2, const uint8_t* valid2, int n2, const DBoW2::FeatureVector& fv2,
                            bool keyframePair, std::vector<int32_t>& match12, std::vector<int32_t>& match21) const {
    std::vector<int32_t> n1v, s1v, f1v, n2v, s2v, f2v;
    flatten(fv1, n1v, s1v, f1v);
    flatten(fv2, n2v, s2v, f2v);
    match12.assign(n1, -1);
    match21.assign(n2, -1);
    int nmatches = 0;
    check(orb_search_by_bow(device_, desc1, angle1, valid1, n1, n1v.data(), s1v.data(), f1v.data(), (int)n1v.size(), desc2, angle2, valid2, n2,
                            n2v.data(), s2v.data(), f2v.data(), (int)n2v.size(), TH_LOW, keyframePair ? 1 : 0, mfNNratio,
                            mbCheckOrientation ? 1 : 0, match12.data(), match21.data(), &nmatches), "orb_search_by_bow");
    return nmatches;
}

int ORBmatcherArrays::SearchForTriangulation(const KeyFrameView& kf1, const KeyFrameView& kf2, const float F12[9], float ex, float ey,
                                       const std::vector<float>& scaleFactors2, const std::vector<float>& levelSigma2_2,
                                       std::vector<std::pair<size_t, size_t> >& vMatchedPairs, bool bOnlyStereo) const {
    std::vector<int32_t> n1v, s1v, f1v, n2v, s2v, f2v, m12(kf1.N, -1);
    flatten(*kf1.featVec, n1v, s1v, f1v);
    flatten(*kf2.featVec, n2v, s2v, f2v);
    int nmatches = 0;
    check(orb_search_for_triangulation(device_, reinterpret_cast<const orb_kp*>(kf1.keysUn), kf1.descriptors, kf1.hasMapPoint, kf1.uRight, kf1.N,
                                       n1v.data(), s1v.data(), f1v.data(), (int)n1v.size(), reinterpret_cast<const orb_kp*>(kf2.keysUn),
                                       kf2.descriptors, kf2.hasMapPoint, kf2.uRight, kf2.N, n2v.data(), s2v.data(), f2v.data(), (int)n2v.size(),
                                       F12, ex, ey, scaleFactors2.data(), levelSigma2_2.data(), (int)scaleFactors2.size(), bOnlyStereo ? 1 : 0,
                                       mbCheckOrientation ? 1 : 0, m12.data(), &nmatches), "orb_search_for_triangulation");
    vMatchedPairs.clear();                                              // ORBmatcherArrays.cc:812-822
    vMatchedPairs.reserve(nmatches);
    for (size_t i = 0; i < m12.size(); ++i)
        if (m12[i] >= 0) vMatchedPairs.push_back(std::make_pair(i, (size_t)m12[i]));
    return nmatches;
}

int ORBmatcherArrays::SearchBySim3(const TargetFrame& kf1, const TargetFrame& kf2, const Queries& q12, const Queries& q21,
                             std::vector<int32_t>& match12) const {
    match12.assign(kf1.N, -1);
    const float b1[4] = {kf1.minX, kf1.minY, kf1.maxX, kf1.maxY}, b2[4] = {kf2.minX, kf2.minY, kf2.maxX, kf2.maxY};
    int nfound = 0;
    check(orb_search_by_sim3(device_, reinterpret_cast<const orb_kp*>(kf1.keysUn), kf1.descriptors, kf1.N, b1,
                             reinterpret_cast<const orb_kp*>(kf2.keysUn), kf2.descriptors, kf2.N, b2, q12.u, q12.v, q12.radius, q12.minLevel,
                             q12.descriptors, q12.valid, q21.u, q21.v, q21.radius, q21.minLevel, q21.descriptors, q21.valid, TH_HIGH,
                             match12.data(), &nfound), "orb_search_by_sim3");
    return nfound;
}

void ORBmatcherArrays::FuseSearch(const TargetFrame& kf, const float* invLevelSigma2, int nlevels, const Queries& q, std::vector<int32_t>& bestIdx,
                            std::vector<int32_t>& bestDist) const {
    bestIdx.assign(q.n, -1);
    bestDist.assign(q.n, 256);
    const float b[4] = {kf.minX, kf.minY, kf.maxX, kf.maxY};
    check(orb_fuse_search(device_, reinterpret_cast<const orb_kp*>(kf.keysUn), kf.descriptors, kf.uRight, kf.N, b, invLevelSigma2, nlevels, q.n,
                          q.u, q.v, q.uR, q.radius, q.minLevel, q.descriptors, q.valid, bestIdx.data(), bestDist.data()), "orb_fuse_search");
}

void ORBmatcherArrays::ComputeDistinctiveDescriptors(const uint8_t* desc, const std::vector<int32_t>& off, std::vector<int32_t>& bestIdx,
                                               std::vector<uint8_t>& bestDesc, int device) {
    const int np = off.empty() ? 0 : (int)off.size() - 1;
    bestIdx.assign(np, -1);
    bestDesc.assign((size_t)np * 32, 0);
    check(orb_distinctive_descriptors(device, desc, off.data(), np, bestIdx.data(), bestDesc.data()), "orb_distinctive_descriptors");
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
