// ORBmatcherArrays.h — array-level form of the loops in orb_slam2/src/ORBmatcher.cc (reference ORBmatcher.h:37-103): what each
// loop reads and writes, without the Frame / KeyFrame / MapPoint objects.  The signature-preserving drop-in (the reference's own
// ORBmatcher.h, unchanged) is host/ORBmatcher.cc.
// The reference methods take Frame / KeyFrame / MapPoint objects; their bodies reduce to the array routines below
// (what each loop reads and writes).  INTEGRATION.md shows the few lines of glue that keep the original
// SearchByProjection(Frame&, ...) signatures on top of them.
#ifndef ORBMATCHER_ARRAYS_H
#define ORBMATCHER_ARRAYS_H

#include <cstdint>
#include <utility>
#include <vector>

#include "ORBVocabulary.h"
#include "cv_compat.h"

namespace ORB_SLAM2 {

class ORBmatcherArrays {
public:
    ORBmatcherArrays(float nnratio = 0.6, bool checkOri = true, int device = 0) : mfNNratio(nnratio), mbCheckOrientation(checkOri), device_(device) {}

    // Computes the Hamming distance between two ORB descriptors (ORBmatcherArrays.cc:1649-1665).
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);

    struct TargetFrame {                       // the fields of the Frame being searched
        const cv::KeyPoint* keysUn; const uint8_t* descriptors; const float* uRight /* NULL: monocular */; int N;
        float minX, minY, maxX, maxY;          // Frame::mnMinX ...
    };
    struct Queries {                           // one entry per projected map point, in the reference's loop order
        int n; const float *u, *v, *radius; const int32_t *minLevel, *maxLevel; const uint8_t* descriptors;
        const float *uR, *erMax, *angle; const uint8_t *valid, *hasObservations;
    };
    // SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono)     ORBmatcherArrays.cc:1330-1472
    int SearchByProjectionLastFrame(const TargetFrame& F, std::vector<uint8_t>& taken, const Queries& q, int thDist,
                                    std::vector<int32_t>& matchOfQuery, std::vector<int32_t>& ownerOfTarget) const;
    // SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, th)           ORBmatcherArrays.cc:45-129
    int SearchByProjectionLocalPoints(const TargetFrame& F, std::vector<uint8_t>& taken, const Queries& q,
                                      std::vector<int32_t>& matchOfQuery, std::vector<int32_t>& ownerOfTarget) const;
    // SearchForInitialization(Frame &F1, Frame &F2, vbPrevMatched, vnMatches12, windowSize)    ORBmatcherArrays.cc:406-521
    // keys1Un / desc1 = F1, F = F2; vbPrevMatched is updated from the matches like the reference does (:515-518)
    int SearchForInitialization(const cv::KeyPoint* keys1Un, const uint8_t* desc1, int n1, const TargetFrame& F2,
                                std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize = 10) const;
    // best / second-best over explicit candidate lists (the gated loops of SearchForTriangulation / SearchBySim3):
    // candidates of query i = candIdx[candOff[i] .. candOff[i+1]) into desc2; returns (bestDist, bestIdx, secondDist) triples
    struct Best2 { int bestDist, secondDist; long long bestIdx, secondIdx; };
    void BestTwoOverCandidates(const uint8_t* desc1, int n1, const uint8_t* desc2, int n2, const std::vector<int32_t>& candOff,
                               const std::vector<int32_t>& candIdx, std::vector<Best2>& out) const;
    // inner loop of SearchByBoW over one vocabulary node (ORBmatcherArrays.cc:196-252)
    int MatchNode(const uint8_t* desc1, const float* angle1, int n1, const uint8_t* desc2, const float* angle2, int n2, int thDist,
                  std::vector<int32_t>& match12) const;

    // SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches) ORBmatcherArrays.cc:160-289 (keyframePair = false, valid2 = NULL) and
    // SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12)     ORBmatcherArrays.cc:524-657 (keyframePair = true): valid1 / valid2 = the
    // keypoint holds a good map point; match12[i] = index in set 2 or -1, match21[j] = index in set 1 or -1
    int SearchByBoW(const uint8_t* desc1, const float* angle1, const uint8_t* valid1, int n1, const DBoW2::FeatureVector& fv1,
                    const uint8_t* desc2, const float* angle2, const uint8_t* valid2, int n2, const DBoW2::FeatureVector& fv2,
                    bool keyframePair, std::vector<int32_t>& match12, std::vector<int32_t>& match21) const;

    // SearchForTriangulation(KeyFrame*, KeyFrame*, F12, vMatchedPairs, bOnlyStereo)  ORBmatcherArrays.cc:659-825.  hasMapPoint:
    // GetMapPoint(idx) != NULL; uRight = mvuRight (NULL: monocular); (ex, ey) = epipole (:665-673)
    struct KeyFrameView {
        const cv::KeyPoint* keysUn; const uint8_t* descriptors; const uint8_t* hasMapPoint; const float* uRight; int N;
        const DBoW2::FeatureVector* featVec;
    };
    int SearchForTriangulation(const KeyFrameView& kf1, const KeyFrameView& kf2, const float F12[9], float ex, float ey,
                               const std::vector<float>& scaleFactors2, const std::vector<float>& levelSigma2_2,
                               std::vector<std::pair<size_t, size_t> >& vMatchedPairs, bool bOnlyStereo) const;
    // SearchBySim3(KeyFrame*, KeyFrame*, vpMatches12, s12, R12, t12, th)  ORBmatcherArrays.cc:1104-1328: the caller projects the map
    // points of each keyframe into the other (the 3x3 float algebra of :1150-1190) and passes the windows as Queries
    // (minLevel / maxLevel unused: level = predicted octave in Queries::minLevel)
    int SearchBySim3(const TargetFrame& kf1, const TargetFrame& kf2, const Queries& q12, const Queries& q21,
                     std::vector<int32_t>& match12) const;
    // The search half of Fuse(KeyFrame*, vpMapPoints, th) (ORBmatcherArrays.cc:827-977) and, with invLevelSigma2 == NULL, of
    // Fuse(KeyFrame*, Scw, vpPoints, th, vpReplacePoint) (:979-1102): q = the projected map points (Queries::uR = u - bf*invz,
    // Queries::minLevel = predicted level); bestIdx / bestDist per point.  Replace / AddObservation stay with the caller.
    void FuseSearch(const TargetFrame& kf, const float* invLevelSigma2, int nlevels, const Queries& q, std::vector<int32_t>& bestIdx,
                    std::vector<int32_t>& bestDist) const;
    // MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:288-361) for many points: observations of point p = rows
    // [off[p], off[p+1]) of desc; bestDesc = npoints x 32
    static void ComputeDistinctiveDescriptors(const uint8_t* desc, const std::vector<int32_t>& off, std::vector<int32_t>& bestIdx,
                                              std::vector<uint8_t>& bestDesc, int device = 0);

    static const int TH_LOW = 50;
    static const int TH_HIGH = 100;
    static const int HISTO_LENGTH = 30;

protected:
    int Search(int mode, const TargetFrame& F, std::vector<uint8_t>& taken, const Queries& q, int thDist,
               std::vector<int32_t>& matchOfQuery, std::vector<int32_t>& ownerOfTarget) const;
    float mfNNratio;
    bool mbCheckOrientation;
    int device_;
};

// Frame::ComputeStereoMatches (Frame.cc:502-676): both extractors must have just processed the left / right image.
class ORBextractor;
int ComputeStereoMatches(ORBextractor& left, ORBextractor& right, const std::vector<cv::KeyPoint>& keysL, const cv::Mat& descL,
                         const std::vector<cv::KeyPoint>& keysR, const cv::Mat& descR, float bf, float b,
                         std::vector<float>& mvuRight, std::vector<float>& mvDepth);

}  // namespace ORB_SLAM2
#endif
