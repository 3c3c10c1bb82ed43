/*
 * slam_stubs.h — stand-ins for the three ORB-SLAM2 classes the hot path's callers hand to ORBmatcher / Frame but whose
 * own translation units cannot be built here (TEST INFRASTRUCTURE ONLY).
 *
 *   MapPoint  (orb_slam2/include/MapPoint.h, src/MapPoint.cc)   needs Map.h + boost::serialization
 *   KeyFrame  (orb_slam2/include/KeyFrame.h, src/KeyFrame.cc)   needs Map.h, KeyFrameDatabase.h + boost::serialization
 *   Converter (orb_slam2/include/Converter.h)                     needs Eigen + g2o
 *
 * This header is force-included (-include) BEFORE the reference's headers and defines their include guards, so that
 * the reference's unmodified ORBmatcher.h / ORBmatcher.cc / Frame.h / Frame.cc / ORBextractor.{h,cc} compile against these
 * classes.  Every member keeps the reference's name, type and meaning; the few method bodies the matcher depends on
 * restate the reference line by line (cited).  Frame is NOT stubbed: Frame.h / Frame.cc are the reference's own.
 */
#ifndef ORB_REF_SLAM_STUBS_H
#define ORB_REF_SLAM_STUBS_H

#define MAPPOINT_H
#define KEYFRAME_H
#define CONVERTER_H

#include <map>
#include <mutex>
#include <set>
#include <vector>

#include <opencv2/core/core.hpp>

#include "Thirdparty/DBoW2/DBoW2/BowVector.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"

namespace ORB_SLAM2 {

using std::pair;   /* the reference's headers rely on `using namespace std` leaking out of its other headers (ORBmatcher.h:70,80) */
using std::vector;
using std::set;

class Frame;
class KeyFrame;
class Map;
class KeyFrameDatabase;

class MapPoint {
public:
    MapPoint() {}
    /* MapPoint.cc:57-63 */
    cv::Mat GetWorldPos() { return mWorldPos.clone(); }
    cv::Mat GetNormal() { return mNormalVector.clone(); }
    /* MapPoint.cc:364-368 */
    cv::Mat GetDescriptor() { return mDescriptor.clone(); }
    int Observations() { return nObs; }
    bool isBad() { return mbBad; }
    /* MapPoint.cc:441-452: 0.8f * mfMinDistance, 1.2f * mfMaxDistance */
    float GetMinDistanceInvariance() { return 0.8f * mfMinDistance; }
    float GetMaxDistanceInvariance() { return 1.2f * mfMaxDistance; }
    /* MapPoint.cc:455-488 */
    int PredictScale(const float& currentDist, KeyFrame* pKF);
    int PredictScale(const float& currentDist, Frame* pF);
    /* MapPoint.cc:370-385 */
    int GetIndexInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) ? (int)mObservations[pKF] : -1; }
    bool IsInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) != 0; }
    /* map mutation of ORBmatcher::Fuse (ORBmatcher.cc:955-973, 1088-1096): recorded for the test, not performed */
    void AddObservation(KeyFrame* pKF, size_t idx);
    void Replace(MapPoint* pMP);
    std::map<KeyFrame*, size_t> GetObservations() { return mObservations; }

public:
    long unsigned int mnId = 0;
    int nObs = 0;
    /* variables used by the tracking (MapPoint.h:93-100) */
    float mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0;
    bool mbTrackInView = false;
    int mnTrackScaleLevel = 0;
    float mTrackViewCos = 0;
    long unsigned int mnTrackReferenceForFrame = 0, mnLastFrameSeen = 0;
    long unsigned int mnFuseCandidateForKF = 0, mnLoopPointForKF = 0;

    /* the reference keeps these protected; the harness fills them directly */
    cv::Mat mWorldPos, mNormalVector, mDescriptor;
    std::map<KeyFrame*, size_t> mObservations;
    bool mbBad = false;
    float mfMinDistance = 0, mfMaxDistance = 0;
    /* what Fuse did to this point (harness read-back) */
    MapPoint* mpReplaced = nullptr;
    std::vector<pair<KeyFrame*, size_t> > mAddedObservations;
};

class KeyFrame {
public:
    /* KeyFrame.cc:37-65 (the fields the matcher reads) */
    explicit KeyFrame(Frame& F);

    void SetPose(const cv::Mat& Tcw);                      /* KeyFrame.cc:80-99 */
    cv::Mat GetRotation() { return Tcw.rowRange(0, 3).colRange(0, 3).clone(); }
    cv::Mat GetTranslation() { return Tcw.rowRange(0, 3).col(3).clone(); }
    cv::Mat GetCameraCenter() { return Ow.clone(); }
    cv::Mat GetPose() { return Tcw.clone(); }

    bool isBad() { return mbBad; }
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    MapPoint* GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
    std::set<MapPoint*> GetMapPoints();                    /* KeyFrame.cc:284-299 */
    void AddMapPoint(MapPoint* pMP, const size_t& idx) { mvpMapPoints[idx] = pMP; mAddedMapPoints.push_back(std::make_pair(pMP, idx)); }
    void ReplaceMapPointMatch(const size_t& idx, MapPoint* pMP) { mvpMapPoints[idx] = pMP; }

    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r) const;   /* KeyFrame.cc:700-739 */
    bool IsInImage(const float& x, const float& y) const { return (x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY); }   /* :742-745 */

public:
    static long unsigned int nNextId;
    long unsigned int mnId;
    const long unsigned int mnFrameId;
    const int mnGridCols, mnGridRows;
    const float mfGridElementWidthInv, mfGridElementHeightInv;
    const float fx, fy, cx, cy, invfx, invfy, mbf, mb, mThDepth;
    const int N;
    const std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    const std::vector<float> mvuRight, mvDepth;
    const cv::Mat mDescriptors;
    DBoW2::BowVector mBowVec;
    DBoW2::FeatureVector mFeatVec;
    const int mnScaleLevels;
    const float mfScaleFactor, mfLogScaleFactor;
    const std::vector<float> mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    const int mnMinX, mnMinY, mnMaxX, mnMaxY;   /* int in the reference too (KeyFrame.h:188-191) */
    const cv::Mat mK;

    std::vector<MapPoint*> mvpMapPoints;
    std::vector<std::vector<std::vector<size_t> > > mGrid;
    bool mbBad = false;
    std::vector<pair<MapPoint*, size_t> > mAddedMapPoints;   /* harness read-back of Fuse */

protected:
    cv::Mat Tcw, Twc, Ow;
};

class Converter {
public:
    /* Converter.cc:28-36 */
    static std::vector<cv::Mat> toDescriptorVector(const cv::Mat& Descriptors) {
        std::vector<cv::Mat> vDesc;
        vDesc.reserve(Descriptors.rows);
        for (int j = 0; j < Descriptors.rows; j++) vDesc.push_back(Descriptors.row(j));
        return vDesc;
    }
};

}  // namespace ORB_SLAM2

#endif
