/* slam_stubs.cpp — bodies of the MapPoint / KeyFrame stand-ins (TEST INFRASTRUCTURE, see slam_stubs.h).  Compiled with
 * the reference's real Frame.h on the include path. */
#include "slam_stubs.h"

#include <cmath>

#include "Frame.h"

namespace ORB_SLAM2 {

long unsigned int KeyFrame::nNextId = 0;

/* MapPoint.cc:455-488 */
int MapPoint::PredictScale(const float& currentDist, KeyFrame* pKF) {
    float ratio = mfMaxDistance / currentDist;
    int nScale = ceil(log(ratio) / pKF->mfLogScaleFactor);
    if (nScale < 0) nScale = 0;
    else if (nScale >= pKF->mnScaleLevels) nScale = pKF->mnScaleLevels - 1;
    return nScale;
}
int MapPoint::PredictScale(const float& currentDist, Frame* pF) {
    float ratio = mfMaxDistance / currentDist;
    int nScale = ceil(log(ratio) / pF->mfLogScaleFactor);
    if (nScale < 0) nScale = 0;
    else if (nScale >= pF->mnScaleLevels) nScale = pF->mnScaleLevels - 1;
    return nScale;
}

void MapPoint::AddObservation(KeyFrame* pKF, size_t idx) {   /* MapPoint.cc:79-93, without the nObs bookkeeping side effects on other points */
    if (mObservations.count(pKF)) return;
    mObservations[pKF] = idx;
    if (pKF->mvuRight[idx] >= 0) nObs += 2; else nObs++;
    mAddedObservations.push_back(std::make_pair(pKF, idx));
}
void MapPoint::Replace(MapPoint* pMP) {                      /* MapPoint.cc:186-228: recorded only */
    if (pMP->mnId == this->mnId) return;
    mpReplaced = pMP;
    mbBad = true;
}

/* KeyFrame.cc:37-65 */
KeyFrame::KeyFrame(Frame& F)
    : mnFrameId(F.mnId), mnGridCols(FRAME_GRID_COLS), mnGridRows(FRAME_GRID_ROWS), mfGridElementWidthInv(F.mfGridElementWidthInv),
      mfGridElementHeightInv(F.mfGridElementHeightInv), fx(F.fx), fy(F.fy), cx(F.cx), cy(F.cy), invfx(F.invfx), invfy(F.invfy), mbf(F.mbf),
      mb(F.mb), mThDepth(F.mThDepth), N(F.N), mvKeys(F.mvKeys), mvKeysUn(F.mvKeysUn), mvuRight(F.mvuRight), mvDepth(F.mvDepth),
      mDescriptors(F.mDescriptors.clone()), mBowVec(F.mBowVec), mFeatVec(F.mFeatVec), mnScaleLevels(F.mnScaleLevels),
      mfScaleFactor(F.mfScaleFactor), mfLogScaleFactor(F.mfLogScaleFactor), mvScaleFactors(F.mvScaleFactors), mvLevelSigma2(F.mvLevelSigma2),
      mvInvLevelSigma2(F.mvInvLevelSigma2), mnMinX(F.mnMinX), mnMinY(F.mnMinY), mnMaxX(F.mnMaxX), mnMaxY(F.mnMaxY), mK(F.mK),
      mvpMapPoints(F.mvpMapPoints) {
    mnId = nNextId++;
    mGrid.resize(mnGridCols);
    for (int i = 0; i < mnGridCols; i++) {
        mGrid[i].resize(mnGridRows);
        for (int j = 0; j < mnGridRows; j++) mGrid[i][j] = F.mGrid[i][j];
    }
    SetPose(F.mTcw);
}

/* KeyFrame.cc:80-99 */
void KeyFrame::SetPose(const cv::Mat& Tcw_) {
    if (Tcw_.empty()) return;
    Tcw_.copyTo(Tcw);
    cv::Mat Rcw = Tcw.rowRange(0, 3).colRange(0, 3);
    cv::Mat tcw = Tcw.rowRange(0, 3).col(3);
    cv::Mat Rwc = Rcw.t();
    Ow = -Rwc * tcw;
    Twc = cv::Mat::eye(4, 4, Tcw.type());
    cv::Mat dR = Twc.rowRange(0, 3).colRange(0, 3), dt = Twc.rowRange(0, 3).col(3);
    Rwc.copyTo(dR);
    Ow.copyTo(dt);
}

/* KeyFrame.cc:284-299 */
std::set<MapPoint*> KeyFrame::GetMapPoints() {
    std::set<MapPoint*> s;
    for (size_t i = 0, iend = mvpMapPoints.size(); i < iend; i++) {
        if (!mvpMapPoints[i]) continue;
        MapPoint* pMP = mvpMapPoints[i];
        if (!pMP->isBad()) s.insert(pMP);
    }
    return s;
}

/* KeyFrame.cc:700-739 */
std::vector<size_t> KeyFrame::GetFeaturesInArea(const float& x, const float& y, const float& r) const {
    std::vector<size_t> vIndices;
    vIndices.reserve(N);
    const int nMinCellX = std::max(0, (int)floor((x - mnMinX - r) * mfGridElementWidthInv));
    if (nMinCellX >= mnGridCols) return vIndices;
    const int nMaxCellX = std::min((int)mnGridCols - 1, (int)ceil((x - mnMinX + r) * mfGridElementWidthInv));
    if (nMaxCellX < 0) return vIndices;
    const int nMinCellY = std::max(0, (int)floor((y - mnMinY - r) * mfGridElementHeightInv));
    if (nMinCellY >= mnGridRows) return vIndices;
    const int nMaxCellY = std::min((int)mnGridRows - 1, (int)ceil((y - mnMinY + r) * mfGridElementHeightInv));
    if (nMaxCellY < 0) return vIndices;
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++) {
        for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
            const std::vector<size_t>& vCell = mGrid[ix][iy];
            for (size_t j = 0, jend = vCell.size(); j < jend; j++) {
                const cv::KeyPoint& kpUn = mvKeysUn[vCell[j]];
                const float distx = kpUn.pt.x - x;
                const float disty = kpUn.pt.y - y;
                if (fabs(distx) < r && fabs(disty) < r) vIndices.push_back(vCell[j]);
            }
        }
    }
    return vIndices;
}

}  // namespace ORB_SLAM2
