// ORBmatcher.cc — drop-in replacement of orb_slam2/src/ORBmatcher.cc.
//
// It implements the class declared by the reference's OWN, unmodified orb_slam2/include/ORBmatcher.h (:37-103): every
// signature stays as it is (Frame&, KeyFrame*, MapPoint*, cv::Mat), so Tracking, LocalMapping and LoopClosing keep calling
// what they call today.  Each routine does, on the host, exactly what the reference does BEFORE and AFTER its Hamming
// loop — the per-map-point projection / frustum / scale gates in the reference's own float arithmetic, and the scatter
// of the result into Frame::mvpMapPoints / vpMatches / the map — and hands the loop itself (window query in
// GetFeaturesInArea order, 256-bit distances, best / second best, the sequential "already matched" rule, ratio test,
// rotation histogram) to liborb_b200.so through the C ABI of include/orb_b200.h.
//
// Parity: tests/test_gpu_dropin.py drives this file and the reference's own ORBmatcher.cc (oracle/_ref) through the same
// harness with the same Frames / KeyFrames / MapPoints and compares every output.
//
// Float arithmetic: per-point camera transforms use `R * p + t` on plain floats, summed left to right with one rounding
// per operation — the arithmetic of cv::gemm's small-matrix path (OpenCV 4.x matmul: 3-term float inner product, then
// one correctly rounded add of t), which is what `Rcw*p3Dw+tcw` evaluates to.  One-off pose algebra (camera centres,
// Sim3 decomposition) is written with the same cv::Mat expressions as the reference so that it follows whatever OpenCV
// the host links.  Build with -ffp-contract=off (the reference builds without -march, hence without FMA).
#include "ORBmatcher.h"

#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

#include "orb_b200.h"

using namespace std;

namespace ORB_SLAM2 {

const int ORBmatcher::TH_HIGH = 100;   // ORBmatcher.cc:37-39
const int ORBmatcher::TH_LOW = 50;
const int ORBmatcher::HISTO_LENGTH = 30;

namespace {

int device() {   // the reference class has no notion of a device: ORB_B200_DEVICE selects it (default 0)
    static const int d = [] { const char* e = getenv("ORB_B200_DEVICE"); return e ? atoi(e) : 0; }();
    return d;
}

void check(int rc, const char* what) {
    if (rc != ORB_OK) throw std::runtime_error(std::string(what) + ": " + orb_last_error());   // OpenCV would throw cv::Exception
}

static_assert(sizeof(cv::KeyPoint) == sizeof(orb_kp), "cv::KeyPoint must be 28 bytes");
inline const orb_kp* kp_ptr(const std::vector<cv::KeyPoint>& v) { return reinterpret_cast<const orb_kp*>(v.data()); }

// rows of a CV_8U N x 32 matrix as one contiguous block (mDescriptors always is; an ROI is repacked)
struct DescBlock {
    std::vector<uint8_t> copy;
    const uint8_t* p = nullptr;
    explicit DescBlock(const cv::Mat& m) {
        if (m.rows == 0) return;
        if (m.isContinuous()) { p = m.data; return; }
        copy.resize((size_t)m.rows * 32);
        for (int i = 0; i < m.rows; ++i) memcpy(&copy[(size_t)i * 32], m.ptr(i), 32);
        p = copy.data();
    }
};

// [R | t] of a camera as plain floats; apply() = the small-matrix cv::gemm arithmetic (see the file header)
struct Rigid {
    float R[9], t[3];
    Rigid(const cv::Mat& Rm, const cv::Mat& tm) {
        for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) R[3 * i + j] = Rm.at<float>(i, j); t[i] = tm.at<float>(i); }
    }
    void apply(const float* p, float* o) const {
        for (int i = 0; i < 3; ++i) {
            float s = R[3 * i] * p[0];
            s = s + R[3 * i + 1] * p[1];
            s = s + R[3 * i + 2] * p[2];
            o[i] = s + t[i];
        }
    }
};
inline void get3(const cv::Mat& m, float* o) { o[0] = m.at<float>(0); o[1] = m.at<float>(1); o[2] = m.at<float>(2); }
// cv::norm(Mat) of a float 3-vector: squares and sum in double, one rounding to float at the assignment
inline float norm3(const float* v) { return (float)std::sqrt((double)v[0] * v[0] + (double)v[1] * v[1] + (double)v[2] * v[2]); }
// cv::Mat::dot on floats: products and sum in double
inline double dot3(const float* a, const float* b) { return (double)a[0] * b[0] + (double)a[1] * b[1] + (double)a[2] * b[2]; }

// one entry per projected map point, in the reference's loop order
struct Queries {
    std::vector<float> u, v, radius, uR, erMax, angle;
    std::vector<int32_t> minLevel, maxLevel;
    std::vector<uint8_t> desc, valid, obs;
    explicit Queries(size_t n) : u(n, 0.f), v(n, 0.f), radius(n, 0.f), uR(n, 0.f), erMax(n, 0.f), angle(n, 0.f), minLevel(n, 0), maxLevel(n, 0),
                                  desc(n * 32, 0), valid(n, 0), obs(n, 1) {}
    size_t size() const { return u.size(); }
    void setDescriptor(size_t i, const cv::Mat& d) { memcpy(&desc[i * 32], d.data, 32); }
};

struct SearchResult { std::vector<int32_t> matchOfQuery, ownerOfTarget; int nmatches = 0; };

// the windowed search on the device (include/orb_b200.h: orb_search_by_projection)
SearchResult search(int mode, int thDist, float nnRatio, bool checkOri, const std::vector<cv::KeyPoint>& keysUn, const cv::Mat& descriptors,
                    const float* uRight, const float bounds[4], std::vector<uint8_t>& taken, const Queries& q, bool stereoGate) {
    const int n = (int)keysUn.size(), nq = (int)q.size();
    SearchResult r;
    r.matchOfQuery.assign(nq, -1);
    r.ownerOfTarget.assign(n, -1);
    if (n == 0 || nq == 0) return r;
    orb_search_params prm;
    prm.mode = mode; prm.th_dist = thDist; prm.nn_ratio = nnRatio; prm.check_orientation = checkOri ? 1 : 0;
    prm.min_x = bounds[0]; prm.min_y = bounds[1]; prm.max_x = bounds[2]; prm.max_y = bounds[3];
    DescBlock d(descriptors);
    check(orb_search_by_projection(device(), &prm, kp_ptr(keysUn), d.p, stereoGate ? uRight : nullptr, n, taken.data(), nq, q.u.data(), q.v.data(),
                                   q.radius.data(), q.minLevel.data(), q.maxLevel.data(), q.desc.data(), stereoGate ? q.uR.data() : nullptr,
                                   stereoGate ? q.erMax.data() : nullptr, q.angle.data(), q.valid.data(), q.obs.data(), r.matchOfQuery.data(),
                                   r.ownerOfTarget.data(), &r.nmatches),
          "orb_search_by_projection");
    return r;
}

// DBoW2::FeatureVector (std::map<NodeId, vector<unsigned>>) -> CSR
struct FeatCsr {
    std::vector<int32_t> node, start, feat;
    explicit FeatCsr(const DBoW2::FeatureVector& fv) {
        start.push_back(0);
        for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it) {
            node.push_back((int32_t)it->first);
            for (size_t k = 0; k < it->second.size(); ++k) feat.push_back((int32_t)it->second[k]);
            start.push_back((int32_t)feat.size());
        }
    }
    int n() const { return (int)node.size(); }
};

std::vector<float> angles(const std::vector<cv::KeyPoint>& k) {
    std::vector<float> a(k.size());
    for (size_t i = 0; i < k.size(); ++i) a[i] = k[i].angle;
    return a;
}

// the scale / viewing gates shared by the keyframe-projection routines (ORBmatcher.cc:335-352, 878-896, 1024-1043):
// distance within the point's scale-invariance range and viewing direction within 60 degrees of the mean normal
bool scaleAndViewGate(MapPoint* pMP, const float* PO, float& dist3D, bool checkNormal) {
    const float maxDistance = pMP->GetMaxDistanceInvariance();
    const float minDistance = pMP->GetMinDistanceInvariance();
    dist3D = norm3(PO);
    if (dist3D < minDistance || dist3D > maxDistance) return false;
    if (checkNormal) {
        float Pn[3];
        get3(pMP->GetNormal(), Pn);
        if (dot3(PO, Pn) < 0.5 * dist3D) return false;
    }
    return true;
}

}  // namespace

ORBmatcher::ORBmatcher(float nnratio, bool checkOri) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}

// ORBmatcher.cc:1649-1665.  One 256-bit popcount stays on the host: MapPoint::ComputeDistinctiveDescriptors calls this in an
// O(N^2) loop (MapPoint.cc:332); the GPU serves the batched loops.
int ORBmatcher::DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {
    const uint8_t* pa = a.ptr<uint8_t>();
    const uint8_t* pb = b.ptr<uint8_t>();
    int dist = 0;
    for (int i = 0; i < 32; i += 8) {
        uint64_t x, y;
        memcpy(&x, pa + i, 8);
        memcpy(&y, pb + i, 8);
        dist += __builtin_popcountll(x ^ y);
    }
    return dist;
}

float ORBmatcher::RadiusByViewingCos(const float& viewCos) { return viewCos > 0.998 ? 2.5 : 4.0; }   // ORBmatcher.cc:131-137

// kept for callers that derive from ORBmatcher (protected in the reference header); the device applies the same test
bool ORBmatcher::CheckDistEpipolarLine(const cv::KeyPoint& kp1, const cv::KeyPoint& kp2, const cv::Mat& F12, const KeyFrame* pKF2) {
    const float a = kp1.pt.x * F12.at<float>(0, 0) + kp1.pt.y * F12.at<float>(1, 0) + F12.at<float>(2, 0);   // ORBmatcher.cc:140-157
    const float b = kp1.pt.x * F12.at<float>(0, 1) + kp1.pt.y * F12.at<float>(1, 1) + F12.at<float>(2, 1);
    const float c = kp1.pt.x * F12.at<float>(0, 2) + kp1.pt.y * F12.at<float>(1, 2) + F12.at<float>(2, 2);
    const float num = a * kp2.pt.x + b * kp2.pt.y + c;
    const float den = a * a + b * b;
    if (den == 0) return false;
    const float dsqr = num * num / den;
    return dsqr < 3.84 * pKF2->mvLevelSigma2[kp2.octave];
}

void ORBmatcher::ComputeThreeMaxima(vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3) {   // ORBmatcher.cc:1603-1644
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = histo[i].size();
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) ind3 = -1;
}

// ---- Tracking::SearchLocalPoints -> SearchByProjection(Frame&, vpMapPoints, th)            (ORBmatcher.cc:45-129) ----
int ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, const float th) {
    const bool bFactor = th != 1.0;
    Queries q(vpMapPoints.size());
    for (size_t iMP = 0; iMP < vpMapPoints.size(); iMP++) {
        MapPoint* pMP = vpMapPoints[iMP];
        if (!pMP->mbTrackInView || pMP->isBad()) continue;                         // :52-57
        const int nPredictedLevel = pMP->mnTrackScaleLevel;
        float r = RadiusByViewingCos(pMP->mTrackViewCos);                          // :62-65
        if (bFactor) r *= th;
        const float window = r * F.mvScaleFactors[nPredictedLevel];
        q.valid[iMP] = 1;
        q.u[iMP] = pMP->mTrackProjX; q.v[iMP] = pMP->mTrackProjY; q.radius[iMP] = window;
        q.minLevel[iMP] = nPredictedLevel - 1; q.maxLevel[iMP] = nPredictedLevel;  // :68
        q.uR[iMP] = pMP->mTrackProjXR; q.erMax[iMP] = window;                      // :91-96
        q.obs[iMP] = pMP->Observations() > 0;                                      // its assignment hides the keypoint from later points only then (:87-89)
        q.setDescriptor(iMP, pMP->GetDescriptor());
    }
    std::vector<uint8_t> taken(F.N, 0);
    for (int j = 0; j < F.N; j++) taken[j] = F.mvpMapPoints[j] && F.mvpMapPoints[j]->Observations() > 0;   // :87-89
    const float bounds[4] = {Frame::mnMinX, Frame::mnMinY, Frame::mnMaxX, Frame::mnMaxY};
    SearchResult r = search(ORB_MODE_LOCAL_POINTS, TH_HIGH, mfNNratio, false, F.mvKeysUn, F.mDescriptors, F.mvuRight.data(), bounds, taken, q, true);
    for (int j = 0; j < F.N; j++)
        if (r.ownerOfTarget[j] >= 0) F.mvpMapPoints[j] = vpMapPoints[r.ownerOfTarget[j]];   // :123
    return r.nmatches;
}

// ---- TrackWithMotionModel -> SearchByProjection(CurrentFrame, LastFrame, th, bMono)        (ORBmatcher.cc:1330-1472) ----
int ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono) {
    const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0, 3).colRange(0, 3);
    const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0, 3).col(3);
    const cv::Mat twc = -Rcw.t() * tcw;
    const cv::Mat Rlw = LastFrame.mTcw.rowRange(0, 3).colRange(0, 3);
    const cv::Mat tlw = LastFrame.mTcw.rowRange(0, 3).col(3);
    const cv::Mat tlc = Rlw * twc + tlw;
    const bool bForward = tlc.at<float>(2) > CurrentFrame.mb && !bMono;           // :1352-1353
    const bool bBackward = -tlc.at<float>(2) > CurrentFrame.mb && !bMono;
    const Rigid Tcw(Rcw, tcw);

    Queries q(LastFrame.N);
    for (int i = 0; i < LastFrame.N; i++) {
        MapPoint* pMP = LastFrame.mvpMapPoints[i];
        if (!pMP || LastFrame.mvbOutlier[i]) continue;
        float x3Dw[3], x3Dc[3];
        get3(pMP->GetWorldPos(), x3Dw);
        Tcw.apply(x3Dw, x3Dc);
        const float xc = x3Dc[0], yc = x3Dc[1];
        const float invzc = 1.0 / x3Dc[2];
        if (invzc < 0) continue;
        const float u = CurrentFrame.fx * xc * invzc + CurrentFrame.cx;
        const float v = CurrentFrame.fy * yc * invzc + CurrentFrame.cy;
        if (u < CurrentFrame.mnMinX || u > CurrentFrame.mnMaxX) continue;
        if (v < CurrentFrame.mnMinY || v > CurrentFrame.mnMaxY) continue;
        const int nLastOctave = LastFrame.mvKeys[i].octave;
        const float radius = th * CurrentFrame.mvScaleFactors[nLastOctave];
        q.valid[i] = 1;
        q.u[i] = u; q.v[i] = v; q.radius[i] = radius;
        if (bForward) { q.minLevel[i] = nLastOctave; q.maxLevel[i] = -1; }          // :1377-1382 (GetFeaturesInArea defaults maxLevel = -1)
        else if (bBackward) { q.minLevel[i] = 0; q.maxLevel[i] = nLastOctave; }
        else { q.minLevel[i] = nLastOctave - 1; q.maxLevel[i] = nLastOctave + 1; }
        q.uR[i] = u - CurrentFrame.mbf * invzc; q.erMax[i] = radius;                // :1409-1414
        q.angle[i] = LastFrame.mvKeysUn[i].angle;
        q.obs[i] = pMP->Observations() > 0;
        q.setDescriptor(i, pMP->GetDescriptor());
    }
    std::vector<uint8_t> taken(CurrentFrame.N, 0);
    for (int j = 0; j < CurrentFrame.N; j++) taken[j] = CurrentFrame.mvpMapPoints[j] && CurrentFrame.mvpMapPoints[j]->Observations() > 0;   // :1405-1407
    const float bounds[4] = {Frame::mnMinX, Frame::mnMinY, Frame::mnMaxX, Frame::mnMaxY};
    SearchResult r = search(ORB_MODE_TRACK_LAST, TH_HIGH, mfNNratio, mbCheckOrientation, CurrentFrame.mvKeysUn, CurrentFrame.mDescriptors,
                            CurrentFrame.mvuRight.data(), bounds, taken, q, true);
    for (int j = 0; j < CurrentFrame.N; j++) {
        if (r.ownerOfTarget[j] >= 0) CurrentFrame.mvpMapPoints[j] = LastFrame.mvpMapPoints[r.ownerOfTarget[j]];   // :1430
        else if (r.ownerOfTarget[j] == -2) CurrentFrame.mvpMapPoints[j] = static_cast<MapPoint*>(NULL);          // :1462-1466
    }
    return r.nmatches;
}

// ---- Relocalization refinement -> SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist)  (ORBmatcher.cc:1474-1601) ----
int ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, const float th, const int ORBdist) {
    const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0, 3).colRange(0, 3);
    const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0, 3).col(3);
    const cv::Mat OwM = -Rcw.t() * tcw;
    float Ow[3];
    get3(OwM, Ow);
    const Rigid Tcw(Rcw, tcw);
    const vector<MapPoint*> vpMPs = pKF->GetMapPointMatches();

    Queries q(vpMPs.size());
    for (size_t i = 0, iend = vpMPs.size(); i < iend; i++) {
        MapPoint* pMP = vpMPs[i];
        if (!pMP || pMP->isBad() || sAlreadyFound.count(pMP)) continue;           // :1493-1497
        float x3Dw[3], x3Dc[3];
        get3(pMP->GetWorldPos(), x3Dw);
        Tcw.apply(x3Dw, x3Dc);
        const float xc = x3Dc[0], yc = x3Dc[1];
        const float invzc = 1.0 / x3Dc[2];
        const float u = CurrentFrame.fx * xc * invzc + CurrentFrame.cx;
        const float v = CurrentFrame.fy * yc * invzc + CurrentFrame.cy;
        if (u < CurrentFrame.mnMinX || u > CurrentFrame.mnMaxX) continue;
        if (v < CurrentFrame.mnMinY || v > CurrentFrame.mnMaxY) continue;
        const float PO[3] = {x3Dw[0] - Ow[0], x3Dw[1] - Ow[1], x3Dw[2] - Ow[2]};
        float dist3D;
        if (!scaleAndViewGate(pMP, PO, dist3D, false)) continue;                  // :1516-1524 (no viewing-angle test here)
        const int nPredictedLevel = pMP->PredictScale(dist3D, &CurrentFrame);
        q.valid[i] = 1;
        q.u[i] = u; q.v[i] = v; q.radius[i] = th * CurrentFrame.mvScaleFactors[nPredictedLevel];
        q.minLevel[i] = nPredictedLevel - 1; q.maxLevel[i] = nPredictedLevel + 1;   // :1532
        q.angle[i] = pKF->mvKeysUn[i].angle;                                       // :1565
        q.setDescriptor(i, pMP->GetDescriptor());
    }
    std::vector<uint8_t> taken(CurrentFrame.N, 0);
    for (int j = 0; j < CurrentFrame.N; j++) taken[j] = CurrentFrame.mvpMapPoints[j] != NULL;   // :1543: ANY map point hides the keypoint
    const float bounds[4] = {Frame::mnMinX, Frame::mnMinY, Frame::mnMaxX, Frame::mnMaxY};
    SearchResult r = search(ORB_MODE_TRACK_LAST, ORBdist, mfNNratio, mbCheckOrientation, CurrentFrame.mvKeysUn, CurrentFrame.mDescriptors, nullptr,
                            bounds, taken, q, false);
    for (int j = 0; j < CurrentFrame.N; j++) {
        if (r.ownerOfTarget[j] >= 0) CurrentFrame.mvpMapPoints[j] = vpMPs[r.ownerOfTarget[j]];     // :1559
        else if (r.ownerOfTarget[j] == -2) CurrentFrame.mvpMapPoints[j] = NULL;                    // :1590-1596
    }
    return r.nmatches;
}

namespace {
// Scw -> Rcw, tcw, Ow exactly as ORBmatcher.cc:301-306 / :988-993 write it
struct Sim3Camera {
    cv::Mat Rcw, tcw, Ow;
    explicit Sim3Camera(const cv::Mat& Scw) {
        cv::Mat sRcw = Scw.rowRange(0, 3).colRange(0, 3);
        const float scw = sqrt(sRcw.row(0).dot(sRcw.row(0)));
        Rcw = sRcw / scw;
        tcw = Scw.rowRange(0, 3).col(3) / scw;
        Ow = -Rcw.t() * tcw;
    }
};

// projection of a map point into a keyframe with the gates of ORBmatcher.cc:322-357 / 846-900 / 1008-1047:
// positive depth, inside the image, scale-invariance range, viewing angle; outputs u, v, invz, predicted level
struct KeyFrameProjection { float u, v, invz; int level; };
bool projectIntoKeyFrame(MapPoint* pMP, KeyFrame* pKF, const Rigid& Tcw, const float* Ow, KeyFrameProjection& o) {
    float p3Dw[3], p3Dc[3];
    get3(pMP->GetWorldPos(), p3Dw);
    Tcw.apply(p3Dw, p3Dc);
    if (p3Dc[2] < 0.0f) return false;
    o.invz = 1 / p3Dc[2];
    const float x = p3Dc[0] * o.invz;
    const float y = p3Dc[1] * o.invz;
    o.u = pKF->fx * x + pKF->cx;
    o.v = pKF->fy * y + pKF->cy;
    if (!pKF->IsInImage(o.u, o.v)) return false;
    const float PO[3] = {p3Dw[0] - Ow[0], p3Dw[1] - Ow[1], p3Dw[2] - Ow[2]};
    float dist3D;
    if (!scaleAndViewGate(pMP, PO, dist3D, true)) return false;
    o.level = pMP->PredictScale(dist3D, pKF);
    return true;
}
}  // namespace

// ---- LoopClosing::ComputeSim3 -> SearchByProjection(pKF, Scw, vpPoints, vpMatched, th)     (ORBmatcher.cc:291-404) ----
int ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, vector<MapPoint*>& vpMatched, int th) {
    const Sim3Camera cam(Scw);
    const Rigid Tcw(cam.Rcw, cam.tcw);
    float Ow[3];
    get3(cam.Ow, Ow);
    set<MapPoint*> spAlreadyFound(vpMatched.begin(), vpMatched.end());            // :309-310
    spAlreadyFound.erase(static_cast<MapPoint*>(NULL));

    Queries q(vpPoints.size());
    for (int iMP = 0, iendMP = vpPoints.size(); iMP < iendMP; iMP++) {
        MapPoint* pMP = vpPoints[iMP];
        if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
        KeyFrameProjection pr;
        if (!projectIntoKeyFrame(pMP, pKF, Tcw, Ow, pr)) continue;
        q.valid[iMP] = 1;
        q.u[iMP] = pr.u; q.v[iMP] = pr.v; q.radius[iMP] = th * pKF->mvScaleFactors[pr.level];
        q.minLevel[iMP] = pr.level - 1; q.maxLevel[iMP] = pr.level;                // :378-381
        q.setDescriptor(iMP, pMP->GetDescriptor());
    }
    std::vector<uint8_t> taken(pKF->N, 0);
    for (int j = 0; j < pKF->N; j++) taken[j] = vpMatched[j] != NULL;            // :376
    const float bounds[4] = {(float)pKF->mnMinX, (float)pKF->mnMinY, (float)pKF->mnMaxX, (float)pKF->mnMaxY};
    // best only, <= TH_LOW, no rotation histogram (:392-397)
    SearchResult r = search(ORB_MODE_TRACK_LAST, TH_LOW, mfNNratio, false, pKF->mvKeysUn, pKF->mDescriptors, nullptr, bounds, taken, q, false);
    for (int j = 0; j < pKF->N; j++)
        if (r.ownerOfTarget[j] >= 0) vpMatched[j] = vpPoints[r.ownerOfTarget[j]];
    return r.nmatches;
}

// ---- Relocalization / loop detection -> SearchByBoW(pKF, F, vpMapPointMatches)             (ORBmatcher.cc:160-289) ----
int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches) {
    const vector<MapPoint*> vpMapPointsKF = pKF->GetMapPointMatches();
    vpMapPointMatches = vector<MapPoint*>(F.N, static_cast<MapPoint*>(NULL));
    const int n1 = pKF->N, n2 = F.N;
    if (n1 == 0 || n2 == 0) return 0;
    std::vector<uint8_t> valid1(n1, 0);
    for (int i = 0; i < n1; i++) valid1[i] = vpMapPointsKF[i] && !vpMapPointsKF[i]->isBad();   // :196-202
    const FeatCsr fv1(pKF->mFeatVec), fv2(F.mFeatVec);
    const std::vector<float> a1 = angles(pKF->mvKeysUn), a2 = angles(F.mvKeys);                 // :239: kp.angle - F.mvKeys[bestIdxF].angle
    DescBlock d1(pKF->mDescriptors), d2(F.mDescriptors);
    std::vector<int32_t> m12(n1, -1), m21(n2, -1);
    int nmatches = 0;
    check(orb_search_by_bow(device(), d1.p, a1.data(), valid1.data(), n1, fv1.node.data(), fv1.start.data(), fv1.feat.data(), fv1.n(), d2.p, a2.data(),
                            nullptr, n2, fv2.node.data(), fv2.start.data(), fv2.feat.data(), fv2.n(), TH_LOW, /*strict*/ 0, mfNNratio,
                            mbCheckOrientation ? 1 : 0, m12.data(), m21.data(), &nmatches),
          "orb_search_by_bow");
    for (int j = 0; j < n2; j++)
        if (m21[j] >= 0) vpMapPointMatches[j] = vpMapPointsKF[m21[j]];                          // :233
    return nmatches;
}

// ---- LoopClosing::ComputeSim3 -> SearchByBoW(pKF1, pKF2, vpMatches12)                       (ORBmatcher.cc:524-657) ----
int ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12) {
    const vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches();
    const vector<MapPoint*> vpMapPoints2 = pKF2->GetMapPointMatches();
    const int n1 = (int)vpMapPoints1.size(), n2 = (int)vpMapPoints2.size();
    vpMatches12 = vector<MapPoint*>(n1, static_cast<MapPoint*>(NULL));
    if (n1 == 0 || n2 == 0) return 0;
    std::vector<uint8_t> valid1(n1, 0), valid2(n2, 0);
    for (int i = 0; i < n1; i++) valid1[i] = vpMapPoints1[i] && !vpMapPoints1[i]->isBad();   // :562-566
    for (int j = 0; j < n2; j++) valid2[j] = vpMapPoints2[j] && !vpMapPoints2[j]->isBad();   // :580-585
    const FeatCsr fv1(pKF1->mFeatVec), fv2(pKF2->mFeatVec);
    const std::vector<float> a1 = angles(pKF1->mvKeysUn), a2 = angles(pKF2->mvKeysUn);
    DescBlock d1(pKF1->mDescriptors), d2(pKF2->mDescriptors);
    std::vector<int32_t> m12(n1, -1), m21(n2, -1);
    int nmatches = 0;
    check(orb_search_by_bow(device(), d1.p, a1.data(), valid1.data(), n1, fv1.node.data(), fv1.start.data(), fv1.feat.data(), fv1.n(), d2.p, a2.data(),
                            valid2.data(), n2, fv2.node.data(), fv2.start.data(), fv2.feat.data(), fv2.n(), TH_LOW, /*strict: '<' at :600*/ 1,
                            mfNNratio, mbCheckOrientation ? 1 : 0, m12.data(), m21.data(), &nmatches),
          "orb_search_by_bow");
    for (int i = 0; i < n1; i++)
        if (m12[i] >= 0) vpMatches12[i] = vpMapPoints2[m12[i]];                                // :604
    return nmatches;
}

// ---- Monocular initialization -> SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize)  (ORBmatcher.cc:406-521) ----
int ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, vector<cv::Point2f>& vbPrevMatched, vector<int>& vnMatches12, int windowSize) {
    const size_t n1 = F1.mvKeysUn.size();
    vnMatches12 = vector<int>(n1, -1);
    Queries q(n1);
    for (size_t i1 = 0; i1 < n1; i1++) {
        q.valid[i1] = F1.mvKeysUn[i1].octave <= 0;                                 // level1 > 0: continue (:423-425)
        q.u[i1] = vbPrevMatched[i1].x; q.v[i1] = vbPrevMatched[i1].y; q.radius[i1] = windowSize;
        q.minLevel[i1] = 0; q.maxLevel[i1] = 0;                                    // GetFeaturesInArea(.., level1, level1) with level1 == 0 (:427)
        q.angle[i1] = F1.mvKeysUn[i1].angle;
        memcpy(&q.desc[i1 * 32], F1.mDescriptors.ptr((int)i1), 32);
    }
    std::vector<uint8_t> taken(F2.mvKeysUn.size(), 0);
    const float bounds[4] = {Frame::mnMinX, Frame::mnMinY, Frame::mnMaxX, Frame::mnMaxY};
    SearchResult r = search(ORB_MODE_INITIALIZATION, TH_LOW, mfNNratio, mbCheckOrientation, F2.mvKeysUn, F2.mDescriptors, nullptr, bounds, taken, q, false);
    for (size_t i1 = 0; i1 < n1; i1++) vnMatches12[i1] = r.matchOfQuery[i1];
    for (size_t i1 = 0; i1 < n1; i1++)
        if (vnMatches12[i1] >= 0) vbPrevMatched[i1] = F2.mvKeysUn[vnMatches12[i1]].pt;   // :515-518
    return r.nmatches;
}

// ---- LocalMapping::CreateNewMapPoints -> SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo)  (ORBmatcher.cc:659-825) ----
int ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, vector<pair<size_t, size_t> >& vMatchedPairs, const bool bOnlyStereo) {
    // epipole of camera 1 in image 2 (:665-673)
    cv::Mat Cw = pKF1->GetCameraCenter();
    cv::Mat R2w = pKF2->GetRotation();
    cv::Mat t2w = pKF2->GetTranslation();
    cv::Mat C2 = R2w * Cw + t2w;
    const float invz = 1.0f / C2.at<float>(2);
    const float ex = pKF2->fx * C2.at<float>(0) * invz + pKF2->cx;
    const float ey = pKF2->fy * C2.at<float>(1) * invz + pKF2->cy;

    vMatchedPairs.clear();
    const int n1 = pKF1->N, n2 = pKF2->N;
    if (n1 == 0 || n2 == 0) return 0;
    std::vector<uint8_t> has1(n1), has2(n2);
    for (int i = 0; i < n1; i++) has1[i] = pKF1->GetMapPoint(i) != NULL;          // :696-700
    for (int j = 0; j < n2; j++) has2[j] = pKF2->GetMapPoint(j) != NULL;          // :718-722
    const FeatCsr fv1(pKF1->mFeatVec), fv2(pKF2->mFeatVec);
    float f12[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) f12[3 * i + j] = F12.at<float>(i, j);
    DescBlock d1(pKF1->mDescriptors), d2(pKF2->mDescriptors);
    std::vector<int32_t> m12(n1, -1);
    int nmatches = 0;
    check(orb_search_for_triangulation(device(), kp_ptr(pKF1->mvKeysUn), d1.p, has1.data(), pKF1->mvuRight.data(), n1, fv1.node.data(), fv1.start.data(),
                                       fv1.feat.data(), fv1.n(), kp_ptr(pKF2->mvKeysUn), d2.p, has2.data(), pKF2->mvuRight.data(), n2, fv2.node.data(),
                                       fv2.start.data(), fv2.feat.data(), fv2.n(), f12, ex, ey, pKF2->mvScaleFactors.data(), pKF2->mvLevelSigma2.data(),
                                       (int)pKF2->mvScaleFactors.size(), bOnlyStereo ? 1 : 0, mbCheckOrientation ? 1 : 0, m12.data(), &nmatches),
          "orb_search_for_triangulation");
    vMatchedPairs.reserve(nmatches);                                                // :812-822
    for (int i = 0; i < n1; i++)
        if (m12[i] >= 0) vMatchedPairs.push_back(make_pair((size_t)i, (size_t)m12[i]));
    return nmatches;
}

// ---- LoopClosing::ComputeSim3 -> SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th)  (ORBmatcher.cc:1104-1328) ----
int ORBmatcher::SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12, const float& s12, const cv::Mat& R12, const cv::Mat& t12,
                             const float th) {
    cv::Mat R1w = pKF1->GetRotation();
    cv::Mat t1w = pKF1->GetTranslation();
    cv::Mat R2w = pKF2->GetRotation();
    cv::Mat t2w = pKF2->GetTranslation();
    cv::Mat sR12 = s12 * R12;                                                       // :1121-1124
    cv::Mat sR21 = (1.0 / s12) * R12.t();
    cv::Mat t21 = -sR21 * t12;
    const Rigid T1w(R1w, t1w), T2w(R2w, t2w), T12(sR12, t12), T21(sR21, t21);

    const vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches();
    const int N1 = vpMapPoints1.size();
    const vector<MapPoint*> vpMapPoints2 = pKF2->GetMapPointMatches();
    const int N2 = vpMapPoints2.size();
    if (N1 == 0 || N2 == 0) return 0;
    vector<bool> vbAlreadyMatched1(N1, false), vbAlreadyMatched2(N2, false);
    for (int i = 0; i < N1; i++) {                                                  // :1135-1145
        MapPoint* pMP = vpMatches12[i];
        if (pMP) {
            vbAlreadyMatched1[i] = true;
            int idx2 = pMP->GetIndexInKeyFrame(pKF2);
            if (idx2 >= 0 && idx2 < N2) vbAlreadyMatched2[idx2] = true;
        }
    }
    // one direction: the map points of `from` transformed into the camera of `to` (first by the camera of `from`, then by
    // the Sim3 `Tto_from`), gates of :1160-1184 / :1233-1257 (no viewing-angle test)
    auto project = [&](const vector<MapPoint*>& pts, const vector<bool>& already, const Rigid& Tfrom, const Rigid& Tto_from, KeyFrame* to, Queries& q) {
        for (size_t i = 0; i < pts.size(); i++) {
            MapPoint* pMP = pts[i];
            if (!pMP || already[i] || pMP->isBad()) continue;
            float p3Dw[3], pa[3], pb[3];
            get3(pMP->GetWorldPos(), p3Dw);
            Tfrom.apply(p3Dw, pa);
            Tto_from.apply(pa, pb);
            if (pb[2] < 0.0) continue;
            const float invz = 1.0 / pb[2];
            const float x = pb[0] * invz;
            const float y = pb[1] * invz;
            const float u = pKF1->fx * x + pKF1->cx;                                // the reference uses pKF1's intrinsics for both directions (:1107-1110)
            const float v = pKF1->fy * y + pKF1->cy;
            if (!to->IsInImage(u, v)) continue;
            float dist3D;
            if (!scaleAndViewGate(pMP, pb, dist3D, false)) continue;
            const int nPredictedLevel = pMP->PredictScale(dist3D, to);
            q.valid[i] = 1;
            q.u[i] = u; q.v[i] = v; q.radius[i] = th * to->mvScaleFactors[nPredictedLevel];
            q.minLevel[i] = nPredictedLevel;                                         // orb_search_by_sim3 searches [level - 1, level]
            q.setDescriptor(i, pMP->GetDescriptor());
        }
    };
    Queries q12(N1), q21(N2);
    project(vpMapPoints1, vbAlreadyMatched1, T1w, T21, pKF2, q12);
    project(vpMapPoints2, vbAlreadyMatched2, T2w, T12, pKF1, q21);

    const float b1[4] = {(float)pKF1->mnMinX, (float)pKF1->mnMinY, (float)pKF1->mnMaxX, (float)pKF1->mnMaxY};
    const float b2[4] = {(float)pKF2->mnMinX, (float)pKF2->mnMinY, (float)pKF2->mnMaxX, (float)pKF2->mnMaxY};
    DescBlock d1(pKF1->mDescriptors), d2(pKF2->mDescriptors);
    std::vector<int32_t> m12(N1, -1);
    int nFound = 0;
    check(orb_search_by_sim3(device(), kp_ptr(pKF1->mvKeysUn), d1.p, N1, b1, kp_ptr(pKF2->mvKeysUn), d2.p, N2, b2, q12.u.data(), q12.v.data(),
                             q12.radius.data(), q12.minLevel.data(), q12.desc.data(), q12.valid.data(), q21.u.data(), q21.v.data(), q21.radius.data(),
                             q21.minLevel.data(), q21.desc.data(), q21.valid.data(), TH_HIGH, m12.data(), &nFound),
          "orb_search_by_sim3");
    for (int i1 = 0; i1 < N1; i1++)
        if (m12[i1] >= 0) vpMatches12[i1] = vpMapPoints2[m12[i1]];                 // :1314-1322
    return nFound;
}

namespace {
// the search half of both Fuse overloads on the device (orb_fuse_search); the map mutation stays with the caller
void fuseSearch(KeyFrame* pKF, bool reprojectionGate, const Queries& q, std::vector<int32_t>& bestIdx, std::vector<int32_t>& bestDist) {
    const int nq = (int)q.size();
    bestIdx.assign(nq, -1);
    bestDist.assign(nq, 256);
    if (nq == 0 || pKF->N == 0) return;
    const float bounds[4] = {(float)pKF->mnMinX, (float)pKF->mnMinY, (float)pKF->mnMaxX, (float)pKF->mnMaxY};
    DescBlock d(pKF->mDescriptors);
    check(orb_fuse_search(device(), kp_ptr(pKF->mvKeysUn), d.p, pKF->mvuRight.data(), pKF->N, bounds, reprojectionGate ? pKF->mvInvLevelSigma2.data() : nullptr,
                          (int)pKF->mvInvLevelSigma2.size(), nq, q.u.data(), q.v.data(), q.uR.data(), q.radius.data(), q.minLevel.data(), q.desc.data(),
                          q.valid.data(), bestIdx.data(), bestDist.data()),
          "orb_fuse_search");
}
}  // namespace

// ---- LocalMapping::SearchInNeighbors -> Fuse(pKF, vpMapPoints, th)                          (ORBmatcher.cc:827-977) ----
int ORBmatcher::Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, const float th) {
    cv::Mat Rcw = pKF->GetRotation();
    cv::Mat tcw = pKF->GetTranslation();
    const float& bf = pKF->mbf;
    float Ow[3];
    get3(pKF->GetCameraCenter(), Ow);
    const Rigid Tcw(Rcw, tcw);
    const int nMPs = vpMapPoints.size();

    Queries q(nMPs);
    for (int i = 0; i < nMPs; i++) {
        MapPoint* pMP = vpMapPoints[i];
        if (!pMP) continue;
        if (pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;                       // re-checked below: the loop mutates the map as it goes
        KeyFrameProjection pr;
        if (!projectIntoKeyFrame(pMP, pKF, Tcw, Ow, pr)) continue;
        q.valid[i] = 1;
        q.u[i] = pr.u; q.v[i] = pr.v; q.uR[i] = pr.u - bf * pr.invz;                // :869
        q.radius[i] = th * pKF->mvScaleFactors[pr.level];
        q.minLevel[i] = pr.level;
        q.setDescriptor(i, pMP->GetDescriptor());
    }
    std::vector<int32_t> bestIdx, bestDist;
    fuseSearch(pKF, true, q, bestIdx, bestDist);

    int nFused = 0;
    for (int i = 0; i < nMPs; i++) {                                                // :955-973, in the reference's order
        if (!q.valid[i] || bestDist[i] > TH_LOW) continue;
        MapPoint* pMP = vpMapPoints[i];
        if (pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;                       // state as of this iteration (:846-847)
        MapPoint* pMPinKF = pKF->GetMapPoint(bestIdx[i]);
        if (pMPinKF) {
            if (!pMPinKF->isBad()) {
                if (pMPinKF->Observations() > pMP->Observations()) pMP->Replace(pMPinKF);
                else pMPinKF->Replace(pMP);
            }
        } else {
            pMP->AddObservation(pKF, bestIdx[i]);
            pKF->AddMapPoint(pMP, bestIdx[i]);
        }
        nFused++;
    }
    return nFused;
}

// ---- LoopClosing::SearchAndFuse -> Fuse(pKF, Scw, vpPoints, th, vpReplacePoint)             (ORBmatcher.cc:979-1102) ----
int ORBmatcher::Fuse(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, float th, vector<MapPoint*>& vpReplacePoint) {
    const Sim3Camera cam(Scw);
    const Rigid Tcw(cam.Rcw, cam.tcw);
    float Ow[3];
    get3(cam.Ow, Ow);
    const set<MapPoint*> spAlreadyFound = pKF->GetMapPoints();                     // :996, evaluated once
    const int nPoints = vpPoints.size();

    Queries q(nPoints);
    for (int iMP = 0; iMP < nPoints; iMP++) {
        MapPoint* pMP = vpPoints[iMP];
        if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
        KeyFrameProjection pr;
        if (!projectIntoKeyFrame(pMP, pKF, Tcw, Ow, pr)) continue;
        q.valid[iMP] = 1;
        q.u[iMP] = pr.u; q.v[iMP] = pr.v;
        q.radius[iMP] = th * pKF->mvScaleFactors[pr.level];
        q.minLevel[iMP] = pr.level;
        q.setDescriptor(iMP, pMP->GetDescriptor());
    }
    std::vector<int32_t> bestIdx, bestDist;
    fuseSearch(pKF, false, q, bestIdx, bestDist);                                   // no reprojection gate in this overload (:1062-1078)

    int nFused = 0;
    for (int iMP = 0; iMP < nPoints; iMP++) {                                       // :1081-1097
        if (!q.valid[iMP] || bestDist[iMP] > TH_LOW) continue;
        MapPoint* pMP = vpPoints[iMP];
        if (pMP->isBad()) continue;
        MapPoint* pMPinKF = pKF->GetMapPoint(bestIdx[iMP]);
        if (pMPinKF) {
            if (!pMPinKF->isBad()) vpReplacePoint[iMP] = pMPinKF;
        } else {
            pMP->AddObservation(pKF, bestIdx[iMP]);
            pKF->AddMapPoint(pMP, bestIdx[iMP]);
        }
        nFused++;
    }
    return nFused;
}

}  // namespace ORB_SLAM2
