/*
 * harness.cpp — flat C entry points over the reference's C++ class surface (TEST INFRASTRUCTURE ONLY).
 *
 * This one source is compiled into two shared libraries that export the same rh_* symbols:
 *
 *   oracle/_ref/liborb_ref.so            harness + the reference's UNMODIFIED ORBextractor.cc, ORBmatcher.cc, Frame.cc and
 *                                        DBoW2 sources (compiled where they lie under /root/reference) over the OpenCV
 *                                        stand-in of cvshim.hpp = the reference's own CPU path ("oracle/_ref")
 *   tests/_build/liborb_dropin.so        harness + the reference's unmodified Frame.cc and headers, but with
 *                                        orb_slam_2_ros_b200/host/{ORBextractor.cc, ORBmatcher.cc, Frame_ComputeStereoMatches.cc}
 *                                        in place of the reference's ORBextractor.cc / ORBmatcher.cc / Frame::ComputeStereoMatches
 *                                        = the drop-in: the same callers, the CUDA library underneath
 *
 * tests/test_ref_pin.py drives the first against the oracle restatement (CPU); tests/test_gpu_dropin.py drives both with
 * the same inputs and compares every output (GPU).  Everything a caller of the reference would touch goes through the
 * reference's own types here: cv::Mat images, Frame constructors, KeyFrame / MapPoint pointers, ORBmatcher methods.
 */
#include <cstdint>
#include <cstring>
#include <map>
#include <new>
#include <set>
#include <string>
#include <vector>

/* Frame keeps AssignFeaturesToGrid / ComputeImageBounds / UndistortKeyPoints private (Frame.h:202-212); the from-arrays
 * frames below need them.  Access specifiers do not change the class layout. */
#define private public
#include "Frame.h"
#undef private
#include "ORBextractor.h"
#include "ORBmatcher.h"

using namespace ORB_SLAM2;

namespace {

struct Kp { float x, y, size, angle, response; int32_t octave, class_id; };
static_assert(sizeof(Kp) == sizeof(cv::KeyPoint), "cv::KeyPoint is 28 bytes");

struct PointSet {
    std::vector<MapPoint*> pts;
    std::map<MapPoint*, int> index;
    ~PointSet() { for (MapPoint* p : pts) delete p; }
    int idx(MapPoint* p) const { if (!p) return -1; auto it = index.find(p); return it == index.end() ? -2 : it->second; }
};

cv::Mat make_K(const float* k4) {
    cv::Mat K = cv::Mat::eye(3, 3, CV_32F);
    K.at<float>(0, 0) = k4[0]; K.at<float>(1, 1) = k4[1]; K.at<float>(0, 2) = k4[2]; K.at<float>(1, 2) = k4[3];
    return K;
}
cv::Mat mat_from(const float* v, int r, int c) {
    cv::Mat m(r, c, CV_32F);
    for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) m.at<float>(i, j) = v[i * c + j];
    return m;
}
cv::Mat image_from(const uint8_t* img, int w, int h, int stride) {
    cv::Mat m(h, w, CV_8UC1);
    for (int y = 0; y < h; ++y) memcpy(m.ptr(y), img + (size_t)y * stride, w);
    return m;
}

}  // namespace

/* oracle/_ref only (mono_alloc.cpp): marks the calls during which quadtree list nodes may come from the monotonic arena */
extern "C" void rh_alloc_scope(int delta) __attribute__((weak));
namespace {
struct ExtractScope {
    ExtractScope() { if (rh_alloc_scope) rh_alloc_scope(+1); }
    ~ExtractScope() { if (rh_alloc_scope) rh_alloc_scope(-1); }
};
}  // namespace

extern "C" {

const char* rh_arm(void) {
#ifdef RH_ARM_NAME
    return RH_ARM_NAME;
#else
    return "unknown";
#endif
}

/* ------------------------------------------------------------ ORBextractor ---------------------------------------- */
void* rh_extractor_create(int nfeatures, float scale, int nlevels, int ini, int mn) { return new ORBextractor(nfeatures, scale, nlevels, ini, mn); }
void rh_extractor_destroy(void* ex) { delete (ORBextractor*)ex; }

void rh_extractor_tables(void* p, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2) {
    ORBextractor* ex = (ORBextractor*)p;
    std::vector<float> a = ex->GetScaleFactors(), b = ex->GetInverseScaleFactors(), c = ex->GetScaleSigmaSquares(), d = ex->GetInverseScaleSigmaSquares();
    for (int i = 0; i < ex->GetLevels(); ++i) { scale[i] = a[i]; inv_scale[i] = b[i]; sigma2[i] = c[i]; inv_sigma2[i] = d[i]; }
}

/* operator()(image, mask, keypoints, descriptors): returns the keypoint count, or -needed when cap is too small */
int rh_extract(void* p, const uint8_t* img, int w, int h, int stride, void* kps, uint8_t* desc32, int cap) {
    ORBextractor* ex = (ORBextractor*)p;
    cv::Mat image = image_from(img, w, h, stride), descriptors;
    std::vector<cv::KeyPoint> keys;
    {
        ExtractScope scope;
        (*ex)(image, cv::Mat(), keys, descriptors);
    }
    const int n = (int)keys.size();
    if (n > cap) return -n;
    if (n) {
        memcpy(kps, keys.data(), (size_t)n * sizeof(cv::KeyPoint));
        for (int i = 0; i < n; ++i) memcpy(desc32 + (size_t)i * 32, descriptors.ptr(i), 32);
    }
    return n;
}

/* mvImagePyramid[level] (public data member, ORBextractor.h:85): the interior ROI and the 19-px border around it */
int rh_level_dims(void* p, int level, int* w, int* h) {
    ORBextractor* ex = (ORBextractor*)p;
    if (level < 0 || level >= (int)ex->mvImagePyramid.size() || ex->mvImagePyramid[level].empty()) return -1;
    *w = ex->mvImagePyramid[level].cols; *h = ex->mvImagePyramid[level].rows;
    return 0;
}
int rh_get_level(void* p, int level, uint8_t* dst_bordered) {
    ORBextractor* ex = (ORBextractor*)p;
    const cv::Mat& m = ex->mvImagePyramid[level];
    const int W = m.cols + 38, H = m.rows + 38;
    const uint8_t* base = m.data - 19 * (ptrdiff_t)m.step - 19;
    for (int y = 0; y < H; ++y) memcpy(dst_bordered + (size_t)y * W, base + (size_t)y * m.step, W);
    return 0;
}

/* --------------------------------------------------------------- Frames ------------------------------------------- */
/* the next constructed Frame recomputes the image bounds / grid pitch / intrinsics (Frame.cc:106-122) */
void rh_reset_calibration(void) { Frame::mbInitialComputations = true; }

/* Storage for a Frame whose `mb` already holds mbf / fx.  The stereo constructor calls ComputeStereoMatches (which reads
 * mb as minZ, Frame.cc:533) BEFORE it assigns mb (Frame.cc:124) and never initialises it: in the running system a Frame is
 * constructed into the storage of the previous one (Tracking.cc:212: mCurrentFrame = Frame(...)), so the value read is the
 * previous frame's mb.  The harness reproduces that state instead of reading uninitialised heap memory. */
static void* frame_storage(const float* k4, float bf) {
    void* mem = ::operator new(sizeof(Frame));
    memset(mem, 0, sizeof(Frame));
    reinterpret_cast<Frame*>(mem)->mb = bf / k4[0];
    return mem;
}
/* Frame(imGray, timeStamp, extractor, voc, K, distCoef, bf, thDepth)  (Frame.cc:178-236) */
void* rh_frame_mono(void* ex, const uint8_t* img, int w, int h, int stride, const float* k4, float bf, float th_depth) {
    cv::Mat K = make_K(k4), dist = cv::Mat::zeros(4, 1, CV_32F), im = image_from(img, w, h, stride);
    ExtractScope scope;
    return new (frame_storage(k4, bf)) Frame(im, 0.0, (ORBextractor*)ex, (ORBVocabulary*)nullptr, K, dist, bf, th_depth);
}
/* Frame(imLeft, imRight, ...)  (Frame.cc:60-128): two extractor threads + ComputeStereoMatches */
void* rh_frame_stereo(void* exl, void* exr, const uint8_t* iml, const uint8_t* imr, int w, int h, int stride, const float* k4, float bf,
                      float th_depth) {
    cv::Mat K = make_K(k4), dist = cv::Mat::zeros(4, 1, CV_32F), l = image_from(iml, w, h, stride), r = image_from(imr, w, h, stride);
    ExtractScope scope;
    return new (frame_storage(k4, bf)) Frame(l, r, 0.0, (ORBextractor*)exl, (ORBextractor*)exr, (ORBVocabulary*)nullptr, K, dist, bf, th_depth);
}
/* a Frame from given (undistorted) keypoints: what the constructors leave behind, without an image.  Scale tables come
 * from the extractor, the bounds are [0,w) x [0,h) as ComputeImageBounds gives for an undistorted camera. */
void* rh_frame_from_arrays(void* exp, const void* kps, const uint8_t* desc32, const float* u_right, const float* depth, int n, int w, int h,
                           const float* k4, float bf, float th_depth) {
    ORBextractor* ex = (ORBextractor*)exp;
    Frame* F = new Frame();
    F->mpORBvocabulary = nullptr; F->mpORBextractorLeft = ex; F->mpORBextractorRight = nullptr; F->mTimeStamp = 0;
    F->mK = make_K(k4); F->mDistCoef = cv::Mat::zeros(4, 1, CV_32F); F->mbf = bf; F->mThDepth = th_depth; F->mpReferenceKF = nullptr;
    F->mnId = Frame::nNextId++;
    F->mnScaleLevels = ex->GetLevels(); F->mfScaleFactor = ex->GetScaleFactor(); F->mfLogScaleFactor = log(F->mfScaleFactor);
    F->mvScaleFactors = ex->GetScaleFactors(); F->mvInvScaleFactors = ex->GetInverseScaleFactors();
    F->mvLevelSigma2 = ex->GetScaleSigmaSquares(); F->mvInvLevelSigma2 = ex->GetInverseScaleSigmaSquares();
    F->N = n;
    F->mvKeys.resize(n);
    if (n) memcpy(F->mvKeys.data(), kps, (size_t)n * sizeof(cv::KeyPoint));
    F->mDescriptors.create(std::max(n, 1), 32, CV_8U);
    if (n) memcpy(F->mDescriptors.data, desc32, (size_t)n * 32);
    F->UndistortKeyPoints();
    F->mvuRight.assign(n, -1.f); F->mvDepth.assign(n, -1.f);
    if (u_right) F->mvuRight.assign(u_right, u_right + n);
    if (depth) F->mvDepth.assign(depth, depth + n);
    F->mvpMapPoints.assign(n, nullptr);
    F->mvbOutlier.assign(n, false);
    if (Frame::mbInitialComputations) {
        cv::Mat im(h, w, CV_8UC1);
        F->ComputeImageBounds(im);
        Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(Frame::mnMaxX - Frame::mnMinX);
        Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(Frame::mnMaxY - Frame::mnMinY);
        Frame::fx = k4[0]; Frame::fy = k4[1]; Frame::cx = k4[2]; Frame::cy = k4[3];
        Frame::invfx = 1.0f / Frame::fx; Frame::invfy = 1.0f / Frame::fy;
        Frame::mbInitialComputations = false;
    }
    F->mb = F->mbf / Frame::fx;
    F->AssignFeaturesToGrid();
    return F;
}
void* rh_frame_copy(void* f) { return new Frame(*(Frame*)f); }
void rh_frame_destroy(void* f) { delete (Frame*)f; }
int rh_frame_n(void* f) { return ((Frame*)f)->N; }
void rh_frame_get(void* fp, void* kps, void* kps_un, uint8_t* desc32, float* u_right, float* depth) {
    Frame* F = (Frame*)fp;
    const int n = F->N;
    if (!n) return;
    if (kps) memcpy(kps, F->mvKeys.data(), (size_t)n * sizeof(cv::KeyPoint));
    if (kps_un) memcpy(kps_un, F->mvKeysUn.data(), (size_t)n * sizeof(cv::KeyPoint));
    if (desc32) for (int i = 0; i < n; ++i) memcpy(desc32 + (size_t)i * 32, F->mDescriptors.ptr(i), 32);
    if (u_right) memcpy(u_right, F->mvuRight.data(), (size_t)n * 4);
    if (depth) memcpy(depth, F->mvDepth.data(), (size_t)n * 4);
}
int rh_frame_right_n(void* f) { return (int)((Frame*)f)->mvKeysRight.size(); }
void rh_frame_get_right(void* fp, void* kps, uint8_t* desc32) {
    Frame* F = (Frame*)fp;
    const int n = (int)F->mvKeysRight.size();
    if (!n) return;
    memcpy(kps, F->mvKeysRight.data(), (size_t)n * sizeof(cv::KeyPoint));
    for (int i = 0; i < n; ++i) memcpy(desc32 + (size_t)i * 32, F->mDescriptorsRight.ptr(i), 32);
}
void rh_frame_bounds(float* b4) { b4[0] = Frame::mnMinX; b4[1] = Frame::mnMinY; b4[2] = Frame::mnMaxX; b4[3] = Frame::mnMaxY; }
void rh_frame_set_pose(void* f, const float* tcw16) { ((Frame*)f)->SetPose(mat_from(tcw16, 4, 4)); }
/* Frame::GetFeaturesInArea (Frame.cc:354-412): returns the count, indices in the reference's order */
int rh_frame_features_in_area(void* f, float x, float y, float r, int min_level, int max_level, int32_t* out, int cap) {
    std::vector<size_t> v = ((Frame*)f)->GetFeaturesInArea(x, y, r, min_level, max_level);
    for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = (int32_t)v[i];
    return (int)v.size();
}
/* DBoW2::FeatureVector of the frame from CSR arrays (node ids ascending; features of node j = feat[start[j] .. start[j+1])) */
void rh_frame_set_featvec(void* fp, const int32_t* node, const int32_t* start, const int32_t* feat, int nnodes) {
    Frame* F = (Frame*)fp;
    F->mFeatVec.clear();
    for (int j = 0; j < nnodes; ++j)
        for (int k = start[j]; k < start[j + 1]; ++k) F->mFeatVec.addFeature((DBoW2::NodeId)node[j], (unsigned)feat[k]);
}

/* ------------------------------------------------------------ ORBVocabulary --------------------------------------- */
/* System.cc:  mpVocabulary->loadFromTextFile(strVocFile)  (TemplatedVocabulary.h:1351-1441) */
void* rh_voc_load_text(const char* path) {
    ORBVocabulary* v = new ORBVocabulary();
    if (!v->loadFromTextFile(path)) { delete v; return nullptr; }
    return v;
}
void rh_voc_destroy(void* v) { delete (ORBVocabulary*)v; }
int rh_voc_size(void* v) { return (int)((ORBVocabulary*)v)->size(); }
/* Frame::ComputeBoW (Frame.cc:428-435): transform(vCurrentDesc, mBowVec, mFeatVec, 4) */
void rh_frame_compute_bow(void* fp, void* voc) {
    Frame* F = (Frame*)fp;
    F->mpORBvocabulary = (ORBVocabulary*)voc;
    F->mBowVec.clear();
    F->mFeatVec.clear();
    F->ComputeBoW();
}
int rh_frame_get_bow(void* fp, int32_t* word, double* value, int cap) {
    Frame* F = (Frame*)fp;
    int n = 0;
    for (DBoW2::BowVector::const_iterator it = F->mBowVec.begin(); it != F->mBowVec.end(); ++it, ++n)
        if (n < cap) { word[n] = (int32_t)it->first; value[n] = it->second; }
    return n;
}
/* CSR: node ids ascending, start[nnodes + 1], feat; returns the node count (-1 when a capacity is too small) */
int rh_frame_get_featvec(void* fp, int32_t* node, int32_t* start, int32_t* feat, int cap_nodes, int cap_feat) {
    Frame* F = (Frame*)fp;
    int n = 0, m = 0;
    if (cap_nodes > 0) start[0] = 0;
    for (DBoW2::FeatureVector::const_iterator it = F->mFeatVec.begin(); it != F->mFeatVec.end(); ++it, ++n) {
        if (n >= cap_nodes || m + (int)it->second.size() > cap_feat) return -1;
        node[n] = (int32_t)it->first;
        for (size_t k = 0; k < it->second.size(); ++k) feat[m++] = (int32_t)it->second[k];
        start[n + 1] = m;
    }
    return n;
}
/* KeyFrameDatabase's similarity: mpVoc->score(v1, v2) (L1Scoring::score, ScoringObject.cpp:23-66) */
double rh_voc_score(void* voc, void* f1, void* f2) { return ((ORBVocabulary*)voc)->score(((Frame*)f1)->mBowVec, ((Frame*)f2)->mBowVec); }

/* ------------------------------------------------------------- MapPoints ------------------------------------------ */
/* n map points: world position, mean viewing direction, representative descriptor, Observations(), isBad(), the scale
 * invariance distances (mfMinDistance / mfMaxDistance).  Any array may be NULL (zeros / nObs = 1 / not bad / [0, 1e9]). */
void* rh_points_create(int n, const float* pos3, const float* normal3, const uint8_t* desc32, const int32_t* nobs, const uint8_t* bad,
                       const float* min_dist, const float* max_dist) {
    PointSet* s = new PointSet();
    s->pts.resize(n);
    for (int i = 0; i < n; ++i) {
        MapPoint* p = new MapPoint();
        p->mnId = (unsigned long)i;
        p->mWorldPos = cv::Mat::zeros(3, 1, CV_32F);
        p->mNormalVector = cv::Mat::zeros(3, 1, CV_32F);
        p->mDescriptor = cv::Mat::zeros(1, 32, CV_8U);
        for (int k = 0; k < 3; ++k) {
            if (pos3) p->mWorldPos.at<float>(k) = pos3[i * 3 + k];
            if (normal3) p->mNormalVector.at<float>(k) = normal3[i * 3 + k];
        }
        if (desc32) memcpy(p->mDescriptor.data, desc32 + (size_t)i * 32, 32);
        p->nObs = nobs ? nobs[i] : 1;
        p->mbBad = bad ? bad[i] != 0 : false;
        p->mfMinDistance = min_dist ? min_dist[i] : 0.f;
        p->mfMaxDistance = max_dist ? max_dist[i] : 1e9f;
        s->pts[i] = p;
        s->index[p] = i;
    }
    return s;
}
void rh_points_destroy(void* s) { delete (PointSet*)s; }
/* Fuse's effect on the points: replaced_by[i] = index of the point that replaced i (-1 none); nobs[i] = Observations() */
void rh_points_state(void* sp, int32_t* replaced_by, int32_t* nobs, uint8_t* bad) {
    PointSet* s = (PointSet*)sp;
    for (size_t i = 0; i < s->pts.size(); ++i) {
        if (replaced_by) replaced_by[i] = s->idx(s->pts[i]->mpReplaced);
        if (nobs) nobs[i] = s->pts[i]->nObs;
        if (bad) bad[i] = s->pts[i]->mbBad;
    }
}
/* Frame::mvpMapPoints <-> indices into the point set (-1 = NULL) */
void rh_frame_set_points(void* fp, void* sp, const int32_t* idx) {
    Frame* F = (Frame*)fp; PointSet* s = (PointSet*)sp;
    for (int i = 0; i < F->N; ++i) F->mvpMapPoints[i] = idx[i] < 0 ? nullptr : s->pts[idx[i]];
}
void rh_frame_get_points(void* fp, void* sp, int32_t* idx) {
    Frame* F = (Frame*)fp; PointSet* s = (PointSet*)sp;
    for (int i = 0; i < F->N; ++i) idx[i] = s->idx(F->mvpMapPoints[i]);
}
void rh_frame_set_outliers(void* fp, const uint8_t* o) { Frame* F = (Frame*)fp; for (int i = 0; i < F->N; ++i) F->mvbOutlier[i] = o[i] != 0; }

/* ------------------------------------------------------------- KeyFrames ------------------------------------------ */
void* rh_keyframe_create(void* frame) { return new KeyFrame(*(Frame*)frame); }
void rh_keyframe_destroy(void* kf) { delete (KeyFrame*)kf; }
void rh_keyframe_set_pose(void* kf, const float* tcw16) { ((KeyFrame*)kf)->SetPose(mat_from(tcw16, 4, 4)); }
void rh_keyframe_set_bad(void* kf, int bad) { ((KeyFrame*)kf)->mbBad = bad != 0; }
void rh_keyframe_set_points(void* kp, void* sp, const int32_t* idx, int observe) {
    KeyFrame* K = (KeyFrame*)kp; PointSet* s = (PointSet*)sp;
    for (int i = 0; i < K->N; ++i) {
        K->mvpMapPoints[i] = idx[i] < 0 ? nullptr : s->pts[idx[i]];
        if (observe && idx[i] >= 0) s->pts[idx[i]]->mObservations[K] = i;   /* IsInKeyFrame / GetIndexInKeyFrame see it; nObs is the caller's */
    }
}
void rh_keyframe_get_points(void* kp, void* sp, int32_t* idx) {
    KeyFrame* K = (KeyFrame*)kp; PointSet* s = (PointSet*)sp;
    for (int i = 0; i < K->N; ++i) idx[i] = s->idx(K->mvpMapPoints[i]);
}

/* ------------------------------------------------------------- ORBmatcher ----------------------------------------- */
int rh_descriptor_distance(const uint8_t* a, const uint8_t* b) {
    cv::Mat ma(1, 32, CV_8U, (void*)a), mb(1, 32, CV_8U, (void*)b);
    return ORBmatcher::DescriptorDistance(ma, mb);
}
void rh_matcher_constants(int32_t* th_low, int32_t* th_high, int32_t* histo) {
    *th_low = ORBmatcher::TH_LOW; *th_high = ORBmatcher::TH_HIGH; *histo = ORBmatcher::HISTO_LENGTH;
}

/* Tracking::SearchLocalPoints (Tracking.cc:949-1003): isInFrustum marks the points, then
 * SearchByProjection(Frame&, vector<MapPoint*>&, th)  (ORBmatcher.cc:45-129) */
int rh_search_local_points(float nnratio, int check_ori, void* fp, void* sp, const int32_t* idx, int n, float th, int run_frustum) {
    Frame* F = (Frame*)fp; PointSet* s = (PointSet*)sp;
    std::vector<MapPoint*> v(n);
    for (int i = 0; i < n; ++i) {
        v[i] = s->pts[idx[i]];
        if (run_frustum) { v[i]->mbTrackInView = false; F->isInFrustum(v[i], 0.5); }
    }
    ORBmatcher m(nnratio, check_ori != 0);
    return m.SearchByProjection(*F, v, th);
}
/* the tracking fields isInFrustum leaves on the points (Frame.cc:286-351) */
void rh_points_track_state(void* sp, uint8_t* in_view, float* proj_x, float* proj_y, float* proj_xr, int32_t* level, float* view_cos) {
    PointSet* s = (PointSet*)sp;
    for (size_t i = 0; i < s->pts.size(); ++i) {
        MapPoint* p = s->pts[i];
        in_view[i] = p->mbTrackInView; proj_x[i] = p->mTrackProjX; proj_y[i] = p->mTrackProjY; proj_xr[i] = p->mTrackProjXR;
        level[i] = p->mnTrackScaleLevel; view_cos[i] = p->mTrackViewCos;
    }
}
/* SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono)  (ORBmatcher.cc:1330-1472) */
int rh_search_last_frame(float nnratio, int check_ori, void* cur, void* last, float th, int mono) {
    ORBmatcher m(nnratio, check_ori != 0);
    return m.SearchByProjection(*(Frame*)cur, *(const Frame*)last, th, mono != 0);
}
/* SearchByProjection(Frame &CurrentFrame, KeyFrame*, sAlreadyFound, th, ORBdist)  (ORBmatcher.cc:1474-1601) */
int rh_search_reloc(float nnratio, int check_ori, void* cur, void* kf, void* sp, const int32_t* found_idx, int nfound, float th, int orb_dist) {
    PointSet* s = (PointSet*)sp;
    std::set<MapPoint*> found;
    for (int i = 0; i < nfound; ++i) found.insert(s->pts[found_idx[i]]);
    ORBmatcher m(nnratio, check_ori != 0);
    return m.SearchByProjection(*(Frame*)cur, (KeyFrame*)kf, found, th, orb_dist);
}
/* SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th)  (ORBmatcher.cc:291-404); matched[kf N] in/out as point indices */
int rh_search_loop(float nnratio, int check_ori, void* kf, const float* scw16, void* sp, const int32_t* idx, int n, int32_t* matched, int th) {
    KeyFrame* K = (KeyFrame*)kf; PointSet* s = (PointSet*)sp;
    std::vector<MapPoint*> v(n), vm(K->N);
    for (int i = 0; i < n; ++i) v[i] = s->pts[idx[i]];
    for (int i = 0; i < K->N; ++i) vm[i] = matched[i] < 0 ? nullptr : s->pts[matched[i]];
    ORBmatcher m(nnratio, check_ori != 0);
    const int r = m.SearchByProjection(K, mat_from(scw16, 4, 4), v, vm, th);
    for (int i = 0; i < K->N; ++i) matched[i] = s->idx(vm[i]);
    return r;
}
/* SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches)  (ORBmatcher.cc:160-289): matches[F.N] = point index or -1 */
int rh_search_bow_kf_frame(float nnratio, int check_ori, void* kf, void* fp, void* sp, int32_t* matches) {
    Frame* F = (Frame*)fp; PointSet* s = (PointSet*)sp;
    std::vector<MapPoint*> vm;
    ORBmatcher m(nnratio, check_ori != 0);
    const int r = m.SearchByBoW((KeyFrame*)kf, *F, vm);
    for (int i = 0; i < F->N; ++i) matches[i] = s->idx(vm[i]);
    return r;
}
/* SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12)  (ORBmatcher.cc:524-657): matches12[kf1 N] = point index (of kf2) or -1 */
int rh_search_bow_kf_kf(float nnratio, int check_ori, void* kf1, void* kf2, void* sp, int32_t* matches12) {
    KeyFrame* K1 = (KeyFrame*)kf1; PointSet* s = (PointSet*)sp;
    std::vector<MapPoint*> vm;
    ORBmatcher m(nnratio, check_ori != 0);
    const int r = m.SearchByBoW(K1, (KeyFrame*)kf2, vm);
    for (int i = 0; i < K1->N; ++i) matches12[i] = s->idx(vm[i]);
    return r;
}
/* SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize)  (ORBmatcher.cc:406-521) */
int rh_search_initialization(float nnratio, int check_ori, void* f1, void* f2, float* prev_xy, int32_t* matches12, int window) {
    Frame* F1 = (Frame*)f1;
    std::vector<cv::Point2f> prev(F1->N);
    for (int i = 0; i < F1->N; ++i) prev[i] = cv::Point2f(prev_xy[2 * i], prev_xy[2 * i + 1]);
    std::vector<int> m12;
    ORBmatcher m(nnratio, check_ori != 0);
    const int r = m.SearchForInitialization(*F1, *(Frame*)f2, prev, m12, window);
    for (int i = 0; i < F1->N; ++i) { matches12[i] = m12[i]; prev_xy[2 * i] = prev[i].x; prev_xy[2 * i + 1] = prev[i].y; }
    return r;
}
/* SearchForTriangulation(KF1, KF2, F12, vMatchedPairs, bOnlyStereo)  (ORBmatcher.cc:659-825): pairs[2k], pairs[2k+1] */
int rh_search_triangulation(float nnratio, int check_ori, void* kf1, void* kf2, const float* f12, int32_t* pairs, int cap, int only_stereo) {
    std::vector<std::pair<size_t, size_t> > vp;
    ORBmatcher m(nnratio, check_ori != 0);
    const int r = m.SearchForTriangulation((KeyFrame*)kf1, (KeyFrame*)kf2, mat_from(f12, 3, 3), vp, only_stereo != 0);
    for (size_t k = 0; k < vp.size() && (int)k < cap; ++k) { pairs[2 * k] = (int32_t)vp[k].first; pairs[2 * k + 1] = (int32_t)vp[k].second; }
    return r;
}
/* SearchBySim3(KF1, KF2, vpMatches12, s12, R12, t12, th)  (ORBmatcher.cc:1104-1328): matches12[kf1 N] in/out as point indices */
int rh_search_sim3(float nnratio, int check_ori, void* kf1, void* kf2, void* sp, int32_t* matches12, float s12, const float* r12, const float* t12,
                   float th) {
    KeyFrame* K1 = (KeyFrame*)kf1; PointSet* s = (PointSet*)sp;
    std::vector<MapPoint*> vm(K1->N);
    for (int i = 0; i < K1->N; ++i) vm[i] = matches12[i] < 0 ? nullptr : s->pts[matches12[i]];
    ORBmatcher m(nnratio, check_ori != 0);
    const int r = m.SearchBySim3(K1, (KeyFrame*)kf2, vm, s12, mat_from(r12, 3, 3), mat_from(t12, 3, 1), th);
    for (int i = 0; i < K1->N; ++i) matches12[i] = s->idx(vm[i]);
    return r;
}
/* Fuse(KeyFrame*, vpMapPoints, th)  (ORBmatcher.cc:827-977); the map mutation is read back with rh_points_state /
 * rh_keyframe_get_points */
int rh_fuse(float nnratio, int check_ori, void* kf, void* sp, const int32_t* idx, int n, float th) {
    PointSet* s = (PointSet*)sp;
    std::vector<MapPoint*> v(n);
    for (int i = 0; i < n; ++i) v[i] = idx[i] < 0 ? nullptr : s->pts[idx[i]];
    ORBmatcher m(nnratio, check_ori != 0);
    return m.Fuse((KeyFrame*)kf, v, th);
}
/* Fuse(KeyFrame*, Scw, vpPoints, th, vpReplacePoint)  (ORBmatcher.cc:979-1102): replace[n] in/out as point indices */
int rh_fuse_sim3(float nnratio, int check_ori, void* kf, const float* scw16, void* sp, const int32_t* idx, int n, float th, int32_t* replace) {
    PointSet* s = (PointSet*)sp;
    std::vector<MapPoint*> v(n), rep(n);
    for (int i = 0; i < n; ++i) { v[i] = s->pts[idx[i]]; rep[i] = replace[i] < 0 ? nullptr : s->pts[replace[i]]; }
    ORBmatcher m(nnratio, check_ori != 0);
    const int r = m.Fuse((KeyFrame*)kf, mat_from(scw16, 4, 4), v, th, rep);
    for (int i = 0; i < n; ++i) replace[i] = s->idx(rep[i]);
    return r;
}

/* ---- OpenCV stand-in probes (pinned against cv2 by tests/test_ref_pin.py) ---- */
void rh_probe_gemm(const float* a, int ar, int ac, const float* b, int br, int bc, const float* c, float* d) {
    cv::Mat A = mat_from(a, ar, ac), B = mat_from(b, br, bc), D;
    if (c) D = A * B + mat_from(c, ar, bc); else D = A * B;
    for (int i = 0; i < ar; ++i) for (int j = 0; j < bc; ++j) d[i * bc + j] = D.at<float>(i, j);
}
double rh_probe_norm(const float* a, int n) { return cv::norm(mat_from(a, n, 1)); }
double rh_probe_dot(const float* a, const float* b, int n) { return mat_from(a, n, 1).dot(mat_from(b, n, 1)); }

}  /* extern "C" */
