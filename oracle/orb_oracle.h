/*
 * orb_oracle.h — C interface of the CPU ORACLE for the ORB front-end hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This library is a CPU restatement of the reference algorithm
 * (wjjcdy/orb_slam_2_ros: orb_slam2/src/ORBextractor.cc, ORBmatcher.cc, Frame.cc) used as the
 * checker for the CUDA product.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product (orb_slam_2_ros_b200/) never links or calls it.
 *
 * Parity pins:
 *   (i)   OpenCV primitive semantics = OpenCV 4.13.0 (resize INTER_LINEAR 8u, FAST 9/16 + NMS,
 *         GaussianBlur 7x7 sigma 2 fixed-point, fastAtan2, cvRound), each checked against the
 *         cv2 4.13.0 wheel by tests/test_oracle_vs_cv2.py and tests/golden/.
 *   (ii)  DistributeOctTree equal-size tie-break = (size, creation sequence) instead of the
 *         reference's allocator-dependent (size, heap pointer)  (ORBextractor.cc:615,705-708).
 *   (iii) descriptor steering uses a=(float)cos((double)angle), b=(float)sin((double)angle).
 * PINNED TO THE REFERENCE'S OWN CODE: the reference ships no tests / golden vectors / fixtures for this path, so
 * the pin is the reference itself — oracle/_ref/liborb_ref.so is the reference's unmodified ORBextractor.cc / ORBmatcher.cc /
 * Frame.cc / DBoW2 compiled over the OpenCV stand-in of oracle/ref_shim (oracle/Makefile), and tests/test_ref_pin.py checks
 * this restatement against it bit for bit (extractor incl. order, stereo, grid, SearchForInitialization, SearchByBoW x2, BoW
 * transform).  Pin (i) is against the third-party module that owns the arithmetic (OpenCV 4.13.0); (ii) is the reference's
 * behaviour under a monotonic node allocator (ref_shim/mono_alloc.cpp); (iii) is confirmed by _ref calling this glibc's cosf.
 */
#ifndef ORB_ORACLE_H
#define ORB_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* bit-compatible with cv::KeyPoint (28 bytes) */
typedef struct {
    float x, y;
    float size;
    float angle;
    float response;
    int32_t octave;
    int32_t class_id;
} orc_kp;

typedef struct {
    int32_t best_dist;   /* 256 if no candidate */
    int32_t best_idx;    /* -1 if no candidate */
    int32_t second_dist; /* 256 if fewer than two candidates */
    int32_t second_idx;  /* index that produced second_dist (-1 if none) */
} orc_top2;

/* ---------------- extractor (ORBextractor.cc) ---------------- */
void* orc_extractor_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th);
void orc_extractor_destroy(void* ex);
/* tables: out arrays must hold nlevels entries */
void orc_extractor_tables(void* ex, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                          int32_t* features_per_level, int32_t* umax16);
/* operator(): returns number of keypoints written (<= cap), or -needed if cap too small.  */
int orc_extract(void* ex, const uint8_t* img, int w, int h, int stride, orc_kp* kps, uint8_t* desc32,
                int cap);
/* stage taps (valid after orc_extract) */
int orc_level_dims(void* ex, int level, int* w, int* h);
int orc_get_level(void* ex, int level, uint8_t* dst_bordered); /* (w+38)*(h+38), border included */
int orc_get_blurred(void* ex, int level, uint8_t* dst);          /* w*h, only if level had keypoints */
int orc_raw_corner_count(void* ex, int level);
int orc_get_raw_corners(void* ex, int level, orc_kp* out, int cap); /* vToDistributeKeys order */
int orc_level_kp_count(void* ex, int level);
int orc_get_level_kps(void* ex, int level, orc_kp* out, int cap);   /* level coords, with angle */
/* per-level statistics: stats[level*4+0]=raw corners, +1=kept, +2=tie_at_cut events, +3=cells retried */
void orc_get_stats(void* ex, int32_t* stats);
/* descriptor taps whose pre-rounding coordinate lies within eps of a .5 boundary (reported, not an error) */
int orc_count_near_half_taps(void* ex, float eps);

/* ---------------- OpenCV-primitive restatements (pinned against cv2 4.13.0) ---------------- */
void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh,
                          int dstride);
void orc_border_reflect101(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int border,
                           int dstride);
int orc_fast9_16(const uint8_t* img, int w, int h, int stride, int threshold, int nms, orc_kp* out, int cap);
void orc_gaussian7x7_s2(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride);
/* cvtColor(.., *2GRAY) on 8-bit interleaved BGR / RGB / BGRA / RGBA (channels = 3 | 4, rgb_order = red first) */
void orc_cvt_gray(const uint8_t* src, int w, int h, int stride, int channels, int rgb_order, uint8_t* dst, int dstride);
float orc_fast_atan2(float y, float x);
int orc_cv_round_f(float v);
float orc_ic_angle(const uint8_t* center, int stride);
void orc_brief_descriptor(const uint8_t* center, int stride, float angle_deg, uint8_t* desc32);

/* pin (iii): steering coefficients for the n consecutive fp32 bit patterns starting at first_bits */
void orc_sincos_range(uint32_t first_bits, long long n, float* a, float* b);

/* ---------------- matcher (ORBmatcher.cc) ---------------- */
int orc_descriptor_distance(const uint8_t* a, const uint8_t* b);
/* brute-force best / second-best of every query against db[0..ndb) in index order
 * (inner loop of ORBmatcher.cc:202-227) */
void orc_hamming_top2(const uint8_t* q, int nq, const uint8_t* db, int ndb, orc_top2* out);
/* same over CSR candidate lists: candidates of query i = cand_idx[cand_off[i]..cand_off[i+1]) in order */
void orc_hamming_top2_csr(const uint8_t* q, int nq, const uint8_t* db, const int32_t* cand_off,
                          const int32_t* cand_idx, orc_top2* out);

/* Frame grid (Frame.cc:239-256, 354-425): 64 x 48 cells */
void* orc_grid_create(const orc_kp* kps_un, int n, float min_x, float min_y, float max_x, float max_y);
void orc_grid_destroy(void* g);
/* GetFeaturesInArea; returns count, writes indices (reference order) */
int orc_grid_query(void* g, float x, float y, float r, int min_level, int max_level, int32_t* out, int cap);

/* matcher modes for orc_search_by_projection.  The two remaining SearchByProjection overloads are parameterisations of
 * ORC_MODE_TRACK_LAST (the same loop body: window, skip taken, best only, threshold, optional rotation histogram):
 *   ORBmatcher.cc:1474-1601 (Cur, KF, sAlreadyFound, th, ORBdist): th_dist = ORBdist, q_valid = pMP && !isBad && !sAlreadyFound.count(pMP)
 *       && inside the image / distance range, levels [l-1, l+1], taken = (mvpMapPoints[i2] != NULL)  (:1543), u_right = NULL,
 *       q_angle = pKF->mvKeysUn[i].angle, q_obs = all;
 *   ORBmatcher.cc:291-404 (KF, Scw, vpPoints, vpMatched, th): th_dist = TH_LOW, check_orientation = 0, levels [l-1, l],
 *       taken = (vpMatched[idx] != NULL)  (:376), q_valid = !isBad && !spAlreadyFound.count(pMP) && the projection gates (:322-357).
 * tests/test_gpu_match.py::test_search_by_projection_reloc_and_loop_parameterisations runs both against the CUDA path, and
 * tests/test_gpu_dropin.py runs the reference's own compiled code for these two overloads against the drop-in. */
enum {
    ORC_MODE_TRACK_LAST = 0,  /* ORBmatcher.cc:1330-1472  (Cur, Last, th, bMono)  best only, <= th_dist, rot-hist */
    ORC_MODE_LOCAL_POINTS = 1, /* ORBmatcher.cc:45-129     (F, vpMapPoints, th) best/second, level ratio test */
    ORC_MODE_INITIALIZATION = 2 /* ORBmatcher.cc:406-521   SearchForInitialization: steal-by-better-distance, ratio, rot-hist */
};

typedef struct {
    int32_t mode;
    int32_t th_dist;           /* TH_HIGH (100) or ORBdist */
    float nn_ratio;            /* mfNNratio */
    int32_t check_orientation; /* mbCheckOrientation (TRACK_LAST only) */
} orc_search_params;

/*
 * Windowed search with the reference's sequential "already matched" dependency.
 * Target frame: kps_un[n], desc[n*32], u_right[n] (may be NULL => all -1), grid g.
 * taken[n]: in/out, nonzero = target already has a map point with Observations()>0.
 * Queries (in reference loop order): q_u,q_v projected position; q_radius window radius;
 * q_min_level/q_max_level for GetFeaturesInArea; q_desc; q_ur = predicted right coordinate
 * (u - bf*invz, or mTrackProjXR); q_er_max = max |q_ur - u_right| (radius / r*scale);
 * q_angle = query keypoint angle (deg) for the rotation histogram; q_valid: 0 => skipped.
 * match_of_query[i] = matched target index or -1 (after rotation filtering).
 * Returns nmatches.
 */
int orc_search_by_projection(const orc_search_params* prm, void* grid, const orc_kp* kps_un,
                             const uint8_t* desc, const float* u_right, int n, uint8_t* taken, int nq,
                             const float* q_u, const float* q_v, const float* q_radius,
                             const int32_t* q_min_level, const int32_t* q_max_level, const uint8_t* q_desc,
                             const float* q_ur, const float* q_er_max, const float* q_angle,
                             const uint8_t* q_valid, int32_t* match_of_query);

/* Extended form: q_obs[i] (NULL = all 1) says whether query i's map point has Observations()>0 — only then
 * does its assignment block later queries (OM:1405-1407, OM:87-89).  target_query[n] (may be NULL) receives
 * the final owner (query index) of every target keypoint after the rotation filter, -1 = none: this is the
 * final state of Frame::mvpMapPoints the caller would observe. */
int orc_search_by_projection_ex(const orc_search_params* prm, void* grid, const orc_kp* kps_un,
                                const uint8_t* desc, const float* u_right, int n, uint8_t* taken, int nq,
                                const float* q_u, const float* q_v, const float* q_radius,
                                const int32_t* q_min_level, const int32_t* q_max_level, const uint8_t* q_desc,
                                const float* q_ur, const float* q_er_max, const float* q_angle,
                                const uint8_t* q_valid, const uint8_t* q_obs, int32_t* match_of_query,
                                int32_t* target_query);

/* brute-force frame-to-frame matching with ratio test + rotation histogram (the per-node inner loop of
 * SearchByBoW, ORBmatcher.cc:196-252, applied to one node holding all keypoints of both frames). */
int orc_match_bruteforce(const uint8_t* desc1, const float* angle1, int n1, const uint8_t* desc2,
                         const float* angle2, int n2, int th_dist, float nn_ratio, int check_orientation,
                         int32_t* match12);

/* ---------------- stereo (Frame.cc:502-676) ---------------- */
typedef struct {
    float bf;  /* mbf */
    float b;   /* mb  */
} orc_stereo_params;
/* exL/exR: oracle extractors holding the pyramids of the left/right image. Outputs u_right[nl], depth[nl]
 * (-1 = no match), best_sad[nl] (-1 if not reached).  Returns number of matches kept. */
int orc_stereo_match(void* ex_left, void* ex_right, const orc_kp* kps_l, const uint8_t* desc_l, int nl,
                     const orc_kp* kps_r, const uint8_t* desc_r, int nr, const orc_stereo_params* prm,
                     float* u_right, float* depth, int32_t* best_sad);

/* SearchByBoW over two FeatureVectors (ORBmatcher.cc:160-289 strict = 0, :524-657 strict = 1); see orb_oracle_match.cpp */
int orc_search_by_bow(const uint8_t* desc1, const float* angle1, const uint8_t* valid1, int n1, const int32_t* fv1_node,
                      const int32_t* fv1_start, const int32_t* fv1_feat, int nfv1, const uint8_t* desc2, const float* angle2,
                      const uint8_t* valid2, int n2, const int32_t* fv2_node, const int32_t* fv2_start, const int32_t* fv2_feat,
                      int nfv2, int th_dist, int strict, float nn_ratio, int check_orientation, int32_t* match12, int32_t* match21);

/* SearchForTriangulation (ORBmatcher.cc:659-825), one direction of SearchBySim3 (:1150-1213) and
 * MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:288-361) on arrays; see orb_oracle_match.cpp */
int orc_search_for_triangulation(const orc_kp* kps1, const uint8_t* desc1, const uint8_t* has_mp1, const float* u_right1, int n1,
                                 const int32_t* fv1_node, const int32_t* fv1_start, const int32_t* fv1_feat, int nfv1,
                                 const orc_kp* kps2, const uint8_t* desc2, const uint8_t* has_mp2, const float* u_right2, int n2,
                                 const int32_t* fv2_node, const int32_t* fv2_start, const int32_t* fv2_feat, int nfv2, const float* F12,
                                 float ex, float ey, const float* scale_factors, const float* level_sigma2, int only_stereo,
                                 int check_orientation, int32_t* match12);
void orc_sim3_search_one_way(void* grid, const orc_kp* kps_un, const uint8_t* desc, int nq, const float* q_u, const float* q_v,
                             const float* q_radius, const int32_t* q_level, const uint8_t* q_desc, const uint8_t* q_valid, int32_t* match);
/* search half of ORBmatcher::Fuse (ORBmatcher.cc:827-977, 979-1102) */
void orc_fuse_search(void* grid, const orc_kp* kps_un, const uint8_t* desc, const float* u_right, const float* inv_level_sigma2,
                     int nq, const float* q_u, const float* q_v, const float* q_ur, const float* q_radius, const int32_t* q_level,
                     const uint8_t* q_desc, const uint8_t* q_valid, int32_t* best_idx, int32_t* best_dist);
void orc_distinctive_descriptors(const uint8_t* desc, const int32_t* off, int npoints, int32_t* best);

/* ---------------- bag of words (DBoW2 TemplatedVocabulary<FORB>, orb_oracle_bow.cpp) ---------------- */
/* nodes in id order (0 = root, TemplatedVocabulary.h:1389-1436); children = nodes with that parent in id order */
void* orc_voc_create(int k, int L, int scoring, int weighting, int n_nodes, const int32_t* parent, const uint8_t* is_leaf,
                     const uint8_t* desc32, const double* weight);
void* orc_voc_load_text(const char* path); /* loadFromTextFile, TemplatedVocabulary.h:1351-1441; NULL on failure */
void orc_voc_destroy(void* voc);
void orc_voc_info(void* voc, int32_t* k, int32_t* L, int32_t* n_nodes, int32_t* n_words, int32_t* scoring, int32_t* weighting);
void orc_voc_export(void* voc, int32_t* parent, uint8_t* is_leaf, uint8_t* desc32, double* weight);
/* transform(feature, id, weight, &nid, levelsup), TemplatedVocabulary.h:1231-1272 */
void orc_bow_transform_features(void* voc, const uint8_t* desc32, int n, int levelsup, int32_t* word_id, double* weight,
                                int32_t* node_id);
/* transform(features, BowVector, FeatureVector, levelsup), TemplatedVocabulary.h:1140-1218 (flat, sorted outputs) */
void orc_bow_transform(void* voc, const uint8_t* desc32, int n, int levelsup, int32_t* n_bow, int32_t* bow_word,
                       double* bow_value, int32_t* n_fv, int32_t* fv_node, int32_t* fv_start, int32_t* fv_feat);
double orc_bow_score_l1(const int32_t* w1, const double* v1, int n1, const int32_t* w2, const double* v2, int n2);

#ifdef __cplusplus
}
#endif
#endif
