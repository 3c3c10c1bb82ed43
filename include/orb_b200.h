/*
 * orb_b200.h — C ABI of liborb_b200.so: the B200-native (sm_100a) ORB front-end of ORB-SLAM2.
 *
 * This is the drop-in boundary for the hot path of wjjcdy/orb_slam_2_ros.  The reference has no FFI
 * layer; the replaced surface is the C++ class surface of two translation units plus one Frame method:
 *
 *   ORBextractor::ORBextractor(...)            orb_slam2/include/ORBextractor.h:51-52   -> orb_create
 *   ORBextractor::operator()(img,mask,kps,desc) ORBextractor.h:59-61, ORBextractor.cc:1083 -> orb_extract
 *   ORBextractor getters                        ORBextractor.h:63-83                      -> orb_get_tables
 *   ORBextractor::mvImagePyramid (public data)  ORBextractor.h:85                         -> orb_pyramid_level
 *   ORBmatcher::DescriptorDistance + best/2nd   ORBmatcher.cc:1649-1665, 202-227          -> orb_hamming_top2*
 *   ORBmatcher::SearchByProjection (Frame/Last) ORBmatcher.cc:45-129, 1330-1472           -> orb_search_by_projection
 *   ORBmatcher::SearchByBoW inner loop          ORBmatcher.cc:196-252                     -> orb_match_bruteforce
 *   ORBmatcher::SearchByBoW (both overloads)    ORBmatcher.cc:160-289, 524-657            -> orb_search_by_bow
 *   ORBmatcher::SearchForTriangulation          ORBmatcher.cc:659-825                     -> orb_search_for_triangulation
 *   ORBmatcher::SearchBySim3                    ORBmatcher.cc:1104-1328                   -> orb_search_by_sim3
 *   ORBmatcher::Fuse x2 (search half)           ORBmatcher.cc:827-977, 979-1102           -> orb_fuse_search
 *   MapPoint::ComputeDistinctiveDescriptors     MapPoint.cc:288-361                       -> orb_distinctive_descriptors
 *   Frame::ComputeStereoMatches                 Frame.cc:502-676                          -> orb_stereo_match
 *   ORBVocabulary::transform / loadFromTextFile DBoW2/TemplatedVocabulary.h:1140-1272, 1351 -> orb_bow_transform*, orb_voc_*
 *
 * INTEGRATION.md shows the C++ shims (ORBextractor.cc / ORBmatcher.cc / Frame.cc replacements) that bind
 * these entry points with the reference's own signatures.
 *
 * Conventions: every function returns an int status (ORB_OK == 0, negative = error, never throws, never
 * aborts); orb_last_error() gives a thread-local message.  Outputs are caller-allocated with explicit
 * capacities.  Pointers are HOST pointers unless the parameter name starts with d_ (device pointer on the
 * context's device).  One CUDA stream per context; distinct contexts may be used concurrently from
 * different threads (the reference runs the left/right extractors on two threads, Frame.cc:79-82); one
 * context must not be used by two threads at once.  There is no CPU fallback: without a CUDA device
 * every compute entry point returns ORB_ERR_NO_DEVICE.
 */
#ifndef ORB_B200_H
#define ORB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    ORB_OK = 0,
    ORB_ERR_INVALID = -1,   /* bad argument (NULL, sizes, unsupported dims)                 */
    ORB_ERR_CUDA = -2,      /* CUDA runtime error, see orb_last_error()                      */
    ORB_ERR_CAPACITY = -3,  /* caller buffer / context capacity too small                    */
    ORB_ERR_NO_DEVICE = -4, /* no CUDA device visible                                        */
    ORB_ERR_TOO_SMALL = -5  /* image too small for the reference's 30-px cell grid           */
};

/* bit-compatible with cv::KeyPoint (28 bytes): pt.x, pt.y, size, angle, response, octave, class_id */
typedef struct {
    float x, y;
    float size;
    float angle;
    float response;
    int32_t octave;
    int32_t class_id;
} orb_kp;

/* best / second-best of one query (ORBmatcher.cc:217-226 update rule). idx == -1, dist == 256 when absent. */
typedef struct {
    int32_t best_dist;
    int32_t second_dist;
    int64_t best_idx;
    int64_t second_idx;
} orb_top2;

typedef struct orb_ctx orb_ctx; /* one ORBextractor instance (tables + device arena + stream) */
typedef struct orb_db orb_db;   /* a device-resident shard of a descriptor database            */

/* ---- library ---------------------------------------------------------------------------------- */
const char* orb_last_error(void);
int orb_device_count(void);
int orb_version(void);

/* ---- extractor (ORBextractor.cc) ---------------------------------------------------------------- */
/* max_batch: frames processed per launch group (arena is sized for it; larger batches are chunked). */
int orb_create(orb_ctx** ctx, int nfeatures, float scale_factor, int nlevels, int ini_th_fast, int min_th_fast,
               int device, int max_batch);
void orb_destroy(orb_ctx* ctx);
/* arrays of nlevels entries (any may be NULL): mvScaleFactor, mvInvScaleFactor, mvLevelSigma2,
 * mvInvLevelSigma2, mnFeaturesPerLevel */
int orb_get_tables(orb_ctx* ctx, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                   int32_t* features_per_level);
int orb_max_keypoints(orb_ctx* ctx); /* upper bound of keypoints per frame (sum of N_l + 3 per level ...) */

/* operator(): one host image (any row stride) -> host keypoints + descriptors. *n_out = count. */
int orb_extract(orb_ctx* ctx, const uint8_t* img, int w, int h, size_t stride, orb_kp* kps, uint8_t* desc32,
                int cap, int* n_out);
/* nframes host images of identical size; outputs are [nframes][cap] / [nframes][cap][32]. */
int orb_extract_batch(orb_ctx* ctx, const uint8_t* imgs, int nframes, int w, int h, size_t row_stride,
                      size_t frame_stride, orb_kp* kps, uint8_t* desc32, int cap, int32_t* n_out);
/* Same with DEVICE pointers, asynchronous on the context stream (no host synchronisation): the form a
 * GPU-resident caller (and bench.py's device-resident leg) uses.  nframes <= max_batch. */
int orb_extract_batch_device(orb_ctx* ctx, const uint8_t* d_imgs, int nframes, int w, int h, size_t row_stride,
                             size_t frame_stride, orb_kp* d_kps, uint8_t* d_desc32, int cap, int32_t* d_n_out);
/* Colour input: Tracking::GrabImage* converts the camera image with cvtColor(.., CV_RGB2GRAY /
 * CV_BGR2GRAY / CV_RGBA2GRAY / CV_BGRA2GRAY) before the extractor sees it (Tracking.cc:181-204, 223-234, 253-264).  These
 * entry points take the interleaved 8-bit colour image and fuse that conversion into the level-0 pyramid copy
 * (OpenCV 4.13.0 arithmetic: (B*3735 + G*19235 + R*9798 + 16384) >> 15).  row_stride / frame_stride are in bytes. */
enum { ORB_PIX_GRAY8 = 0, ORB_PIX_BGR8 = 1, ORB_PIX_RGB8 = 2, ORB_PIX_BGRA8 = 3, ORB_PIX_RGBA8 = 4 };
int orb_extract_batch_pix(orb_ctx* ctx, const uint8_t* imgs, int pixel_format, int nframes, int w, int h, size_t row_stride,
                          size_t frame_stride, orb_kp* kps, uint8_t* desc32, int cap, int32_t* n_out);
int orb_extract_batch_device_pix(orb_ctx* ctx, const uint8_t* d_imgs, int pixel_format, int nframes, int w, int h,
                                 size_t row_stride, size_t frame_stride, orb_kp* d_kps, uint8_t* d_desc32, int cap,
                                 int32_t* d_n_out);

/* stream plumbing: adopt an external cudaStream_t (e.g. torch's current stream) / wait for the context */
int orb_set_stream(orb_ctx* ctx, void* cuda_stream);
int orb_sync(orb_ctx* ctx);
/* kernels launched by this context so far (the bench's gpu_launches counter) */
int64_t orb_launch_count(orb_ctx* ctx);

/* per-stage device timers (CUDA events on the context stream, recorded around every stage of every extract
 * call while enabled; reading never blocks the stream).  stage_ms[ORB_NUM_STAGES] = accumulated milliseconds of
 * {pyramid, FAST cells, quadtree, blur, orientation+descriptors}; *calls / *frames = batches / frames covered. */
#define ORB_NUM_STAGES 5
int orb_profile_enable(orb_ctx* ctx, int enable);
int orb_profile_read(orb_ctx* ctx, double* stage_ms, int64_t* calls, int64_t* frames, int reset);

/* mvImagePyramid[level] of frame `frame` of the last extract call: interior size via orb_level_dims;
 * orb_pyramid_level copies the BORDERED buffer ((w+38) x (h+38), reflect-101 border, like the Mat the
 * reference's ROI lives in) to host memory with row stride dst_stride. */
int orb_level_dims(orb_ctx* ctx, int level, int* w, int* h);
int orb_pyramid_level(orb_ctx* ctx, int frame, int level, uint8_t* dst, size_t dst_stride);

/* All levels of one frame with a single synchronisation (what a host caller that reads mvImagePyramid wants after every
 * operator()): the bordered buffer of level l — (h_l + 38) rows of (w_l + 38) bytes, the reference's own layout — is copied to
 * dst + offset[l], pitch[l] = w_l + 38, so that the ROI (19, 19, w_l, h_l) of Mat(h_l + 38, w_l + 38, CV_8U, dst + offset[l], pitch[l])
 * is mvImagePyramid[l] with step = w_l + 38 (ORBextractor.cc:1161-1165).  dst == NULL: only *needed (bytes) is set. */
int orb_pyramid_levels(orb_ctx* ctx, int frame, uint8_t* dst, size_t dst_bytes, size_t* offset, size_t* pitch, size_t* needed);

/* stage taps of the last call, for parity tests (host outputs) */
int orb_debug_blurred(orb_ctx* ctx, int frame, int level, uint8_t* dst, size_t dst_stride); /* w x h */
/* corners before the quadtree, level coordinates relative to the 16-px border (vToDistributeKeys,
 * ORBextractor.cc:856-859), sorted in the reference's (cell row, cell col, y, x) order. returns count via *n */
int orb_debug_raw_corners(orb_ctx* ctx, int frame, int level, float* xyr /* [cap][3] */, int cap, int* n);
/* per-level count of quadtree cuts that fell inside a group of equal-size nodes (where the reference's own
 * result depends on heap addresses, ORBextractor.cc:705-708,752-755; see DESIGN.md pin (ii)) */
int orb_debug_tie_counts(orb_ctx* ctx, int frame, int32_t* ties /* [nlevels] */);

/* ---- Hamming search (ORBmatcher.cc) --------------------------------------------------------------- */
/* brute force, host buffers: every query against db[0..ndb) */
int orb_hamming_top2(int device, const uint8_t* q, int nq, const uint8_t* db, int64_t ndb, orb_top2* out);

/* the same over explicit candidate lists: candidates of query i = cand_idx[cand_off[i] .. cand_off[i+1]) (indices into
 * db), ties keep the first candidate in list order.  The building block for the matcher loops whose gates (epipolar
 * line, Sim3 / reprojection window, BoW node) are evaluated by the caller: SearchForTriangulation, SearchBySim3,
 * SearchByProjection(KeyFrame, Scw, ...) (ORBmatcher.cc:659-825, 1104-1328, 291-404). */
int orb_hamming_top2_csr(int device, const uint8_t* q, int nq, const uint8_t* db, int64_t ndb, const int32_t* cand_off,
                         const int32_t* cand_idx, orb_top2* out);

/* device-resident database shard (config 5: map-wide relocalisation-scale search).
 * index_base = global index of the shard's first row; results carry global indices. */
int orb_db_create(orb_db** db, int device, int64_t capacity_rows, int64_t index_base);
void orb_db_destroy(orb_db* db);
int orb_db_add(orb_db* db, const uint8_t* desc32, int64_t nrows);          /* host rows, appended */
int orb_db_add_device(orb_db* db, const uint8_t* d_desc32, int64_t nrows); /* device rows, appended */
int64_t orb_db_size(orb_db* db);
int orb_db_set_stream(orb_db* db, void* cuda_stream);
int orb_db_query_top2(orb_db* db, const uint8_t* q, int nq, orb_top2* out);              /* host q / out */
int orb_db_query_top2_device(orb_db* db, const uint8_t* d_q, int nq, orb_top2* d_out);   /* async */
int64_t orb_db_launch_count(orb_db* db);
/* device timers of the search and merge kernels (accumulated ms over *calls queries), like orb_profile_read */
int orb_db_profile_enable(orb_db* db, int enable);
int orb_db_profile_read(orb_db* db, double* search_ms, double* merge_ms, int64_t* calls, int reset);
/* exact merge of per-shard results (after an all-gather): parts is [nparts][nq]; best = min distance,
 * lowest global index on ties; second = second smallest of the union. */
int orb_top2_merge(const orb_top2* parts, int nparts, int nq, orb_top2* out);
/* the same on the device, asynchronous on cuda_stream: the step after ncclAllGather of the per-shard results */
int orb_top2_merge_device(int device, const orb_top2* d_parts, int nparts, int nq, orb_top2* d_out, void* cuda_stream);

/* Sharded database (BASELINE config 5): one orb_db per rank / GPU holding a contiguous slice of the
 * rows (index_base = global index of its first row); every rank asks the same queries and receives the exact global answer.
 * Inside: per-shard search -> ncclAllGather of the nq x 24-byte per-shard results on the shard's stream -> device merge
 * (best = minimum distance, lowest global index on ties; second = second smallest of the union — each shard's two best
 * suffice).  NCCL is bound at run time (dlopen of libnccl.so.2); the unique id is created on rank 0 and handed to the
 * other ranks by whatever transport the host program has (MPI, torch.distributed, a file): 128 bytes.
 * orb_db_create_sharded is collective (ncclCommInitRank); world == 1 degenerates to orb_db_create. */
#define ORB_SHARD_ID_BYTES 128
int orb_shard_unique_id(uint8_t* id128);
int orb_db_create_sharded(orb_db** db, int device, int64_t capacity_rows, int64_t index_base, int rank, int world, const uint8_t* id128);
int orb_db_query_top2_sharded(orb_db* db, const uint8_t* d_q, int nq, orb_top2* d_out);          /* device pointers, async */
int orb_db_query_top2_sharded_host(orb_db* db, const uint8_t* q, int nq, orb_top2* out);          /* host pointers, blocking */

/* ---- windowed search with the sequential "already matched" rule ------------------------------------ */
enum {
    ORB_MODE_TRACK_LAST = 0,  /* ORBmatcher.cc:1330-1472: best only, <= th_dist, rotation histogram      */
    ORB_MODE_LOCAL_POINTS = 1, /* ORBmatcher.cc:45-129: best/second, same-octave ratio test, no histogram */
    ORB_MODE_INITIALIZATION = 2 /* ORBmatcher.cc:406-521 SearchForInitialization: queries = F1 keypoints (q_valid = octave 0,
                                  q_u/q_v = vbPrevMatched, q_radius = windowSize, levels 0/0), targets = F2; a target is
                                  re-assigned to a strictly better match and the previous owner loses it; accept
                                  best <= th_dist && best < second * nn_ratio; rotation histogram.  match_of_query =
                                  vnMatches12, target_query = vnMatches21; `taken` is ignored */
};
typedef struct {
    int32_t mode;
    int32_t th_dist;           /* TH_HIGH (100) or ORBdist                                  */
    float nn_ratio;            /* mfNNratio                                                 */
    int32_t check_orientation; /* mbCheckOrientation                                        */
    float min_x, min_y, max_x, max_y; /* Frame::mnMinX.. (image bounds, Frame.cc:472-500)   */
} orb_search_params;

/*
 * Target frame: kps_un[n] (Frame::mvKeysUn), desc[n][32], u_right[n] (Frame::mvuRight; NULL = mono).
 * taken[n] in/out: target already holds a map point with Observations()>0.
 * Queries in the reference's loop order: projected position (q_u,q_v), window radius, octave band for
 * GetFeaturesInArea (Frame.cc:354-412), descriptor, predicted right coordinate + tolerance (NULL when
 * u_right is NULL), angle for the rotation histogram, q_valid (NULL = all), q_obs (query's map point has
 * Observations()>0, NULL = all).  Outputs: match_of_query[nq] (target index or -1), target_query[n]
 * (final owner of each target = final Frame::mvpMapPoints, may be NULL: query index, -1 = untouched, -2 = matched and
 * then removed by the rotation filter, i.e. the reference leaves NULL there, ORBmatcher.cc:1462-1466), *nmatches.
 */
int orb_search_by_projection(int device, const orb_search_params* prm, const orb_kp* kps_un, const uint8_t* desc,
                             const float* u_right, int n, uint8_t* taken, int nq, const float* q_u,
                             const float* q_v, const float* q_radius, const int32_t* q_min_level,
                             const int32_t* q_max_level, const uint8_t* q_desc, const float* q_ur,
                             const float* q_er_max, const float* q_angle, const uint8_t* q_valid,
                             const uint8_t* q_obs, int32_t* match_of_query, int32_t* target_query, int* nmatches);

/* The same search for npairs (target frame, query set) pairs at once, everything device-resident and asynchronous on
 * cuda_stream (per-pair matching batches shard across GPUs like frames; BASELINE config 2).  The target
 * arrays are what orb_extract_batch_device left on the device ([npairs][cap_n] keypoints, [npairs][cap_n][32] descriptors,
 * [npairs] counts), the query arrays are [npairs][cap_q]; a pointer that may be NULL in orb_search_by_projection may be NULL
 * here.  Modes TRACK_LAST and LOCAL_POINTS.  d_nmatches[p] = -1 when pair p's candidate arena (160 candidates per query on
 * average) overflowed: re-run that pair through orb_search_by_projection.  One stream per calling thread. */
typedef struct {
    const orb_kp* d_kps_un; const uint8_t* d_desc; const float* d_u_right; const int32_t* d_n; int32_t cap_n;
    uint8_t* d_taken;                       /* [npairs][cap_n] in/out */
    const int32_t* d_nq; int32_t cap_q;
    const float *d_q_u, *d_q_v, *d_q_radius; const int32_t *d_q_min_level, *d_q_max_level; const uint8_t* d_q_desc;
    const float *d_q_ur, *d_q_er_max, *d_q_angle; const uint8_t *d_q_valid, *d_q_obs;
    int32_t* d_match_of_query;              /* [npairs][cap_q] */
    int32_t* d_target_query;                /* [npairs][cap_n] */
    int32_t* d_nmatches;                    /* [npairs] */
} orb_search_batch;
int orb_search_by_projection_batch_device(int device, const orb_search_params* prm, int npairs, const orb_search_batch* batch, void* cuda_stream);

/* SearchByBoW inner loop (ORBmatcher.cc:196-252) over one node holding all keypoints of both frames:
 * queries = frame 1 in order, skip already-matched targets, best <= th_dist and best < ratio*second,
 * rotation histogram + three-maxima filter.  match12[n1] = index in frame 2 or -1. */
int orb_match_bruteforce(int device, const uint8_t* desc1, const float* angle1, int n1, const uint8_t* desc2,
                         const float* angle2, int n2, int th_dist, float nn_ratio, int check_orientation,
                         int32_t* match12, int* nmatches);

/* orb_match_bruteforce for npairs frame pairs at once, device-resident (the outputs of two orb_extract_batch_device calls:
 * keypoints carry the angles), asynchronous on cuda_stream.  d_match12 [npairs][cap1] (-1 in the unused slots), d_nmatches
 * [npairs]. */
int orb_match_bruteforce_batch_device(int device, int npairs, const orb_kp* d_kps1, const uint8_t* d_desc1, const int32_t* d_n1, int cap1,
                                      const orb_kp* d_kps2, const uint8_t* d_desc2, const int32_t* d_n2, int cap2, int th_dist, float nn_ratio,
                                      int check_orientation, int32_t* d_match12, int32_t* d_nmatches, void* cuda_stream);

/* SearchByBoW over two FeatureVectors (ORBmatcher.cc:160-289 KeyFrame -> Frame: strict = 0, valid2 = NULL;
 * ORBmatcher.cc:524-657 KeyFrame -> KeyFrame: strict = 1).  Feature vectors in the CSR form orb_bow_transform returns
 * for one frame: node ids ascending, features of node j = fvX_feat[fvX_start[j] .. fvX_start[j+1]).  valid1[i] / valid2[j]
 * (NULL = all): the keypoint holds a good map point.  Per shared node the queries are walked in list order, targets
 * already matched are skipped, a match needs best <= th_dist (strict: <) and best < nn_ratio * second; then the
 * rotation histogram filter.  match12[n1] = target index or -1, match21[n2] = query index or -1. */
int orb_search_by_bow(int device, const uint8_t* desc1, const float* angle1, const uint8_t* valid1, int n1, const int32_t* fv1_node,
                      const int32_t* fv1_start, const int32_t* fv1_feat, int nfv1, const uint8_t* desc2, const float* angle2,
                      const uint8_t* valid2, int n2, const int32_t* fv2_node, const int32_t* fv2_start, const int32_t* fv2_feat, int nfv2,
                      int th_dist, int strict, float nn_ratio, int check_orientation, int32_t* match12, int32_t* match21, int* nmatches);

/* ORBmatcher::SearchForTriangulation (ORBmatcher.cc:659-825) on arrays.  has_mpX[i] != 0: the keypoint already holds a map
 * point (skipped); u_rightX (NULL = monocular): bStereo = u_right >= 0; F12 row-major 3x3; (ex, ey) = epipole of camera 1 in
 * image 2 (ORBmatcher.cc:665-673); scale_factors / level_sigma2 = pKF2->mvScaleFactors / mvLevelSigma2 (nlevels entries).
 * Per shared vocabulary node every query keeps the target with the smallest distance <= TH_LOW that passes the epipole and
 * epipolar-line gates (CheckDistEpipolarLine, :140-157), the last one in list order on equal distances (:741); no target is
 * ever marked as taken (the reference never sets vbMatched2); rotation histogram filter.  match12[n1] = vMatches12. */
int orb_search_for_triangulation(int device, const orb_kp* kps1, const uint8_t* desc1, const uint8_t* has_mp1, const float* u_right1, int n1,
                                 const int32_t* fv1_node, const int32_t* fv1_start, const int32_t* fv1_feat, int nfv1, const orb_kp* kps2,
                                 const uint8_t* desc2, const uint8_t* has_mp2, const float* u_right2, int n2, const int32_t* fv2_node,
                                 const int32_t* fv2_start, const int32_t* fv2_feat, int nfv2, const float* F12, float ex, float ey,
                                 const float* scale_factors, const float* level_sigma2, int nlevels, int only_stereo, int check_orientation,
                                 int32_t* match12, int* nmatches);

/* ORBmatcher::SearchBySim3 (ORBmatcher.cc:1104-1328) on arrays.  boundsX = {mnMinX, mnMinY, mnMaxX, mnMaxY} of keyframe X.
 * q12_*[n1]: the map point of keypoint i of keyframe 1 projected into keyframe 2 — (u, v), radius = th * scale[pred], predicted
 * level, descriptor; q12_valid[i] = 0 for the entries the reference skips (no / bad / already matched map point, negative
 * depth, outside the image or the scale-invariance range; NULL = all valid).  q21_*[n2] likewise.  Each direction keeps the
 * best candidate of GetFeaturesInArea with octave in [pred-1, pred] when its distance <= th_dist (TH_HIGH); match12[i] is set
 * when both directions agree (:1282-1299). */
int orb_search_by_sim3(int device, const orb_kp* kps1_un, const uint8_t* desc1, int n1, const float* bounds1, const orb_kp* kps2_un,
                       const uint8_t* desc2, int n2, const float* bounds2, const float* q12_u, const float* q12_v, const float* q12_radius,
                       const int32_t* q12_level, const uint8_t* q12_desc, const uint8_t* q12_valid, const float* q21_u, const float* q21_v,
                       const float* q21_radius, const int32_t* q21_level, const uint8_t* q21_desc, const uint8_t* q21_valid, int th_dist,
                       int32_t* match12, int* nfound);

/* The search half of ORBmatcher::Fuse (ORBmatcher.cc:827-977; with inv_level_sigma2 == NULL the Sim3 overload :979-1102) on
 * arrays: query i = a map point projected into the keyframe — (q_u, q_v), right coordinate q_ur = u - bf * invz, radius =
 * th * mvScaleFactors[pred], predicted level, descriptor; q_valid[i] = 0 for the points the reference skips before the
 * search (:846-886).  Candidates = GetFeaturesInArea(u, v, radius) with octave in [pred-1, pred] that pass the reprojection
 * gate e2 * mvInvLevelSigma2[octave] <= 7.8 (keypoint with mvuRight >= 0, e2 includes er) / <= 5.99 (monocular) (:914-940);
 * best_idx / best_dist = the first candidate with the smallest distance (-1 / 256 when none).  The caller fuses when
 * best_dist <= TH_LOW; Replace / AddObservation / AddMapPoint (:955-973) mutate the map and stay on the host. */
int orb_fuse_search(int device, const orb_kp* kps_un, const uint8_t* desc, const float* u_right, int n, const float* bounds,
                    const float* inv_level_sigma2, int nlevels, int nq, const float* q_u, const float* q_v, const float* q_ur,
                    const float* q_radius, const int32_t* q_level, const uint8_t* q_desc, const uint8_t* q_valid, int32_t* best_idx,
                    int32_t* best_dist);

/* MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:288-361) for npoints map points at once: the observed descriptors of
 * point p are rows [off[p], off[p+1]) (off[0] == 0); best_idx[p] = the row (within the point) with the least median distance to
 * the others, first on ties, -1 without observations; best_desc32 (may be NULL) receives that descriptor. */
int orb_distinctive_descriptors(int device, const uint8_t* desc32, const int32_t* off, int npoints, int32_t* best_idx, uint8_t* best_desc32);

/* ---- stereo (Frame::ComputeStereoMatches, Frame.cc:502-676) ------------------------------------------ */
/* ctx_left / ctx_right hold the pyramids of the last single-frame extract of the left / right image.
 * Outputs u_right[nl] (mvuRight), depth[nl] (mvDepth), -1 = no match. */
int orb_stereo_match(orb_ctx* ctx_left, orb_ctx* ctx_right, const orb_kp* kps_l, const uint8_t* desc_l, int nl,
                     const orb_kp* kps_r, const uint8_t* desc_r, int nr, float bf, float b, float* u_right,
                     float* depth, int* nmatches);

/* The same for npairs stereo pairs at once, device-resident (per-pair batches shard across GPUs like frames):
 * pair p = frame p of the last orb_extract_batch_device call of each context (their pyramids are still in the arenas);
 * d_kps_* / d_desc_* / d_n_* are exactly the outputs of those calls ([npairs][cap], [npairs][cap][32], [npairs]).  Outputs
 * d_u_right / d_depth [npairs][cap] (-1 = no match, also in the unused slots), d_nmatches [npairs].  Asynchronous on the LEFT
 * context's stream (the right context's stream is joined by an event): no host copy, no host synchronisation. */
int orb_stereo_match_batch_device(orb_ctx* ctx_left, orb_ctx* ctx_right, int npairs, const orb_kp* d_kps_l, const uint8_t* d_desc_l,
                                  const int32_t* d_n_l, const orb_kp* d_kps_r, const uint8_t* d_desc_r, const int32_t* d_n_r, int cap,
                                  float bf, float b, float* d_u_right, float* d_depth, int32_t* d_nmatches);

/* ---- bag of words: ORBVocabulary = DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB> ------------------------- */
/* (orb_slam2/include/ORBVocabulary.h:31; Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h)                                */
typedef struct orb_voc orb_voc;
/* the in-memory form of loadFromTextFile (TemplatedVocabulary.h:1389-1436): nodes in id order, node 0 = root,
 * parent[i] < i, children of a node = the nodes naming it as parent, in id order; word ids are given to the leaves
 * (is_leaf != 0) in id order.  scoring / weighting = DBoW2::ScoringType / WeightingType (BowVector.h:36-53). */
int orb_voc_create(orb_voc** voc, int device, int k, int L, int scoring, int weighting, int n_nodes, const int32_t* parent,
                   const uint8_t* is_leaf, const uint8_t* desc32, const double* weight);
/* TemplatedVocabulary::loadFromTextFile (TemplatedVocabulary.h:1351-1441), the ORBvoc.txt format */
int orb_voc_load_text(orb_voc** voc, int device, const char* path);
void orb_voc_destroy(orb_voc* voc);
int orb_voc_info(orb_voc* voc, int* k, int* L, int* n_nodes, int* n_words, int* scoring, int* weighting);
/* transform(feature, word_id, weight, &nid, levelsup) for n descriptors (TemplatedVocabulary.h:1231-1272) */
int orb_bow_transform_features(orb_voc* voc, const uint8_t* desc32, int n, int levelsup, int32_t* word_id, double* weight,
                               int32_t* node_id);
int orb_bow_transform_features_device(orb_voc* voc, const uint8_t* d_desc32, int n, int levelsup, int32_t* d_word_id,
                                      double* d_weight, int32_t* d_node_id, void* cuda_stream);
/* transform(features, BowVector&, FeatureVector&, levelsup) (TemplatedVocabulary.h:1140-1218) for nframes descriptor
 * sets at once; frame f = descriptor rows [desc_off[f], desc_off[f+1]) (desc_off[0] == 0, at most 8192 rows per frame).
 * Outputs use the frames' own row ranges as capacity (o = desc_off[f]):
 *   BowVector f     = (bow_word[o + j], bow_value[o + j]), j < bow_n[f], ascending word id, normalised as the
 *                     vocabulary's scoring type demands;
 *   FeatureVector f = nodes fv_node[o + j], j < fv_n[f], ascending; the features of node j are
 *                     fv_feat[o + fv_start[o + f + j] .. o + fv_start[o + f + j + 1]) (indices inside the frame, ascending).
 * Capacities: bow_word / bow_value / fv_node / fv_feat hold desc_off[nframes] entries, fv_start desc_off[nframes] + nframes. */
int orb_bow_transform(orb_voc* voc, const uint8_t* desc32, const int32_t* desc_off, int nframes, int levelsup, int32_t* bow_n,
                      int32_t* bow_word, double* bow_value, int32_t* fv_n, int32_t* fv_node, int32_t* fv_start, int32_t* fv_feat);
/* the same with every pointer on the vocabulary's device, asynchronous on cuda_stream; d_scratch holds 16 bytes per
 * descriptor; max_frame = the largest frame's row count */
int orb_bow_transform_device(orb_voc* voc, const uint8_t* d_desc32, const int32_t* d_desc_off, int nframes, int n_total,
                             int max_frame, int levelsup, void* d_scratch, int32_t* d_bow_n, int32_t* d_bow_word,
                             double* d_bow_value, int32_t* d_fv_n, int32_t* d_fv_node, int32_t* d_fv_start, int32_t* d_fv_feat,
                             void* cuda_stream);

/* ---- map file record payloads (reference BoostArchiver.h:46-91, KeyFrame.cc:858-864) --------------
 * System::SaveMap writes a boost::archive::binary_oarchive with no_header (System.cc:627): primitives are stored raw in
 * native byte order.  These entry points read / write the two PAYLOADS of that file that feed the GPU descriptor
 * database — nothing else of the archive (class-id / object-tracking preambles, pointers, the Map graph) is interpreted:
 *   cv::Mat record     (BoostArchiver.h:61-91):  int32 cols, int32 rows, u64 elemSize, u64 type, rows*cols*elemSize bytes
 *   cv::KeyPoint record (BoostArchiver.h:46-58): f32 angle, i32 class_id, i32 octave, f32 response, f32 response (the
 *                       reference serialises `response` twice and never `size`), f32 pt.x, f32 pt.y  = 28 bytes; a
 *                       loaded KeyPoint therefore has size = 0 (cv::KeyPoint's default) — kept.
 * Parity: the layout follows the reference's serialize() bodies; no boost is available in this image, so it is not
 * pinned against a file written by the reference binary. */
int orb_mat_record_bytes(int rows, int cols, size_t elem_size, size_t* bytes);
int orb_mat_record_encode(const uint8_t* data, int rows, int cols, size_t elem_size, size_t elem_type, uint8_t* out, size_t cap,
                          size_t* written);
/* *data points INTO buf (no copy); consumed = bytes of the record */
int orb_mat_record_decode(const uint8_t* buf, size_t len, int* rows, int* cols, size_t* elem_size, size_t* elem_type,
                          const uint8_t** data, size_t* consumed);
int orb_keypoint_records_encode(const orb_kp* kps, int n, uint8_t* out /* 28 * n bytes */);
int orb_keypoint_records_decode(const uint8_t* buf, int n, orb_kp* kps);
/* decode one KeyFrame::mDescriptors record (N x 32, CV_8UC1) and append its rows to the device shard; rows_added = N */
int orb_db_add_mat_record(orb_db* db, const uint8_t* buf, size_t len, size_t* consumed, int64_t* rows_added);

/* ---- diagnostics (used by bench.py and the parity tests; not part of the replaced surface) -----------------------------
 * orb_bench_issue_rate: register-only issue-rate microbenchmarks on `device`; kind 0: POPC -> *gops = 1e9 popc/s; kind 1:
 * a 256-bit compare with 8 POPC + the top-2 update -> 1e9 compares/s; kind 2: the same with the search kernel's carry-save
 * (4 POPC) popcount.  orb_debug_sincos_range: the steering coefficients a = (float)cos, b = (float)sin the descriptor kernel
 * computes for the n consecutive fp32 angle bit patterns starting at first_bits (parity pin (iii)). */
int orb_bench_issue_rate(int device, int kind, int iters, double* gops);
int orb_debug_sincos_range(int device, uint32_t first_bits, long long n, float* a, float* b);

#ifdef __cplusplus
}
#endif
#endif /* ORB_B200_H */
