/*
 * orb_oracle_match.cpp — CPU ORACLE (test infrastructure, never shipped / never on the product path).
 *
 * Restates the matcher and stereo parts of the hot path:
 *     /root/reference/orb_slam2/src/ORBmatcher.cc   (cited as OM:<line>)
 *     /root/reference/orb_slam2/src/Frame.cc        (cited as FR:<line>)
 * The reference routines walk Frame / MapPoint objects; the oracle takes the same data as flat arrays
 * (what each loop actually reads), keeps the loop order, the comparison operators (< vs <=), the
 * int/float promotion of every test and the sequential "already matched" dependency.
 */
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>

#include "orb_oracle.h"

extern "C" {
const uint8_t* orc__level_ptr(void* p, int l, int* w, int* h, int* stride);
int orc__nlevels(void* p);
float orc__scale(void* p, int l);
float orc__inv_scale(void* p, int l);
}

namespace {

const int TH_HIGH = 100;     // OM:37
const int TH_LOW = 50;       // OM:38
const int HISTO_LENGTH = 30; // OM:39
const int GRID_COLS = 64;    // Frame.h:38
const int GRID_ROWS = 48;    // Frame.h:37

// OM:1649-1665 — 256-bit Hamming distance, SWAR population count on 8 x 32-bit words
inline int descriptor_distance(const uint8_t* a, const uint8_t* b) {
    uint32_t wa[8], wb[8];
    memcpy(wa, a, 32);
    memcpy(wb, b, 32);
    int dist = 0;
    for (int i = 0; i < 8; ++i) {
        uint32_t v = wa[i] ^ wb[i];
        v -= (v >> 1) & 0x55555555u;                          // 2-bit sums
        v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);     // 4-bit sums
        v = (v + (v >> 4)) & 0x0F0F0F0Fu;                     // 8-bit sums
        dist += (int)((v * 0x01010101u) >> 24);               // add the four bytes
    }
    return dist;
}

struct Grid {
    float min_x, min_y, max_x, max_y, inv_w, inv_h;
    const orc_kp* kps; int n;
    std::vector<int32_t> cell[GRID_COLS][GRID_ROWS];
};

// OM:1603-1644
void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; ++i) {
        const int s = (int)histo[i].size();
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

inline int rot_bin(float a_query, float a_target) {  // OM:1436-1441
    const float factor = 1.0f / HISTO_LENGTH;
    float rot = a_query - a_target;
    if (rot < 0.0) rot += 360.0f;
    int bin = (int)roundf(rot * factor);
    if (bin == HISTO_LENGTH) bin = 0;
    return bin;
}

}  // namespace

extern "C" {

int orc_descriptor_distance(const uint8_t* a, const uint8_t* b) { return descriptor_distance(a, b); }

void orc_hamming_top2(const uint8_t* q, int nq, const uint8_t* db, int ndb, orc_top2* out) {
    for (int i = 0; i < nq; ++i) {
        int best = 256, best_idx = -1, second = 256, second_idx = -1;
        const uint8_t* dq = q + (size_t)i * 32;
        for (int j = 0; j < ndb; ++j) {
            const int d = descriptor_distance(dq, db + (size_t)j * 32);
            if (d < best) { second = best; second_idx = best_idx; best = d; best_idx = j; }
            else if (d < second) { second = d; second_idx = j; }
        }
        out[i] = {best, best_idx, second, second_idx};
    }
}

void orc_hamming_top2_csr(const uint8_t* q, int nq, const uint8_t* db, const int32_t* off, const int32_t* idx,
                          orc_top2* out) {
    for (int i = 0; i < nq; ++i) {
        int best = 256, best_idx = -1, second = 256, second_idx = -1;
        const uint8_t* dq = q + (size_t)i * 32;
        for (int c = off[i]; c < off[i + 1]; ++c) {
            const int j = idx[c];
            const int d = descriptor_distance(dq, db + (size_t)j * 32);
            if (d < best) { second = best; second_idx = best_idx; best = d; best_idx = j; }
            else if (d < second) { second = d; second_idx = j; }
        }
        out[i] = {best, best_idx, second, second_idx};
    }
}

// FR:239-256, 415-425
void* orc_grid_create(const orc_kp* kps, int n, float min_x, float min_y, float max_x, float max_y) {
    Grid* g = new Grid;
    g->min_x = min_x; g->min_y = min_y; g->max_x = max_x; g->max_y = max_y;
    g->inv_w = (float)GRID_COLS / (max_x - min_x);  // FR:162
    g->inv_h = (float)GRID_ROWS / (max_y - min_y);  // FR:163
    g->kps = kps; g->n = n;
    for (int i = 0; i < n; ++i) {
        int px = (int)roundf((kps[i].x - min_x) * g->inv_w);
        int py = (int)roundf((kps[i].y - min_y) * g->inv_h);
        if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) continue;
        g->cell[px][py].push_back(i);
    }
    return g;
}
void orc_grid_destroy(void* g) { delete (Grid*)g; }

// FR:354-412
static void grid_query(const Grid* g, float x, float y, float r, int minLevel, int maxLevel, std::vector<int32_t>& out) {
    out.clear();
    const int nMinCellX = std::max(0, (int)floorf((x - g->min_x - r) * g->inv_w));
    if (nMinCellX >= GRID_COLS) return;
    const int nMaxCellX = std::min(GRID_COLS - 1, (int)ceilf((x - g->min_x + r) * g->inv_w));
    if (nMaxCellX < 0) return;
    const int nMinCellY = std::max(0, (int)floorf((y - g->min_y - r) * g->inv_h));
    if (nMinCellY >= GRID_ROWS) return;
    const int nMaxCellY = std::min(GRID_ROWS - 1, (int)ceilf((y - g->min_y + r) * g->inv_h));
    if (nMaxCellY < 0) return;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ++ix)
        for (int iy = nMinCellY; iy <= nMaxCellY; ++iy)
            for (int32_t id : g->cell[ix][iy]) {
                const orc_kp& kp = g->kps[id];
                if (bCheckLevels) {
                    if (kp.octave < minLevel) continue;
                    if (maxLevel >= 0 && kp.octave > maxLevel) continue;
                }
                const float dx = kp.x - x, dy = kp.y - y;
                if (fabsf(dx) < r && fabsf(dy) < r) out.push_back(id);
            }
}

int orc_grid_query(void* g, float x, float y, float r, int minLevel, int maxLevel, int32_t* out, int cap) {
    std::vector<int32_t> v;
    grid_query((Grid*)g, x, y, r, minLevel, maxLevel, v);
    int n = std::min((int)v.size(), cap);
    memcpy(out, v.data(), (size_t)n * 4);
    return (int)v.size();
}

/* OM:1330-1472 (TRACK_LAST) and OM:45-129 (LOCAL_POINTS); q_obs[i] = query map point has Observations()>0
 * (only then does its assignment block later queries, OM:1405-1407 / OM:87-89).
 * target_query[n] (out): final owner (query index) of each target keypoint, -1 = none, -2 = matched and then removed by
 * the rotation filter (the reference sets that keypoint's map point to NULL, OM:1462-1466). */
int orc_search_by_projection_ex(const orc_search_params* prm, void* grid, const orc_kp* kps_un,
                                const uint8_t* desc, const float* u_right, int n, uint8_t* taken, int nq,
                                const float* q_u, const float* q_v, const float* q_radius,
                                const int32_t* q_min_level, const int32_t* q_max_level, const uint8_t* q_desc,
                                const float* q_ur, const float* q_er_max, const float* q_angle,
                                const uint8_t* q_valid, const uint8_t* q_obs, int32_t* match_of_query,
                                int32_t* target_query) {
    const Grid* g = (const Grid*)grid;
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    std::vector<int32_t> owner(n, -1);
    std::vector<int32_t> cand;
    if (prm->mode == ORC_MODE_INITIALIZATION) {
        // ORBmatcher::SearchForInitialization, OM:406-521.  Queries = F1 keypoints of octave 0 (q_valid), window centre =
        // vbPrevMatched (q_u, q_v), radius = windowSize, levels (0, 0); a target may be stolen by a strictly better
        // match (vMatchedDistance, OM:445) and the previous owner loses it (OM:464-468).
        std::vector<int> vMatchedDistance(n, INT_MAX);
        for (int i = 0; i < nq; ++i) match_of_query[i] = -1;          // vnMatches12
        for (int i1 = 0; i1 < nq; ++i1) {
            if (q_valid && !q_valid[i1]) continue;                     // level1 > 0 (OM:423-425)
            grid_query(g, q_u[i1], q_v[i1], q_radius[i1], q_min_level[i1], q_max_level[i1], cand);
            if (cand.empty()) continue;
            const uint8_t* d1 = q_desc + (size_t)i1 * 32;
            int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
            for (int32_t i2 : cand) {
                const int dist = descriptor_distance(d1, desc + (size_t)i2 * 32);
                if (vMatchedDistance[i2] <= dist) continue;
                if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
                else if (dist < bestDist2) { bestDist2 = dist; }
            }
            if (bestDist <= prm->th_dist) {
                if (bestDist < (float)bestDist2 * prm->nn_ratio) {
                    if (owner[bestIdx2] >= 0) { match_of_query[owner[bestIdx2]] = -1; nmatches--; }
                    match_of_query[i1] = bestIdx2;
                    owner[bestIdx2] = i1;
                    vMatchedDistance[bestIdx2] = bestDist;
                    nmatches++;
                    if (prm->check_orientation) rotHist[rot_bin(q_angle[i1], kps_un[bestIdx2].angle)].push_back(i1);
                }
            }
        }
        if (prm->check_orientation) {
            int ind1 = -1, ind2 = -1, ind3 = -1;
            three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
            for (int b = 0; b < HISTO_LENGTH; ++b) {
                if (b == ind1 || b == ind2 || b == ind3) continue;
                for (int idx1 : rotHist[b])
                    if (match_of_query[idx1] >= 0) { match_of_query[idx1] = -1; nmatches--; }
            }
        }
        if (target_query) memcpy(target_query, owner.data(), (size_t)n * 4);   // vnMatches21 (not touched by the filter)
        return nmatches;
    }
    for (int i = 0; i < nq; ++i) {
        match_of_query[i] = -1;
        if (q_valid && !q_valid[i]) continue;
        grid_query(g, q_u[i], q_v[i], q_radius[i], q_min_level[i], q_max_level[i], cand);
        if (cand.empty()) continue;
        const uint8_t* dq = q_desc + (size_t)i * 32;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int32_t idx : cand) {
            if (taken[idx]) continue;
            if (u_right && u_right[idx] > 0) {
                const float er = fabsf(q_ur[i] - u_right[idx]);
                if (er > q_er_max[i]) continue;
            }
            const int dist = descriptor_distance(dq, desc + (size_t)idx * 32);
            if (prm->mode == ORC_MODE_LOCAL_POINTS) {
                if (dist < bestDist) {
                    bestDist2 = bestDist; bestDist = dist;
                    bestLevel2 = bestLevel; bestLevel = kps_un[idx].octave;
                    bestIdx = idx;
                } else if (dist < bestDist2) {
                    bestLevel2 = kps_un[idx].octave;
                    bestDist2 = dist;
                }
            } else {
                if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
            }
        }
        if (bestDist <= prm->th_dist) {
            if (prm->mode == ORC_MODE_LOCAL_POINTS) {
                if (bestLevel == bestLevel2 && (float)bestDist > prm->nn_ratio * (float)bestDist2) continue;
            }
            match_of_query[i] = bestIdx;
            owner[bestIdx] = i;
            if (!q_obs || q_obs[i]) taken[bestIdx] = 1;
            nmatches++;
            if (prm->mode == ORC_MODE_TRACK_LAST && prm->check_orientation)
                rotHist[rot_bin(q_angle[i], kps_un[bestIdx].angle)].push_back(bestIdx);
        }
    }
    if (prm->mode == ORC_MODE_TRACK_LAST && prm->check_orientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int b = 0; b < HISTO_LENGTH; ++b) {
            if (b == ind1 || b == ind2 || b == ind3) continue;
            for (int idx : rotHist[b]) { owner[idx] = -2; nmatches--; }   /* OM:1462-1466: NULLed, not "left alone" */
        }
    }
    if (target_query) memcpy(target_query, owner.data(), (size_t)n * 4);
    // per-query view after the rotation filter
    for (int i = 0; i < nq; ++i)
        if (match_of_query[i] >= 0 && owner[match_of_query[i]] != i) {
            // removed by the rotation filter, or overwritten by a later query (only when q_obs==0)
            if (owner[match_of_query[i]] == -2) match_of_query[i] = -1;
        }
    return nmatches;
}

int orc_search_by_projection(const orc_search_params* prm, void* grid, const orc_kp* kps_un, const uint8_t* desc,
                             const float* u_right, int n, uint8_t* taken, int nq, const float* q_u,
                             const float* q_v, const float* q_radius, const int32_t* q_min_level,
                             const int32_t* q_max_level, const uint8_t* q_desc, const float* q_ur,
                             const float* q_er_max, const float* q_angle, const uint8_t* q_valid,
                             int32_t* match_of_query) {
    return orc_search_by_projection_ex(prm, grid, kps_un, desc, u_right, n, taken, nq, q_u, q_v, q_radius,
                                       q_min_level, q_max_level, q_desc, q_ur, q_er_max, q_angle, q_valid,
                                       nullptr, match_of_query, nullptr);
}

/* OM:196-252 applied to a single vocabulary node that holds every keypoint of both frames:
 * queries = frame 1 in index order, candidates = frame 2 in index order, already-matched targets skipped
 * (OM:210-211), accept best<=th_dist && (float)best < ratio*(float)second (OM:229-231), rotation
 * histogram on angle1-angle2 (OM:239-248), three-maxima filter (OM:268-285).
 * match12[i] = index in frame 2 or -1.  Returns nmatches. */
int orc_match_bruteforce(const uint8_t* desc1, const float* angle1, int n1, const uint8_t* desc2,
                         const float* angle2, int n2, int th_dist, float nn_ratio, int check_orientation,
                         int32_t* match12) {
    std::vector<int32_t> owner(n2, -1);
    std::vector<int> rotHist[HISTO_LENGTH];
    int nmatches = 0;
    for (int i = 0; i < n1; ++i) {
        match12[i] = -1;
        int best1 = 256, bestIdx = -1, best2 = 256;
        const uint8_t* d1 = desc1 + (size_t)i * 32;
        for (int j = 0; j < n2; ++j) {
            if (owner[j] >= 0) continue;
            const int dist = descriptor_distance(d1, desc2 + (size_t)j * 32);
            if (dist < best1) { best2 = best1; best1 = dist; bestIdx = j; }
            else if (dist < best2) { best2 = dist; }
        }
        if (best1 <= th_dist) {
            if ((float)best1 < nn_ratio * (float)best2) {
                owner[bestIdx] = i;
                match12[i] = bestIdx;
                if (check_orientation) rotHist[rot_bin(angle1[i], angle2[bestIdx])].push_back(bestIdx);
                nmatches++;
            }
        }
    }
    if (check_orientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int b = 0; b < HISTO_LENGTH; ++b) {
            if (b == ind1 || b == ind2 || b == ind3) continue;
            for (int idx : rotHist[b]) { match12[owner[idx]] = -1; owner[idx] = -1; nmatches--; }
        }
    }
    return nmatches;
}

/* SearchByBoW over FeatureVectors (OM:160-289 KeyFrame -> Frame, strict = 0; OM:524-657 KeyFrame -> KeyFrame, strict = 1).
 * Feature vectors in CSR form: node ids ascending, features of node j = fvX_feat[fvX_start[j] .. fvX_start[j+1]).
 * valid1[i]: the query keypoint holds a good map point (OM:196-202 / 562-566); valid2[j] (NULL = all): the target may be
 * matched at all (OM:580-584, KeyFrame variant only).  A target already matched is skipped (OM:210 / 580).  Accept
 * best <= th_dist (strict: <, OM:600) and (float)best < ratio * (float)second; rotation histogram on angle1 - angle2 and
 * three-maxima filter.  match12[n1] = target index or -1; match21[n2] = query index or -1 (the final vpMapPointMatches
 * of the Frame variant).  Returns nmatches. */
int orc_search_by_bow(const uint8_t* desc1, const float* angle1, const uint8_t* valid1, int n1, const int32_t* fv1_node,
                      const int32_t* fv1_start, const int32_t* fv1_feat, int nfv1, const uint8_t* desc2, const float* angle2,
                      const uint8_t* valid2, int n2, const int32_t* fv2_node, const int32_t* fv2_start, const int32_t* fv2_feat,
                      int nfv2, int th_dist, int strict, float nn_ratio, int check_orientation, int32_t* match12, int32_t* match21) {
    for (int i = 0; i < n1; ++i) match12[i] = -1;
    for (int j = 0; j < n2; ++j) match21[j] = -1;
    std::vector<int> rotHist[HISTO_LENGTH];
    int nmatches = 0;
    int a = 0, b = 0;
    while (a < nfv1 && b < nfv2) {
        if (fv1_node[a] == fv2_node[b]) {
            for (int i1 = fv1_start[a]; i1 < fv1_start[a + 1]; ++i1) {
                const int idx1 = fv1_feat[i1];
                if (valid1 && !valid1[idx1]) continue;
                const uint8_t* d1 = desc1 + (size_t)idx1 * 32;
                int best1 = 256, bestIdx2 = -1, best2 = 256;
                for (int i2 = fv2_start[b]; i2 < fv2_start[b + 1]; ++i2) {
                    const int idx2 = fv2_feat[i2];
                    if (match21[idx2] >= 0 || (valid2 && !valid2[idx2])) continue;
                    const int dist = descriptor_distance(d1, desc2 + (size_t)idx2 * 32);
                    if (dist < best1) { best2 = best1; best1 = dist; bestIdx2 = idx2; }
                    else if (dist < best2) { best2 = dist; }
                }
                if (strict ? (best1 < th_dist) : (best1 <= th_dist)) {
                    if ((float)best1 < nn_ratio * (float)best2) {
                        match12[idx1] = bestIdx2;
                        match21[bestIdx2] = idx1;
                        if (check_orientation) rotHist[rot_bin(angle1[idx1], angle2[bestIdx2])].push_back(idx1);
                        nmatches++;
                    }
                }
            }
            ++a; ++b;
        } else if (fv1_node[a] < fv2_node[b]) {
            while (a < nfv1 && fv1_node[a] < fv2_node[b]) ++a;   /* lower_bound */
        } else {
            while (b < nfv2 && fv2_node[b] < fv1_node[a]) ++b;
        }
    }
    if (check_orientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int h = 0; h < HISTO_LENGTH; ++h) {
            if (h == ind1 || h == ind2 || h == ind3) continue;
            for (int idx1 : rotHist[h]) { match21[match12[idx1]] = -1; match12[idx1] = -1; nmatches--; }
        }
    }
    return nmatches;
}

/* SearchForTriangulation (OM:659-825) on arrays.  has_mp[i]: the keypoint already holds a map point (skipped, OM:703-705 /
 * 726-728); u_right (NULL = monocular): bStereo = u_right[i] >= 0; F12 row-major; (ex, ey) epipole in image 2 (OM:665-673);
 * scale_factors / level_sigma2 = pKF2->mvScaleFactors / mvLevelSigma2.  NOTE: the reference never sets vbMatched2, so two
 * queries may pick the same target; candidates are gated `dist > TH_LOW || dist > bestDist` (OM:741), so among equal
 * distances the LAST gate-passing candidate in list order wins.  match12[n1] = vMatches12.  Returns nmatches. */
int orc_search_for_triangulation(const orc_kp* kps1, const uint8_t* desc1, const uint8_t* has_mp1, const float* u_right1, int n1,
                                 const int32_t* fv1_node, const int32_t* fv1_start, const int32_t* fv1_feat, int nfv1,
                                 const orc_kp* kps2, const uint8_t* desc2, const uint8_t* has_mp2, const float* u_right2, int n2,
                                 const int32_t* fv2_node, const int32_t* fv2_start, const int32_t* fv2_feat, int nfv2, const float* F12,
                                 float ex, float ey, const float* scale_factors, const float* level_sigma2, int only_stereo,
                                 int check_orientation, int32_t* match12) {
    int nmatches = 0;
    std::vector<bool> vbMatched2(n2, false);
    for (int i = 0; i < n1; ++i) match12[i] = -1;
    std::vector<int> rotHist[HISTO_LENGTH];
    int a = 0, b = 0;
    while (a < nfv1 && b < nfv2) {
        if (fv1_node[a] == fv2_node[b]) {
            for (int i1 = fv1_start[a]; i1 < fv1_start[a + 1]; ++i1) {
                const int idx1 = fv1_feat[i1];
                if (has_mp1 && has_mp1[idx1]) continue;
                const bool bStereo1 = u_right1 && u_right1[idx1] >= 0;
                if (only_stereo && !bStereo1) continue;
                const orc_kp& kp1 = kps1[idx1];
                const uint8_t* d1 = desc1 + (size_t)idx1 * 32;
                int bestDist = TH_LOW, bestIdx2 = -1;
                for (int i2 = fv2_start[b]; i2 < fv2_start[b + 1]; ++i2) {
                    const int idx2 = fv2_feat[i2];
                    if (vbMatched2[idx2] || (has_mp2 && has_mp2[idx2])) continue;
                    const bool bStereo2 = u_right2 && u_right2[idx2] >= 0;
                    if (only_stereo && !bStereo2) continue;
                    const int dist = descriptor_distance(d1, desc2 + (size_t)idx2 * 32);
                    if (dist > TH_LOW || dist > bestDist) continue;
                    const orc_kp& kp2 = kps2[idx2];
                    if (!bStereo1 && !bStereo2) {
                        const float distex = ex - kp2.x, distey = ey - kp2.y;
                        if (distex * distex + distey * distey < 100 * scale_factors[kp2.octave]) continue;
                    }
                    /* CheckDistEpipolarLine, OM:140-157 */
                    const float la = kp1.x * F12[0] + kp1.y * F12[3] + F12[6];
                    const float lb = kp1.x * F12[1] + kp1.y * F12[4] + F12[7];
                    const float lc = kp1.x * F12[2] + kp1.y * F12[5] + F12[8];
                    const float num = la * kp2.x + lb * kp2.y + lc;
                    const float den = la * la + lb * lb;
                    if (den == 0) continue;
                    const float dsqr = num * num / den;
                    if (dsqr < 3.84 * level_sigma2[kp2.octave]) { bestIdx2 = idx2; bestDist = dist; }
                }
                if (bestIdx2 >= 0) {
                    match12[idx1] = bestIdx2;
                    nmatches++;
                    if (check_orientation) rotHist[rot_bin(kp1.angle, kps2[bestIdx2].angle)].push_back(idx1);
                }
            }
            ++a; ++b;
        } else if (fv1_node[a] < fv2_node[b]) {
            while (a < nfv1 && fv1_node[a] < fv2_node[b]) ++a;
        } else {
            while (b < nfv2 && fv2_node[b] < fv1_node[a]) ++b;
        }
    }
    if (check_orientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int h = 0; h < HISTO_LENGTH; ++h) {
            if (h == ind1 || h == ind2 || h == ind3) continue;
            for (int idx1 : rotHist[h]) { match12[idx1] = -1; nmatches--; }
        }
    }
    return nmatches;
}

/* One direction of SearchBySim3 (OM:1150-1213 / 1216-1279) on arrays: query i = a map point of the other keyframe projected
 * to (q_u, q_v) with radius th * scale[pred] and predicted level q_level (queries the reference skips have q_valid = 0);
 * candidates = GetFeaturesInArea(u, v, radius) (KeyFrame.cc:700-739, no level filter), octave in [pred-1, pred], best
 * distance with strict '<', accepted when <= TH_HIGH.  match[i] = target index or -1. */
void orc_sim3_search_one_way(void* grid, const orc_kp* kps_un, const uint8_t* desc, int nq, const float* q_u, const float* q_v,
                             const float* q_radius, const int32_t* q_level, const uint8_t* q_desc, const uint8_t* q_valid, int32_t* match) {
    std::vector<int32_t> cand;
    for (int i = 0; i < nq; ++i) {
        match[i] = -1;
        if (q_valid && !q_valid[i]) continue;
        grid_query((Grid*)grid, q_u[i], q_v[i], q_radius[i], -1, -1, cand);
        int bestDist = INT_MAX, bestIdx = -1;
        for (int32_t idx : cand) {
            if (kps_un[idx].octave < q_level[i] - 1 || kps_un[idx].octave > q_level[i]) continue;
            const int dist = descriptor_distance(q_desc + (size_t)i * 32, desc + (size_t)idx * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
        }
        if (bestDist <= TH_HIGH) match[i] = bestIdx;
    }
}

/* The search half of ORBmatcher::Fuse (OM:827-977; OM:979-1102 when inv_level_sigma2 == NULL) on arrays: query i = a map
 * point projected into the keyframe at (q_u, q_v) with right coordinate q_ur, radius th * scale[pred], predicted level;
 * candidates = GetFeaturesInArea(u, v, radius), octave in [pred-1, pred], reprojection gate e2 * invSigma2 > 7.8 (stereo
 * keypoint, with er) / > 5.99 (monocular) (OM:914-940), best distance with strict '<'.  best_idx / best_dist receive the
 * winner (-1 / 256 when none); the caller fuses when best_dist <= TH_LOW (the map mutation stays on the host). */
void orc_fuse_search(void* grid, const orc_kp* kps_un, const uint8_t* desc, const float* u_right, const float* inv_level_sigma2,
                     int nq, const float* q_u, const float* q_v, const float* q_ur, const float* q_radius, const int32_t* q_level,
                     const uint8_t* q_desc, const uint8_t* q_valid, int32_t* best_idx, int32_t* best_dist) {
    std::vector<int32_t> cand;
    for (int i = 0; i < nq; ++i) {
        best_idx[i] = -1; best_dist[i] = 256;
        if (q_valid && !q_valid[i]) continue;
        const float u = q_u[i], v = q_v[i];
        grid_query((Grid*)grid, u, v, q_radius[i], -1, -1, cand);
        int bestDist = 256, bestIdx = -1;
        for (int32_t idx : cand) {
            const orc_kp& kp = kps_un[idx];
            const int kpLevel = kp.octave;
            if (kpLevel < q_level[i] - 1 || kpLevel > q_level[i]) continue;
            if (inv_level_sigma2) {
                if (u_right && u_right[idx] >= 0) {
                    const float ex = u - kp.x, ey = v - kp.y, er = q_ur[i] - u_right[idx];
                    const float e2 = ex * ex + ey * ey + er * er;
                    if (e2 * inv_level_sigma2[kpLevel] > 7.8) continue;
                } else {
                    const float ex = u - kp.x, ey = v - kp.y;
                    const float e2 = ex * ex + ey * ey;
                    if (e2 * inv_level_sigma2[kpLevel] > 5.99) continue;
                }
            }
            const int dist = descriptor_distance(q_desc + (size_t)i * 32, desc + (size_t)idx * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
        }
        best_idx[i] = bestIdx; best_dist[i] = bestDist;
    }
}

/* MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:288-361) for npoints map points at once: the observed descriptors of
 * point p are rows [off[p], off[p+1]); best[p] = index (within the point) of the descriptor with the least median distance
 * to the others (first on ties), -1 for a point without observations. */
void orc_distinctive_descriptors(const uint8_t* desc, const int32_t* off, int npoints, int32_t* best) {
    for (int p = 0; p < npoints; ++p) {
        const int N = off[p + 1] - off[p];
        best[p] = -1;
        if (N <= 0) continue;
        const uint8_t* d = desc + (size_t)off[p] * 32;
        std::vector<float> Distances((size_t)N * N);
        for (int i = 0; i < N; i++) {
            Distances[(size_t)i * N + i] = 0;
            for (int j = i + 1; j < N; j++) {
                const int distij = descriptor_distance(d + (size_t)i * 32, d + (size_t)j * 32);
                Distances[(size_t)i * N + j] = (float)distij;
                Distances[(size_t)j * N + i] = (float)distij;
            }
        }
        int BestMedian = INT_MAX, BestIdx = 0;
        for (int i = 0; i < N; i++) {
            std::vector<int> vDists(Distances.begin() + (size_t)i * N, Distances.begin() + (size_t)(i + 1) * N);
            std::sort(vDists.begin(), vDists.end());
            const int median = vDists[(size_t)(0.5 * (N - 1))];
            if (median < BestMedian) { BestMedian = median; BestIdx = i; }
        }
        best[p] = BestIdx;
    }
}

/* FR:502-676 */
int orc_stereo_match(void* exL, void* exR, const orc_kp* kpsL, const uint8_t* descL, int N, const orc_kp* kpsR,
                     const uint8_t* descR, int Nr, const orc_stereo_params* prm, float* mvuRight, float* mvDepth,
                     int32_t* best_sad) {
    for (int i = 0; i < N; ++i) { mvuRight[i] = -1.0f; mvDepth[i] = -1.0f; if (best_sad) best_sad[i] = -1; }
    const int thOrbDist = (TH_HIGH + TH_LOW) / 2;
    int w0, h0, s0;
    orc__level_ptr(exL, 0, &w0, &h0, &s0);
    const int nRows = h0;
    const float mbf = prm->bf, mb = prm->b;

    std::vector<std::vector<int>> vRowIndices(nRows);
    for (int iR = 0; iR < Nr; ++iR) {
        const float kpY = kpsR[iR].y;
        const float r = 2.0f * orc__scale(exL, kpsR[iR].octave);
        const int maxr = (int)ceilf(kpY + r);
        const int minr = (int)floorf(kpY - r);
        for (int yi = minr; yi <= maxr; ++yi)
            if (yi >= 0 && yi < nRows) vRowIndices[yi].push_back(iR);  // guard: the reference indexes unchecked
    }
    const float minZ = mb, minD = 0, maxD = mbf / minZ;
    std::vector<std::pair<int, int>> vDistIdx;
    for (int iL = 0; iL < N; ++iL) {
        const orc_kp& kpL = kpsL[iL];
        const int levelL = kpL.octave;
        const float vL = kpL.y, uL = kpL.x;
        const int row = (int)vL;
        if (row < 0 || row >= nRows) continue;
        const std::vector<int>& vCandidates = vRowIndices[row];
        if (vCandidates.empty()) continue;
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = TH_HIGH;
        int bestIdxR = 0;
        const uint8_t* dL = descL + (size_t)iL * 32;
        for (int iR : vCandidates) {
            const orc_kp& kpR = kpsR[iR];
            if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
            const float uR = kpR.x;
            if (uR >= minU && uR <= maxU) {
                const int dist = descriptor_distance(dL, descR + (size_t)iR * 32);
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        if (bestDist < thOrbDist) {
            const float uR0 = kpsR[bestIdxR].x;
            const float scaleFactor = orc__inv_scale(exL, kpL.octave);
            const float scaleduL = roundf(kpL.x * scaleFactor);
            const float scaledvL = roundf(kpL.y * scaleFactor);
            const float scaleduR0 = roundf(uR0 * scaleFactor);
            const int w = 5;
            int lw, lh, ls, rw, rh, rs;
            const uint8_t* PL = orc__level_ptr(exL, kpL.octave, &lw, &lh, &ls);
            const uint8_t* PR = orc__level_ptr(exR, kpL.octave, &rw, &rh, &rs);
            const int y0 = (int)(scaledvL - w), xl0 = (int)(scaleduL - w);
            float IL[11][11];
            {
                const float c = (float)PL[(size_t)(y0 + w) * ls + xl0 + w];
                for (int yy = 0; yy < 11; ++yy)
                    for (int xx = 0; xx < 11; ++xx) IL[yy][xx] = (float)PL[(ptrdiff_t)(y0 + yy) * ls + xl0 + xx] - c;
            }
            int bestDistS = INT_MAX;
            int bestincR = 0;
            const int L = 5;
            float vDists[2 * 5 + 1];
            const float iniu = scaleduR0 + L - w;
            const float endu = scaleduR0 + L + w + 1;
            if (iniu < 0 || endu >= rw) continue;
            for (int incR = -L; incR <= +L; ++incR) {
                const int xr0 = (int)(scaleduR0 + incR - w);
                const float c = (float)PR[(ptrdiff_t)(y0 + w) * rs + xr0 + w];
                double acc = 0;  // cv::norm(NORM_L1) accumulates CV_32F in double; all terms are small integers
                for (int yy = 0; yy < 11; ++yy)
                    for (int xx = 0; xx < 11; ++xx) {
                        const float ir = (float)PR[(ptrdiff_t)(y0 + yy) * rs + xr0 + xx] - c;
                        acc += fabs((double)(IL[yy][xx] - ir));
                    }
                float dist = (float)acc;
                if (dist < bestDistS) { bestDistS = (int)dist; bestincR = incR; }
                vDists[L + incR] = dist;
            }
            if (best_sad) best_sad[iL] = bestDistS;
            if (bestincR == -L || bestincR == L) continue;
            const float dist1 = vDists[L + bestincR - 1];
            const float dist2 = vDists[L + bestincR];
            const float dist3 = vDists[L + bestincR + 1];
            const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
            if (deltaR < -1 || deltaR > 1) continue;
            float bestuR = orc__scale(exL, kpL.octave) * ((float)scaleduR0 + (float)bestincR + deltaR);
            float disparity = (uL - bestuR);
            if (disparity >= minD && disparity < maxD) {
                if (disparity <= 0) {
                    disparity = 0.01;
                    bestuR = uL - 0.01;
                }
                mvDepth[iL] = mbf / disparity;
                mvuRight[iL] = bestuR;
                vDistIdx.push_back(std::pair<int, int>(bestDistS, iL));
            }
        }
    }
    if (vDistIdx.empty()) return 0;  // guard: the reference reads vDistIdx[0] of an empty vector here
    std::sort(vDistIdx.begin(), vDistIdx.end());
    const float median = vDistIdx[vDistIdx.size() / 2].first;
    const float thDist = 1.5f * 1.4f * median;
    int kept = (int)vDistIdx.size();
    for (int i = (int)vDistIdx.size() - 1; i >= 0; --i) {
        if (vDistIdx[i].first < thDist) break;
        mvuRight[vDistIdx[i].second] = -1;
        mvDepth[vDistIdx[i].second] = -1;
        kept--;
    }
    return kept;
}

}  // extern "C"
