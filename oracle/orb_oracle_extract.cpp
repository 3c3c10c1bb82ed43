/*
 * orb_oracle_extract.cpp — CPU ORACLE (test infrastructure, never shipped / never on the product path).
 *
 * Restates, step by step, the reference extractor
 *     /root/reference/orb_slam2/src/ORBextractor.cc   (cited below as OE:<line>)
 * with every OpenCV primitive it calls replaced by an explicit integer / IEEE-float recipe that
 * reproduces OpenCV 4.13.0 bit for bit (pinned by tests/test_oracle_vs_cv2.py against the cv2 wheel).
 * See orb_oracle.h for the three parity pins.  Written from the algorithm's description, no reference
 * code is copied; the 1024-integer rBRIEF pattern is data (orb_pattern_31.inc).
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <list>
#include <utility>
#include <vector>

#include "orb_oracle.h"

namespace {

const int kPatch = 31;      // PATCH_SIZE            OE:70
const int kHalfPatch = 15;  // HALF_PATCH_SIZE       OE:71
const int kEdge = 19;       // EDGE_THRESHOLD        OE:72

static const int kPattern[1024] = {
#include "orb_pattern_31.inc"
};

// ---- OpenCV scalar helpers -------------------------------------------------------------------
// cvRound: round-half-to-even (SSE cvtss2si / lrint in the default rounding mode).
inline int cv_round(float v) { return (int)lrintf(v); }
inline int cv_round(double v) { return (int)lrint(v); }
inline int cv_floor(float v) { return (int)floorf(v); }

inline int reflect101(int i, int n) {
    // gfedcb|abcdefgh|gfedcba ; a single reflection is enough for |border| < n
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        else i = 2 * n - 2 - i;
    }
    return i;
}

// ---- resize(INTER_LINEAR) on CV_8UC1 (OpenCV imgproc resize.cpp fixed-point path) -----------
struct LinTab {
    std::vector<int> ofs;       // source index of the left/top tap
    std::vector<short> c0, c1;  // Q11 weights
};

// horizontal axis table: taps are clamped and the weight collapses to (2048,0) at the borders
LinTab make_tab_x(int sn, int dn) {
    LinTab t;
    t.ofs.resize(dn); t.c0.resize(dn); t.c1.resize(dn);
    const double scale = 1.0 / ((double)dn / (double)sn);
    for (int d = 0; d < dn; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = cv_floor(f);
        f -= (float)s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= sn - 1) { s = sn - 1; f = 0.f; }
        t.ofs[d] = s;
        t.c0[d] = (short)cv_round((1.f - f) * 2048.f);
        t.c1[d] = (short)cv_round(f * 2048.f);
    }
    return t;
}
// vertical axis table: the row indices are clipped, the weights are NOT collapsed
LinTab make_tab_y(int sn, int dn) {
    LinTab t;
    t.ofs.resize(dn); t.c0.resize(dn); t.c1.resize(dn);
    const double scale = 1.0 / ((double)dn / (double)sn);
    for (int d = 0; d < dn; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = cv_floor(f);
        f -= (float)s;
        t.ofs[d] = s;
        t.c0[d] = (short)cv_round((1.f - f) * 2048.f);
        t.c1[d] = (short)cv_round(f * 2048.f);
    }
    return t;
}

void resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh,
                      int dstride) {
    if (sw == dw && sh == dh) {
        for (int y = 0; y < sh; ++y) memcpy(dst + (size_t)y * dstride, src + (size_t)y * sstride, sw);
        return;
    }
    LinTab tx = make_tab_x(sw, dw), ty = make_tab_y(sh, dh);
    std::vector<int> r0(dw), r1(dw);
    auto hrow = [&](int sy, std::vector<int>& out) {
        const uint8_t* S = src + (size_t)sy * sstride;
        for (int d = 0; d < dw; ++d) {
            int s = tx.ofs[d];
            int s1 = std::min(s + 1, sw - 1);
            out[d] = (int)S[s] * tx.c0[d] + (int)S[s1] * tx.c1[d];
        }
    };
    for (int y = 0; y < dh; ++y) {
        int s0 = std::min(std::max(ty.ofs[y], 0), sh - 1);
        int s1 = std::min(std::max(ty.ofs[y] + 1, 0), sh - 1);
        hrow(s0, r0);
        hrow(s1, r1);
        const int b0 = ty.c0[y], b1 = ty.c1[y];
        uint8_t* D = dst + (size_t)y * dstride;
        for (int d = 0; d < dw; ++d) {
            int v = (((b0 * (r0[d] >> 4)) >> 16) + ((b1 * (r1[d] >> 4)) >> 16) + 2) >> 2;
            D[d] = (uint8_t)std::min(std::max(v, 0), 255);
        }
    }
}

// copyMakeBorder(BORDER_REFLECT_101): dst is (w+2b)x(h+2b); src may alias the interior of dst.
void border_reflect101(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int b, int dstride) {
    // interior first (memmove: src may be the interior of dst)
    for (int y = 0; y < h; ++y) {
        uint8_t* drow = dst + (size_t)(y + b) * dstride + b;
        const uint8_t* srow = src + (size_t)y * sstride;
        if (drow != srow) memmove(drow, srow, w);
    }
    // left/right
    for (int y = 0; y < h; ++y) {
        uint8_t* row = dst + (size_t)(y + b) * dstride;
        for (int x = 0; x < b; ++x) {
            row[x] = row[b + reflect101(x - b, w)];
            row[b + w + x] = row[b + reflect101(w + x, w)];
        }
    }
    // top/bottom (full rows, already containing left/right borders)
    const int W = w + 2 * b;
    for (int y = 0; y < b; ++y) {
        memcpy(dst + (size_t)y * dstride, dst + (size_t)(b + reflect101(y - b, h)) * dstride, W);
        memcpy(dst + (size_t)(b + h + y) * dstride, dst + (size_t)(b + reflect101(h + y, h)) * dstride, W);
    }
}

// ---- FAST 9/16 with 3x3 non-max suppression (OpenCV features2d fast.cpp semantics) ------------
const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

// score = max( max_arc min_k (c - r_k), max_arc min_k (r_k - c) ) - 1 over the 16 arcs of 9 contiguous
// ring pixels; a pixel is a corner at threshold t  <=>  score >= t.
inline int fast_score(const uint8_t* p, const int* off) {
    int c = p[0];
    int r[25];
    for (int k = 0; k < 16; ++k) r[k] = p[off[k]];
    for (int k = 16; k < 25; ++k) r[k] = r[k - 16];
    int best_dark = -256, best_bright = -256;  // max over arcs of min(c-r) / min(r-c)
    for (int k = 0; k < 16; ++k) {
        int mx = r[k], mn = r[k];
        for (int j = 1; j < 9; ++j) {
            mx = std::max(mx, r[k + j]);
            mn = std::min(mn, r[k + j]);
        }
        best_dark = std::max(best_dark, c - mx);
        best_bright = std::max(best_bright, mn - c);
    }
    return std::max(best_dark, best_bright) - 1;
}

int fast9_16(const uint8_t* img, int w, int h, int stride, int threshold, bool nms, std::vector<orc_kp>& out) {
    out.clear();
    if (w < 7 || h < 7) return 0;
    int off[16];
    for (int k = 0; k < 16; ++k) off[k] = kRingDy[k] * stride + kRingDx[k];
    std::vector<int> score((size_t)w * h, 0);  // 0 = not a corner at this threshold
    std::vector<std::pair<int, int>> corners;
    const int t = threshold;
    for (int y = 3; y < h - 3; ++y) {
        const uint8_t* row = img + (size_t)y * stride;
        for (int x = 3; x < w - 3; ++x) {
            const uint8_t* p = row + x;
            const int c = p[0];
            const int lo = c - t, hi = c + t;
            // quick reject: every opposite pair (k,k+8) must contain a darker (resp. brighter) pixel
            int a = p[off[0]], b = p[off[8]];
            bool dk = (a < lo) | (b < lo), br = (a > hi) | (b > hi);
            if (!(dk | br)) continue;
            a = p[off[4]]; b = p[off[12]];
            dk &= (a < lo) | (b < lo); br &= (a > hi) | (b > hi);
            if (!(dk | br)) continue;
            a = p[off[2]]; b = p[off[10]];
            dk &= (a < lo) | (b < lo); br &= (a > hi) | (b > hi);
            if (!(dk | br)) continue;
            a = p[off[6]]; b = p[off[14]];
            dk &= (a < lo) | (b < lo); br &= (a > hi) | (b > hi);
            if (!(dk | br)) continue;
            int s = fast_score(p, off);
            if (s >= t) {
                score[(size_t)y * w + x] = s;
                corners.emplace_back(x, y);
            }
        }
    }
    for (auto& xy : corners) {
        int x = xy.first, y = xy.second;
        int s = score[(size_t)y * w + x];
        if (nms) {
            const int* sc = &score[(size_t)y * w + x];
            if (!(s > sc[-1] && s > sc[1] && s > sc[-w - 1] && s > sc[-w] && s > sc[-w + 1] && s > sc[w - 1] &&
                  s > sc[w] && s > sc[w + 1]))
                continue;
        }
        orc_kp kp;
        kp.x = (float)x; kp.y = (float)y; kp.size = 7.f; kp.angle = -1.f; kp.response = (float)s;
        kp.octave = 0; kp.class_id = -1;
        out.push_back(kp);
    }
    return (int)out.size();
}

// ---- GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) on CV_8UC1, OpenCV >= 4 fixed-point path ----
void gaussian7x7_s2(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride) {
    static const int K[7] = {18, 34, 48, 56, 48, 34, 18};  // Q8, sum 256
    std::vector<uint16_t> hbuf((size_t)w * h);
    for (int y = 0; y < h; ++y) {
        const uint8_t* S = src + (size_t)y * sstride;
        for (int x = 0; x < w; ++x) {
            int acc = 0;
            for (int k = 0; k < 7; ++k) acc += K[k] * S[reflect101(x + k - 3, w)];
            hbuf[(size_t)y * w + x] = (uint16_t)acc;  // <= 255*256
        }
    }
    for (int y = 0; y < h; ++y) {
        uint8_t* D = dst + (size_t)y * dstride;
        for (int x = 0; x < w; ++x) {
            uint32_t acc = 0;
            for (int k = 0; k < 7; ++k) acc += (uint32_t)K[k] * hbuf[(size_t)reflect101(y + k - 3, h) * w + x];
            D[x] = (uint8_t)((acc + 32768u) >> 16);
        }
    }
}

// ---- fastAtan2 (OpenCV core mathfuncs_core, float32 polynomial, degrees) ----------------------
float fast_atan2(float y, float x) {
    const float p1 = 0.9997878412794807f * (float)(180 / M_PI);
    const float p3 = -0.3258083974640975f * (float)(180 / M_PI);
    const float p5 = 0.1555786518463281f * (float)(180 / M_PI);
    const float p7 = -0.04432655554792128f * (float)(180 / M_PI);
    const float eps = (float)2.2204460492503131e-16;  // (float)DBL_EPSILON
    // built with -ffp-contract=off: every * and + below is a separately rounded fp32 operation
    float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + eps);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + eps);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// ---- IC_Angle (OE:77-104) ----------------------------------------------------------------------
struct Umax { int v[16]; };
Umax make_umax() {  // OE:463-478
    Umax u;
    int vmax = (int)floorf(kHalfPatch * sqrtf(2.f) / 2 + 1);
    int vmin = (int)ceilf(kHalfPatch * sqrtf(2.f) / 2);
    const double hp2 = kHalfPatch * kHalfPatch;
    for (int v = 0; v <= vmax; ++v) u.v[v] = cv_round(sqrt(hp2 - v * v));
    for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
        while (u.v[v0] == u.v[v0 + 1]) ++v0;
        u.v[v] = v0;
        ++v0;
    }
    return u;
}
const Umax kUmax = make_umax();

float ic_angle(const uint8_t* center, int step) {
    int m_01 = 0, m_10 = 0;
    for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m_10 += u * center[u];
    for (int v = 1; v <= kHalfPatch; ++v) {
        int v_sum = 0;
        int d = kUmax.v[v];
        for (int u = -d; u <= d; ++u) {
            int val_plus = center[u + v * step], val_minus = center[u - v * step];
            v_sum += (val_plus - val_minus);
            m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
    }
    return fast_atan2((float)m_01, (float)m_10);
}

// ---- computeOrbDescriptor (OE:106-147) -----------------------------------------------------------
const float kFactorPI = (float)(M_PI / 180.f);

struct TapStat { long near_half; float eps; };

void brief_descriptor(const uint8_t* center, int step, float angle_deg, uint8_t* desc, TapStat* st) {
    float angle = angle_deg * kFactorPI;
    // pin (iii): double-precision libm, rounded once to float
    float a = (float)cos((double)angle), b = (float)sin((double)angle);
    auto tap = [&](int idx) -> int {
        const float px = (float)kPattern[2 * idx], py = (float)kPattern[2 * idx + 1];
        float fy = px * b + py * a, fx = px * a - py * b;  // -ffp-contract=off: no FMA
        if (st) {
            float ry = fabsf(fy - floorf(fy) - 0.5f), rx = fabsf(fx - floorf(fx) - 0.5f);
            if (ry < st->eps || rx < st->eps) st->near_half++;
        }
        return center[cv_round(fy) * step + cv_round(fx)];
    };
    for (int i = 0; i < 32; ++i) {
        int val = 0;
        for (int k = 0; k < 8; ++k) {
            int t0 = tap(16 * i + 2 * k), t1 = tap(16 * i + 2 * k + 1);
            val |= (t0 < t1) << k;
        }
        desc[i] = (uint8_t)val;
    }
}

// ---- quadtree (OE:498-787) -------------------------------------------------------------------------
struct P2i { int x, y; };
struct Node {
    std::vector<orc_kp> keys;
    P2i UL, UR, BL, BR;
    std::list<Node>::iterator lit;
    bool no_more = false;
    int seq = 0;  // creation sequence: pin (ii)
};

void divide_node(const Node& n, Node& n1, Node& n2, Node& n3, Node& n4) {
    const int halfX = (int)ceilf((float)(n.UR.x - n.UL.x) / 2);
    const int halfY = (int)ceilf((float)(n.BR.y - n.UL.y) / 2);
    n1.UL = n.UL;                         n1.UR = {n.UL.x + halfX, n.UL.y};
    n1.BL = {n.UL.x, n.UL.y + halfY};     n1.BR = {n.UL.x + halfX, n.UL.y + halfY};
    n2.UL = n1.UR;  n2.UR = n.UR;  n2.BL = n1.BR;  n2.BR = {n.UR.x, n.UL.y + halfY};
    n3.UL = n1.BL;  n3.UR = n1.BR; n3.BL = n.BL;   n3.BR = {n1.BR.x, n.BL.y};
    n4.UL = n3.UR;  n4.UR = n2.BR; n4.BL = n3.BR;  n4.BR = n.BR;
    for (const orc_kp& kp : n.keys) {
        if (kp.x < (float)n1.UR.x) {
            if (kp.y < (float)n1.BR.y) n1.keys.push_back(kp);
            else n3.keys.push_back(kp);
        } else if (kp.y < (float)n1.BR.y) n2.keys.push_back(kp);
        else n4.keys.push_back(kp);
    }
    if (n1.keys.size() == 1) n1.no_more = true;
    if (n2.keys.size() == 1) n2.no_more = true;
    if (n3.keys.size() == 1) n3.no_more = true;
    if (n4.keys.size() == 1) n4.no_more = true;
}

struct SizeSeq {
    int size; int seq; Node* node;
    bool operator<(const SizeSeq& o) const { return size != o.size ? size < o.size : seq < o.seq; }
};

std::vector<orc_kp> distribute_quadtree(const std::vector<orc_kp>& to_dist, int minX, int maxX, int minY,
                                         int maxY, int N, int* tie_at_cut) {
    const int nIni = (int)roundf((float)(maxX - minX) / (maxY - minY));
    const float hX = (float)(maxX - minX) / nIni;
    std::list<Node> nodes;
    std::vector<Node*> ini(nIni);
    int seq = 0;
    for (int i = 0; i < nIni; ++i) {
        Node ni;
        ni.UL = {(int)(hX * (float)i), 0};
        ni.UR = {(int)(hX * (float)(i + 1)), 0};
        ni.BL = {ni.UL.x, maxY - minY};
        ni.BR = {ni.UR.x, maxY - minY};
        ni.seq = seq++;
        nodes.push_back(ni);
        ini[i] = &nodes.back();
    }
    for (const orc_kp& kp : to_dist) ini[(int)(kp.x / hX)]->keys.push_back(kp);
    for (auto it = nodes.begin(); it != nodes.end();) {
        if (it->keys.size() == 1) { it->no_more = true; ++it; }
        else if (it->keys.empty()) it = nodes.erase(it);
        else ++it;
    }
    bool finish = false;
    std::vector<SizeSeq> cand;
    auto push_child = [&](Node& c, int* n_expand) {
        if (c.keys.empty()) return;
        c.seq = seq++;
        nodes.push_front(c);
        if (c.keys.size() > 1) {
            if (n_expand) ++*n_expand;
            cand.push_back({(int)c.keys.size(), nodes.front().seq, &nodes.front()});
            nodes.front().lit = nodes.begin();
        }
    };
    while (!finish) {
        int prev = (int)nodes.size();
        int n_expand = 0;
        cand.clear();
        for (auto it = nodes.begin(); it != nodes.end();) {
            if (it->no_more) { ++it; continue; }
            Node n1, n2, n3, n4;
            divide_node(*it, n1, n2, n3, n4);
            push_child(n1, &n_expand); push_child(n2, &n_expand);
            push_child(n3, &n_expand); push_child(n4, &n_expand);
            it = nodes.erase(it);
        }
        if ((int)nodes.size() >= N || (int)nodes.size() == prev) {
            finish = true;
        } else if ((int)nodes.size() + n_expand * 3 > N) {
            while (!finish) {
                prev = (int)nodes.size();
                std::vector<SizeSeq> prev_cand = cand;
                cand.clear();
                std::sort(prev_cand.begin(), prev_cand.end());
                for (int j = (int)prev_cand.size() - 1; j >= 0; --j) {
                    Node n1, n2, n3, n4;
                    divide_node(*prev_cand[j].node, n1, n2, n3, n4);
                    push_child(n1, nullptr); push_child(n2, nullptr);
                    push_child(n3, nullptr); push_child(n4, nullptr);
                    nodes.erase(prev_cand[j].node->lit);
                    if ((int)nodes.size() >= N) {
                        if (j > 0 && prev_cand[j - 1].size == prev_cand[j].size && tie_at_cut) ++*tie_at_cut;
                        break;
                    }
                }
                if ((int)nodes.size() >= N || (int)nodes.size() == prev) finish = true;
            }
        }
    }
    std::vector<orc_kp> result;
    result.reserve(nodes.size());
    for (Node& n : nodes) {
        const orc_kp* best = &n.keys[0];
        float max_resp = best->response;
        for (size_t k = 1; k < n.keys.size(); ++k)
            if (n.keys[k].response > max_resp) { best = &n.keys[k]; max_resp = n.keys[k].response; }
        result.push_back(*best);
    }
    return result;
}

// ---- the extractor object ---------------------------------------------------------------------------
struct Level {
    int w = 0, h = 0, stride = 0;      // interior size, bordered stride (w + 38)
    std::vector<uint8_t> buf;          // (w+38)*(h+38)
    std::vector<uint8_t> blurred;      // w*h (only when the level has keypoints)
    bool has_blur = false;
    std::vector<orc_kp> raw, kps;
    int tie_at_cut = 0, retried = 0;
    uint8_t* interior() { return buf.data() + (size_t)kEdge * stride + kEdge; }
};

struct Extractor {
    int nfeatures; double scaleFactor; int nlevels, iniTh, minTh;
    std::vector<float> scale, inv_scale, sigma2, inv_sigma2;
    std::vector<int> per_level;
    std::vector<Level> lv;
    long near_half = 0;

    Extractor(int nf, float sf, int nl, int ini, int mn)
        : nfeatures(nf), scaleFactor(sf), nlevels(nl), iniTh(ini), minTh(mn) {
        // OE:416-455
        scale.resize(nl); sigma2.resize(nl); inv_scale.resize(nl); inv_sigma2.resize(nl);
        scale[0] = 1.f; sigma2[0] = 1.f;
        for (int i = 1; i < nl; ++i) {
            scale[i] = (float)(scale[i - 1] * scaleFactor);  // float * double -> float
            sigma2[i] = scale[i] * scale[i];
        }
        for (int i = 0; i < nl; ++i) { inv_scale[i] = 1.f / scale[i]; inv_sigma2[i] = 1.f / sigma2[i]; }
        per_level.resize(nl);
        float factor = (float)(1.0f / scaleFactor);
        float desired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
        int sum = 0;
        for (int l = 0; l < nl - 1; ++l) {
            per_level[l] = cv_round(desired);
            sum += per_level[l];
            desired *= factor;
        }
        per_level[nl - 1] = std::max(nfeatures - sum, 0);
        lv.resize(nl);
    }

    // OE:1152-1185
    void compute_pyramid(const uint8_t* img, int w, int h, int stride) {
        for (int l = 0; l < nlevels; ++l) {
            Level& L = lv[l];
            L.w = cv_round((float)w * inv_scale[l]);
            L.h = cv_round((float)h * inv_scale[l]);
            L.stride = L.w + 2 * kEdge;
            L.buf.assign((size_t)L.stride * (L.h + 2 * kEdge), 0);
            L.has_blur = false;
            if (l != 0) {
                Level& P = lv[l - 1];
                resize_linear_u8(P.interior(), P.w, P.h, P.stride, L.interior(), L.w, L.h, L.stride);
                border_reflect101(L.interior(), L.w, L.h, L.stride, L.buf.data(), kEdge, L.stride);
            } else {
                border_reflect101(img, w, h, stride, L.buf.data(), kEdge, L.stride);
            }
        }
    }

    // OE:790-892
    int compute_keypoints() {
        const float W = 30;
        for (int l = 0; l < nlevels; ++l) {
            Level& L = lv[l];
            L.raw.clear(); L.kps.clear(); L.tie_at_cut = 0; L.retried = 0;
            const int minBX = kEdge - 3, minBY = minBX;
            const int maxBX = L.w - kEdge + 3, maxBY = L.h - kEdge + 3;
            const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
            const int nCols = (int)(width / W), nRows = (int)(height / W);
            if (nCols <= 0 || nRows <= 0) return -1;  // the reference divides by zero here
            const int wCell = (int)ceilf(width / nCols), hCell = (int)ceilf(height / nRows);
            std::vector<orc_kp> cell;
            for (int i = 0; i < nRows; ++i) {
                const float iniY = (float)(minBY + i * hCell);
                float maxY = iniY + hCell + 6;
                if (iniY >= maxBY - 3) continue;
                if (maxY > maxBY) maxY = (float)maxBY;
                for (int j = 0; j < nCols; ++j) {
                    const float iniX = (float)(minBX + j * wCell);
                    float maxX = iniX + wCell + 6;
                    if (iniX >= maxBX - 6) continue;
                    if (maxX > maxBX) maxX = (float)maxBX;
                    const uint8_t* sub = L.interior() + (size_t)(int)iniY * L.stride + (int)iniX;
                    const int cw = (int)maxX - (int)iniX, ch = (int)maxY - (int)iniY;
                    fast9_16(sub, cw, ch, L.stride, iniTh, true, cell);
                    if (cell.empty()) {
                        fast9_16(sub, cw, ch, L.stride, minTh, true, cell);
                        L.retried++;
                    }
                    for (orc_kp& kp : cell) {
                        kp.x += j * wCell;
                        kp.y += i * hCell;
                        L.raw.push_back(kp);
                    }
                }
            }
            if (maxBY - minBY <= 0) return -1;
            const int nIni = (int)roundf((float)(maxBX - minBX) / (maxBY - minBY));
            if (nIni <= 0) return -1;  // the reference divides by zero here
            if (!L.raw.empty())
                L.kps = distribute_quadtree(L.raw, minBX, maxBX, minBY, maxBY, per_level[l], &L.tie_at_cut);
            const int scaledPatch = (int)(kPatch * scale[l]);
            for (orc_kp& kp : L.kps) {
                kp.x += minBX; kp.y += minBY;
                kp.octave = l;
                kp.size = (float)scaledPatch;
            }
        }
        for (int l = 0; l < nlevels; ++l) {
            Level& L = lv[l];
            for (orc_kp& kp : L.kps) {
                const uint8_t* c = L.interior() + (size_t)cv_round(kp.y) * L.stride + cv_round(kp.x);
                kp.angle = ic_angle(c, L.stride);
            }
        }
        return 0;
    }

    // OE:1083-1149
    int run(const uint8_t* img, int w, int h, int stride, orc_kp* out, uint8_t* desc, int cap) {
        if (!img || w <= 0 || h <= 0) return 0;
        compute_pyramid(img, w, h, stride);
        if (compute_keypoints() != 0) return -1000000;
        int total = 0;
        for (int l = 0; l < nlevels; ++l) total += (int)lv[l].kps.size();
        if (total > cap) return -total;
        near_half = 0;
        TapStat st{0, 1e-4f};
        int off = 0;
        for (int l = 0; l < nlevels; ++l) {
            Level& L = lv[l];
            if (L.kps.empty()) continue;
            L.blurred.assign((size_t)L.w * L.h, 0);
            gaussian7x7_s2(L.interior(), L.w, L.h, L.stride, L.blurred.data(), L.w);
            L.has_blur = true;
            for (size_t i = 0; i < L.kps.size(); ++i) {
                const orc_kp& kp = L.kps[i];
                const uint8_t* c = L.blurred.data() + (size_t)cv_round(kp.y) * L.w + cv_round(kp.x);
                brief_descriptor(c, L.w, kp.angle, desc + (size_t)(off + i) * 32, &st);
            }
            for (size_t i = 0; i < L.kps.size(); ++i) {
                orc_kp kp = L.kps[i];
                if (l != 0) { kp.x *= scale[l]; kp.y *= scale[l]; }
                out[off + i] = kp;
            }
            off += (int)L.kps.size();
        }
        near_half = st.near_half;
        return total;
    }
};

}  // namespace

// ================================ C interface =====================================================
extern "C" {

void* orc_extractor_create(int nf, float sf, int nl, int ini, int mn) { return new Extractor(nf, sf, nl, ini, mn); }
void orc_extractor_destroy(void* ex) { delete (Extractor*)ex; }

void orc_extractor_tables(void* p, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                          int32_t* per_level, int32_t* umax16) {
    Extractor* ex = (Extractor*)p;
    for (int i = 0; i < ex->nlevels; ++i) {
        if (scale) scale[i] = ex->scale[i];
        if (inv_scale) inv_scale[i] = ex->inv_scale[i];
        if (sigma2) sigma2[i] = ex->sigma2[i];
        if (inv_sigma2) inv_sigma2[i] = ex->inv_sigma2[i];
        if (per_level) per_level[i] = ex->per_level[i];
    }
    if (umax16) for (int i = 0; i < 16; ++i) umax16[i] = kUmax.v[i];
}

int orc_extract(void* p, const uint8_t* img, int w, int h, int stride, orc_kp* kps, uint8_t* desc, int cap) {
    return ((Extractor*)p)->run(img, w, h, stride, kps, desc, cap);
}

int orc_level_dims(void* p, int l, int* w, int* h) {
    Extractor* ex = (Extractor*)p;
    if (l < 0 || l >= ex->nlevels) return -1;
    *w = ex->lv[l].w; *h = ex->lv[l].h;
    return 0;
}
int orc_get_level(void* p, int l, uint8_t* dst) {
    Extractor* ex = (Extractor*)p;
    if (l < 0 || l >= ex->nlevels) return -1;
    memcpy(dst, ex->lv[l].buf.data(), ex->lv[l].buf.size());
    return 0;
}
int orc_get_blurred(void* p, int l, uint8_t* dst) {
    Extractor* ex = (Extractor*)p;
    if (l < 0 || l >= ex->nlevels || !ex->lv[l].has_blur) return -1;
    memcpy(dst, ex->lv[l].blurred.data(), ex->lv[l].blurred.size());
    return 0;
}
int orc_raw_corner_count(void* p, int l) { return (int)((Extractor*)p)->lv[l].raw.size(); }
int orc_get_raw_corners(void* p, int l, orc_kp* out, int cap) {
    auto& v = ((Extractor*)p)->lv[l].raw;
    int n = std::min((int)v.size(), cap);
    memcpy(out, v.data(), (size_t)n * sizeof(orc_kp));
    return n;
}
int orc_level_kp_count(void* p, int l) { return (int)((Extractor*)p)->lv[l].kps.size(); }
int orc_get_level_kps(void* p, int l, orc_kp* out, int cap) {
    auto& v = ((Extractor*)p)->lv[l].kps;
    int n = std::min((int)v.size(), cap);
    memcpy(out, v.data(), (size_t)n * sizeof(orc_kp));
    return n;
}
void orc_get_stats(void* p, int32_t* stats) {
    Extractor* ex = (Extractor*)p;
    for (int l = 0; l < ex->nlevels; ++l) {
        stats[l * 4 + 0] = (int)ex->lv[l].raw.size();
        stats[l * 4 + 1] = (int)ex->lv[l].kps.size();
        stats[l * 4 + 2] = ex->lv[l].tie_at_cut;
        stats[l * 4 + 3] = ex->lv[l].retried;
    }
}
int orc_count_near_half_taps(void* p, float) { return (int)((Extractor*)p)->near_half; }

void orc_resize_linear_u8(const uint8_t* s, int sw, int sh, int ss, uint8_t* d, int dw, int dh, int ds) {
    resize_linear_u8(s, sw, sh, ss, d, dw, dh, ds);
}
void orc_border_reflect101(const uint8_t* s, int w, int h, int ss, uint8_t* d, int b, int ds) {
    border_reflect101(s, w, h, ss, d, b, ds);
}
int orc_fast9_16(const uint8_t* img, int w, int h, int stride, int th, int nms, orc_kp* out, int cap) {
    std::vector<orc_kp> v;
    fast9_16(img, w, h, stride, th, nms != 0, v);
    int n = std::min((int)v.size(), cap);
    memcpy(out, v.data(), (size_t)n * sizeof(orc_kp));
    return (int)v.size();
}
void orc_gaussian7x7_s2(const uint8_t* s, int w, int h, int ss, uint8_t* d, int ds) { gaussian7x7_s2(s, w, h, ss, d, ds); }
float orc_fast_atan2(float y, float x) { return fast_atan2(y, x); }
int orc_cv_round_f(float v) { return cv_round(v); }
float orc_ic_angle(const uint8_t* c, int stride) { return ic_angle(c, stride); }
void orc_brief_descriptor(const uint8_t* c, int stride, float angle, uint8_t* d) { brief_descriptor(c, stride, angle, d, nullptr); }

// pin (iii) checker: a=(float)cos((double)x), b=(float)sin((double)x) for n consecutive fp32 bit patterns
void orc_sincos_range(uint32_t first_bits, long long n, float* a, float* b) {
    for (long long i = 0; i < n; ++i) {
        uint32_t bits = first_bits + (uint32_t)i;
        float x;
        memcpy(&x, &bits, 4);
        a[i] = (float)cos((double)x);
        b[i] = (float)sin((double)x);
    }
}

// accessors used by the stereo oracle (orb_oracle_match.cpp)
const uint8_t* orc__level_ptr(void* p, int l, int* w, int* h, int* stride) {
    Extractor* ex = (Extractor*)p;
    *w = ex->lv[l].w; *h = ex->lv[l].h; *stride = ex->lv[l].stride;
    return ex->lv[l].interior();
}
int orc__nlevels(void* p) { return ((Extractor*)p)->nlevels; }
float orc__scale(void* p, int l) { return ((Extractor*)p)->scale[l]; }
float orc__inv_scale(void* p, int l) { return ((Extractor*)p)->inv_scale[l]; }

/* cvtColor(src, dst, CV_BGR2GRAY / CV_RGB2GRAY / CV_BGRA2GRAY / CV_RGBA2GRAY) on CV_8U, the conversion Tracking::GrabImage*
 * applies before the extractor (Tracking.cc:181-204, 223-234, 253-264).  OpenCV 4.13.0 arithmetic (checked against the cv2
 * wheel by tests/test_oracle_vs_cv2.py): gray = (B*3735 + G*19235 + R*9798 + (1 << 14)) >> 15. */
void orc_cvt_gray(const uint8_t* src, int w, int h, int stride, int channels, int rgb_order, uint8_t* dst, int dstride) {
    for (int y = 0; y < h; ++y) {
        const uint8_t* s = src + (size_t)y * stride;
        uint8_t* d = dst + (size_t)y * dstride;
        for (int x = 0; x < w; ++x, s += channels) {
            const int b = rgb_order ? s[2] : s[0], g = s[1], r = rgb_order ? s[0] : s[2];
            d[x] = (uint8_t)((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15);
        }
    }
}

}  // extern "C"
