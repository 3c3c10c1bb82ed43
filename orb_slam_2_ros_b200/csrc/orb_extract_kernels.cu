// orb_extract_kernels.cu — sm_100a kernels of the ORB extractor (the replacement of
// ORBextractor::operator(), reference orb_slam2/src/ORBextractor.cc:1083-1149).
//
// Stages (one launch each, batched over frames; blockIdx.z / blockIdx.y = frame):
//   K1  pyr_level0_kernel / pyr_resize_kernel   ComputePyramid              ORBextractor.cc:1152-1185
//   K2  fast_cells_kernel                       per-cell FAST + retry       ORBextractor.cc:820-863
//   K3  quadtree_kernel                         DistributeOctTree           ORBextractor.cc:561-787
//   K5  blur_kernel                             GaussianBlur 7x7 sigma 2    ORBextractor.cc:1129-1130
//   K4+K6 orient_describe_kernel                IC_Angle + rBRIEF + output  ORBextractor.cc:77-147,1134-1147
//
// All arithmetic is integer or separately-rounded IEEE fp32 (file is compiled with -fmad=false) so the
// results are bit-identical to the CPU reference semantics (OpenCV 4.13.0 primitives).
#include "orb_internal.cuh"

namespace {

__constant__ int c_pattern[1024] = {
#include "orb_pattern_31.inc"
};
__constant__ int c_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

__device__ __forceinline__ int reflect101(int i, int n) {
    // |i| < n guaranteed for a 19-px border on levels >= 20 px; loop keeps tiny levels correct
    while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * n - 2 - i;
    return i;
}

// ======================================================================================================
// K1a: level 0 = copyMakeBorder(image, 19, BORDER_REFLECT_101)
// ======================================================================================================
__global__ void __launch_bounds__(256)
pyr_level0_kernel(const uint8_t* __restrict__ in, size_t row_stride, size_t frame_stride, uint8_t* __restrict__ pyr,
                  const __grid_constant__ Geometry g) {
    const LevelGeom& L = g.lv[0];
    const int bx4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int by = blockIdx.y * blockDim.y + threadIdx.y;
    const int f = blockIdx.z;
    if (bx4 >= L.pitch || by >= L.rows) return;
    const uint8_t* src = in + (size_t)f * frame_stride + (size_t)reflect101(by - ORB_EDGE, L.h) * row_stride;
    unsigned v = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int bx = bx4 + i;
        unsigned b = 0;
        if (bx < L.w + 2 * ORB_EDGE) b = src[reflect101(bx - ORB_EDGE, L.w)];
        v |= b << (8 * i);
    }
    *reinterpret_cast<unsigned*>(pyr + L.base + (long long)f * L.frame_stride + (long long)by * L.pitch + bx4) = v;
}

// ======================================================================================================
// K1b: level l = resize(level l-1, INTER_LINEAR) + copyMakeBorder(REFLECT_101), fused:
// each thread produces 4 bordered output bytes; a border pixel recomputes the interior pixel it mirrors.
// Fixed-point recipe (OpenCV resize.cpp 8u path): Q11 taps horizontally, then
//   out = (((b0*(H0>>4))>>16) + ((b1*(H1>>4))>>16) + 2) >> 2
// ======================================================================================================
__global__ void __launch_bounds__(256)
pyr_resize_kernel(uint8_t* __restrict__ pyr, const ResizeTap* __restrict__ taps, int level,
                  const __grid_constant__ Geometry g) {
    const LevelGeom& L = g.lv[level];
    const LevelGeom& P = g.lv[level - 1];
    const int bx4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int by = blockIdx.y * blockDim.y + threadIdx.y;
    const int f = blockIdx.z;
    if (bx4 >= L.pitch || by >= L.rows) return;
    const ResizeTap ty = taps[L.ytab + reflect101(by - ORB_EDGE, L.h)];
    const uint8_t* S = pyr + P.base + (long long)f * P.frame_stride + (long long)ORB_EDGE * P.pitch + ORB_EDGE;
    const uint8_t* r0 = S + (long long)ty.s0 * P.pitch;
    const uint8_t* r1 = S + (long long)ty.s1 * P.pitch;
    const int b0 = ty.c0, b1 = ty.c1;
    unsigned v = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int bx = bx4 + i;
        unsigned o = 0;
        if (bx < L.w + 2 * ORB_EDGE) {
            const ResizeTap tx = taps[L.xtab + reflect101(bx - ORB_EDGE, L.w)];
            const int h0 = (int)__ldcg(r0 + tx.s0) * tx.c0 + (int)__ldcg(r0 + tx.s1) * tx.c1;
            const int h1 = (int)__ldcg(r1 + tx.s0) * tx.c0 + (int)__ldcg(r1 + tx.s1) * tx.c1;
            o = (unsigned)((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2) & 0xFFu;
        }
        v |= o << (8 * i);
    }
    *reinterpret_cast<unsigned*>(pyr + L.base + (long long)f * L.frame_stride + (long long)by * L.pitch + bx4) = v;
}

// ======================================================================================================
// tile loader: copies `rows` rows of `width` bytes starting at g (arbitrary alignment, row pitch gp a
// multiple of 4 and base 4-aligned) into smem as aligned 32-bit words.  Returns the byte shift: global byte
// (row r, col x) lands at s[r*sp + shift + x].
// ======================================================================================================
__device__ __forceinline__ int load_tile_u8(uint8_t* s, int sp, const uint8_t* gsrc, long long gp, int width, int rows) {
    const int shift = (int)((uintptr_t)gsrc & 3);
    const uint8_t* g0 = gsrc - shift;
    const int words = (width + shift + 3) >> 2;
    for (int i = threadIdx.x; i < rows * words; i += blockDim.x) {
        const int r = i / words, wd = i - r * words;
        const unsigned v = __ldg(reinterpret_cast<const unsigned*>(g0 + (long long)r * gp) + wd);
        reinterpret_cast<unsigned*>(s + r * sp)[wd] = v;
    }
    return shift;
}

// ======================================================================================================
// K2: per-cell FAST-9/16 + 3x3 NMS + iniThFAST/minThFAST retry.  One CTA per (cell, frame).
//   score(p) = max( max_arc min_k (c - r_k), max_arc min_k (r_k - c) ) - 1  (16 arcs of 9 ring pixels);
//   corner at threshold t <=> score >= t; NMS is confined to the cell's evaluated area, exactly like a
//   cv::FAST call on the cell sub-image; the cell emits its NMS maxima with score >= iniThFAST if there is
//   at least one, else those with score >= minThFAST (the reference's second FAST call).
// ======================================================================================================
#define FAST_TP 72        // smem pitch of the pixel tile  (cell sub-image width <= 66, + alignment shift)
#define FAST_SP 64        // smem pitch of the score tile  (evaluated width <= 60, + 2 apron)
#define FAST_THREADS 128

__device__ __forceinline__ int fast_score16(int c, const int (&r)[16]) {
    int mx3[16], mn3[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        mx3[k] = __vimax3_s32(r[k], r[(k + 1) & 15], r[(k + 2) & 15]);
        mn3[k] = __vimin3_s32(r[k], r[(k + 1) & 15], r[(k + 2) & 15]);
    }
    int amin = 255, bmax = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        amin = min(amin, __vimax3_s32(mx3[k], mx3[(k + 3) & 15], mx3[(k + 6) & 15]));
        bmax = max(bmax, __vimin3_s32(mn3[k], mn3[(k + 3) & 15], mn3[(k + 6) & 15]));
    }
    return max(c - amin, bmax - c) - 1;
}

__global__ void __launch_bounds__(FAST_THREADS)
fast_cells_kernel(const uint8_t* __restrict__ pyr, unsigned long long* __restrict__ corners,
                  int* __restrict__ corner_count, const __grid_constant__ Geometry g) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int cell = blockIdx.x, f = blockIdx.y;
    int l = 0;
    while (l + 1 < g.nlevels && cell >= g.lv[l + 1].cell_base) ++l;
    const LevelGeom& L = g.lv[l];
    const int ci = cell - L.cell_base;
    const int i = ci / L.nCols, j = ci - i * L.nCols;
    // reference ORBextractor.cc:822-837 (all values are integers held in floats there)
    const int iniY = ORB_MINB + i * L.hCell, iniX = ORB_MINB + j * L.wCell;
    if (iniY >= L.maxBY - 3 || iniX >= L.maxBX - 6) return;
    const int maxY = min(iniY + L.hCell + 6, L.maxBY), maxX = min(iniX + L.wCell + 6, L.maxBX);
    const int cw = maxX - iniX, ch = maxY - iniY;
    if (cw < 7 || ch < 7) return;  // cv::FAST returns nothing on such a sub-image
    const int ew = cw - 6, eh = ch - 6, npx = ew * eh;

    // smem carve-up
    uint8_t* tile = smem;                                    // 66 x FAST_TP
    uint8_t* score = tile + 66 * FAST_TP;                    // 62 x FAST_SP (1-px zero apron)
    unsigned short* surv = reinterpret_cast<unsigned short*>(score + 62 * FAST_SP);  // <= 3600
    unsigned long long* outl = reinterpret_cast<unsigned long long*>(surv + 3600);   // <= 900
    __shared__ int s_nsurv, s_nout, s_base, s_any_ini;
    if (threadIdx.x == 0) { s_nsurv = 0; s_nout = 0; s_any_ini = 0; }

    const uint8_t* src = pyr + L.base + (long long)f * L.frame_stride + (long long)(ORB_EDGE + iniY) * L.pitch +
                         ORB_EDGE + iniX;
    const int shift = load_tile_u8(tile, FAST_TP, src, L.pitch, cw, ch);
    for (int k = threadIdx.x; k < (eh + 2) * (FAST_SP / 4); k += blockDim.x) reinterpret_cast<unsigned*>(score)[k] = 0;
    __syncthreads();

    const int tmin = g.min_th, tini = g.ini_th;
    // pass 1: quick reject at minThFAST — every opposite ring pair must hold a darker (brighter) pixel
    for (int p = threadIdx.x; p < npx; p += blockDim.x) {
        const int y = p / ew, x = p - y * ew;
        const uint8_t* t = tile + (y + 3) * FAST_TP + shift + x + 3;
        const int c = t[0], lo = c - tmin, hi = c + tmin;
        int a = t[3 * FAST_TP], b = t[-3 * FAST_TP];               // ring 0, 8
        bool dk = (a < lo) | (b < lo), br = (a > hi) | (b > hi);
        if (dk | br) {
            a = t[3]; b = t[-3];                                   // ring 4, 12
            dk &= (a < lo) | (b < lo); br &= (a > hi) | (b > hi);
            if (dk | br) {
                a = t[2 * FAST_TP + 2]; b = t[-2 * FAST_TP - 2];   // ring 2, 10
                dk &= (a < lo) | (b < lo); br &= (a > hi) | (b > hi);
                a = t[-2 * FAST_TP + 2]; b = t[2 * FAST_TP - 2];   // ring 6, 14
                dk &= (a < lo) | (b < lo); br &= (a > hi) | (b > hi);
                if (dk | br) surv[atomicAdd(&s_nsurv, 1)] = (unsigned short)p;
            }
        }
    }
    __syncthreads();
    const int nsurv = s_nsurv;
    // pass 2: exact score of the survivors
    for (int s = threadIdx.x; s < nsurv; s += blockDim.x) {
        const int p = surv[s];
        const int y = p / ew, x = p - y * ew;
        const uint8_t* t = tile + (y + 3) * FAST_TP + shift + x + 3;
        int r[16];
        r[0] = t[3 * FAST_TP];       r[1] = t[3 * FAST_TP + 1];   r[2] = t[2 * FAST_TP + 2];   r[3] = t[FAST_TP + 3];
        r[4] = t[3];                 r[5] = t[-FAST_TP + 3];      r[6] = t[-2 * FAST_TP + 2];  r[7] = t[-3 * FAST_TP + 1];
        r[8] = t[-3 * FAST_TP];      r[9] = t[-3 * FAST_TP - 1];  r[10] = t[-2 * FAST_TP - 2]; r[11] = t[-FAST_TP - 3];
        r[12] = t[-3];               r[13] = t[FAST_TP - 3];      r[14] = t[2 * FAST_TP - 2];  r[15] = t[3 * FAST_TP - 1];
        const int sc = fast_score16(t[0], r);
        if (sc >= tmin) score[(y + 1) * FAST_SP + x + 1] = (uint8_t)sc;
    }
    __syncthreads();
    // pass 3: 3x3 NMS (strict >, neighbours outside the evaluated area count as 0)
    unsigned maxmask = 0;  // bit k: survivor threadIdx.x + k*blockDim.x is an NMS maximum
    bool any_ini = false;
    for (int s = threadIdx.x, k = 0; s < nsurv; s += blockDim.x, ++k) {
        const int p = surv[s];
        const int y = p / ew, x = p - y * ew;
        const uint8_t* q = score + (y + 1) * FAST_SP + x + 1;
        const int sc = q[0];
        if (sc != 0 && sc > q[-1] && sc > q[1] && sc > q[-FAST_SP - 1] && sc > q[-FAST_SP] && sc > q[-FAST_SP + 1] &&
            sc > q[FAST_SP - 1] && sc > q[FAST_SP] && sc > q[FAST_SP + 1]) {
            maxmask |= 1u << k;
            any_ini |= (sc >= tini);
        }
    }
    if (any_ini) s_any_ini = 1;
    __syncthreads();
    const int th = s_any_ini ? tini : tmin;
    for (int s = threadIdx.x, k = 0; s < nsurv; s += blockDim.x, ++k) {
        if (!((maxmask >> k) & 1u)) continue;
        const int p = surv[s];
        const int y = p / ew, x = p - y * ew;
        const int sc = score[(y + 1) * FAST_SP + x + 1];
        if (sc < th) continue;
        // keypoint position in vToDistributeKeys coordinates (ORBextractor.cc:856-857)
        outl[atomicAdd(&s_nout, 1)] = corner_pack(x + 3 + j * L.wCell, y + 3 + i * L.hCell, sc, (ci << 12) | (y << 6) | x);
    }
    __syncthreads();
    const int nout = s_nout;
    if (nout == 0) return;
    if (threadIdx.x == 0) s_base = atomicAdd(&corner_count[f * g.nlevels + l], nout);
    __syncthreads();
    unsigned long long* dst = corners + L.corner_base + (long long)f * L.corner_cap + s_base;
    for (int k = threadIdx.x; k < nout; k += blockDim.x) dst[k] = outl[k];
}

// ======================================================================================================
// K3: DistributeOctTree as block-wide scans.  One CTA per (level, frame).
// Nodes keep stable ids; `list` holds the ids in std::list order.  One iteration expands the nodes of a
// processing order S (phase 1: every multi-point node in list order, ORBextractor.cc:630-689; phase 2: the
// previous iteration's multi-point children sorted by (size, creation index) and walked from the back with
// the size>=N cut, ORBextractor.cc:700-761).  New list = reverse(children in creation order) ++ (old list
// minus expanded parents), exactly what push_front / erase produce.
// ======================================================================================================
#define QT_THREADS 256

struct QtNode { short x0, y0, x1, y1; };

__device__ __forceinline__ int block_exclusive_scan(int* data, int n, int* s_warp /*[8]*/, int* s_carry) {
    // in-place exclusive scan of data[0..n), returns the total; all threads must call
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) *s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += QT_THREADS) {
        const int i = base + threadIdx.x;
        const int v = (i < n) ? data[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[wid] = x;
        __syncthreads();
        int woff = 0;
        for (int k = 0; k < wid; ++k) woff += s_warp[k];
        const int carry = *s_carry;
        if (i < n) data[i] = carry + woff + x - v;
        __syncthreads();
        if (threadIdx.x == QT_THREADS - 1) *s_carry = carry + woff + x;
        __syncthreads();
    }
    return *s_carry;
}

__global__ void __launch_bounds__(QT_THREADS)
quadtree_kernel(const unsigned long long* __restrict__ corners, const int* __restrict__ corner_count,
                unsigned short* __restrict__ node_of_key, unsigned long long* __restrict__ kept,
                int* __restrict__ kept_count, int* __restrict__ tie_count, const __grid_constant__ Geometry g) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int l = blockIdx.x, f = blockIdx.y;
    const LevelGeom& L = g.lv[l];
    const int cap = g.max_node_cap;
    const int K = min(corner_count[f * g.nlevels + l], L.corner_cap);
    const int N = L.quota;
    if (K == 0) {
        if (threadIdx.x == 0) kept_count[f * g.nlevels + l] = 0;
        return;
    }
    const unsigned long long* keys = corners + L.corner_base + (long long)f * L.corner_cap;
    unsigned short* nok = node_of_key + L.corner_base + (long long)f * L.corner_cap;

    // smem carve-up (cap entries each unless noted)
    unsigned long long* best = reinterpret_cast<unsigned long long*>(smem);         // 8*cap
    int* cnt = reinterpret_cast<int*>(best + cap);                                    // 4*cap
    int* childcnt = cnt + cap;                                                        // 16*cap
    int* scan_a = childcnt + 4 * cap;                                                 // 4*cap
    int* scan_b = scan_a + cap;                                                       // 4*cap
    int* cand_ci = scan_b + cap;                                                      // 4*cap
    int* cand_ci2 = cand_ci + cap;                                                    // 4*cap
    QtNode* nodes = reinterpret_cast<QtNode*>(cand_ci2 + cap);                        // 8*cap
    unsigned short* childid = reinterpret_cast<unsigned short*>(nodes + cap);         // 8*cap
    unsigned short* list0 = childid + 4 * cap;                                        // 2*cap
    unsigned short* list1 = list0 + cap;                                              // 2*cap
    unsigned short* S = list1 + cap;                                                  // 2*cap
    unsigned short* cand_id = S + cap;                                                // 2*cap
    unsigned short* cand_id2 = cand_id + cap;                                         // 2*cap
    short2* split = reinterpret_cast<short2*>(cand_id2 + cap);                        // 4*cap
    uint8_t* expanding = reinterpret_cast<uint8_t*>(split + cap);                     // cap
    __shared__ int s_warp[8], s_carry, s_L, s_nS, s_nExp, s_ncand, s_ncand2, s_nexp_children, s_nextid;

    const int W = L.maxBX - ORB_MINB, H = L.maxBY - ORB_MINB;
    const float hX = L.hX;
    const int nIni = L.nIni;
    // ---- roots (ORBextractor.cc:566-606) ----
    for (int i = threadIdx.x; i < cap; i += blockDim.x) { cnt[i] = 0; expanding[i] = 0; }
    __syncthreads();
    for (int i = threadIdx.x; i < nIni; i += blockDim.x) {
        QtNode n;
        n.x0 = (short)(int)(hX * (float)i); n.x1 = (short)(int)(hX * (float)(i + 1));
        n.y0 = 0; n.y1 = (short)H;
        nodes[i] = n;
    }
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        int root = (int)((float)corner_x(keys[k]) / hX);
        root = min(root, nIni - 1);
        atomicAdd(&cnt[root], 1);
        nok[k] = (unsigned short)root;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int n = 0;
        for (int i = 0; i < nIni; ++i)
            if (cnt[i] > 0) list0[n++] = (unsigned short)i;
        s_L = n; s_nextid = nIni; s_ncand = 0;
    }
    __syncthreads();
    (void)W;

    unsigned short* list = list0;
    unsigned short* nlist = list1;
    int phase = 1;
    bool finish = false;
    while (!finish) {
        const int Lcur = s_L;
        const int nextid = s_nextid;
        // ---- 1. processing order S ----
        if (phase == 1) {
            for (int p = threadIdx.x; p < Lcur; p += blockDim.x) scan_a[p] = cnt[list[p]] > 1;
            __syncthreads();
            const int nS = block_exclusive_scan(scan_a, Lcur, s_warp, &s_carry);
            for (int p = threadIdx.x; p < Lcur; p += blockDim.x)
                if (cnt[list[p]] > 1) S[scan_a[p]] = list[p];
            if (threadIdx.x == 0) s_nS = nS;
        } else {
            const int nc = s_ncand;
            for (int a = threadIdx.x; a < nc; a += blockDim.x) {
                const int ca = cnt[cand_id[a]], ia = cand_ci[a];
                int rank = 0;  // number of candidates processed before a: larger (size, creation index)
                for (int b = 0; b < nc; ++b) {
                    const int cb = cnt[cand_id[b]], ib = cand_ci[b];
                    rank += (cb > ca) || (cb == ca && ib > ia);
                }
                S[rank] = cand_id[a];
            }
            if (threadIdx.x == 0) s_nS = nc;
        }
        __syncthreads();
        const int nS = s_nS;
        if (nS == 0) break;  // nothing can be divided: list size unchanged -> finish (ORBextractor.cc:693)
        // ---- 2. split points, flags ----
        for (int r = threadIdx.x; r < nS; r += blockDim.x) {
            const int id = S[r];
            const QtNode n = nodes[id];
            const int halfX = (n.x1 - n.x0 + 1) >> 1, halfY = (n.y1 - n.y0 + 1) >> 1;  // ceil(d/2)
            split[id] = make_short2((short)(n.x0 + halfX), (short)(n.y0 + halfY));
            expanding[id] = 1;
            childcnt[4 * id + 0] = 0; childcnt[4 * id + 1] = 0; childcnt[4 * id + 2] = 0; childcnt[4 * id + 3] = 0;
        }
        __syncthreads();
        // ---- 3. key pass A: children sizes (DivideNode, ORBextractor.cc:529-543) ----
        for (int k = threadIdx.x; k < K; k += blockDim.x) {
            const int id = nok[k];
            if (expanding[id]) {
                const unsigned long long rec = keys[k];
                const short2 sp = split[id];
                const int c = (corner_x(rec) >= sp.x ? 1 : 0) + (corner_y(rec) >= sp.y ? 2 : 0);
                atomicAdd(&childcnt[4 * id + c], 1);
            }
        }
        __syncthreads();
        // ---- 4. how many of S are expanded (phase 2: cut at size >= N), creation offsets ----
        for (int r = threadIdx.x; r < nS; r += blockDim.x) {
            const int id = S[r];
            const int nc = (childcnt[4 * id] > 0) + (childcnt[4 * id + 1] > 0) + (childcnt[4 * id + 2] > 0) +
                           (childcnt[4 * id + 3] > 0);
            scan_a[r] = nc;       // -> creation offset
            scan_b[r] = nc - 1;   // -> id offset / size gain
        }
        __syncthreads();
        block_exclusive_scan(scan_b, nS, s_warp, &s_carry);
        if (threadIdx.x == 0) s_nExp = nS;
        __syncthreads();
        if (phase == 2) {
            // size after expanding S[0..r] = Lcur + scan_b[r] + gain_r ; first r reaching N ends the walk
            for (int r = threadIdx.x; r < nS; r += blockDim.x) {
                const int after = Lcur + scan_b[r] + (scan_a[r] - 1);
                const int before = Lcur + scan_b[r];
                if (after >= N && before < N) s_nExp = r + 1;
            }
            __syncthreads();
        }
        const int nExp = s_nExp;
        if (phase == 2 && nExp < nS && threadIdx.x == 0 && cnt[S[nExp]] == cnt[S[nExp - 1]])
            atomicAdd(&tie_count[f * g.nlevels + l], 1);  // the cut fell inside an equal-size group (pin (ii))
        const int T = block_exclusive_scan(scan_a, nExp, s_warp, &s_carry);  // total children created
        if (threadIdx.x == 0) { s_ncand2 = 0; s_nexp_children = 0; }
        __syncthreads();
        // ---- 5. create children ----
        for (int r = threadIdx.x; r < nExp; r += blockDim.x) {
            const int id = S[r];
            const QtNode n = nodes[id];
            const short2 sp = split[id];
            int q = 0;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int cc = childcnt[4 * id + c];
                if (cc == 0) continue;
                int cid = (q == 0) ? id : nextid + scan_b[r] + (q - 1);
                if (cid >= cap) cid = id;  // cannot happen (node_cap bound); keeps indices in range
                QtNode ch;
                ch.x0 = (c & 1) ? sp.x : n.x0; ch.x1 = (c & 1) ? n.x1 : sp.x;
                ch.y0 = (c & 2) ? sp.y : n.y0; ch.y1 = (c & 2) ? n.y1 : sp.y;
                const int ci = scan_a[r] + q;           // creation index within this iteration
                if (cid < cap) {
                    nodes[cid] = ch;
                    cnt[cid] = cc;
                    nlist[T - 1 - ci] = (unsigned short)cid;
                    if (cc > 1) {
                        const int slot = atomicAdd(&s_ncand2, 1);
                        cand_id2[slot] = (unsigned short)cid;
                        cand_ci2[slot] = ci;
                    }
                }
                childid[4 * id + c] = (unsigned short)cid;
                ++q;
            }
        }
        // nodes of S beyond the cut stay in the list unexpanded
        for (int r = nExp + threadIdx.x; r < nS; r += blockDim.x) expanding[S[r]] = 0;
        __syncthreads();
        // ---- 6. old list minus expanded parents keeps its order behind the new children ----
        for (int p = threadIdx.x; p < Lcur; p += blockDim.x) scan_a[p] = expanding[list[p]] ? 0 : 1;
        __syncthreads();
        const int nkeep = block_exclusive_scan(scan_a, Lcur, s_warp, &s_carry);
        for (int p = threadIdx.x; p < Lcur; p += blockDim.x) {
            const int id = list[p];
            if (!expanding[id] && T + scan_a[p] < cap) nlist[T + scan_a[p]] = (unsigned short)id;
        }
        // ---- 7. key pass B: move keys into the children ----
        for (int k = threadIdx.x; k < K; k += blockDim.x) {
            const int id = nok[k];
            if (expanding[id]) {
                const unsigned long long rec = keys[k];
                const short2 sp = split[id];
                const int c = (corner_x(rec) >= sp.x ? 1 : 0) + (corner_y(rec) >= sp.y ? 2 : 0);
                nok[k] = childid[4 * id + c];
            }
        }
        __syncthreads();
        for (int r = threadIdx.x; r < nExp; r += blockDim.x) expanding[S[r]] = 0;
        const int Lnew = min(T + nkeep, cap);
        const int ncand2 = s_ncand2;
        __syncthreads();
        // candidates of the next phase-2 round
        for (int a = threadIdx.x; a < ncand2; a += blockDim.x) { cand_id[a] = cand_id2[a]; cand_ci[a] = cand_ci2[a]; }
        if (threadIdx.x == 0) { s_L = Lnew; s_nextid = nextid + (Lnew - Lcur); s_ncand = ncand2; }
        { unsigned short* t = list; list = nlist; nlist = t; }
        __syncthreads();
        // ---- 8. termination (ORBextractor.cc:693-699, 758-759) ----
        if (Lnew >= N || Lnew == Lcur) finish = true;
        else if (phase == 1 && Lnew + 3 * ncand2 > N) phase = 2;
    }
    __syncthreads();
    // ---- keep the best-response key of every node (first in reference order on ties), list order ----
    const int Lfin = s_L;
    for (int i = threadIdx.x; i < cap; i += blockDim.x) best[i] = 0ull;
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) atomicMax(&best[nok[k]], keys[k]);
    __syncthreads();
    unsigned long long* out = kept + (long long)f * g.total_kp_slots + L.kp_base;
    for (int p = threadIdx.x; p < Lfin; p += blockDim.x) out[p] = best[list[p]];
    if (threadIdx.x == 0) kept_count[f * g.nlevels + l] = Lfin;
}

// ======================================================================================================
// K5: GaussianBlur(7x7, sigma 2) — OpenCV >= 4 fixed-point path: Q8 kernel [18 34 48 56 48 34 18],
// horizontal pass exact in 16 bits, vertical pass rounded (V + 32768) >> 16.  Reads the bordered pyramid
// (whose 19-px reflect-101 border IS the blur's BORDER_REFLECT_101), writes the w x h blurred level.
// ======================================================================================================
#define BL_TW 64
#define BL_TH 32
#define BL_TP 76   // (BL_TW + 6 + 3 alignment) rounded up to a multiple of 4
#define BL_THREADS 256

__global__ void __launch_bounds__(BL_THREADS)
blur_kernel(const uint8_t* __restrict__ pyr, uint8_t* __restrict__ blur, const int* __restrict__ tile_level_base,
            const __grid_constant__ Geometry g) {
    __shared__ __align__(16) uint8_t tile[(BL_TH + 6) * BL_TP];
    __shared__ __align__(16) unsigned short hbuf[(BL_TH + 6) * BL_TW];
    const int t = blockIdx.x, f = blockIdx.y;
    int l = 0;
    while (l + 1 < g.nlevels && t >= tile_level_base[l + 1]) ++l;
    const LevelGeom& L = g.lv[l];
    const int tl = t - tile_level_base[l];
    const int tiles_x = (L.w + BL_TW - 1) / BL_TW;
    const int ty = tl / tiles_x, tx = tl - ty * tiles_x;
    const int x0 = tx * BL_TW, y0 = ty * BL_TH;
    const int tw = min(BL_TW, L.w - x0), th = min(BL_TH, L.h - y0);
    const uint8_t* src = pyr + L.base + (long long)f * L.frame_stride + (long long)(ORB_EDGE + y0 - 3) * L.pitch +
                         ORB_EDGE + x0 - 3;
    const int shift = load_tile_u8(tile, BL_TP, src, L.pitch, tw + 6, th + 6);
    __syncthreads();
    for (int i = threadIdx.x; i < (th + 6) * tw; i += blockDim.x) {
        const int r = i / tw, x = i - r * tw;
        const uint8_t* p = tile + r * BL_TP + shift + x;
        hbuf[r * BL_TW + x] = (unsigned short)(18 * (p[0] + p[6]) + 34 * (p[1] + p[5]) + 48 * (p[2] + p[4]) + 56 * p[3]);
    }
    __syncthreads();
    uint8_t* dst = blur + L.bbase + (long long)f * L.bframe_stride + (long long)y0 * L.bpitch + x0;
    const int tw4 = (tw + 3) >> 2;
    for (int i = threadIdx.x; i < th * tw4; i += blockDim.x) {
        const int r = i / tw4, x4 = (i - r * tw4) * 4;
        unsigned v = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned short* h = hbuf + r * BL_TW + x4 + k;   // columns >= tw hold stale data, masked by pitch
            const unsigned acc = 18u * (h[0] + h[6 * BL_TW]) + 34u * (h[BL_TW] + h[5 * BL_TW]) +
                                 48u * (h[2 * BL_TW] + h[4 * BL_TW]) + 56u * h[3 * BL_TW];
            v |= ((acc + 32768u) >> 16) << (8 * k);
        }
        *reinterpret_cast<unsigned*>(dst + (long long)r * L.bpitch + x4) = v;  // bpitch, x0 multiples of 4
    }
}

// ======================================================================================================
// K4 + K6: one warp per kept keypoint: IC_Angle (integer moments + fastAtan2 polynomial), rBRIEF-256 on
// the blurred level, coordinate scaling and the final KeyPoint / descriptor rows in level-major order.
// ======================================================================================================
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    // OpenCV fastAtan2 (core/mathfuncs_core): fp32, every operation rounded separately
    const float rad2deg = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * rad2deg, p3 = -0.3258083974640975f * rad2deg;
    const float p5 = 0.1555786518463281f * rad2deg, p7 = -0.04432655554792128f * rad2deg;
    const float eps = (float)2.2204460492503131e-16;  // (float)DBL_EPSILON
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

#define OD_WARPS 8
#define OD_KPW 4                       // keypoints per warp
#define OD_KPB (OD_WARPS * OD_KPW)     // keypoints per CTA (== 32: one lane of warp 0 per keypoint in phase 2)

// Three phases per CTA of 32 keypoint slots:
//   1. every warp accumulates the patch moments of its 4 keypoints (lane = patch column, 31 row loads in flight)
//   2. warp 0, one LANE per keypoint: fastAtan2 + the double-precision sin/cos of pin (iii) (thread-parallel, so the
//      long fp64 sequence is issued once per 32 keypoints instead of once per keypoint)
//   3. every warp builds the 4 descriptors (lane = descriptor byte); the 512-point pattern sits in shared memory as
//      float4 [8][32] so that a warp-wide read is conflict-free (a per-lane index into __constant__ serialises)
__global__ void __launch_bounds__(OD_WARPS * 32)
orient_describe_kernel(const uint8_t* __restrict__ pyr, const uint8_t* __restrict__ blur,
                       const unsigned long long* __restrict__ kept, const int* __restrict__ kept_count,
                       orb_kp* __restrict__ kps_out, uint8_t* __restrict__ desc_out, int cap,
                       int* __restrict__ n_out, const __grid_constant__ Geometry g) {
    __shared__ float4 s_pat[8 * 32];
    __shared__ int s_m[OD_KPB][2];
    __shared__ float s_ang[OD_KPB], s_a[OD_KPB], s_b[OD_KPB];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = blockIdx.y;
    const int* kc = kept_count + f * g.nlevels;
    // pattern: byte `i` of the descriptor uses points 16 i .. 16 i + 15; s_pat[k * 32 + i] = (x0, y0, x1, y1) of bit k
    {
        const int k = threadIdx.x >> 5, i = threadIdx.x & 31;
        const int* p = c_pattern + i * 32 + 4 * k;   // one-off divergent constant reads, 4 per thread
        s_pat[k * 32 + i] = make_float4((float)p[0], (float)p[1], (float)p[2], (float)p[3]);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int tot = 0;
        for (int l = 0; l < g.nlevels; ++l) tot += kc[l];
        n_out[f] = tot;
    }
    // ---- slot -> (level, index, output row) for this warp's 4 keypoints ----
    int lv[OD_KPW], px[OD_KPW], py[OD_KPW], off[OD_KPW], sc[OD_KPW];
#pragma unroll
    for (int q = 0; q < OD_KPW; ++q) {
        const int slot = blockIdx.x * OD_KPB + warp * OD_KPW + q;
        lv[q] = -1; px[q] = py[q] = off[q] = sc[q] = 0;
        if (slot < g.total_kp_slots) {
            int l = 0;
            while (l + 1 < g.nlevels && slot >= g.lv[l + 1].kp_base) ++l;
            const int i = slot - g.lv[l].kp_base;
            if (i < kc[l]) {
                int o = i;
                for (int k = 0; k < l; ++k) o += kc[k];
                if (o < cap) {
                    const unsigned long long rec = kept[(long long)f * g.total_kp_slots + slot];
                    lv[q] = l; off[q] = o; sc[q] = corner_score(rec);
                    px[q] = corner_x(rec) + ORB_MINB; py[q] = corner_y(rec) + ORB_MINB;   // ORBextractor.cc:881-882
                }
            }
        }
    }
    // ---- phase 1: IC_Angle moments (ORBextractor.cc:77-104) ----
#pragma unroll
    for (int q = 0; q < OD_KPW; ++q) {
        int m10 = 0, m01 = 0;
        if (lv[q] >= 0 && lane < 31) {
            const LevelGeom& L = g.lv[lv[q]];
            const int pitch = L.pitch;
            const uint8_t* center = pyr + L.base + (long long)f * L.frame_stride + (ORB_EDGE + py[q]) * pitch + ORB_EDGE + px[q];
            const int u = lane - ORB_HALF_PATCH;
            const int au = abs(u);
            int col = 0;
#pragma unroll
            for (int v = -ORB_HALF_PATCH; v <= ORB_HALF_PATCH; ++v) {
                const int um = (v < 0 ? -v : v);
                // umax = 15 15 15 15 14 14 14 13 13 12 11 10 9 8 6 3 (ORBextractor.cc:463-478), folded at compile time
                const int lim = um <= 3 ? 15 : um <= 6 ? 14 : um <= 8 ? 13 : um == 9 ? 12 : um == 10 ? 11 : um == 11 ? 10 :
                                um == 12 ? 9 : um == 13 ? 8 : um == 14 ? 6 : 3;
                if (au <= lim) {
                    const int val = center[v * pitch + u];
                    col += val;
                    m01 += v * val;
                }
            }
            m10 = u * col;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            m10 += __shfl_xor_sync(0xffffffffu, m10, o);
            m01 += __shfl_xor_sync(0xffffffffu, m01, o);
        }
        if (lane == 0) { s_m[warp * OD_KPW + q][0] = m01; s_m[warp * OD_KPW + q][1] = m10; }
    }
    __syncthreads();
    // ---- phase 2: angle and steering coefficients, one lane per keypoint ----
    if (warp == 0) {
        const float angle = fast_atan2_deg((float)s_m[lane][0], (float)s_m[lane][1]);
        const float factorPI = (float)(3.1415926535897932384626433832795 / 180.0);  // (float)(CV_PI/180.f)
        const float ang = __fmul_rn(angle, factorPI);
        double sn, cs;
        sincos((double)ang, &sn, &cs);   // pin (iii): a = (float)cos((double)ang), b = (float)sin((double)ang)
        s_ang[lane] = angle; s_a[lane] = (float)cs; s_b[lane] = (float)sn;
    }
    __syncthreads();
    // ---- phase 3: computeOrbDescriptor (ORBextractor.cc:106-147): lane i builds byte i ----
#pragma unroll
    for (int q = 0; q < OD_KPW; ++q) {
        if (lv[q] < 0) continue;   // warp-uniform
        const int l = lv[q];
        const LevelGeom& L = g.lv[l];
        const int kq = warp * OD_KPW + q;
        const float a = s_a[kq], b = s_b[kq];
        const int bp = L.bpitch;
        const uint8_t* bc = blur + L.bbase + (long long)f * L.bframe_stride + py[q] * bp + px[q];
        unsigned val = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float4 pt = s_pat[k * 32 + lane];
            const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(pt.x, b), __fmul_rn(pt.y, a)));
            const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(pt.x, a), __fmul_rn(pt.y, b)));
            const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(pt.z, b), __fmul_rn(pt.w, a)));
            const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(pt.z, a), __fmul_rn(pt.w, b)));
            const int t0 = bc[r0 * bp + c0], t1 = bc[r1 * bp + c1];
            val |= (unsigned)(t0 < t1) << k;
        }
        // 32-byte descriptor row: gather 4 lanes' bytes into one word, 8 lanes store 8 words (one 32-B sector)
        unsigned w = val;
        w |= __shfl_down_sync(0xffffffffu, val, 1) << 8;
        w |= __shfl_down_sync(0xffffffffu, val, 2) << 16;
        w |= __shfl_down_sync(0xffffffffu, val, 3) << 24;
        const long long o = (long long)f * cap + off[q];
        if ((lane & 3) == 0) reinterpret_cast<unsigned*>(desc_out + o * 32)[lane >> 2] = w;
        if (lane == 0) {
            orb_kp kp;
            kp.x = (l != 0) ? __fmul_rn((float)px[q], L.scale) : (float)px[q];   // ORBextractor.cc:1139-1145
            kp.y = (l != 0) ? __fmul_rn((float)py[q], L.scale) : (float)py[q];
            kp.size = L.size;
            kp.angle = s_ang[kq];
            kp.response = (float)sc[q];
            kp.octave = l;
            kp.class_id = -1;
            kps_out[o] = kp;
        }
    }
}

// debug: the steering coefficients a=(float)cos((double)x), b=(float)sin((double)x) of computeOrbDescriptor for
// the n consecutive fp32 bit patterns starting at first_bits (exhaustive parity check against the host libm)
__global__ void sincos_range_kernel(unsigned first_bits, long long n, float* __restrict__ a, float* __restrict__ b) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = __uint_as_float(first_bits + (unsigned)i);
    double sn, cs;
    sincos((double)x, &sn, &cs);   // the same call as orient_describe_kernel phase 2
    a[i] = (float)cs;
    b[i] = (float)sn;
}

}  // namespace

extern "C" int orb_debug_sincos_range(int device, unsigned first_bits, long long n, float* a, float* b) {
    if (n <= 0 || !a || !b) return ORB_ERR_INVALID;
    if (orb_device_count() <= 0) { orb_set_error("no CUDA device visible"); return ORB_ERR_NO_DEVICE; }
    ORB_CUDA(cudaSetDevice(device));
    float *da = nullptr, *db = nullptr;
    ORB_CUDA(cudaMalloc(&da, sizeof(float) * n));
    ORB_CUDA(cudaMalloc(&db, sizeof(float) * n));
    sincos_range_kernel<<<(unsigned)((n + 255) / 256), 256>>>(first_bits, n, da, db);
    ORB_CUDA(cudaMemcpy(a, da, sizeof(float) * n, cudaMemcpyDeviceToHost));
    ORB_CUDA(cudaMemcpy(b, db, sizeof(float) * n, cudaMemcpyDeviceToHost));
    cudaFree(da); cudaFree(db);
    return ORB_OK;
}

// ======================================================================================================
// host-side launch sequence of one batch (asynchronous on c->stream)
// ======================================================================================================
int orb_launch_extract(orb_ctx* c, const uint8_t* d_imgs, int F, size_t row_stride, size_t frame_stride,
                       orb_kp* d_kps, uint8_t* d_desc, int cap, int* d_n_out) {
    const Geometry& g = c->g;
    cudaStream_t st = c->stream;
    cudaEvent_t* ev = nullptr;
    if (c->profile) {
        const int slot = c->prof_head;
        if (c->prof_pending[slot]) { int rc = orb_profile_harvest(c, slot); if (rc != ORB_OK) return rc; }
        ev = c->prof_ev[slot];
        c->prof_frames[slot] = F;
        c->prof_pending[slot] = true;
        c->prof_head = (slot + 1) % ORB_PROF_RING;
    }
#define ORB_STAGE_MARK(i) do { if (ev) ORB_CUDA(cudaEventRecord(ev[i], st)); } while (0)
    ORB_CUDA(cudaMemsetAsync(c->d_corner_count, 0, sizeof(int) * 2 * (size_t)c->max_batch * g.nlevels, st));
    int* d_tie = c->d_corner_count + (size_t)c->max_batch * g.nlevels;  // second half: tie-at-cut counters
    ORB_STAGE_MARK(0);
    {   // K1
        const LevelGeom& L = g.lv[0];
        dim3 blk(64, 4), grd((L.pitch / 4 + 63) / 64, (L.rows + 3) / 4, F);
        pyr_level0_kernel<<<grd, blk, 0, st>>>(d_imgs, row_stride, frame_stride, c->d_pyr, g);
        c->launches++;
        for (int l = 1; l < g.nlevels; ++l) {
            const LevelGeom& Ll = g.lv[l];
            dim3 grd2((Ll.pitch / 4 + 63) / 64, (Ll.rows + 3) / 4, F);
            pyr_resize_kernel<<<grd2, blk, 0, st>>>(c->d_pyr, c->d_taps, l, g);
            c->launches++;
        }
    }
    ORB_STAGE_MARK(1);
    {   // K2
        const size_t smem = 66 * FAST_TP + 62 * FAST_SP + 3600 * 2 + 900 * 8;
        fast_cells_kernel<<<dim3(g.total_cells, F), FAST_THREADS, smem, st>>>(c->d_pyr, c->d_corners, c->d_corner_count, g);
        c->launches++;
    }
    ORB_STAGE_MARK(2);
    {   // K3
        const size_t smem = (size_t)g.max_node_cap * 80;
        quadtree_kernel<<<dim3(g.nlevels, F), QT_THREADS, smem, st>>>(c->d_corners, c->d_corner_count, c->d_node_of_key,
                                                                       c->d_kept, c->d_kept_count, d_tie, g);
        c->launches++;
    }
    ORB_STAGE_MARK(3);
    {   // K5
        blur_kernel<<<dim3(c->blur_tiles, F), BL_THREADS, 0, st>>>(c->d_pyr, c->d_blur, c->d_blur_tile_base, g);
        c->launches++;
    }
    ORB_STAGE_MARK(4);
    {   // K4 + K6
        orient_describe_kernel<<<dim3((g.total_kp_slots + OD_KPB - 1) / OD_KPB, F), OD_WARPS * 32, 0, st>>>(
            c->d_pyr, c->d_blur, c->d_kept, c->d_kept_count, d_kps, d_desc, cap, d_n_out, g);
        c->launches++;
    }
    ORB_STAGE_MARK(5);
#undef ORB_STAGE_MARK
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

int orb_blur_tile_bases(const Geometry& g, int* bases) {
    int tot = 0;
    for (int l = 0; l < g.nlevels; ++l) {
        bases[l] = tot;
        tot += ((g.lv[l].w + BL_TW - 1) / BL_TW) * ((g.lv[l].h + BL_TH - 1) / BL_TH);
    }
    return tot;
}

int orb_fast_smem_bytes() { return 66 * FAST_TP + 62 * FAST_SP + 3600 * 2 + 900 * 8; }
