// orb_fast.cu — K2: per-cell FAST-9/16 + 3x3 NMS + iniThFAST/minThFAST retry
// (reference orb_slam2/src/ORBextractor.cc:790-863; cv::FAST semantics of OpenCV 4.13.0, DESIGN.md pin (i)).
//
// Semantics kept from the reference (each cell is an independent cv::FAST call on the cell sub-image):
//   * cell (i,j) covers interior pixels [iniX, maxX) x [iniY, maxY), iniX = 16 + j*wCell, maxX = min(iniX+wCell+6,
//     maxBorderX); corners are evaluated on the sub-image minus a 3-px rim, so the evaluated areas of neighbouring
//     cells tile the level without overlap;
//   * score(p) = max(max_arc min_k (c - r_k), max_arc min_k (r_k - c)) - 1 over the 16 arcs of 9 ring pixels; p is a
//     corner at threshold t  <=>  score >= t; NMS keeps p iff score > all 8 neighbours, neighbours outside the
//     cell's evaluated area (or below t) count as 0;
//   * a cell emits its NMS maxima with score >= iniThFAST if it has any, else those with score >= minThFAST.
//
// Kernel shape: one CTA = a strip of up to fast_G consecutive cells of one cell row (tile <= 256 x 66 px in shared
// memory).  Threads work on aligned 32-bit words of 4 pixels:
//   pass 1a direction-aware quick reject, 2 pixels per instruction on s16x2 lanes (VIMNMX.S16x2): a bright (dark)
//           9-arc holds one pixel of each antipodal ring pair, so  min over the 4 axis/diagonal pairs of
//           max(r_k, r_k+8) > c + t  (resp. max of min < c - t) is necessary; survivors are compacted into a
//           shared-memory candidate list
//   pass 1b exact 16-pixel score (VIMNMX3 trees) of the candidates, all lanes busy -> u8 score tile
//   pass 2  cells decide their threshold: any NMS maximum with score >= iniThFAST?
//   pass 3  NMS maxima with score >= the cell's threshold -> shared-memory record list -> one global atomicAdd per
//           CTA -> coalesced copy into the level's corner list (record = corner_pack(x, y, score, order key)).
#include "orb_internal.cuh"

namespace {

#define FS_THREADS 256
#define FS_PITCH 272       // bytes per shared-memory row: 7 (alignment shift) + 256 (tile) + word slack, multiple of 16
#define FS_ROWS 66         // cell sub-image height <= hCell + 6 <= 66
#define FS_SROWS 62        // evaluated rows <= 60, plus a zero row above and below
#define FS_MAXG 8          // 250 / 30
#define FS_CAND 9216       // candidate list entries (u16): 36 rows x 256 px per chunk
#define FS_TILE_BYTES (FS_ROWS * FS_PITCH)
#define FS_SCORE_BYTES (FS_SROWS * FS_PITCH)
#define FS_SMEM (FS_TILE_BYTES + FS_CAND * 2 + FS_SCORE_BYTES)
#define FS_OUT_CAP ((FS_TILE_BYTES + FS_CAND * 2) / 8)   // records that fit the (dead) tile + candidate list: 4548 >= 3750

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

// bits 0/1: pixel 2h / 2h+1 of the word passes the directional quick test.  All operands hold two u8 pixels
// widened to s16x2.
template <int H>
__device__ __forceinline__ unsigned quick2(unsigned c, unsigned r0, unsigned r8, unsigned r4, unsigned r12, unsigned r2,
                                           unsigned r10, unsigned r6, unsigned r14, unsigned tp1, unsigned ntp1) {
    constexpr unsigned SEL = H ? 0x4342u : 0x4140u;
    const unsigned C = __byte_perm(c, 0u, SEL);
    const unsigned a0 = __byte_perm(r0, 0u, SEL), a8 = __byte_perm(r8, 0u, SEL);
    const unsigned a4 = __byte_perm(r4, 0u, SEL), a12 = __byte_perm(r12, 0u, SEL);
    const unsigned a2 = __byte_perm(r2, 0u, SEL), a10 = __byte_perm(r10, 0u, SEL);
    const unsigned a6 = __byte_perm(r6, 0u, SEL), a14 = __byte_perm(r14, 0u, SEL);
    const unsigned minmax = __vmins2(__vimin3_s16x2(__vmaxs2(a0, a8), __vmaxs2(a4, a12), __vmaxs2(a2, a10)), __vmaxs2(a6, a14));
    const unsigned maxmin = __vmaxs2(__vimax3_s16x2(__vmins2(a0, a8), __vmins2(a4, a12), __vmins2(a2, a10)), __vmins2(a6, a14));
    bool bh, bl, dh, dl;
    __vibmax_s16x2(minmax, C + tp1, &bh, &bl);               // minmax >= c + t + 1
    __vibmax_s16x2(__vadd2(C, ntp1), maxmin, &dh, &dl);      // c - t - 1 >= maxmin
    return ((bl | dl) ? 1u : 0u) | ((bh | dh) ? 2u : 0u);
}

__global__ void __launch_bounds__(FS_THREADS)
fast_strip_kernel(const uint8_t* __restrict__ pyr, unsigned long long* __restrict__ corners,
                  int* __restrict__ corner_count, const __grid_constant__ Geometry g) {
    extern __shared__ __align__(16) uint8_t fs_smem[];
    uint8_t* tile = fs_smem;                                                        // FS_ROWS x FS_PITCH pixels
    unsigned short* cand = reinterpret_cast<unsigned short*>(fs_smem + FS_TILE_BYTES);   // FS_CAND entries
    uint8_t* score = fs_smem + FS_TILE_BYTES + FS_CAND * 2;                          // FS_SROWS x FS_PITCH scores
    unsigned long long* outl = reinterpret_cast<unsigned long long*>(fs_smem);      // pass 3: aliases tile + cand
    __shared__ int s_any[FS_MAXG];
    __shared__ int s_ncand, s_nout, s_base;

    const int cta = blockIdx.x, f = blockIdx.y;
    int l = 0;
    while (l + 1 < g.nlevels && cta >= g.lv[l + 1].fast_cta_base) ++l;
    const LevelGeom& L = g.lv[l];
    const int ci = cta - L.fast_cta_base;
    const int i = ci / L.fast_groups, gi = ci - i * L.fast_groups;
    // reference ORBextractor.cc:822-837 (all values are integers held in floats there)
    const int iniY = ORB_MINB + i * L.hCell;
    if (iniY >= L.maxBY - 3) return;
    const int ch = min(iniY + L.hCell + 6, L.maxBY) - iniY;
    if (ch < 7) return;                                    // cv::FAST returns nothing on such a sub-image
    const int j0 = gi * L.fast_G;
    int ncell = 0, X1 = 0;                                 // valid cells of this strip, right end of the last one
    for (int j = j0; j < min(j0 + L.fast_G, L.nCols); ++j) {
        const int iniX = ORB_MINB + j * L.wCell;
        if (iniX >= L.maxBX - 6) break;
        const int maxX = min(iniX + L.wCell + 6, L.maxBX);
        if (maxX - iniX < 7) break;
        ncell = j - j0 + 1; X1 = maxX;
    }
    if (ncell == 0) return;
    const int X0 = ORB_MINB + j0 * L.wCell;
    const int tw = X1 - X0;                                // tile width in pixels (<= 256)
    const int ew = tw - 6, eh = ch - 6;                    // evaluated area

    // ---- load the tile as aligned words: tile pixel (x, y) lands at tile[y * FS_PITCH + a + x] ----
    const uint8_t* src = pyr + L.base + (long long)f * L.frame_stride + L.ioff + iniY * L.pitch + X0;
    // (one extra word on the left, so that word index wd - 1 of the first evaluated word stays inside the row)
    const int a = (int)((uintptr_t)src & 3) + 4;
    const unsigned* src_w = reinterpret_cast<const unsigned*>(src - a);
    const int lw = (a + tw + 3) >> 2;                      // words per row to load (<= 66)
    const unsigned inv_lw = 0xFFFFFFFFu / (unsigned)lw + 1u;
    const int gpw = L.pitch >> 2;
    for (int k = threadIdx.x; k < ch * lw; k += FS_THREADS) {
        const int r = (int)__umulhi((unsigned)k, inv_lw), wd = k - r * lw;
        reinterpret_cast<unsigned*>(tile + r * FS_PITCH)[wd] = __ldg(src_w + r * gpw + wd);
    }
    if (threadIdx.x < FS_MAXG) s_any[threadIdx.x] = 0;
    if (threadIdx.x == 0) { s_ncand = 0; s_nout = 0; }
    for (int k = threadIdx.x; k < (eh + 2) * (FS_PITCH / 16); k += FS_THREADS)
        reinterpret_cast<uint4*>(score)[k] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();

    // ---- pass 1: scores.  Evaluated pixels are the shared-memory bytes [sb_lo, sb_hi) of tile rows [3, 3+eh) ----
    const int tmin = g.min_th, tini = g.ini_th;
    const unsigned tp1 = (unsigned)(tmin + 1) * 0x10001u, ntp1 = ((unsigned)(-(tmin + 1)) & 0xFFFFu) * 0x10001u;
    const int sb_lo = a + 3, sb_hi = a + 3 + ew;
    const int wlo = sb_lo >> 2, whi = (sb_hi - 1) >> 2;
    const int nw = whi - wlo + 1;                          // words holding evaluated pixels (<= 64)
    const unsigned inv_nw = 0xFFFFFFFFu / (unsigned)nw + 1u;
    const unsigned vfirst = 0xFu << (sb_lo & 3), vlast = 0xFu >> (3 - ((sb_hi - 1) & 3));
    const int chunk_rows = min(eh, FS_CAND / (4 * nw));
    const int lane = threadIdx.x & 31;
    for (int rc0 = 0; rc0 < eh; rc0 += chunk_rows) {
        const int rows = min(chunk_rows, eh - rc0);
        const int total = rows * nw;
        // 1a: quick test + compaction (uniform trip count: the ballots need the whole warp)
        for (int k0 = 0; k0 < total; k0 += FS_THREADS) {
            const int k = k0 + threadIdx.x;
            unsigned m = 0;
            int r = 0, wd = 0;
            if (k < total) {
                r = (int)__umulhi((unsigned)k, inv_nw);
                wd = wlo + (k - r * nw);
                r += rc0;
                const unsigned* row = reinterpret_cast<const unsigned*>(tile + (r + 3) * FS_PITCH) + wd;   // centre row
                constexpr int P = FS_PITCH / 4;
                const unsigned c = row[0], wl = row[-1], wr = row[1];
                const unsigned r0 = row[3 * P], r8 = row[-3 * P];                                    // (0,+3) (0,-3)
                const unsigned r4 = __funnelshift_r(c, wr, 24), r12 = __funnelshift_r(wl, c, 8);     // (+3,0) (-3,0)
                const unsigned* rp = row + 2 * P;
                const unsigned* rm = row - 2 * P;
                const unsigned r2 = __funnelshift_r(rp[0], rp[1], 16), r14 = __funnelshift_r(rp[-1], rp[0], 16);  // (+2,+2) (-2,+2)
                const unsigned r6 = __funnelshift_r(rm[0], rm[1], 16), r10 = __funnelshift_r(rm[-1], rm[0], 16);  // (+2,-2) (-2,-2)
                m = quick2<0>(c, r0, r8, r4, r12, r2, r10, r6, r14, tp1, ntp1) |
                    (quick2<1>(c, r0, r8, r4, r12, r2, r10, r6, r14, tp1, ntp1) << 2);
                if (wd == wlo) m &= vfirst;
                if (wd == whi) m &= vlast;
            }
            const unsigned b0 = __ballot_sync(0xffffffffu, m & 1u), b1 = __ballot_sync(0xffffffffu, m & 2u);
            const unsigned b2 = __ballot_sync(0xffffffffu, m & 4u), b3 = __ballot_sync(0xffffffffu, m & 8u);
            const int tot = __popc(b0) + __popc(b1) + __popc(b2) + __popc(b3);
            if (tot) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&s_ncand, tot);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (m) {
                    const unsigned below = (1u << lane) - 1u;
                    int o = base + __popc(b0 & below) + __popc(b1 & below) + __popc(b2 & below) + __popc(b3 & below);
                    const int code = r * FS_PITCH + (wd << 2);
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if (m & (1u << b)) cand[o++] = (unsigned short)(code + b);
                }
            }
        }
        __syncthreads();
        // 1b: exact score of the candidates
        const int ncand = s_ncand;
        for (int k = threadIdx.x; k < ncand; k += FS_THREADS) {
            const int code = cand[k];
            const int r = (int)__umulhi((unsigned)code, 0xFFFFFFFFu / FS_PITCH + 1u);
            const int sb = code - r * FS_PITCH;
            const uint8_t* t = tile + (r + 3) * FS_PITCH + sb;
            int rr[16];
            rr[0] = t[3 * FS_PITCH];       rr[1] = t[3 * FS_PITCH + 1];   rr[2] = t[2 * FS_PITCH + 2];   rr[3] = t[FS_PITCH + 3];
            rr[4] = t[3];                  rr[5] = t[-FS_PITCH + 3];      rr[6] = t[-2 * FS_PITCH + 2];  rr[7] = t[-3 * FS_PITCH + 1];
            rr[8] = t[-3 * FS_PITCH];      rr[9] = t[-3 * FS_PITCH - 1];  rr[10] = t[-2 * FS_PITCH - 2]; rr[11] = t[-FS_PITCH - 3];
            rr[12] = t[-3];                rr[13] = t[FS_PITCH - 3];      rr[14] = t[2 * FS_PITCH - 2];  rr[15] = t[3 * FS_PITCH - 1];
            const int sc = fast_score16(t[0], rr);
            if (sc >= tmin) score[(r + 1) * FS_PITCH + sb] = (uint8_t)sc;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_ncand = 0;
        // (the next chunk's first atomicAdd on s_ncand comes after its own quick-test work; the barrier below orders it)
        __syncthreads();
    }

    // ---- pass 2 / 3 scan the score tile in 16-byte chunks and share the NMS test ----
    const int wCell = L.wCell;
    const unsigned inv_wc = 0xFFFFFFFFu / (unsigned)wCell + 1u;
    auto nms_max = [&](int r, int sb, int sc, int& jj, int& xr) -> bool {
        const int xe = sb - sb_lo;                         // column inside the strip's evaluated area
        jj = (int)__umulhi((unsigned)xe, inv_wc);
        xr = xe - jj * wCell;
        const bool left = (xr == 0), right = (xr == wCell - 1) || (xe == ew - 1);   // cell edges: neighbours beyond are 0
        const uint8_t* q = score + (r + 1) * FS_PITCH + sb;
        bool ok = sc > q[-FS_PITCH] && sc > q[FS_PITCH];
        if (!left) ok = ok && sc > q[-1] && sc > q[-FS_PITCH - 1] && sc > q[FS_PITCH - 1];
        if (!right) ok = ok && sc > q[1] && sc > q[-FS_PITCH + 1] && sc > q[FS_PITCH + 1];
        return ok;
    };
    const int qlo = sb_lo >> 4, nq = ((sb_hi - 1) >> 4) - qlo + 1;      // 16-byte chunks per row (<= 17)
    const unsigned inv_nq = 0xFFFFFFFFu / (unsigned)nq + 1u;
    for (int k = threadIdx.x; k < eh * nq; k += FS_THREADS) {
        const int r = (int)__umulhi((unsigned)k, inv_nq);
        const int qd = qlo + (k - r * nq);
        const uint4 v = reinterpret_cast<const uint4*>(score + (r + 1) * FS_PITCH)[qd];
        if ((v.x | v.y | v.z | v.w) == 0u) continue;
        const unsigned wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int w4 = 0; w4 < 4; ++w4) {
            if (wv[w4] == 0u) continue;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int sc = (wv[w4] >> (8 * b)) & 0xFF;
                if (sc >= tini) {
                    int jj, xr;
                    if (nms_max(r, (qd << 4) + (w4 << 2) + b, sc, jj, xr)) s_any[jj] = 1;
                }
            }
        }
    }
    __syncthreads();   // also: every read of tile / cand is done, outl may overwrite them
    for (int k = threadIdx.x; k < eh * nq; k += FS_THREADS) {
        const int r = (int)__umulhi((unsigned)k, inv_nq);
        const int qd = qlo + (k - r * nq);
        const uint4 v = reinterpret_cast<const uint4*>(score + (r + 1) * FS_PITCH)[qd];
        if ((v.x | v.y | v.z | v.w) == 0u) continue;
        const unsigned wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int w4 = 0; w4 < 4; ++w4) {
            if (wv[w4] == 0u) continue;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int sc = (wv[w4] >> (8 * b)) & 0xFF;
                if (sc == 0) continue;
                int jj, xr;
                const int sb = (qd << 4) + (w4 << 2) + b;
                if (!nms_max(r, sb, sc, jj, xr)) continue;
                if (sc < (s_any[jj] ? tini : tmin)) continue;
                const int xe = sb - sb_lo;
                // vToDistributeKeys coordinates (ORBextractor.cc:856-857) and the reference's visiting order key
                const int cell = i * L.nCols + j0 + jj;
                const int o = atomicAdd(&s_nout, 1);
                if (o < FS_OUT_CAP)
                    outl[o] = corner_pack(j0 * wCell + 3 + xe, i * L.hCell + 3 + r, sc, (cell << 12) | (r << 6) | xr);
            }
        }
    }
    __syncthreads();
    const int nout = min(s_nout, FS_OUT_CAP);
    if (nout == 0) return;
    if (threadIdx.x == 0) s_base = atomicAdd(corner_count + f * g.nlevels + l, nout);
    __syncthreads();
    const int base = s_base;
    unsigned long long* dst = corners + L.corner_base + (long long)f * L.corner_cap;
    for (int k = threadIdx.x; k < nout; k += FS_THREADS)
        if (base + k < L.corner_cap) dst[base + k] = outl[k];
}

}  // namespace

int orb_launch_fast(orb_ctx* c, int F) {
    const Geometry& g = c->g;
    if (!c->fast_attr_set) {   // > 48 KB of dynamic shared memory needs the opt-in, once per context (= per device)
        ORB_CUDA(cudaFuncSetAttribute(fast_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FS_SMEM));
        c->fast_attr_set = true;
    }
    fast_strip_kernel<<<dim3(g.fast_ctas, F), FS_THREADS, FS_SMEM, c->stream>>>(c->d_pyr, c->d_corners, c->d_corner_count, g);
    c->launches++;
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}
