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
//   pass 1a direction-aware quick reject, 2 pixels per instruction on u16x2 lanes whose high bytes hold the pixels
//           (VIMNMX.U16x2, word pairs = 8 pixels per item): a bright (dark) 9-arc holds one pixel of each antipodal ring
//           pair, so  min over the 4 axis/diagonal pairs of max(r_k, r_k+8) > c + t  (resp. max of min < c - t) is
//           necessary; both are tested at once as  max(minmax - c, c - maxmin) >= t << 8  on the clamped margins.
//           Every lane keeps a 64-bit mask of its candidates; each WARP compacts its masks once into a private
//           shared-memory segment.  Stage A = the whole strip at iniThFAST; stage B = the cells that have no NMS maximum
//           >= iniThFAST, at minThFAST — either a second walk over those cells (template DUAL = false) or from a second
//           mask stage A built in the same walk (DUAL = true); orb_launch_fast picks the variant from the share of empty
//           cells the kernel reports
//   pass 1b exact 16-pixel score of the warp's candidates, all lanes busy; both polarities in one s16x2 tree on
//           (r, 255 - r) pairs -> u8 score tile + CTA-wide list of scored corners (score >= minThFAST)
//   pass 2  NMS over the scored list (dense); cells decide their threshold: any NMS maximum >= iniThFAST?
//   pass 3  NMS maxima with score >= the cell's threshold -> shared-memory record list -> one global atomicAdd per
//           CTA -> coalesced copy into the level's corner list (record = corner_pack(x, y, score, order key)).
// Capacity overflows (pathological images only) fall back to slower but identical-result paths.
#include "orb_internal.cuh"
#include <type_traits>

namespace {

#define FS_THREADS 256
#ifndef FS_QUICK
#define FS_QUICK 2         // stage A's quick test: 2 = word pairs on u16 high-byte lanes, 1 = single words on widened s16 lanes
#endif
#define FS_WARPS (FS_THREADS / 32)
#define FS_PITCH 256       // bytes per shared-memory row = one TMA box row: alignment shift (<= 19) + tile (<= 237)
#define FS_ROWS 66         // cell sub-image height <= hCell + 6 <= 66
#define FS_SROWS 62        // evaluated rows <= 60, plus a zero row above and below
#define FS_MAXG 8          // 250 / 30
#define FS_WCAP 448        // candidate entries (u16) per warp segment
#define FS_SCAP 2048       // scored-corner list entries (u16)
#define FS_CAND_BYTES (FS_WARPS * FS_WCAP * 2)
#define FS_SCORED_BYTES (FS_SCAP * 2)
// The tile and the score tile are sized by the tallest cell sub-image of the geometry (Geometry::fast_rows: 40 rows at
// 640x480 instead of the worst case 66), which is what decides how many CTAs fit an SM: 30 KB -> 6-7 CTAs, 44 KB -> 5.
#define FS_TILE_BYTES(rows) ((rows) * FS_PITCH)
#define FS_SCORE_BYTES(rows) (((rows) - 4) * FS_PITCH)      // evaluated rows (rows - 6) + a zero row above and below
#define FS_CODE_BYTES (8 * FS_THREADS * 2)                   // stage A: tile code of every (item, thread), for the candidate decode
#define FS_SMEM(rows) (FS_TILE_BYTES(rows) + FS_CAND_BYTES + FS_SCORED_BYTES + FS_SCORE_BYTES(rows) + FS_CODE_BYTES + 256)   // 128 for the alignment + 128 readable bytes in front of the tile
#define FS_OUT_CAP(rows) ((FS_TILE_BYTES(rows) + FS_CAND_BYTES) / 8)   // records that fit the (dead) tile + candidate segments; anything beyond goes straight to global

// k / d by the host-built magic number inv = floor((2^32 - 1) / d) + 1 (exact for the small k used here); d == 1 has no 32-bit magic
// number (it would be 2^32), and it does occur: a strip whose last cell is 7 px wide evaluates ONE column there
__device__ __forceinline__ int fs_div(int k, int d, unsigned inv) { return d == 1 ? k : (int)__umulhi((unsigned)k, inv); }

// exact FAST score of the pixel at t (shared-memory tile): both polarities in one s16x2 min/max tree on the packed
// pairs (r_k, 255 - r_k):  lo -> max_arc min r = bmax,  hi -> max_arc min (255 - r) = 255 - min_arc max r = 255 - amin
__device__ __forceinline__ int fast_score_at(const uint8_t* t) {
    unsigned p[16];
    const int o[16] = {3 * FS_PITCH,      3 * FS_PITCH + 1,  2 * FS_PITCH + 2,  FS_PITCH + 3,  3,  -FS_PITCH + 3, -2 * FS_PITCH + 2,
                       -3 * FS_PITCH + 1, -3 * FS_PITCH,     -3 * FS_PITCH - 1, -2 * FS_PITCH - 2, -FS_PITCH - 3, -3, FS_PITCH - 3,
                       2 * FS_PITCH - 2,  3 * FS_PITCH - 1};
#pragma unroll
    for (int k = 0; k < 16; ++k) p[k] = (unsigned)t[o[k]] * 0xFFFF0001u + 0x00FF0000u;   // lo = r, hi = 255 - r
    unsigned m3[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) m3[k] = __vimin3_s16x2(p[k], p[(k + 1) & 15], p[(k + 2) & 15]);
    unsigned best = 0u;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        const unsigned a = __vimin3_s16x2(m3[k], m3[(k + 3) & 15], m3[(k + 6) & 15]);
        const unsigned b = __vimin3_s16x2(m3[k + 1], m3[(k + 4) & 15], m3[(k + 7) & 15]);
        best = __vimax3_s16x2(best, a, b);
    }
    const int c = t[0];
    const int bmax = (int)(best & 0xFFFFu), amin = 255 - (int)(best >> 16);
    return max(c - amin, bmax - c) - 1;
}

// bits 0/1: pixel 2h / 2h+1 of the word passes the directional quick test.  All operands hold two u8 pixels
// widened to s16x2.
// ---- TMA (cp.async.bulk.tensor) + mbarrier plumbing ----------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int x, int y, int z, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}

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

// The same test on UNSIGNED 16-bit lanes whose HIGH bytes hold the pixels and whose low bytes are don't-care: u16 order is
// lexicographic in (pixel, low byte), so the high byte of every min / max is the min / max of the pixels and a raw tile
// word serves its odd pixels as it is (its even pixels after << 8) — no byte-to-halfword widening (18 PRMT per word in
// quick2).  With K = t << 8 the comparisons  minmax >= C + K  and  C - K >= maxmin  can only err towards "pass" (the low
// bytes act below the pixel difference t + 1 the true condition demands): false candidates are scored exactly like any
// other, a true corner is never lost.  The centre is clamped so that C + K / C - K cannot wrap in a lane (saturated
// regions would otherwise pass wholesale).  BLO / BHI: the result bits of the low / high lane.
template <unsigned BLO, unsigned BHI>
__device__ __forceinline__ unsigned quick_u16(unsigned C, unsigned a0, unsigned a8, unsigned a4, unsigned a12, unsigned a2, unsigned a10,
                                              unsigned a6, unsigned a14, unsigned K, unsigned NK, unsigned LIMB) {
    const unsigned minmax = __vminu2(__vimin3_u16x2(__vmaxu2(a0, a8), __vmaxu2(a4, a12), __vmaxu2(a2, a10)), __vmaxu2(a6, a14));
    const unsigned maxmin = __vmaxu2(__vimax3_u16x2(__vminu2(a0, a8), __vminu2(a4, a12), __vminu2(a2, a10)), __vminu2(a6, a14));
    bool bh, bl, dh, dl;
    __vibmax_u16x2(minmax, __vadd2(__vminu2(C, LIMB), K), &bh, &bl);    // minmax >= min(C, 0xFFFF - K) + K
    __vibmax_u16x2(__vadd2(__vmaxu2(C, K), NK), maxmin, &dh, &dl);      // max(C, K) - K >= maxmin
    return ((bl | dl) ? BLO : 0u) | ((bh | dh) ? BHI : 0u);
}
// 4 candidate bits of one tile word: odd pixels from the raw words, even pixels from the words << 8
__device__ __forceinline__ unsigned quick_word(unsigned c, unsigned r0, unsigned r8, unsigned r4, unsigned r12, unsigned r2, unsigned r10,
                                               unsigned r6, unsigned r14, unsigned K, unsigned NK, unsigned LIMB) {
    return quick_u16<2u, 8u>(c, r0, r8, r4, r12, r2, r10, r6, r14, K, NK, LIMB) |
           quick_u16<1u, 4u>(c << 8, r0 << 8, r8 << 8, r4 << 8, r12 << 8, r2 << 8, r10 << 8, r6 << 8, r14 << 8, K, NK, LIMB);
}

// Both thresholds in one evaluation (FS_DUAL): the margins  tb = minmax - C  (bright) and  td = C - maxmin  (dark), clamped at 0,
// are compared with K = t << 8 for iniThFAST and for minThFAST.  tq >= K is the same test as quick_u16's (and drops its two
// wrap-around false positives); the minThFAST bits are only used later, for the cells that turn out empty at iniThFAST, so
// those cells need no second walk over the tile.
template <unsigned BLO, unsigned BHI>
__device__ __forceinline__ unsigned quick_u16_dual(unsigned C, unsigned a0, unsigned a8, unsigned a4, unsigned a12, unsigned a2, unsigned a10,
                                                   unsigned a6, unsigned a14, unsigned KI, unsigned KM, unsigned& mlow) {
    const unsigned minmax = __vminu2(__vimin3_u16x2(__vmaxu2(a0, a8), __vmaxu2(a4, a12), __vmaxu2(a2, a10)), __vmaxu2(a6, a14));
    const unsigned maxmin = __vmaxu2(__vimax3_u16x2(__vminu2(a0, a8), __vminu2(a4, a12), __vminu2(a2, a10)), __vminu2(a6, a14));
    const unsigned tq = __vmaxu2(__vsub2(__vmaxu2(minmax, C), C), __vsub2(C, __vminu2(C, maxmin)));
    bool ih, il, mh, ml;
    __vibmax_u16x2(tq, KI, &ih, &il);                        // tq >= KI
    __vibmax_u16x2(tq, KM, &mh, &ml);                        // tq >= KM
    mlow |= (ml ? BLO : 0u) | (mh ? BHI : 0u);
    return (il ? BLO : 0u) | (ih ? BHI : 0u);
}
// one threshold, same margin form: one compare (two predicates) per lane pair instead of a bright and a dark one
template <unsigned BLO, unsigned BHI>
__device__ __forceinline__ unsigned quick_u16_margin(unsigned C, unsigned a0, unsigned a8, unsigned a4, unsigned a12, unsigned a2, unsigned a10,
                                                     unsigned a6, unsigned a14, unsigned K) {
    const unsigned minmax = __vminu2(__vimin3_u16x2(__vmaxu2(a0, a8), __vmaxu2(a4, a12), __vmaxu2(a2, a10)), __vmaxu2(a6, a14));
    const unsigned maxmin = __vmaxu2(__vimax3_u16x2(__vminu2(a0, a8), __vminu2(a4, a12), __vminu2(a2, a10)), __vminu2(a6, a14));
    const unsigned tq = __vmaxu2(__vsub2(__vmaxu2(minmax, C), C), __vsub2(C, __vminu2(C, maxmin)));
    bool h, l;
    __vibmax_u16x2(tq, K, &h, &l);                           // tq >= K
    return (l ? BLO : 0u) | (h ? BHI : 0u);
}
__device__ __forceinline__ unsigned quick_word_margin(unsigned c, unsigned r0, unsigned r8, unsigned r4, unsigned r12, unsigned r2, unsigned r10,
                                                      unsigned r6, unsigned r14, unsigned K) {
    return quick_u16_margin<2u, 8u>(c, r0, r8, r4, r12, r2, r10, r6, r14, K) |
           quick_u16_margin<1u, 4u>(c << 8, r0 << 8, r8 << 8, r4 << 8, r12 << 8, r2 << 8, r10 << 8, r6 << 8, r14 << 8, K);
}
__device__ __forceinline__ unsigned quick_word_dual(unsigned c, unsigned r0, unsigned r8, unsigned r4, unsigned r12, unsigned r2, unsigned r10,
                                                    unsigned r6, unsigned r14, unsigned KI, unsigned KM, unsigned& mlow) {
    return quick_u16_dual<2u, 8u>(c, r0, r8, r4, r12, r2, r10, r6, r14, KI, KM, mlow) |
           quick_u16_dual<1u, 4u>(c << 8, r0 << 8, r8 << 8, r4 << 8, r12 << 8, r2 << 8, r10 << 8, r6 << 8, r14 << 8, KI, KM, mlow);
}

#ifndef FS_MARGIN
#define FS_MARGIN 1         // single-threshold quick test in the margin form (one compare per lane pair) instead of a bright and a dark compare
#endif
#ifndef FS_STAGEB_WORDS
#define FS_STAGEB_WORDS 0   // 1: the single-threshold variant re-walks the empty cells word by word (s16 lanes) instead of by word pairs
#endif
template <bool TMA, bool DUAL>
#ifndef FS_MINB
#define FS_MINB 6
#endif
__global__ void __launch_bounds__(FS_THREADS, FS_MINB)
fast_strip_kernel(const uint8_t* __restrict__ pyr, const FastStrip* __restrict__ strips, unsigned long long* __restrict__ corners,
                  int* __restrict__ corner_count, const __grid_constant__ Geometry g, const FastTmaps* __restrict__ tm, int f0,
                  int* __restrict__ stats) {
    extern __shared__ __align__(1024) uint8_t fs_smem_raw[];
    // the TMA destination must be 128-byte aligned: align by hand (static shared variables precede the dynamic window)
    uint8_t* fs_smem = fs_smem_raw + ((128u - (smem_u32(fs_smem_raw) & 127u)) & 127u) + 128;   // (the pair walk of stage A may read the word left of the tile's first one)
    uint8_t* tile = fs_smem;                                                                     // FS_ROWS x FS_PITCH pixels
    const int tile_bytes = FS_TILE_BYTES(g.fast_rows), out_cap = FS_OUT_CAP(g.fast_rows);
    unsigned short* cand = reinterpret_cast<unsigned short*>(fs_smem + tile_bytes);               // per-warp segments
    unsigned short* scored = reinterpret_cast<unsigned short*>(fs_smem + tile_bytes + FS_CAND_BYTES);
    uint8_t* score = fs_smem + tile_bytes + FS_CAND_BYTES + FS_SCORED_BYTES;                      // (fast_rows - 4) x FS_PITCH scores
    unsigned short* codes = reinterpret_cast<unsigned short*>(score + FS_SCORE_BYTES(g.fast_rows));   // [8][FS_THREADS]
    unsigned long long* outl = reinterpret_cast<unsigned long long*>(fs_smem);                   // pass 3: aliases tile + cand
    __shared__ int s_any[FS_MAXG + 1];                   // (+1: the pair walk may look one cell beyond the strip)
    __shared__ int s_nscored, s_nout, s_base, s_nempty;
    __shared__ int s_bpre[FS_MAXG + 1], s_bw0[FS_MAXG], s_bnw[FS_MAXG];   // stage B: per empty cell word-item prefix, first word, words per row
    __shared__ unsigned s_binv[FS_MAXG], s_bmf[FS_MAXG], s_bml[FS_MAXG];
    __shared__ __align__(8) unsigned long long s_mbar;
    __shared__ uint8_t s_pvm[40];                          // stage A: valid-pixel mask of the word pairs of a row

    const int f = blockIdx.y;
    const FastStrip S = strips[blockIdx.x];               // host-precomputed (uniform loads)
    // DUAL: stage A evaluates the quick test at iniThFAST and at minThFAST together (+5 % of the kernel) so that cells found empty
    // at iniThFAST need no second walk over the tile (up to +30 %).  Which one is faster depends on the image content; both give the
    // same corners.  The host picks the variant per launch from the share of empty cells the previous launches reported in `stats`
    // (orb_launch_fast).
    const int l = S.level;
    const LevelGeom& L = g.lv[l];
    const int i = S.i, j0 = S.j0;
    const int ch = S.ch, tw = S.tw;
    const int ew = tw - 6, eh = ch - 6;                    // evaluated area

    // ---- load the tile: tile pixel (x, y) lands at tile[y * FS_PITCH + a + x] ----
    // (a in [4, 19]: the tile starts at a 16-byte aligned global address — the TMA unit faults otherwise — with at
    // least one spare word on the left, so that word index wd - 1 of the first evaluated word exists)
    const int a = S.a;
    if (TMA) {
        // one elected thread arms the mbarrier and issues ONE bulk tensor copy (box 256 B x fast_rows rows) of the strip with
        // its halo; rows / bytes beyond the level are zero-filled by the TMA unit and never evaluated
        if (threadIdx.x == 0) mbar_init(&s_mbar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(&s_mbar, (unsigned)(ORB_TMA_BOX_W * g.fast_rows));
            tma_load_3d(tile, &tm->m[l], ORB_XOFF + S.X0 - a, ORB_EDGE + S.iniY, f0 + f, &s_mbar);
        }
    } else {
        const unsigned* src_w = reinterpret_cast<const unsigned*>(pyr + L.base + (long long)f * L.frame_stride + L.ioff +
                                                                  S.iniY * L.pitch + S.X0 - a);
        const int lw = S.lw;                               // words per row to load (<= 64)
        const unsigned inv_lw = S.inv_lw;
        const int gpw = L.pitch >> 2;
        for (int k = threadIdx.x; k < ch * lw; k += FS_THREADS) {
            const int r = (int)__umulhi((unsigned)k, inv_lw), wd = k - r * lw;
            reinterpret_cast<unsigned*>(tile + r * FS_PITCH)[wd] = __ldg(src_w + r * gpw + wd);
        }
    }
    if (threadIdx.x <= FS_MAXG) s_any[threadIdx.x] = 0;
    if (threadIdx.x >= 64 && threadIdx.x < 64 + S.np) {   // valid pixels of pair q: smem bytes [sb_lo, sb_hi) inside [4 * (w0p + 2q), + 8)
        const int b0 = 4 * (S.w0p + 2 * ((int)threadIdx.x - 64));
        const int lo = max(S.a + 3 - b0, 0), hi = min(S.a + 3 + (S.tw - 6) - b0, 8);
        s_pvm[threadIdx.x - 64] = (uint8_t)(hi > lo ? ((1u << hi) - 1u) & ~((1u << lo) - 1u) : 0u);
    }
    if (threadIdx.x == 0) { s_nscored = 0; s_nout = 0; }
    for (int k = threadIdx.x; k < (eh + 2) * (FS_PITCH / 16); k += FS_THREADS)
        reinterpret_cast<uint4*>(score)[k] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    if (TMA) mbar_wait(&s_mbar, 0);   // the tile has landed (async-proxy writes are visible after the wait)

    // ---- pass 1: scores.  Evaluated pixels are the shared-memory bytes [sb_lo, sb_hi) of tile rows [3, 3+eh) ----
    const int tmin = g.min_th, tini = g.ini_th;
    const int sb_lo = a + 3, sb_hi = a + 3 + ew;
    const int wlo = S.wlo, nw = S.nw, whi = wlo + nw - 1;  // words holding evaluated pixels (<= 64)
    const unsigned inv_nw = S.inv_nw;
    const unsigned vfirst = (0xFu << (sb_lo & 3)) & 0xFu, vlast = 0xFu >> (3 - ((sb_hi - 1) & 3));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // code = r * FS_PITCH + sb addresses the tile (3 rows down: the halo), the score tile (one zero row down) and the list alike
    // SKIP (stage B): a pixel that already has a score (stage A scored it) is in the list already
    auto score_one = [&](int code, const bool skip) {
        const int sc = fast_score_at(tile + 3 * FS_PITCH + code);
        if (sc >= tmin) {
            if (skip && score[FS_PITCH + code] != 0) return;
            score[FS_PITCH + code] = (uint8_t)sc;
            const int o = atomicAdd(&s_nscored, 1);
            if (o < FS_SCAP) scored[o] = (unsigned short)code;
        }
    };
    // warp-level compaction of the per-lane candidate masks + exact scores.  decode(bit) -> tile code (row * FS_PITCH + byte) of
    // candidate bit `bit` of this lane's mask
    auto compact_score = [&](unsigned long long cmask, auto&& decode, const bool skip) {
        const int cnt = __popcll(cmask);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        const int wtotal = __shfl_sync(0xffffffffu, incl, 31);
        unsigned short* seg = cand + warp * FS_WCAP;
        if (wtotal <= FS_WCAP) {
            // every lane appends its own candidates (set bits of its mask) at its prefix offset; at iniThFAST a lane
            // holds only a few, so walking the set bits is cheaper than a uniform walk over all words
            int o = incl - cnt;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                unsigned mm = half ? (unsigned)(cmask >> 32) : (unsigned)cmask;
                while (mm) {
                    const int bit = __ffs((int)mm) - 1;
                    mm &= mm - 1u;
                    seg[o++] = (unsigned short)decode(bit + 32 * half);
                }
            }
            __syncwarp();
            // 1b: exact score of the warp's candidates
            for (int k = lane; k < wtotal; k += 32) score_one(seg[k], skip);
            __syncwarp();
        } else {
            // segment overflow (a warp with > 448 candidates): score in place, lane by lane
            while (cmask) {
                const int bit = __ffsll((long long)cmask) - 1;
                cmask &= cmask - 1ull;
                score_one(decode(bit), skip);
            }
        }
    };
    // One candidate pass over `total` word items (<= 60 * 64: at most 15 iterations of 256 threads); map(k) gives the
    // item's row, word and valid-pixel mask.  Quick test at threshold th -> per-lane 64-bit candidate mask -> warp-local
    // compaction -> exact scores.  SKIP: pixels that already have a score (an earlier pass) are not candidates.
    auto cand_pass = [&](const int total, auto&& map, const int th, const bool skip) {
        const unsigned tp1 = (unsigned)(th + 1) * 0x10001u, ntp1 = ((unsigned)(-(th + 1)) & 0xFFFFu) * 0x10001u;
        // 1a: quick test, 4 candidate bits per iteration per lane
        unsigned long long cmask = 0ull;
        int it = 0;
        for (int k = threadIdx.x; k < total; k += FS_THREADS, ++it) {
            int r, wd;
            unsigned vm;
            map(k, r, wd, vm);
            const unsigned* row = reinterpret_cast<const unsigned*>(tile + (r + 3) * FS_PITCH) + wd;   // centre row
            constexpr int P = FS_PITCH / 4;
            const unsigned c = row[0], wl = row[-1], wr = row[1];
            const unsigned r0 = row[3 * P], r8 = row[-3 * P];                                    // (0,+3) (0,-3)
            const unsigned r4 = __funnelshift_r(c, wr, 24), r12 = __funnelshift_r(wl, c, 8);     // (+3,0) (-3,0)
            const unsigned* rp = row + 2 * P;
            const unsigned* rm = row - 2 * P;
            const unsigned r2 = __funnelshift_r(rp[0], rp[1], 16), r14 = __funnelshift_r(rp[-1], rp[0], 16);  // (+2,+2) (-2,+2)
            const unsigned r6 = __funnelshift_r(rm[0], rm[1], 16), r10 = __funnelshift_r(rm[-1], rm[0], 16);  // (+2,-2) (-2,-2)
            unsigned m = quick2<0>(c, r0, r8, r4, r12, r2, r10, r6, r14, tp1, ntp1) |
                         (quick2<1>(c, r0, r8, r4, r12, r2, r10, r6, r14, tp1, ntp1) << 2);
            m &= vm;
            if (skip) {
                const unsigned sw = reinterpret_cast<const unsigned*>(score + (r + 1) * FS_PITCH)[wd];
                const unsigned nz = __vcmpne4(sw, 0u);     // 0xFF per pixel that already has a score
                m &= ~((nz & 1u) | ((nz >> 7) & 2u) | ((nz >> 14) & 4u) | ((nz >> 21) & 8u));
            }
            cmask |= (unsigned long long)m << (4 * it);
        }
        compact_score(cmask, [&](int bit) {
            int r, wd;
            unsigned vm;
            map(threadIdx.x + (bit >> 2) * FS_THREADS, r, wd, vm);
            return r * FS_PITCH + (wd << 2) + (bit & 3);
        }, false);
    };
    // Stage A's candidate pass over WORD PAIRS (8 pixels of one row per item, <= 8 items per thread): the 18 tile words an
    // item needs are 11 loads (4 of them LDS.64) and 10 funnel shifts instead of 2 x (11 + 6), and the u16 high-byte form of
    // the quick test (quick_u16) needs no widening.  Pair q of a row = tile words w0p + 2q, w0p + 2q + 1 (w0p even).
    unsigned long long cmask_low = 0ull;                   // FS_DUAL: this lane's stage-A pixels that pass the quick test at minThFAST but not at iniThFAST
    (void)cmask_low;
    // map(k) gives item k's row, pair and valid-pixel mask; `total` <= eh * pairs per row <= 8 * FS_THREADS
    auto pair_pass = [&](const int th, const int total, auto&& map, const bool skip, auto both_tag) {
        constexpr bool both = decltype(both_tag)::value;
        const unsigned K = ((unsigned)th << 8) * 0x10001u, NK = ((0x10000u - ((unsigned)th << 8)) & 0xFFFFu) * 0x10001u, LIMB = 0xFFFFFFFFu - K;
        const unsigned KM = ((unsigned)tmin << 8) * 0x10001u;
        (void)NK; (void)LIMB; (void)KM;
        const int w0p = S.w0p;
        unsigned long long cmask = 0ull;
        int it = 0;
        for (int k = threadIdx.x; k < total; k += FS_THREADS, ++it) {
            int r, q;
            unsigned vm;
            map(k, r, q, vm);
            constexpr int P = FS_PITCH / 4;
            const int code0 = r * FS_PITCH + ((w0p + 2 * q) << 2);          // tile code of the item's first pixel, kept for the candidate decode
            codes[it * FS_THREADS + threadIdx.x] = (unsigned short)code0;
            const unsigned* row = reinterpret_cast<const unsigned*>(tile + 3 * FS_PITCH + code0);   // centre row, 8-byte aligned
            const uint2 c = *reinterpret_cast<const uint2*>(row);
            const unsigned wl = row[-1], wr = row[2];
            const uint2 u = *reinterpret_cast<const uint2*>(row + 3 * P), d = *reinterpret_cast<const uint2*>(row - 3 * P);   // (0,+3) (0,-3)
            const unsigned* rp = row + 2 * P;
            const unsigned* rm = row - 2 * P;
            const uint2 pp = *reinterpret_cast<const uint2*>(rp), mm = *reinterpret_cast<const uint2*>(rm);
            const unsigned pl = rp[-1], pr = rp[2], ml = rm[-1], mr = rm[2];
            const unsigned pmid = __funnelshift_r(pp.x, pp.y, 16), mmid = __funnelshift_r(mm.x, mm.y, 16);   // (+2,.) of word 0 = (-2,.) of word 1
            unsigned m;
            if constexpr (both) {
            unsigned l0 = 0u, l1 = 0u;
            const unsigned m0 = quick_word_dual(c.x, u.x, d.x, __funnelshift_r(c.x, c.y, 24), __funnelshift_r(wl, c.x, 8),      // (+3,0) (-3,0)
                                                pmid, __funnelshift_r(ml, mm.x, 16),                                           // (+2,+2) (-2,-2)
                                                mmid, __funnelshift_r(pl, pp.x, 16), K, KM, l0);                               // (+2,-2) (-2,+2)
            const unsigned m1 = quick_word_dual(c.y, u.y, d.y, __funnelshift_r(c.y, wr, 24), __funnelshift_r(c.x, c.y, 8),
                                                __funnelshift_r(pp.y, pr, 16), mmid,
                                                __funnelshift_r(mm.y, mr, 16), pmid, K, KM, l1);
            m = (m0 | (m1 << 4)) & vm;
            cmask_low |= (unsigned long long)((l0 | (l1 << 4)) & vm & ~m) << (8 * it);   // passes at minThFAST only
            } else {
#if FS_MARGIN
            const unsigned m0 = quick_word_margin(c.x, u.x, d.x, __funnelshift_r(c.x, c.y, 24), __funnelshift_r(wl, c.x, 8),    // (+3,0) (-3,0)
                                                  pmid, __funnelshift_r(ml, mm.x, 16),                                         // (+2,+2) (-2,-2)
                                                  mmid, __funnelshift_r(pl, pp.x, 16), K);                                     // (+2,-2) (-2,+2)
            const unsigned m1 = quick_word_margin(c.y, u.y, d.y, __funnelshift_r(c.y, wr, 24), __funnelshift_r(c.x, c.y, 8),
                                                  __funnelshift_r(pp.y, pr, 16), mmid,
                                                  __funnelshift_r(mm.y, mr, 16), pmid, K);
#else
            const unsigned m0 = quick_word(c.x, u.x, d.x, __funnelshift_r(c.x, c.y, 24), __funnelshift_r(wl, c.x, 8),           // (+3,0) (-3,0)
                                           pmid, __funnelshift_r(ml, mm.x, 16),                                                // (+2,+2) (-2,-2)
                                           mmid, __funnelshift_r(pl, pp.x, 16), K, NK, LIMB);                                  // (+2,-2) (-2,+2)
            const unsigned m1 = quick_word(c.y, u.y, d.y, __funnelshift_r(c.y, wr, 24), __funnelshift_r(c.x, c.y, 8),
                                           __funnelshift_r(pp.y, pr, 16), mmid,
                                           __funnelshift_r(mm.y, mr, 16), pmid, K, NK, LIMB);
#endif
            m = (m0 | (m1 << 4)) & vm;
            }
            cmask |= (unsigned long long)m << (8 * it);
        }
        compact_score(cmask, [&](int bit) { return (int)codes[(bit >> 3) * FS_THREADS + threadIdx.x] + (bit & 7); }, skip);
    };
    // stage A: the whole strip at iniThFAST (a cell that has an NMS maximum >= iniThFAST never needs anything lower)
#if FS_QUICK == 2
    {
        const int np = S.np;
        const unsigned inv_np = S.inv_np;
        auto rows = [&](int k, int& r, int& q, unsigned& vm) {
            r = (int)__umulhi((unsigned)k, inv_np);          // np >= 2 (host)
            q = k - r * np;
            vm = s_pvm[q];
        };
        pair_pass(tini, eh * np, rows, false, std::integral_constant<bool, DUAL>{});
    }
#else
    cand_pass(eh * nw,
              [&](int k, int& r, int& wd, unsigned& vm) {
                  r = fs_div(k, nw, inv_nw);
                  wd = wlo + (k - r * nw);
                  vm = (wd == wlo ? vfirst : 0xFu) & (wd == whi ? vlast : 0xFu);
              },
              tini, false);
#endif
    __syncthreads();

    // ---- pass 2: NMS (strict >, neighbours outside the cell's evaluated area count as 0) of the scores >= iniThFAST;
    //      cells decide their threshold: any NMS maximum >= iniThFAST? ----
    const int wCell = L.wCell;
    const unsigned inv_wc = S.inv_wc;
    auto nms_max = [&](int r, int sb, int sc, int& jj, int& xr) -> bool {
        const int xe = sb - sb_lo;                         // column inside the strip's evaluated area
        jj = (int)__umulhi((unsigned)xe, inv_wc);
        xr = xe - jj * wCell;
        const bool left = (xr == 0), right = (xr == wCell - 1) || (xe == ew - 1);   // cell edges: neighbours beyond are 0
        const uint8_t* q = score + (r + 1) * FS_PITCH + sb;
        // branch-free: all 8 neighbours are loaded, the columns beyond a cell edge are masked to 0
        const int l3 = __vimax3_s32(q[-1], q[-FS_PITCH - 1], q[FS_PITCH - 1]);
        const int r3 = __vimax3_s32(q[1], q[-FS_PITCH + 1], q[FS_PITCH + 1]);
        const int m = __vimax3_s32(max((int)q[-FS_PITCH], (int)q[FS_PITCH]), left ? 0 : l3, right ? 0 : r3);
        return sc > m;
    };
    auto emit = [&](int r, int sb, int sc, int jj, int xr) {
        // vToDistributeKeys coordinates (ORBextractor.cc:856-857) and the reference's visiting order key
        const int cell = i * L.nCols + j0 + jj;
        const unsigned long long rec = corner_pack(j0 * wCell + 3 + (sb - sb_lo), i * L.hCell + 3 + r, sc, (cell << 12) | (r << 6) | xr);
        const int o = atomicAdd(&s_nout, 1);
        if (o < out_cap) {
            outl[o] = rec;
        } else {   // staging list full (> 3000 NMS maxima in one strip): append to the level's list directly
            const int gi = atomicAdd(corner_count + f * g.nlevels + l, 1);
            if (gi < L.corner_cap) corners[L.corner_base + (long long)f * L.corner_cap + gi] = rec;
        }
    };
    const int n1 = s_nscored;
    if (n1 <= FS_SCAP) {
        // dense: one scored corner per thread; bit 15 of the entry remembers "is an NMS maximum" (scores >= iniThFAST
        // only: a neighbour below iniThFAST cannot suppress them, so later passes never change these flags)
        for (int k = threadIdx.x; k < n1; k += FS_THREADS) {
            const int code = scored[k];
            const int r = (int)__umulhi((unsigned)code, 0xFFFFFFFFu / FS_PITCH + 1u);
            const int sb = code - r * FS_PITCH;
            const int sc = score[(r + 1) * FS_PITCH + sb];
            int jj, xr;
            if (sc >= tini && nms_max(r, sb, sc, jj, xr)) {
                scored[k] = (unsigned short)(code | 0x8000);
                s_any[jj] = 1;
            }
        }
    } else {
        // scored-list overflow: scan the score tile instead (same result, slower)
        for (int k = threadIdx.x; k < eh * nw; k += FS_THREADS) {
            const int r = fs_div(k, nw, inv_nw);
            const int wd = wlo + (k - r * nw);
            const unsigned sw = reinterpret_cast<const unsigned*>(score + (r + 1) * FS_PITCH)[wd];
            for (int b = 0; b < 4; ++b) {
                const int sc = (sw >> (8 * b)) & 0xFF;
                int jj, xr;
                if (sc >= tini && nms_max(r, (wd << 2) + b, sc, jj, xr)) s_any[jj] = 1;
            }
        }
    }
    __syncthreads();

    // ---- stage B (ORBextractor.cc:846-850): the cells without any maximum >= iniThFAST are searched again at
    //      minThFAST; their word items are enumerated cell by cell through a small prefix table ----
#if FS_QUICK == 2
    if (stats != nullptr && threadIdx.x == 0) {            // share of empty cells, for the host's choice of the variant
        int cells = 0, empties = 0;
        for (int jj = 0; jj < FS_MAXG && jj * wCell < ew; ++jj) { ++cells; empties += s_any[jj] ? 0 : 1; }
        atomicAdd(stats + 2 * (blockIdx.x & 63), empties);
        atomicAdd(stats + 2 * (blockIdx.x & 63) + 1, cells);
    }
    // (dual: stage A has already run the quick test at minThFAST: keep, of this lane's low-threshold bits, the pixels of the cells
    // that are empty, and score them; an item's 8 pixels lie in at most two cells)
    if constexpr (DUAL) {
        const int total = eh * S.np;
        unsigned long long keep = 0ull;
        int it = 0;
        for (int k = threadIdx.x; k < total; k += FS_THREADS, ++it) {
            const int b0 = (int)(codes[it * FS_THREADS + threadIdx.x] & (FS_PITCH - 1));   // shared-memory byte of the item's first pixel
            const int xe = max(b0 - sb_lo, 0);
            const int jj = (int)__umulhi((unsigned)xe, inv_wc);
            const int nb = min(max(sb_lo + (jj + 1) * wCell - b0, 0), 8);                  // pixels [0, nb) of the item lie in cell jj, the rest in jj + 1
            const unsigned lowbits = (1u << nb) - 1u;
            const unsigned km = (s_any[jj] ? 0u : lowbits) | (s_any[min(jj + 1, FS_MAXG)] ? 0u : (0xFFu & ~lowbits));
            keep |= (unsigned long long)km << (8 * it);
        }
        cmask_low &= keep;
        if (__syncthreads_or(cmask_low != 0ull))
            compact_score(cmask_low, [&](int bit) { return (int)codes[(bit >> 3) * FS_THREADS + threadIdx.x] + (bit & 7); }, true);
    } else {
#if FS_STAGEB_WORDS
    if (threadIdx.x == 0) {
        int ne = 0, pre = 0;
        for (int jj = 0; jj < FS_MAXG && jj * wCell < ew; ++jj) {
            if (s_any[jj]) continue;
            const int lo = sb_lo + jj * wCell, hi = min(lo + wCell, sb_hi);   // shared-memory bytes of the cell's evaluated columns
            const int w0 = lo >> 2, n = ((hi - 1) >> 2) - w0 + 1;
            s_bw0[ne] = w0; s_bnw[ne] = n; s_binv[ne] = 0xFFFFFFFFu / (unsigned)n + 1u;
            s_bmf[ne] = (0xFu << (lo & 3)) & 0xFu; s_bml[ne] = 0xFu >> (3 - ((hi - 1) & 3));
            s_bpre[ne] = pre; pre += n * eh; ++ne;
        }
        s_bpre[ne] = pre; s_nempty = ne;
    }
    __syncthreads();
    const int nempty = s_nempty;
    if (nempty > 0 && tmin < tini) {
        cand_pass(s_bpre[nempty],
                  [&](int k, int& r, int& wd, unsigned& vm) {
                      int e = 0;
                      while (e + 1 < nempty && k >= s_bpre[e + 1]) ++e;
                      const int kk = k - s_bpre[e], n = s_bnw[e];
                      r = fs_div(kk, n, s_binv[e]);
                      const int cw = kk - r * n;
                      wd = s_bw0[e] + cw;
                      vm = (cw == 0 ? s_bmf[e] : 0xFu) & (cw == n - 1 ? s_bml[e] : 0xFu);
                  },
                  tmin, true);
    }
#else
    // (word PAIRS like stage A; runs of adjacent empty cells are enumerated as one range, so that a pair straddling two empty
    // cells is visited once and the item count stays <= eh * pairs per row; interior cells are wider than a pair, so two
    // runs never share one)
    if (threadIdx.x == 0) {
        int ne = 0, pre = 0, start = -1;
        const int pb0 = S.w0p << 2;                        // shared-memory byte of pair 0's first pixel (<= sb_lo)
        for (int jj = 0; jj <= FS_MAXG; ++jj) {
            const bool cell = jj < FS_MAXG && jj * wCell < ew;
            const bool empty = cell && !s_any[jj];
            if (empty && start < 0) start = jj;
            if (!empty && start >= 0) {                    // cells [start, jj) are empty
                const int lo = sb_lo + start * wCell, hi = min(sb_lo + jj * wCell, sb_hi);   // shared-memory bytes of the run's evaluated columns
                const int q0 = (lo - pb0) >> 3, n = ((hi - 1 - pb0) >> 3) - q0 + 1;
                s_bw0[ne] = q0; s_bnw[ne] = n; s_binv[ne] = 0xFFFFFFFFu / (unsigned)n + 1u;
                s_bmf[ne] = (0xFFu << ((lo - pb0) & 7)) & 0xFFu; s_bml[ne] = 0xFFu >> (7 - ((hi - 1 - pb0) & 7));
                s_bpre[ne] = pre; pre += n * eh; ++ne;
                start = -1;
            }
            if (!cell) break;
        }
        s_bpre[ne] = pre; s_nempty = ne;
    }
    __syncthreads();
    const int nempty = s_nempty;
    if (nempty > 0 && tmin < tini) {
        pair_pass(tmin, s_bpre[nempty],
                  [&](int k, int& r, int& q, unsigned& vm) {
                      int e = 0;
                      while (e + 1 < nempty && k >= s_bpre[e + 1]) ++e;
                      const int kk = k - s_bpre[e], n = s_bnw[e];
                      r = fs_div(kk, n, s_binv[e]);
                      const int cp = kk - r * n;
                      q = s_bw0[e] + cp;
                      vm = s_pvm[q] & (cp == 0 ? s_bmf[e] : 0xFFu) & (cp == n - 1 ? s_bml[e] : 0xFFu);
                  },
                  true, std::false_type{});
    }
#endif
    }
#else
    if (threadIdx.x == 0) {
        int ne = 0, pre = 0;
        for (int jj = 0; jj < FS_MAXG && jj * wCell < ew; ++jj) {
            if (s_any[jj]) continue;
            const int lo = sb_lo + jj * wCell, hi = min(lo + wCell, sb_hi);   // shared-memory bytes of the cell's evaluated columns
            const int w0 = lo >> 2, n = ((hi - 1) >> 2) - w0 + 1;
            s_bw0[ne] = w0; s_bnw[ne] = n; s_binv[ne] = 0xFFFFFFFFu / (unsigned)n + 1u;
            s_bmf[ne] = (0xFu << (lo & 3)) & 0xFu; s_bml[ne] = 0xFu >> (3 - ((hi - 1) & 3));
            s_bpre[ne] = pre; pre += n * eh; ++ne;
        }
        s_bpre[ne] = pre; s_nempty = ne;
    }
    __syncthreads();
    const int nempty = s_nempty;
    if (nempty > 0 && tmin < tini) {
        cand_pass(s_bpre[nempty],
                  [&](int k, int& r, int& wd, unsigned& vm) {
                      int e = 0;
                      while (e + 1 < nempty && k >= s_bpre[e + 1]) ++e;
                      const int kk = k - s_bpre[e], n = s_bnw[e];
                      r = fs_div(kk, n, s_binv[e]);
                      const int cw = kk - r * n;
                      wd = s_bw0[e] + cw;
                      vm = (cw == 0 ? s_bmf[e] : 0xFu) & (cw == n - 1 ? s_bml[e] : 0xFu);
                  },
                  tmin, true);
    }
#endif
    __syncthreads();   // also: every read of tile / cand is done, outl may overwrite them

    // ---- pass 3: emit.  Cells with a maximum >= iniThFAST emit their flagged maxima; the others emit every NMS
    //      maximum >= minThFAST (all of their scores are in the tile after stage B) ----
    const int nscored = s_nscored;
    if (nscored <= FS_SCAP) {
        for (int k = threadIdx.x; k < nscored; k += FS_THREADS) {
            const int code = scored[k] & 0x7FFF;
            const bool flagged = (scored[k] & 0x8000) != 0;
            const int r = (int)__umulhi((unsigned)code, 0xFFFFFFFFu / FS_PITCH + 1u);
            const int sb = code - r * FS_PITCH;
            const int sc = score[(r + 1) * FS_PITCH + sb];
            const int xe = sb - sb_lo;
            int jj = (int)__umulhi((unsigned)xe, inv_wc), xr = xe - jj * wCell;
            if (sc >= tini) {
                if (!flagged) continue;
            } else {
                if (s_any[jj] || !nms_max(r, sb, sc, jj, xr)) continue;
            }
            emit(r, sb, sc, jj, xr);
        }
    } else {
        for (int k = threadIdx.x; k < eh * nw; k += FS_THREADS) {
            const int r = fs_div(k, nw, inv_nw);
            const int wd = wlo + (k - r * nw);
            const unsigned sw = reinterpret_cast<const unsigned*>(score + (r + 1) * FS_PITCH)[wd];
            for (int b = 0; b < 4; ++b) {
                const int sc = (sw >> (8 * b)) & 0xFF;
                int jj, xr;
                if (sc == 0 || !nms_max(r, (wd << 2) + b, sc, jj, xr)) continue;
                if (sc < (s_any[jj] ? tini : tmin)) continue;
                emit(r, (wd << 2) + b, sc, jj, xr);
            }
        }
    }
    __syncthreads();
    const int nout = min(s_nout, out_cap);
    if (nout == 0) return;
    if (threadIdx.x == 0) s_base = atomicAdd(corner_count + f * g.nlevels + l, nout);
    __syncthreads();
    const int base = s_base;
    unsigned long long* dst = corners + L.corner_base + (long long)f * L.corner_cap;
    for (int k = threadIdx.x; k < nout; k += FS_THREADS)
        if (base + k < L.corner_cap) dst[base + k] = outl[k];
}

}  // namespace


// experiment / tuning knob: one shared-memory carve-out for every kernel of the chain (ORB_B200_CARVEOUT, percent of the
// maximum) so that kernels of different chunks can share an SM without the SM draining to re-partition L1 / shared memory
void orb_carveout_fast(int pct) {
    cudaFuncSetAttribute(fast_strip_kernel<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(fast_strip_kernel<true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(fast_strip_kernel<false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(fast_strip_kernel<false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}

int orb_launch_fast(orb_ctx* c, const Geometry& g, int* d_corner_count, int F, int f0, cudaStream_t st, int level_begin, int level_end) {
    if (g.fast_rows > FS_ROWS) { orb_set_error("FAST strip of %d rows exceeds the tile", g.fast_rows); return ORB_ERR_INVALID; }
    if (!c->fast_attr_set) {   // > 48 KB of dynamic shared memory needs the opt-in, once per context (= per device)
        ORB_CUDA(cudaFuncSetAttribute(fast_strip_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FS_SMEM(FS_ROWS)));
        ORB_CUDA(cudaFuncSetAttribute(fast_strip_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FS_SMEM(FS_ROWS)));
        ORB_CUDA(cudaFuncSetAttribute(fast_strip_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FS_SMEM(FS_ROWS)));
        ORB_CUDA(cudaFuncSetAttribute(fast_strip_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FS_SMEM(FS_ROWS)));
        c->fast_attr_set = true;
    }
    // strips are stored level by level: a level range is a contiguous range of CTAs
    level_end = min(level_end, g.nlevels);
    const int first = g.lv[level_begin].fast_cta_base, last = level_end < g.nlevels ? g.lv[level_end].fast_cta_base : g.fast_ctas;
    if (last <= first) return ORB_OK;
    // Variant choice (fast_dual_mode 2 = by content): every few launches the kernel also reports how many of its cells were empty at
    // iniThFAST; the counters come back through a pinned buffer and are looked at, without waiting, by a later launch.  Above ~1/4
    // empty cells the two-threshold variant is faster (hysteresis 0.18 / 0.28).  Nothing of this happens inside a stream capture.
    int* stats = nullptr;
    if (c->fast_dual_mode == 2 && c->d_fast_stats) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(st, &cap);
        if (cap == cudaStreamCaptureStatusNone) {
            const cudaError_t ready = c->fast_stats_pending ? cudaEventQuery(c->ev_fast_stats) : cudaErrorNotReady;
            if (c->fast_stats_pending && ready == cudaErrorNotReady) (void)cudaGetLastError();   // "not ready" is an answer, not a fault of the launch below
            if (c->fast_stats_pending && ready == cudaSuccess) {
                long long empties = 0, cells = 0;
                for (int k = 0; k < 64; ++k) { empties += c->h_fast_stats[2 * k]; cells += c->h_fast_stats[2 * k + 1]; }
                if (cells > 0) {
                    const double share = (double)empties / (double)cells;
                    if (share > 0.28) c->fast_dual_now = true;
                    else if (share < 0.18) c->fast_dual_now = false;
                }
                c->fast_stats_pending = false;
            }
            if (!c->fast_stats_pending && (c->fast_launch_no++ & 3) == 0) stats = c->d_fast_stats;
        }
    }
    const bool dual = c->fast_dual_mode == 1 || (c->fast_dual_mode == 2 && c->fast_dual_now);
    const dim3 grid(last - first, F);
    const size_t smem = FS_SMEM(g.fast_rows);
    if (c->use_tma) {
        if (dual) fast_strip_kernel<true, true><<<grid, FS_THREADS, smem, st>>>(c->d_pyr, c->d_strips + first, c->d_corners, d_corner_count, g, c->d_tmaps, f0, stats);
        else fast_strip_kernel<true, false><<<grid, FS_THREADS, smem, st>>>(c->d_pyr, c->d_strips + first, c->d_corners, d_corner_count, g, c->d_tmaps, f0, stats);
    } else {
        if (dual) fast_strip_kernel<false, true><<<grid, FS_THREADS, smem, st>>>(c->d_pyr, c->d_strips + first, c->d_corners, d_corner_count, g, c->d_tmaps, f0, stats);
        else fast_strip_kernel<false, false><<<grid, FS_THREADS, smem, st>>>(c->d_pyr, c->d_strips + first, c->d_corners, d_corner_count, g, c->d_tmaps, f0, stats);
    }
    c->launches++;
    ORB_CUDA(cudaGetLastError());
    if (stats) {
        ORB_CUDA(cudaMemcpyAsync(c->h_fast_stats, c->d_fast_stats, 128 * sizeof(int), cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaMemsetAsync(c->d_fast_stats, 0, 128 * sizeof(int), st));
        ORB_CUDA(cudaEventRecord(c->ev_fast_stats, st));
        c->fast_stats_pending = true;
    }
    return ORB_OK;
}
