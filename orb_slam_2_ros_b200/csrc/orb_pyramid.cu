// orb_pyramid.cu — K1: ORBextractor::ComputePyramid (reference orb_slam2/src/ORBextractor.cc:1152-1185) for sm_100a.
//
//   level 0      = copyMakeBorder(image, 19, BORDER_REFLECT_101)                        ORBextractor.cc:1180
//   level l >= 1 = resize(level l-1, INTER_LINEAR) + copyMakeBorder(19, REFLECT_101)    ORBextractor.cc:1171-1175
//
// Data layout: every level of every frame is a bordered u8 image of (h + 38) rows with a
// 64-byte-multiple pitch; interior pixel (0,0) sits at byte ORB_XOFF = 32 of row 19, so interior rows are
// 16-byte aligned and all kernels work on aligned 32-bit words of 4 pixels.
//
// Launch chain per batch: pyr_level0 (bordered level 0 from the caller's frame) -> 7 x pyr_level (bordered level l from the
// stored u8 interior of level l-1: a true dependency, OpenCV's fixed-point result is defined on the rounded level).  The
// 19-px reflect-101 frame is written by the pass that computes the level: side borders are extra words of the row whose
// taps are those of the reflected columns, top / bottom border rows are second stores of the interior rows they reflect.
//
// Resize arithmetic = OpenCV resize.cpp 8u INTER_LINEAR: Q11 column taps (c0,c1), Q11 row taps (b0,b1),
//   H = S[s]*c0 + S[s+1]*c1 ;  out = (((b0*(H0>>4))>>16) + ((b1*(H1>>4))>>16) + 2) >> 2
// The fast path computes one output word (4 px) x ORB_RESIZE_ROWS rows per thread: 3 aligned source words per
// source row, the two taps of a column picked with a funnel shift + byte permute and multiplied with IDP.2A (__dp2a_lo),
// and the horizontal result of the lower source row is reused as the upper row of the next output row when they coincide.
#include "orb_internal.cuh"

namespace {

#ifndef RESIZE_UNROLL
#define RESIZE_UNROLL 2   // row pairs per unrolled iteration of the staged kernel's full-strip loop
#endif

// ---- level 0: the bordered image straight from the caller's frame ----------------------------------------------------
// One thread = one 32-bit word of a bordered row (words 3 .. 3 + border_words of the row: bytes 12 ..).  Interior words of
// a 4-byte-aligned gray frame are plain word copies; border words are byte-reversed unaligned windows of the same input
// row (reflect-101: gfedcb|abcdefgh|gfedcba), the word that straddles interior | right border takes its interior bytes
// from the row and the rest from the reflection.  Rows 1..19 and h-20..h-2 are stored a second time as the top / bottom
// border rows they reflect into, so the whole 19-px frame is written by the pass that produces the level.
__device__ __forceinline__ int reflect1(int i, int n) {   // a single reflection suffices: 19 < w, h (smaller levels are rejected
    i = i < 0 ? -i : i;                                   // at geometry build time because the reference's 30-px cell grid does
    return i >= n ? 2 * n - 2 - i : i;                    // not exist there either)
}

// store `v` at word D of bordered row y + 19 and at the border rows that reflect interior row y
template <typename T>
__device__ __forceinline__ void store_mirrored(T* __restrict__ D, int dp, int y, int h, T v) {
    D[(y + ORB_EDGE) * dp] = v;
    if ((unsigned)(y - 1) < (unsigned)ORB_EDGE) D[(ORB_EDGE - y) * dp] = v;                         // y in [1, 19]  -> row 19 - y
    if ((unsigned)(h - 2 - y) < (unsigned)ORB_EDGE) D[(2 * h + ORB_EDGE - 2 - y) * dp] = v;          // y in [h-20, h-2] -> row 2h + 17 - y
}

// CH = 1: gray; 3 / 4: interleaved colour, cvtColor(.., *2GRAY) of Tracking::GrabImage* fused into the pass — OpenCV 4.13.0
// 8-bit arithmetic (pin (i)):  gray = (B*3735 + G*19235 + R*9798 + 16384) >> 15;  RGB = the red channel comes first.
// ALIGN (gray only): 16 = the frame's rows are 16-byte aligned (interior vectors are one LDG.128), 4 = 4-byte aligned (the
// input is read as aligned words), 1 = bytes.  One thread = one 16-byte vector of a bordered row.
template <int CH, bool RGB, int ALIGN>
__global__ void __launch_bounds__(256)
pyr_level0_kernel(const uint8_t* __restrict__ in, size_t row_stride, size_t frame_stride, uint8_t* __restrict__ pyr,
                  const __grid_constant__ Geometry g) {
    const LevelGeom& L = g.lv[0];
    const int vpr = L.pitch >> 4;                             // 16-byte vectors per bordered row
    // item space: first the vectors that lie wholly inside the interior (l0_ni per row: plain copies for aligned gray input),
    // then, from a warp boundary on, the other vpr - l0_ni vectors of every row — so no warp mixes the two kinds
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    int y, v;
    bool inner;
    if (item < g.l0_border_first) {
        if (item >= g.l0_ni * L.h) return;
        y = (int)__umulhi((unsigned)item, g.l0_inv_ni); v = 2 + item - y * g.l0_ni;
        inner = true;
    } else {
        const int it = item - g.l0_border_first, nb = vpr - g.l0_ni;
        if (it >= nb * L.h) return;
        y = (int)__umulhi((unsigned)it, g.l0_inv_nb);
        v = it - y * nb;
        v = v < 2 ? v : v + g.l0_ni;                          // vectors 0, 1 (left border), then those from the last interior vector on
        inner = false;
    }
    const int f = blockIdx.y;
    const int xv = 16 * v - ORB_XOFF;                         // interior column of the vector's first byte
    const uint8_t* src = in + (size_t)f * frame_stride + (size_t)y * row_stride;
    uint4 out;
    if (CH == 1 && ALIGN == 16 && inner) {
        out = __ldg(reinterpret_cast<const uint4*>(src + xv));
    } else {
        unsigned wv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int x0 = xv + 4 * q;                        // word q: columns x0 .. x0 + 3 (bytes left of column -19 and right of w + 18 are dead)
            unsigned r;
            if (CH == 1 && ALIGN >= 4) {
                const unsigned* sw = reinterpret_cast<const unsigned*>(src);
                const int lastw = (L.w - 1) >> 2;
                if (inner || (x0 >= 0 && x0 + 4 <= L.w)) {
                    r = __ldg(sw + (x0 >> 2));
                } else {
                    // reflect-101 of the 4 columns = the byte-reversed window [s0, s0 + 3] of the row: left s0 = -x0 - 3, right s0 = 2w - 5 - x0
                    const bool left = x0 < 0;
                    const int s0 = min(max(left ? -x0 - 3 : 2 * L.w - 5 - x0, 0), L.w - 1);
                    const int w0 = s0 >> 2, w1 = min(w0 + 1, lastw);
                    const unsigned refl = __byte_perm(__funnelshift_r(__ldg(sw + w0), __ldg(sw + w1), (s0 & 3) * 8), 0u, 0x0123);
                    // the straddling word keeps its n = w - x0 interior bytes
                    const int n = left ? 0 : min(max(L.w - x0, 0), 4);
                    const unsigned m = n >= 4 ? 0xFFFFu : ((1u << (4 * n)) - 1u);
                    r = __byte_perm(__ldg(sw + min(max(x0, 0) >> 2, lastw)), refl, (0x3210u & m) | (0x7654u & ~m));
                }
            } else {
                r = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int x = inner ? x0 + k : min(max(reflect1(x0 + k, L.w), 0), L.w - 1);   // dead bytes repeat an edge pixel
                    unsigned px;
                    if (CH == 1) {
                        px = __ldg(src + x);
                    } else {
                        const unsigned c0 = __ldg(src + x * CH), c1 = __ldg(src + x * CH + 1), c2 = __ldg(src + x * CH + 2);
                        const unsigned b = RGB ? c2 : c0, rr = RGB ? c0 : c2;
                        px = (b * 3735u + c1 * 19235u + rr * 9798u + 16384u) >> 15;
                    }
                    r |= px << (8 * k);
                }
            }
            wv[q] = r;
        }
        out = make_uint4(wv[0], wv[1], wv[2], wv[3]);
    }
    store_mirrored(reinterpret_cast<uint4*>(pyr + L.base + (long long)f * L.frame_stride) + v, vpr, y, L.h, out);
}

// ---- levels >= 1: resize(level l-1, INTER_LINEAR) + copyMakeBorder(19, REFLECT_101) in ONE pass ----------------------
// One work item = one word of the BORDERED row (bytes 12 ..) x ROWS interior rows.  A border column is the resized value of
// its reflected column, so the per-word tap table (ResizeWord, built on the host for all border_words words of a row)
// simply lists the reflected columns' taps: border words cost what interior words cost and no pass ever re-reads the
// level to frame it.  Border ROWS are second stores of the interior rows they reflect (store_mirrored).
//
// The ALU pipe (LOP3 / SHF / PRMT / IADD3: 2 warp-instructions per clock per SM, tools/pipe_rates.cu) and the FMA pipe
// (IMAD / IDP, also 2 per clock) share the work: the 8 tap bytes of the 4 columns come from 2 funnel shifts + 2 PRMT with
// host-built selectors, the horizontal blend is IDP.2A, the vertical blend IMAD.HI with a zero addend (an addend would be
// a 64-bit register pair that the compiler fills with two moves per multiply).  The two horizontal results (upper / lower
// source row) live in two register sets whose roles swap from row to row (the row loop is unrolled by two), so the lower
// row of one output row becomes the upper row of the next without register moves.
template <bool LDG>
__device__ __forceinline__ void load3(const unsigned* __restrict__ srow, unsigned (&w)[3]) {
    // LDG = false: the source level was written earlier by this same launch (fused tail levels): coherent loads
    w[0] = LDG ? __ldg(srow) : srow[0]; w[1] = LDG ? __ldg(srow + 1) : srow[1]; w[2] = LDG ? __ldg(srow + 2) : srow[2];
}
__device__ __forceinline__ void hcalc4(const unsigned (&w)[3], const ResizeWord& t, unsigned (&h)[4]) {
    const unsigned A = __funnelshift_r(w[0], w[1], t.sh0), B = __funnelshift_r(w[1], w[2], t.sh0);   // 8 source bytes from the leftmost tap
    const unsigned p01 = __byte_perm(A, B, t.sel01), p23 = __byte_perm(A, B, t.sel23);              // (S[s], S[s+1]) pairs
    // (c0*S[s] + c1*S[s+1]) >> 4
    h[0] = __dp2a_lo(t.cc[0], p01, 0u) >> 4;
    h[1] = __dp2a_hi(t.cc[1], p01, 0u) >> 4;
    h[2] = __dp2a_lo(t.cc[2], p23, 0u) >> 4;
    h[3] = __dp2a_hi(t.cc[3], p23, 0u) >> 4;
}

// out = (((b0*(H0>>4))>>16) + ((b1*(H1>>4))>>16) + 2) >> 2 for the 4 columns, packed into one word
__device__ __forceinline__ unsigned vblend4(const unsigned (&hu)[4], const unsigned (&hl)[4], unsigned by) {
    // ((b*(h>>4))>>16) == umulhi(b<<16, h>>4): 0 <= b <= 2048, h>>4 < 2^15
    const unsigned b0 = by << 16, b1 = by & 0xFFFF0000u;
    unsigned s[4];                                        // 4 * out + (0..3) - 2, 10 bits
#pragma unroll
    for (int p = 0; p < 4; ++p) s[p] = __umulhi(b0, hu[p]) + __umulhi(b1, hl[p]);
    // the rounding constant is added to two columns at a time (no carry between the 16-bit lanes: s + 2 < 2^11)
    const unsigned E = __byte_perm(s[0], s[2], 0x5410) + 0x00020002u;   // s0 | s2 << 16
    const unsigned O = __byte_perm(s[1], s[3], 0x5410) + 0x00020002u;   // s1 | s3 << 16
    unsigned r;
    asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(r) : "r"(0x00FF00FFu), "r"(E >> 2), "r"(O << 6));   // mask ? E >> 2 : O << 6
    return r;
}

// MIRROR = false: the strip holds no row that reflects into the top / bottom border (the usual case)
template <int ROWS, bool LDG, bool MIRROR>
__device__ __forceinline__ void level_rows(const unsigned* __restrict__ Sbase, unsigned soff, unsigned ppw, unsigned* __restrict__ Dbase,
                                           unsigned doff, unsigned dpw, const uint2* __restrict__ ytab, const ResizeWord& t, int y0, int y1,
                                           int h) {
    // The kernel is bounded by the latency of its dependent loads (table entry -> source words -> blend -> store; ncu: the
    // long-scoreboard stall dominates, issue slots 45 % busy), so the loop is software-pipelined by hand: the table entry
    // and the lower source row of output row y + 1 are requested before row y is computed.
    unsigned ha[4], hb[4];                                // horizontal results >> 4 of two source rows
    unsigned wa[3], wb[3];                                // raw words of the lower source row: current / prefetched
    int tag = -1;                                         // source row held by the set that was "lower" in the previous output row
    uint2 ty = LDG ? __ldg(ytab + y0) : ytab[y0];
    load3<LDG>(Sbase + (soff + (ty.x >> 16) * ppw), wa);
    // one output row: `up` is the set that was lower in the previous row (holds source row `tag`), `lo` the other one;
    // `wcur` holds the words of this row's lower source row, `wnext` receives those of the next row
    auto row = [&](int y, unsigned (&up)[4], unsigned (&lo)[4], unsigned (&wcur)[3], unsigned (&wnext)[3]) {
        const unsigned s0 = ty.x & 0xFFFFu, s1 = ty.x >> 16, by = ty.y;
        if (y + 1 < y1) {
            ty = LDG ? __ldg(ytab + y + 1) : ytab[y + 1];
            load3<LDG>(Sbase + (soff + (ty.x >> 16) * ppw), wnext);
        }
        if ((int)s0 != tag) {                             // first row of the strip, or the source rows advanced by two
            unsigned wt[3];
            load3<LDG>(Sbase + (soff + s0 * ppw), wt);
            hcalc4(wt, t, up);
        }
        hcalc4(wcur, t, lo);
        tag = (int)s1;
        const unsigned v = vblend4(up, lo, by);
        Dbase[doff + (unsigned)(y + ORB_EDGE) * dpw] = v;
        if (MIRROR) {
            if ((unsigned)(y - 1) < (unsigned)ORB_EDGE) Dbase[doff + (unsigned)(ORB_EDGE - y) * dpw] = v;                      // y in [1, 19] -> row 19 - y
            if ((unsigned)(h - 2 - y) < (unsigned)ORB_EDGE) Dbase[doff + (unsigned)(2 * h + ORB_EDGE - 2 - y) * dpw] = v;       // y in [h-20, h-2] -> row 2h + 17 - y
        }
    };
    for (int y = y0; y < y1; y += 2) {
        row(y, ha, hb, wa, wb);
        if (y + 1 < y1) row(y + 1, hb, ha, wb, wa);
    }
}

template <int ROWS, bool LDG>
__device__ __forceinline__ void level_item(uint8_t* __restrict__ pyr, const ResizeTap* __restrict__ taps,
                                           const ResizeWord* __restrict__ wtaps, int level, const Geometry& g, int f, int item) {
    const LevelGeom& L = g.lv[level];
    const LevelGeom& P = g.lv[level - 1];
    const int nwb = L.border_words;
    const int strip = (int)__umulhi((unsigned)item, L.inv_wpr), bw = item - strip * nwb;
    const ResizeWord t = wtaps[L.xwtab + bw];
    // CTA-uniform 64-bit bases + 32-bit word offsets (one IMAD.WIDE per address instead of a 64-bit add chain)
    const unsigned* Sbase = reinterpret_cast<const unsigned*>(pyr + P.base + (long long)f * P.frame_stride + P.ioff);   // source interior (0,0), 16-byte aligned
    unsigned* Dbase = reinterpret_cast<unsigned*>(pyr + L.base + (long long)f * L.frame_stride);                       // bordered row 0
    const int y0 = strip * ROWS, y1 = min(y0 + ROWS, L.h);
    const uint2* ytab = reinterpret_cast<const uint2*>(taps + L.ytab);   // ResizeTap = {u16 s0, u16 s1, s16 c0, s16 c1}
    const unsigned ppw = (unsigned)P.pitch >> 2, dpw = (unsigned)L.pitch >> 2;
    if (y0 <= ORB_EDGE || y1 > L.h - ORB_EDGE - 2)
        level_rows<ROWS, LDG, true>(Sbase, (unsigned)t.wb, ppw, Dbase, 3u + (unsigned)bw, dpw, ytab, t, y0, y1, L.h);
    else
        level_rows<ROWS, LDG, false>(Sbase, (unsigned)t.wb, ppw, Dbase, 3u + (unsigned)bw, dpw, ytab, t, y0, y1, L.h);
}

#ifndef RESIZE_MINB
#define RESIZE_MINB 5
#endif
template <int ROWS>
__global__ void __launch_bounds__(256, RESIZE_MINB)
pyr_level_kernel(uint8_t* __restrict__ pyr, const ResizeTap* __restrict__ taps, const ResizeWord* __restrict__ wtaps,
                 int level, const __grid_constant__ Geometry g) {
    const LevelGeom& L = g.lv[level];
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= L.border_words * ((L.h + ROWS - 1) / ROWS)) return;
    level_item<ROWS, true>(pyr, taps, wtaps, level, g, blockIdx.y, item);
}

// ---- throughput shape: the source rows of a strip staged in shared memory by bulk async copies ------------------------
// With one output word per thread the direct-load kernel above keeps only 12 bytes per thread in flight: ncu shows the
// long-scoreboard stall dominating and ~2 TB/s per launch.  Here one CTA = ROWS output rows x one segment of <= 256
// bordered words; ONE thread requests every source row the strip touches (<= ROWS * 4/3 + 2 rows of the segment's source
// columns) with cp.async.bulk — the copy engine keeps the whole tile in flight, no registers or warps are tied up —
// and all threads wait on the mbarrier, then run the same row loop out of shared memory.
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// row loop of the staged kernel: everything a row needs besides the pixels comes from ONE LDS.128 of a per-CTA row record
// {byte offset of the upper staged source row, of the lower one, b0 << 16, b1 << 16}; full strips run fully unrolled
template <int ROWS, bool MIRROR, bool FULL>
__device__ __forceinline__ void staged_rows(const unsigned char* __restrict__ sthread, const uint4* __restrict__ s_row, unsigned* __restrict__ Dbase,
                                            unsigned doff, unsigned dpw, const ResizeWord& t, int y0, int nrows, int h) {
    unsigned ha[4], hb[4];                                // horizontal results >> 4 of two source rows
    unsigned tag = 0xFFFFFFFFu;                           // staged row (byte offset) held by the set that was "lower" in the previous output row
    auto row = [&](int i, unsigned (&up)[4], unsigned (&lo)[4]) {
        const uint4 r = s_row[i];
        unsigned w[3];
        if (r.x != tag) {                                 // first row of the strip, or the source rows advanced by two
            const unsigned* p = reinterpret_cast<const unsigned*>(sthread + r.x);
            w[0] = p[0]; w[1] = p[1]; w[2] = p[2];
            hcalc4(w, t, up);
        }
        const unsigned* p = reinterpret_cast<const unsigned*>(sthread + r.y);
        w[0] = p[0]; w[1] = p[1]; w[2] = p[2];
        hcalc4(w, t, lo);
        tag = r.y;
        unsigned s[4];                                    // 4 * out + (0..3) - 2, 10 bits
#pragma unroll
        for (int q = 0; q < 4; ++q) s[q] = __umulhi(r.z, up[q]) + __umulhi(r.w, lo[q]);
        const unsigned E = __byte_perm(s[0], s[2], 0x5410) + 0x00020002u;   // s0 | s2 << 16, rounding constant for both lanes
        const unsigned O = __byte_perm(s[1], s[3], 0x5410) + 0x00020002u;   // s1 | s3 << 16
        unsigned v;
        asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(v) : "r"(0x00FF00FFu), "r"(E >> 2), "r"(O << 6));   // mask ? E >> 2 : O << 6
        const int y = y0 + i;
        Dbase[doff + (unsigned)(y + ORB_EDGE) * dpw] = v;
        if (MIRROR) {
            if ((unsigned)(y - 1) < (unsigned)ORB_EDGE) Dbase[doff + (unsigned)(ORB_EDGE - y) * dpw] = v;                      // y in [1, 19] -> row 19 - y
            if ((unsigned)(h - 2 - y) < (unsigned)ORB_EDGE) Dbase[doff + (unsigned)(2 * h + ORB_EDGE - 2 - y) * dpw] = v;       // y in [h-20, h-2] -> row 2h + 17 - y
        }
    };
    if (FULL) {
        constexpr int kUnroll = RESIZE_UNROLL;
#pragma unroll kUnroll
        for (int i = 0; i < ROWS; i += 2) { row(i, ha, hb); row(i + 1, hb, ha); }
    } else {
        for (int i = 0; i < nrows; i += 2) {
            row(i, ha, hb);
            if (i + 1 < nrows) row(i + 1, hb, ha);
        }
    }
}

template <int ROWS>
__global__ void __launch_bounds__(256, RESIZE_MINB)
pyr_level_staged_kernel(uint8_t* __restrict__ pyr, const ResizeTap* __restrict__ taps, const ResizeWord* __restrict__ wtaps,
                        const ResizeSeg* __restrict__ segs, int level, const __grid_constant__ Geometry g) {
    extern __shared__ __align__(128) unsigned char stage_raw[];
    __shared__ __align__(16) uint4 s_row[ROWS];
    __shared__ __align__(8) unsigned long long s_mbar;
    __shared__ unsigned s_span[2];
    const LevelGeom& L = g.lv[level];
    const LevelGeom& P = g.lv[level - 1];
    const int f = blockIdx.y, tid = threadIdx.x;
    const int strip = blockIdx.x / L.nseg, seg = blockIdx.x - strip * L.nseg;
    const ResizeSeg sg = segs[L.seg_base + seg];
    const int y0 = strip * ROWS, nrows = min(ROWS, L.h - y0);
    const uint2* ytab = reinterpret_cast<const uint2*>(taps + L.ytab);   // ResizeTap = {u16 s0, u16 s1, s16 c0, s16 c1}
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&s_mbar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // source rows s0(first row) .. s1(last row): the taps ascend with y
        const unsigned s_first = __ldg(ytab + y0).x & 0xFFFFu, s_last = __ldg(ytab + y0 + nrows - 1).x >> 16;
        const unsigned n = s_last - s_first + 1u;
        s_span[0] = s_first;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&s_mbar)), "r"(n * (unsigned)sg.nbytes) : "memory");
        const uint8_t* src = pyr + P.base + (long long)f * P.frame_stride + P.ioff + (long long)s_first * P.pitch + 4 * sg.w0;
        for (unsigned r = 0; r < n; ++r)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_addr(stage_raw + r * sg.nbytes)), "l"(src + (size_t)r * P.pitch), "r"(sg.nbytes), "r"(smem_addr(&s_mbar)) : "memory");
    }
    __syncthreads();
    if (tid < nrows) {
        const uint2 ty = __ldg(ytab + y0 + tid);
        const unsigned s_first = s_span[0];
        s_row[tid] = make_uint4(((ty.x & 0xFFFFu) - s_first) * (unsigned)sg.nbytes, ((ty.x >> 16) - s_first) * (unsigned)sg.nbytes,
                                ty.y << 16, ty.y & 0xFFFF0000u);   // ((b*(h>>4))>>16) == umulhi(b<<16, h>>4): 0 <= b <= 2048, h>>4 < 2^15
    }
    const bool active = tid < sg.nw;
    const int bw = sg.bw0 + (active ? tid : 0);
    const ResizeWord t = wtaps[L.xwtab + bw];
    __syncthreads();
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "STAGE_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra STAGE_DONE;\n"
        "bra STAGE_WAIT;\n"
        "STAGE_DONE:\n"
        "}\n" ::"r"(smem_addr(&s_mbar)), "r"(0) : "memory");
    if (!active) return;
    unsigned* Dbase = reinterpret_cast<unsigned*>(pyr + L.base + (long long)f * L.frame_stride);   // bordered row 0
    const unsigned dpw = (unsigned)L.pitch >> 2;
    const unsigned char* sthread = stage_raw + 4 * (t.wb - sg.w0);   // this thread's first staged word of staged row 0
    const bool mirror = y0 <= ORB_EDGE || y0 + nrows > L.h - ORB_EDGE - 2;
    if (mirror) staged_rows<ROWS, true, false>(sthread, s_row, Dbase, 3u + (unsigned)bw, dpw, t, y0, nrows, L.h);
    else if (nrows == ROWS) staged_rows<ROWS, false, true>(sthread, s_row, Dbase, 3u + (unsigned)bw, dpw, t, y0, nrows, L.h);
    else staged_rows<ROWS, false, false>(sthread, s_row, Dbase, 3u + (unsigned)bw, dpw, t, y0, nrows, L.h);
}

// ---- fused tail: levels [first, nlevels) of one frame in ONE launch --------------------------------------------------
// The upper levels are tiny (<= 71 k pixels each at 640x480) and strictly dependent: as separate launches of a small batch
// each costs a launch gap and a latency-bound wave.  Here a thread-block CLUSTER of 8 CTAs owns one frame, computes a
// level with all its threads and meets at the hardware cluster barrier (release / acquire at cluster scope) before the
// next one.  Used for chunks of the host pipeline and the single-frame path; large resident batches keep one launch per level.
#ifndef PYR_TAIL_MAXF
#define PYR_TAIL_MAXF 128
#endif
#define PYR_TAIL_CLUSTER 8
#define PYR_TAIL_THREADS 512
__global__ void __cluster_dims__(PYR_TAIL_CLUSTER, 1, 1) __launch_bounds__(PYR_TAIL_THREADS)
pyr_level_tail_kernel(uint8_t* __restrict__ pyr, const ResizeTap* __restrict__ taps, const ResizeWord* __restrict__ wtaps,
                      int first_level, const __grid_constant__ Geometry g) {
    const int f = blockIdx.y;
    const int tid = blockIdx.x * PYR_TAIL_THREADS + threadIdx.x, nthreads = PYR_TAIL_CLUSTER * PYR_TAIL_THREADS;
    for (int level = first_level; level < g.nlevels; ++level) {
        const LevelGeom& L = g.lv[level];
        const int items = L.border_words * ((L.h + 1) >> 1);
        for (int item = tid; item < items; item += nthreads) level_item<2, false>(pyr, taps, wtaps, level, g, f, item);
        if (level + 1 < g.nlevels) {
            asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        }
    }
}

// ---- generic resize (any scale factor): one bordered output word per thread, per-pixel taps of the reflected columns ----
__global__ void __launch_bounds__(256)
pyr_level_generic_kernel(uint8_t* __restrict__ pyr, const ResizeTap* __restrict__ taps, int level,
                         const __grid_constant__ Geometry g) {
    const LevelGeom& L = g.lv[level];
    const LevelGeom& P = g.lv[level - 1];
    const int nwb = L.border_words;
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= nwb * L.h) return;
    const int y = item / nwb, bw = item - y * nwb;
    const int x0 = 4 * bw - (ORB_XOFF - 12);
    const int f = blockIdx.y;
    const ResizeTap ty = taps[L.ytab + y];
    const uint8_t* S = pyr + P.base + (long long)f * P.frame_stride + P.ioff;
    const uint8_t* r0 = S + (int)ty.s0 * P.pitch;
    const uint8_t* r1 = S + (int)ty.s1 * P.pitch;
    const int b0 = ty.c0, b1 = ty.c1;
    unsigned v = 0;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const ResizeTap tx = taps[L.xtab + min(reflect1(x0 + p, L.w), L.w - 1)];
        const int h0 = (int)r0[tx.s0] * tx.c0 + (int)r0[tx.s1] * tx.c1;
        const int h1 = (int)r1[tx.s0] * tx.c0 + (int)r1[tx.s1] * tx.c1;
        v |= (unsigned)((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2) << (8 * p);
    }
    store_mirrored(reinterpret_cast<unsigned*>(pyr + L.base + (long long)f * L.frame_stride) + 3 + bw, L.pitch >> 2, y, L.h, v);
}

}  // namespace


// experiment / tuning knob: one shared-memory carve-out for every kernel of the chain (ORB_B200_CARVEOUT, percent of the
// maximum) so that kernels of different chunks can share an SM without the SM draining to re-partition L1 / shared memory
void orb_carveout_pyramid(int pct) {
    cudaFuncSetAttribute(pyr_level0_kernel<1, false, 16>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(pyr_level0_kernel<1, false, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(pyr_level0_kernel<1, false, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(pyr_level_staged_kernel<ORB_RESIZE_ROWS>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(pyr_level_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(pyr_level_tail_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(pyr_level_generic_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}

// phase 0: the whole chain; 1: everything before the fused tail launch; 2: the fused tail launch only.  *tail_first_out = the
// first level the tail launch computes (nlevels when there is none).
int orb_launch_pyramid(orb_ctx* c, const Geometry& g, const uint8_t* d_imgs, int pixel_format, int F, size_t row_stride,
                       size_t frame_stride, cudaStream_t st, int phase, int* tail_first_out) {
    // small batches: the levels from the first one with <= 100 k pixels on run as ONE cluster launch (see pyr_level_tail_kernel)
    int tail_first = g.nlevels;
    if (F <= PYR_TAIL_MAXF) {
        for (int l = g.nlevels - 1; l >= 2; --l) {
            if (!g.lv[l].fast_resize || g.lv[l].w * g.lv[l].h > 100000) break;
            tail_first = l;
        }
        if (g.nlevels - tail_first < 2) tail_first = g.nlevels;
    }
    if (tail_first_out) *tail_first_out = tail_first;
    if (phase < 0) return ORB_OK;   // query only
    if (phase == 2) {
        if (tail_first < g.nlevels) {
            pyr_level_tail_kernel<<<dim3(PYR_TAIL_CLUSTER, F), PYR_TAIL_THREADS, 0, st>>>(c->d_pyr, c->d_taps, c->d_wtaps, tail_first, g);
            c->launches++;
        }
        ORB_CUDA(cudaGetLastError());
        return ORB_OK;
    }
    {
        const LevelGeom& L = g.lv[0];
        const int items = g.l0_border_first + ((L.pitch >> 4) - g.l0_ni) * L.h;
        dim3 grd((items + 255) / 256, F);
        const uintptr_t al = ((uintptr_t)d_imgs) | row_stride | frame_stride;
        switch (pixel_format) {
            case ORB_PIX_GRAY8:
                if ((al & 15) == 0) pyr_level0_kernel<1, false, 16><<<grd, 256, 0, st>>>(d_imgs, row_stride, frame_stride, c->d_pyr, g);
                else if ((al & 3) == 0) pyr_level0_kernel<1, false, 4><<<grd, 256, 0, st>>>(d_imgs, row_stride, frame_stride, c->d_pyr, g);
                else pyr_level0_kernel<1, false, 1><<<grd, 256, 0, st>>>(d_imgs, row_stride, frame_stride, c->d_pyr, g);
                break;
            case ORB_PIX_BGR8: pyr_level0_kernel<3, false, 1><<<grd, 256, 0, st>>>(d_imgs, row_stride, frame_stride, c->d_pyr, g); break;
            case ORB_PIX_RGB8: pyr_level0_kernel<3, true, 1><<<grd, 256, 0, st>>>(d_imgs, row_stride, frame_stride, c->d_pyr, g); break;
            case ORB_PIX_BGRA8: pyr_level0_kernel<4, false, 1><<<grd, 256, 0, st>>>(d_imgs, row_stride, frame_stride, c->d_pyr, g); break;
            case ORB_PIX_RGBA8: pyr_level0_kernel<4, true, 1><<<grd, 256, 0, st>>>(d_imgs, row_stride, frame_stride, c->d_pyr, g); break;
            default: orb_set_error("unknown pixel format %d", pixel_format); return ORB_ERR_INVALID;
        }
        c->launches++;
    }
    for (int l = 1; l < tail_first; ++l) {
        const LevelGeom& L = g.lv[l];
        const int nwb = L.border_words;
        if (L.fast_resize && F >= 8) {
            // throughput shape: 8 output rows per thread (the lower source row is reused 4 times out of 5), source rows staged in
            // shared memory by bulk copies
            const int strips = (L.h + ORB_RESIZE_ROWS - 1) / ORB_RESIZE_ROWS;
            pyr_level_staged_kernel<ORB_RESIZE_ROWS><<<dim3(L.nseg * strips, F), L.seg_threads, L.stage_bytes, st>>>(c->d_pyr, c->d_taps, c->d_wtaps,
                                                                                                                 c->d_rsegs, l, g);
        } else if (L.fast_resize) {
            // latency shape (a few frames): 2 rows per thread = 4x the threads and a 4x shorter dependent-load chain
            const int items = nwb * ((L.h + 1) / 2);
            pyr_level_kernel<2><<<dim3((items + 255) / 256, F), 256, 0, st>>>(c->d_pyr, c->d_taps, c->d_wtaps, l, g);
        } else {
            const int items = nwb * L.h;
            pyr_level_generic_kernel<<<dim3((items + 255) / 256, F), 256, 0, st>>>(c->d_pyr, c->d_taps, l, g);
        }
        c->launches++;
    }
    if (phase == 0 && tail_first < g.nlevels) {
        pyr_level_tail_kernel<<<dim3(PYR_TAIL_CLUSTER, F), PYR_TAIL_THREADS, 0, st>>>(c->d_pyr, c->d_taps, c->d_wtaps, tail_first, g);
        c->launches++;
    }
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}
