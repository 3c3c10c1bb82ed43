// orb_pyramid.cu — K1: ORBextractor::ComputePyramid (reference orb_slam2/src/ORBextractor.cc:1152-1185) for sm_100a.
//
//   level 0      = copyMakeBorder(image, 19, BORDER_REFLECT_101)                        ORBextractor.cc:1180
//   level l >= 1 = resize(level l-1, INTER_LINEAR) + copyMakeBorder(19, REFLECT_101)    ORBextractor.cc:1171-1175
//
// Data layout: every level of every frame is a bordered u8 image of (h + 38) rows with a
// 64-byte-multiple pitch; interior pixel (0,0) sits at byte ORB_XOFF = 32 of row 19, so interior rows are
// 16-byte aligned and all kernels work on aligned 32-bit words of 4 pixels.
//
// Launch chain per batch: pyr_copy0 (interior of level 0) -> 7 x pyr_resize (interior of level l from the stored
// u8 interior of level l-1: a true dependency, OpenCV's fixed-point result is defined on the rounded level) ->
// pyr_border (the 19-px reflect-101 frame of ALL levels in one launch; nothing in the chain reads a border).
//
// Resize arithmetic = OpenCV resize.cpp 8u INTER_LINEAR: Q11 column taps (c0,c1), Q11 row taps (b0,b1),
//   H = S[s]*c0 + S[s+1]*c1 ;  out = (((b0*(H0>>4))>>16) + ((b1*(H1>>4))>>16) + 2) >> 2
// The fast path computes one output word (4 px) x ORB_RESIZE_ROWS rows per thread: 3 aligned source words per
// source row, the two taps of a column picked with a funnel shift and multiplied with IDP.2A (__dp2a_lo), and the
// horizontal result of the lower source row is reused as the upper row of the next output row when they coincide.
#include "orb_internal.cuh"

#ifndef RESIZE_PACK
#define RESIZE_PACK 2   // how the 4 results of a resize word are packed: 0 = multiply-adds (FMA pipe), 1 / 2 = part of it on the ALU pipe
#endif

namespace {

__device__ __forceinline__ int reflect101(int i, int n) {
    // |i| < n guaranteed for a 19-px border on levels >= 20 px; the loop keeps tiny levels correct
    while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * n - 2 - i;
    return i;
}

// ---- level 0 interior: 16 bytes per thread -------------------------------------------------------------
template <bool ALIGNED>
__global__ void __launch_bounds__(256)
pyr_copy0_kernel(const uint8_t* __restrict__ in, size_t row_stride, size_t frame_stride, uint8_t* __restrict__ pyr,
                 const __grid_constant__ Geometry g) {
    const LevelGeom& L = g.lv[0];
    const int vpr = (L.w + 15) >> 4;                       // 16-byte vectors per row
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= vpr * L.h) return;
    const int y = item / vpr, x = (item - y * vpr) << 4;
    const int f = blockIdx.y;
    const uint8_t* src = in + (size_t)f * frame_stride + (size_t)y * row_stride + x;
    uint4 v;
    if (ALIGNED && x + 16 <= L.w) {
        v = __ldg(reinterpret_cast<const uint4*>(src));
    } else {
        unsigned wv[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if (x + k < L.w) wv[k >> 2] |= (unsigned)__ldg(src + k) << (8 * (k & 3));
        v = make_uint4(wv[0], wv[1], wv[2], wv[3]);
    }
    // bytes past w land in the right border and are rewritten by pyr_border_kernel
    *reinterpret_cast<uint4*>(pyr + L.base + (long long)f * L.frame_stride + L.ioff + y * L.pitch + x) = v;
}

// ---- level 0 interior from an interleaved colour image: cvtColor(.., *2GRAY) fused into the copy -----------------
// OpenCV 4.13.0 8-bit arithmetic (pin (i)):  gray = (B*3735 + G*19235 + R*9798 + 16384) >> 15.  One output word (4 px)
// per thread; CH = 3 or 4 interleaved channels, RGB = the red channel comes first.
template <int CH, bool RGB>
__global__ void __launch_bounds__(256)
pyr_copy0_color_kernel(const uint8_t* __restrict__ in, size_t row_stride, size_t frame_stride, uint8_t* __restrict__ pyr,
                       const __grid_constant__ Geometry g) {
    const LevelGeom& L = g.lv[0];
    const int wpr = (L.w + 3) >> 2;
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= wpr * L.h) return;
    const int y = item / wpr, x = (item - y * wpr) << 2;
    const int f = blockIdx.y;
    const uint8_t* src = in + (size_t)f * frame_stride + (size_t)y * row_stride + (size_t)x * CH;
    unsigned v = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (x + k < L.w) {
            const unsigned c0 = __ldg(src + k * CH), c1 = __ldg(src + k * CH + 1), c2 = __ldg(src + k * CH + 2);
            const unsigned b = RGB ? c2 : c0, r = RGB ? c0 : c2;
            v |= ((b * 3735u + c1 * 19235u + r * 9798u + 16384u) >> 15) << (8 * k);
        }
    }
    *reinterpret_cast<unsigned*>(pyr + L.base + (long long)f * L.frame_stride + L.ioff + y * L.pitch + x) = v;
}

// ---- fast resize: one output word x ORB_RESIZE_ROWS rows per thread ------------------------------------
// The ALU pipe (LOP3 / SHF / PRMT / SEL / IADD3: 2 warp-instructions per clock per SM, tools/pipe_rates.cu) bounds this
// kernel, the FMA pipe (IMAD / IDP, also 2 per clock) runs next to it.  So: the 8 tap bytes of the 4 columns come from
// 2 funnel shifts + 2 PRMT with host-built selectors, and the byte packing of the 4 results is phrased as
// multiply-adds (mad.lo by 2^16 = insert into the upper half) so that it issues on the FMA pipe.  (Measured: also moving
// the >> 4 of the horizontal pass to mul.hi overloads the FMA pipe; 38 registers -> 6 CTAs per SM.)
__device__ __forceinline__ unsigned madhi_u32(unsigned a, unsigned b, unsigned c) {
    unsigned d;
    asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ unsigned madlo_u32(unsigned a, unsigned b, unsigned c) {
    unsigned d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

template <bool LDG>
__device__ __forceinline__ void hrow4(const unsigned* __restrict__ srow, const ResizeWord& t, unsigned (&h)[4]) {
    // LDG = false: the source level was written earlier by this same launch (fused tail levels): coherent loads
    const unsigned w0 = LDG ? __ldg(srow) : srow[0], w1 = LDG ? __ldg(srow + 1) : srow[1], w2 = LDG ? __ldg(srow + 2) : srow[2];
    const unsigned A = __funnelshift_r(w0, w1, t.sh0), B = __funnelshift_r(w1, w2, t.sh0);   // 8 source bytes from column 0's left tap
    const unsigned p01 = __byte_perm(A, B, t.sel01), p23 = __byte_perm(A, B, t.sel23);      // (S[s], S[s+1]) pairs
    // (c0*S[s] + c1*S[s+1]) >> 4
    h[0] = __dp2a_lo(t.cc[0], p01, 0u) >> 4;
    h[1] = __dp2a_hi(t.cc[1], p01, 0u) >> 4;
    h[2] = __dp2a_lo(t.cc[2], p23, 0u) >> 4;
    h[3] = __dp2a_hi(t.cc[3], p23, 0u) >> 4;
}

// one work item = output word wc x rows [strip * ROWS, +ROWS) of frame f, level `level`
template <int ROWS, bool LDG>
__device__ __forceinline__ void resize_item(uint8_t* __restrict__ pyr, const ResizeTap* __restrict__ taps,
                                            const ResizeWord* __restrict__ wtaps, int level, const Geometry& g, int f, int item) {
    const LevelGeom& L = g.lv[level];
    const LevelGeom& P = g.lv[level - 1];
    const int wpr = (L.w + 3) >> 2;
    const int strip = item / wpr, wc = item - strip * wpr;
    const ResizeWord t = wtaps[L.xwtab + wc];
    // multipliers the compiler cannot see through (it would turn them back into ALU-pipe shifts)
    const unsigned k16 = (unsigned)g.one << 16, k22 = (unsigned)g.one << 22, k6 = (unsigned)g.one << 6;
    const unsigned* S = reinterpret_cast<const unsigned*>(pyr + P.base + (long long)f * P.frame_stride + P.ioff) + t.wb;   // 16-byte aligned + wb words
    uint8_t* D = pyr + L.base + (long long)f * L.frame_stride + L.ioff + 4 * wc;
    const int y0 = strip * ROWS, y1 = min(y0 + ROWS, L.h);
    const uint2* ytab = reinterpret_cast<const uint2*>(taps + L.ytab);   // ResizeTap = {u16 s0, u16 s1, s16 c0, s16 c1}
    const int ppw = P.pitch >> 2;
    unsigned h0[4], h1[4];                                // horizontal results >> 4 of the two source rows
    int have1 = -1;                                       // source row whose horizontal pass sits in h1
    for (int y = y0; y < y1; ++y) {
        const uint2 ty = __ldg(ytab + y);
        const int s0 = (int)(ty.x & 0xFFFFu), s1 = (int)(ty.x >> 16);
        if (s0 == have1) {
#pragma unroll
            for (int p = 0; p < 4; ++p) h0[p] = h1[p];
        } else {
            hrow4<LDG>(S + s0 * ppw, t, h0);
        }
        if (s1 == s0) {
#pragma unroll
            for (int p = 0; p < 4; ++p) h1[p] = h0[p];
        } else {
            hrow4<LDG>(S + s1 * ppw, t, h1);
        }
        have1 = s1;
        // ((b*(h>>4))>>16) == umulhi(b<<16, h>>4): 0 <= b <= 2048, h>>4 < 2^15
        const unsigned b0 = (ty.y & 0xFFFFu) << 16, b1 = ty.y & 0xFFFF0000u;
        unsigned s[4];                                    // 4 * out + (0..3), 10 bits
#pragma unroll
        for (int p = 0; p < 4; ++p) s[p] = madhi_u32(b1, h1[p], madhi_u32(b0, h0[p], 2u));
        // out = s >> 2, packed: even columns in the 16-bit lanes of E, odd columns (shifted to their byte) in O
#if RESIZE_PACK == 0
        const unsigned E = madlo_u32(s[2], k16, s[0]);                       // s0 | s2 << 16
        const unsigned O = madlo_u32(s[3], k22, s[1] * k6);                  // (s1 | s3 << 16) << 6
#elif RESIZE_PACK == 1
        const unsigned E = __byte_perm(s[0], s[2], 0x5410);                  // s0 | s2 << 16 (ALU)
        const unsigned O = madlo_u32(s[3], k22, s[1] * k6);                  // (s1 | s3 << 16) << 6
#else
        const unsigned E = __byte_perm(s[0], s[2], 0x5410);                  // s0 | s2 << 16 (ALU)
        const unsigned O = madlo_u32(s[3], k22, s[1] << 6);                  // (s1 | s3 << 16) << 6
#endif
        *reinterpret_cast<unsigned*>(D + y * L.pitch) = ((E >> 2) & 0x00FF00FFu) | (O & 0xFF00FF00u);
    }
}

#ifndef RESIZE_MINB
#define RESIZE_MINB 8
#endif
template <int ROWS>
__global__ void __launch_bounds__(256, RESIZE_MINB)
pyr_resize_fast_kernel(uint8_t* __restrict__ pyr, const ResizeTap* __restrict__ taps, const ResizeWord* __restrict__ wtaps,
                       int level, const __grid_constant__ Geometry g) {
    const LevelGeom& L = g.lv[level];
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= ((L.w + 3) >> 2) * ((L.h + ROWS - 1) / ROWS)) return;
    resize_item<ROWS, true>(pyr, taps, wtaps, level, g, blockIdx.y, item);
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
pyr_resize_tail_kernel(uint8_t* __restrict__ pyr, const ResizeTap* __restrict__ taps, const ResizeWord* __restrict__ wtaps,
                       int first_level, const __grid_constant__ Geometry g) {
    const int f = blockIdx.y;
    const int tid = blockIdx.x * PYR_TAIL_THREADS + threadIdx.x, nthreads = PYR_TAIL_CLUSTER * PYR_TAIL_THREADS;
    for (int level = first_level; level < g.nlevels; ++level) {
        const LevelGeom& L = g.lv[level];
        const int items = ((L.w + 3) >> 2) * ((L.h + 1) >> 1);
        for (int item = tid; item < items; item += nthreads) resize_item<2, false>(pyr, taps, wtaps, level, g, f, item);
        if (level + 1 < g.nlevels) {
            asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        }
    }
}

// ---- generic resize (any scale factor): one output word per thread, per-pixel taps -----------------------
__global__ void __launch_bounds__(256)
pyr_resize_generic_kernel(uint8_t* __restrict__ pyr, const ResizeTap* __restrict__ taps, int level,
                          const __grid_constant__ Geometry g) {
    const LevelGeom& L = g.lv[level];
    const LevelGeom& P = g.lv[level - 1];
    const int wpr = (L.w + 3) >> 2;
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= wpr * L.h) return;
    const int y = item / wpr, x4 = (item - y * wpr) << 2;
    const int f = blockIdx.y;
    const ResizeTap ty = taps[L.ytab + y];
    const uint8_t* S = pyr + P.base + (long long)f * P.frame_stride + P.ioff;
    const uint8_t* r0 = S + (int)ty.s0 * P.pitch;
    const uint8_t* r1 = S + (int)ty.s1 * P.pitch;
    const int b0 = ty.c0, b1 = ty.c1;
    unsigned v = 0;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const ResizeTap tx = taps[L.xtab + min(x4 + p, L.w - 1)];
        const int h0 = (int)r0[tx.s0] * tx.c0 + (int)r0[tx.s1] * tx.c1;
        const int h1 = (int)r1[tx.s0] * tx.c0 + (int)r1[tx.s1] * tx.c1;
        v |= (unsigned)((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2) << (8 * p);
    }
    *reinterpret_cast<unsigned*>(pyr + L.base + (long long)f * L.frame_stride + L.ioff + y * L.pitch + x4) = v;
}

// ---- the 19-px BORDER_REFLECT_101 frame of every level, one launch ---------------------------------------
// Two kinds of work items, one thread each:
//   side  : ONE border word of one bordered row (16 thread slots per row: 5 left words = bytes 12..31, up to 6 right
//           words, the first of which may straddle interior | border); each is a byte-reversed unaligned window of
//           the reflected interior row, built with one funnel shift + one byte permute, no branches
//   copy  : 16 bytes (or one trailing word) of a top/bottom bordered row — a plain aligned copy of the reflected row
// A single reflection suffices (19 < w, h: smaller levels are rejected at geometry build time because the
// reference's 30-px cell grid does not exist there either).
__device__ __forceinline__ int reflect1(int i, int n) {
    i = i < 0 ? -i : i;
    return i >= n ? 2 * n - 2 - i : i;
}

#define ORB_BORDER_SLOTS 16

__global__ void __launch_bounds__(256)
pyr_border_kernel(uint8_t* __restrict__ pyr, const __grid_constant__ Geometry g) {
    const int item0 = blockIdx.x * 256;                    // CTA-uniform: a CTA never straddles two levels or the two item kinds
    const int f = blockIdx.y;
    if (item0 < g.border_items * ORB_BORDER_SLOTS) {
        const int ritem0 = item0 / ORB_BORDER_SLOTS;
        int l = 0;
        while (l + 1 < g.nlevels && ritem0 >= g.lv[l + 1].border_base) ++l;
        const LevelGeom& L = g.lv[l];
        const int slot = threadIdx.x % ORB_BORDER_SLOTS;
        const int row = ritem0 + threadIdx.x / ORB_BORDER_SLOTS - L.border_base;   // 0 .. h + 37 (+ padding rows)
        if (row >= L.h + 2 * ORB_EDGE) return;
        uint8_t* img = pyr + L.base + (long long)f * L.frame_stride;
        const unsigned* sw = reinterpret_cast<const unsigned*>(img + L.ioff + reflect1(row - ORB_EDGE, L.h) * L.pitch);
        unsigned* drow = reinterpret_cast<unsigned*>(img + row * L.pitch);
        // word `word` of the bordered row holds interior columns x0 .. x0+3 (x0 < 0: left border, x0 + 3 >= w: right).
        // Reflect-101 of those 4 columns = the byte-reversed window [s0, s0+3] of the interior row:
        //   left:  s0 = -x0 - 3        (gfedcb|abcdefgh; byte 12 of the row is a dead byte)
        //   right: s0 = 2w - 5 - x0    (abcdefgh|gfedcba)
        const int first = (ORB_XOFF + L.w) >> 2, end = (ORB_XOFF - ORB_EDGE) / 4 + L.border_words;
        const bool left = slot < 5;
        const int word = left ? (ORB_XOFF - ORB_EDGE) / 4 + slot : first + slot - 5;
        if (word >= end) return;
        const int x0 = 4 * word - ORB_XOFF;
        const int s0 = left ? -x0 - 3 : 2 * L.w - 5 - x0;
        const unsigned refl = __byte_perm(__funnelshift_r(sw[s0 >> 2], sw[(s0 >> 2) + 1], (s0 & 3) * 8), 0u, 0x0123);
        // the straddling word keeps its n = w - x0 interior bytes (taken from the reflected source row: in a top/bottom
        // row the copy items of this launch do not write the partial word)
        const int n = left ? 0 : min(max(L.w - x0, 0), 4);
        const unsigned m = n >= 4 ? 0xFFFFu : ((1u << (4 * n)) - 1u);
        drow[word] = __byte_perm(sw[max(x0, 0) >> 2], refl, (0x3210u & m) | (0x7654u & ~m));
    } else {
        const int it0 = item0 - g.border_items * ORB_BORDER_SLOTS;
        int l = 0;
        while (l + 1 < g.nlevels && it0 >= g.lv[l + 1].copy_base) ++l;
        const LevelGeom& L = g.lv[l];
        const int it = it0 - L.copy_base + threadIdx.x;
        if (it >= L.copy_items) return;
        const int nvec = L.w >> 4, nrem = (L.w >> 2) - 4 * nvec, per_row = nvec + nrem;   // 16-byte vectors + trailing words
        const int r = (int)__umulhi((unsigned)it, L.inv_wpr), u = it - r * per_row;
        const int row = r < ORB_EDGE ? r : L.h + r;                            // r in [19, 38) -> rows h+19 .. h+37
        uint8_t* img = pyr + L.base + (long long)f * L.frame_stride;
        const uint8_t* srow = img + L.ioff + reflect1(row - ORB_EDGE, L.h) * L.pitch;
        uint8_t* drow = img + row * L.pitch + ORB_XOFF;
        if (u < nvec) reinterpret_cast<uint4*>(drow)[u] = reinterpret_cast<const uint4*>(srow)[u];
        else reinterpret_cast<unsigned*>(drow)[4 * nvec + (u - nvec)] = reinterpret_cast<const unsigned*>(srow)[4 * nvec + (u - nvec)];
    }
}

}  // namespace


// experiment / tuning knob: one shared-memory carve-out for every kernel of the chain (ORB_B200_CARVEOUT, percent of the
// maximum) so that kernels of different chunks can share an SM without the SM draining to re-partition L1 / shared memory
void orb_carveout_pyramid(int pct) {
    cudaFuncSetAttribute(pyr_copy0_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(pyr_copy0_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(pyr_resize_fast_kernel<ORB_RESIZE_ROWS>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(pyr_resize_fast_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(pyr_resize_tail_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(pyr_resize_generic_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(pyr_border_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}

// phase 0: the whole chain; 1: everything before the fused tail launch; 2: the fused tail launch only.  *tail_first_out = the
// first level the tail launch computes (nlevels when there is none).
int orb_launch_pyramid(orb_ctx* c, const Geometry& g, const uint8_t* d_imgs, int pixel_format, int F, size_t row_stride,
                       size_t frame_stride, cudaStream_t st, int phase, int* tail_first_out) {
    // small batches: the levels from the first one with <= 100 k pixels on run as ONE cluster launch (see pyr_resize_tail_kernel)
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
            pyr_resize_tail_kernel<<<dim3(PYR_TAIL_CLUSTER, F), PYR_TAIL_THREADS, 0, st>>>(c->d_pyr, c->d_taps, c->d_wtaps, tail_first, g);
            c->launches++;
        }
        ORB_CUDA(cudaGetLastError());
        return ORB_OK;
    }
    if (pixel_format != ORB_PIX_GRAY8) {
        const LevelGeom& L = g.lv[0];
        const int items = ((L.w + 3) >> 2) * L.h;
        dim3 grd((items + 255) / 256, F);
        switch (pixel_format) {
            case ORB_PIX_BGR8: pyr_copy0_color_kernel<3, false><<<grd, 256, 0, st>>>(d_imgs, row_stride, frame_stride, c->d_pyr, g); break;
            case ORB_PIX_RGB8: pyr_copy0_color_kernel<3, true><<<grd, 256, 0, st>>>(d_imgs, row_stride, frame_stride, c->d_pyr, g); break;
            case ORB_PIX_BGRA8: pyr_copy0_color_kernel<4, false><<<grd, 256, 0, st>>>(d_imgs, row_stride, frame_stride, c->d_pyr, g); break;
            case ORB_PIX_RGBA8: pyr_copy0_color_kernel<4, true><<<grd, 256, 0, st>>>(d_imgs, row_stride, frame_stride, c->d_pyr, g); break;
            default: orb_set_error("unknown pixel format %d", pixel_format); return ORB_ERR_INVALID;
        }
        c->launches++;
    } else {
        const LevelGeom& L = g.lv[0];
        const int items = ((L.w + 15) >> 4) * L.h;
        const bool aligned = ((((uintptr_t)d_imgs) | row_stride | frame_stride) & 15) == 0;
        dim3 grd((items + 255) / 256, F);
        if (aligned) pyr_copy0_kernel<true><<<grd, 256, 0, st>>>(d_imgs, row_stride, frame_stride, c->d_pyr, g);
        else pyr_copy0_kernel<false><<<grd, 256, 0, st>>>(d_imgs, row_stride, frame_stride, c->d_pyr, g);
        c->launches++;
    }
    for (int l = 1; l < tail_first; ++l) {
        const LevelGeom& L = g.lv[l];
        const int wpr = (L.w + 3) >> 2;
        if (L.fast_resize && F >= 8) {
            // throughput shape: 8 output rows per thread (the lower source row is reused 4 times out of 5)
            const int items = wpr * ((L.h + ORB_RESIZE_ROWS - 1) / ORB_RESIZE_ROWS);
            pyr_resize_fast_kernel<ORB_RESIZE_ROWS><<<dim3((items + 255) / 256, F), 256, 0, st>>>(c->d_pyr, c->d_taps, c->d_wtaps, l, g);
        } else if (L.fast_resize) {
            // latency shape (a few frames): 2 rows per thread = 4x the threads and a 4x shorter dependent-load chain
            const int items = wpr * ((L.h + 1) / 2);
            pyr_resize_fast_kernel<2><<<dim3((items + 255) / 256, F), 256, 0, st>>>(c->d_pyr, c->d_taps, c->d_wtaps, l, g);
        } else {
            const int items = wpr * L.h;
            pyr_resize_generic_kernel<<<dim3((items + 255) / 256, F), 256, 0, st>>>(c->d_pyr, c->d_taps, l, g);
        }
        c->launches++;
    }
    if (phase == 0 && tail_first < g.nlevels) {
        pyr_resize_tail_kernel<<<dim3(PYR_TAIL_CLUSTER, F), PYR_TAIL_THREADS, 0, st>>>(c->d_pyr, c->d_taps, c->d_wtaps, tail_first, g);
        c->launches++;
    }
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

int orb_launch_border(orb_ctx* c, const Geometry& g, int F, cudaStream_t st) {
    pyr_border_kernel<<<dim3((g.border_items * ORB_BORDER_SLOTS + g.border_copy_items) / 256, F), 256, 0, st>>>(c->d_pyr, g);
    c->launches++;
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}
