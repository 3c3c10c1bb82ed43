// orb_blur.cu — K5: GaussianBlur(level, 7x7, sigma 2, BORDER_REFLECT_101) of every pyramid level
// (reference orb_slam2/src/ORBextractor.cc:1129-1130; OpenCV >= 4 fixed-point path, DESIGN.md pin (i)):
//   Q8 kernel k = [18 34 48 56 48 34 18]; H = sum k_i I (exact, 16 bit); V = sum k_j H_j (exact, 24 bit);
//   out = (V + 32768) >> 16.
// The 19-px reflect-101 border of the pyramid buffers IS the blur's border (3 px are read), so the kernel reads
// the bordered level directly and writes the w x h blurred level.
//
// No shared memory: one thread owns one aligned output word (4 px) and ORB_BLUR_ROWS consecutive rows.  Per input
// row it loads the 3 aligned words around its column and gets the 4 horizontal sums with 8 IDP.4A (__dp4a): the two outer
// columns straight from the aligned words with shifted coefficient vectors, the two inner ones from funnel-shifted
// windows (4 SHF; all-coefficient-shifted would be 10 IDP.4A and no shift: the FMA pipe is the scarcer one here).
// Vertical pass: the horizontal sums (<= 65280, 16 bit) of two consecutive input rows share a register (even row in the
// low half), so the 7 taps of an output pixel are 4 IDP.2A (u16 pair . u8 pair) chained through the accumulator, which
// starts at the rounding constant — instead of 3 adds + 4 multiply-adds on unpacked sums; the register window is 4 row
// pairs x 4 columns.  The 4 results (byte 2 of each accumulator) are packed with 3 PRMT.
#include "orb_internal.cuh"

namespace {

// Q8 taps (18, 34, 48, 56, 48, 34, 18) as byte vectors (byte 0 = lowest address = first row / leftmost pixel)
#define BLUR_K0123 0x38302212u   // (18, 34, 48, 56)
#define BLUR_K456_ 0x00122230u   // (48, 34, 18,  0)
#define BLUR_K_012 0x30221200u   // ( 0, 18, 34, 48)
#define BLUR_K3456 0x12223038u   // (56, 48, 34, 18)

// horizontal sums of the 4 pixels of the word at p (pixels x .. x+3) of one input row
template <bool LDG>
__device__ __forceinline__ void blur_hrow(const unsigned* __restrict__ p, unsigned (&h)[4]) {
    const unsigned w0 = LDG ? __ldg(p - 1) : p[-1], w1 = LDG ? __ldg(p) : p[0], w2 = LDG ? __ldg(p + 1) : p[1];   // pixels x-4..x-1 | x..x+3 | x+4..x+7
    // pixel x:   bytes x-3..x-1 of w0 (its bytes 1..3) . (18,34,48)  +  w1 . (56,48,34,18)
    h[0] = __dp4a(w0, BLUR_K_012, __dp4a(w1, BLUR_K3456, 0u));
    // pixels x+1, x+2: bytes [x+q-3, x+q] . (18,34,48,56) + bytes [x+q+1, x+q+4] . (48,34,18,0)
    h[1] = __dp4a(__funnelshift_r(w0, w1, 16), BLUR_K0123, __dp4a(__funnelshift_r(w1, w2, 16), BLUR_K456_, 0u));
    h[2] = __dp4a(__funnelshift_r(w0, w1, 24), BLUR_K0123, __dp4a(__funnelshift_r(w1, w2, 24), BLUR_K456_, 0u));
    // pixel x+3: w1 . (18,34,48,56) + w2 . (48,34,18,0)
    h[3] = __dp4a(w1, BLUR_K0123, __dp4a(w2, BLUR_K456_, 0u));
}

// one output word from 4 row pairs; EVEN: the 7 input rows start at the first row of pair a, else at its second row
template <bool EVEN>
__device__ __forceinline__ unsigned blur_vword(const unsigned (&a)[4], const unsigned (&b)[4], const unsigned (&c)[4], const unsigned (&d)[4]) {
    unsigned acc[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (EVEN)   // rows (a.lo a.hi b.lo b.hi c.lo c.hi d.lo) x (18 34 48 56 48 34 18)
            acc[q] = __dp2a_hi(d[q], BLUR_K456_, __dp2a_lo(c[q], BLUR_K456_, __dp2a_hi(b[q], BLUR_K0123, __dp2a_lo(a[q], BLUR_K0123, 32768u))));
        else        // rows (a.hi b.lo b.hi c.lo c.hi d.lo d.hi) x (18 34 48 56 48 34 18)
            acc[q] = __dp2a_hi(d[q], BLUR_K3456, __dp2a_lo(c[q], BLUR_K3456, __dp2a_hi(b[q], BLUR_K_012, __dp2a_lo(a[q], BLUR_K_012, 32768u))));
    }
    // out = acc >> 16 = byte 2 of the accumulator (acc < 2^24)
    return __byte_perm(__byte_perm(acc[0], acc[1], 0x0062), __byte_perm(acc[2], acc[3], 0x0062), 0x5410);
}

template <bool FULL, bool LDG>
__device__ __forceinline__ void blur_rows(const unsigned* __restrict__ src, unsigned* __restrict__ dst, long long sstep, long long dstep,
                                          int rows, unsigned (&win)[4][4]) {
    // on entry win[0..2] hold the row pairs (0,1) (2,3) (4,5) of the strip's input rows (input row i = output row i - 3)
    auto next_src = [&]() { src = reinterpret_cast<const unsigned*>(reinterpret_cast<const char*>(src) + sstep); };
    auto next_dst = [&]() { dst = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(dst) + dstep); };
#pragma unroll
    for (int r = 0; r < ORB_BLUR_ROWS; r += 2) {   // output rows r, r + 1 need input rows r .. r + 7 = pairs r/2 .. r/2 + 3
        if (FULL || r < rows) {
            unsigned he[4], ho[4];
            blur_hrow<LDG>(src, he); next_src();                   // input row r + 6
            if (FULL || r + 1 < rows) { blur_hrow<LDG>(src, ho); next_src(); }   // input row r + 7 (only output row r + 1 reads it)
            else { ho[0] = ho[1] = ho[2] = ho[3] = 0u; }
            unsigned (&pa)[4] = win[(r / 2) % 4], (&pb)[4] = win[(r / 2 + 1) % 4], (&pc)[4] = win[(r / 2 + 2) % 4], (&pd)[4] = win[(r / 2 + 3) % 4];
#pragma unroll
            for (int q = 0; q < 4; ++q) pd[q] = __byte_perm(he[q], ho[q], 0x5410);   // even row in the low half
            *dst = blur_vword<true>(pa, pb, pc, pd); next_dst();   // columns >= w of the last word are padding inside bpitch
            if (FULL || r + 1 < rows) { *dst = blur_vword<false>(pa, pb, pc, pd); next_dst(); }
        }
    }
}

#ifndef BLUR_MINB
#define BLUR_MINB 4
#endif
__global__ void __launch_bounds__(256, BLUR_MINB)
blur_kernel(const uint8_t* __restrict__ pyr, uint8_t* __restrict__ blur, const __grid_constant__ Geometry g) {
    // every level starts at a CTA boundary (blur_base is a multiple of 256): level and pitches are warp-uniform
    const int cta0 = blockIdx.x * 256;
    int l = 0;
    while (l + 1 < g.nlevels && cta0 >= g.lv[l + 1].blur_base) ++l;
    const LevelGeom& L = g.lv[l];
    const int f = blockIdx.y;
    const int it = cta0 - L.blur_base + threadIdx.x;
    const int strip = it / L.blur_wpr, wc = it - strip * L.blur_wpr;
    const int y0 = strip * ORB_BLUR_ROWS;
    if (y0 >= L.h) return;
    const int rows = min(ORB_BLUR_ROWS, L.h - y0);
    const int pw = L.pitch >> 2;
    // word holding interior pixels 4wc .. 4wc+3 of row y0 - 3 (rows -3..-1 and h..h+2 are border rows)
    const unsigned* src = reinterpret_cast<const unsigned*>(pyr + L.base + (long long)f * L.frame_stride + L.ioff) +
                          (y0 - 3) * pw + wc;
    unsigned* dst = reinterpret_cast<unsigned*>(blur + L.bbase + (long long)f * L.bframe_stride) + y0 * (L.bpitch >> 2) + wc;
    const int bpw = L.bpitch >> 2;
    const long long sstep = L.pitch, dstep = L.bpitch;   // byte strides, widened once
    unsigned win[4][4];   // horizontal sums of 4 input-row pairs (even row in the low half) x 4 columns
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        unsigned he[4], ho[4];
        blur_hrow<true>(src, he); src = reinterpret_cast<const unsigned*>(reinterpret_cast<const char*>(src) + sstep);
        blur_hrow<true>(src, ho); src = reinterpret_cast<const unsigned*>(reinterpret_cast<const char*>(src) + sstep);
#pragma unroll
        for (int q = 0; q < 4; ++q) win[j][q] = __byte_perm(he[q], ho[q], 0x5410);
    }
    // full strips run without per-row guards, so that the compiler is free to issue the loads of the next rows early
    if (rows == ORB_BLUR_ROWS) blur_rows<true, true>(src, dst, sstep, dstep, rows, win);
    else blur_rows<false, true>(src, dst, sstep, dstep, rows, win);
}

}  // namespace


// experiment / tuning knob: one shared-memory carve-out for every kernel of the chain (ORB_B200_CARVEOUT, percent of the
// maximum) so that kernels of different chunks can share an SM without the SM draining to re-partition L1 / shared memory
void orb_carveout_blur(int pct) {
    cudaFuncSetAttribute(blur_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}

int orb_launch_blur(orb_ctx* c, const Geometry& g, int F, cudaStream_t st) {
    // (a variant that stages the strip's input rows in shared memory with cp.async.bulk, like the pyramid kernels, was
    // measured at the same 0.31-0.32 ms per 512 frames: the kernel is bound by instruction issue on the FMA pipe — IDP.4A /
    // IDP.2A, 78 % busy, issue slots 83 % — not by load latency)
    blur_kernel<<<dim3(g.blur_items / 256, F), 256, 0, st>>>(c->d_pyr, c->d_blur, g);
    c->launches++;
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}
