// orb_blur.cu — K5: GaussianBlur(level, 7x7, sigma 2, BORDER_REFLECT_101) of every pyramid level
// (reference orb_slam2/src/ORBextractor.cc:1129-1130; OpenCV >= 4 fixed-point path, DESIGN.md pin (i)):
//   Q8 kernel k = [18 34 48 56 48 34 18]; H = sum k_i I (exact, 16 bit); V = sum k_j H_j (exact, 24 bit);
//   out = (V + 32768) >> 16.
// The 19-px reflect-101 border of the pyramid buffers IS the blur's border (3 px are read), so the kernel reads
// the bordered level directly and writes the w x h blurred level.
//
// No shared memory: one thread owns one aligned output word (4 px) and ORB_BLUR_ROWS consecutive rows.  Per input
// row it loads the 3 aligned words around its column, forms the 8 byte windows with funnel shifts and gets the 4
// horizontal sums with 8 IDP.4A (__dp4a); the vertical pass slides over a 7-row register window (symmetric taps:
// 3 adds + 4 multiply-adds per pixel).  Neighbouring lanes read overlapping words, which L1 serves.
#include "orb_internal.cuh"

namespace {

#define BLUR_KA 0x38302212u   // bytes (18, 34, 48, 56): taps -3..0 (byte 0 = lowest address)
#define BLUR_KB 0x00122230u   // bytes (48, 34, 18,  0): taps +1..+3

__device__ __forceinline__ void blur_hrow(const unsigned* __restrict__ p, unsigned (&h)[4]) {
    const unsigned w0 = __ldg(p - 1), w1 = __ldg(p), w2 = __ldg(p + 1);   // pixels x-4..x-1 | x..x+3 | x+4..x+7
    // pixel x+q: bytes [x+q-3, x+q] . KA + bytes [x+q+1, x+q+4] . KB
    h[0] = __dp4a(__funnelshift_r(w0, w1, 8), BLUR_KA, __dp4a(__funnelshift_r(w1, w2, 8), BLUR_KB, 0u));
    h[1] = __dp4a(__funnelshift_r(w0, w1, 16), BLUR_KA, __dp4a(__funnelshift_r(w1, w2, 16), BLUR_KB, 0u));
    h[2] = __dp4a(__funnelshift_r(w0, w1, 24), BLUR_KA, __dp4a(__funnelshift_r(w1, w2, 24), BLUR_KB, 0u));
    h[3] = __dp4a(w1, BLUR_KA, __dp4a(w2, BLUR_KB, 0u));
}

template <bool FULL>
__device__ __forceinline__ void blur_rows(const unsigned* __restrict__ src, unsigned* __restrict__ dst, long long sstep, long long dstep,
                                          int rows, unsigned (&win)[7][4]) {
#pragma unroll
    for (int r = 0; r < ORB_BLUR_ROWS; ++r, src = reinterpret_cast<const unsigned*>(reinterpret_cast<const char*>(src) + sstep),
                                         dst = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(dst) + dstep)) {   // running pointers
        if (FULL || r < rows) {
            blur_hrow(src, win[(r + 6) % 7]);
            unsigned v = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const unsigned acc = 18u * (win[r % 7][q] + win[(r + 6) % 7][q]) + 34u * (win[(r + 1) % 7][q] + win[(r + 5) % 7][q]) +
                                     48u * (win[(r + 2) % 7][q] + win[(r + 4) % 7][q]) + 56u * win[(r + 3) % 7][q] + 32768u;
                v |= (acc >> 16) << (8 * q);
            }
            *dst = v;   // columns >= w of the last word are padding inside bpitch
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
    unsigned win[7][4];   // horizontal sums of the 7 rows around the current output row
#pragma unroll
    for (int j = 0; j < 6; ++j, src = reinterpret_cast<const unsigned*>(reinterpret_cast<const char*>(src) + sstep)) blur_hrow(src, win[j]);
    // full strips run without per-row guards, so that the compiler is free to issue the loads of the next rows early
    if (rows == ORB_BLUR_ROWS) blur_rows<true>(src, dst, sstep, dstep, rows, win);
    else blur_rows<false>(src, dst, sstep, dstep, rows, win);
}

}  // namespace


// experiment / tuning knob: one shared-memory carve-out for every kernel of the chain (ORB_B200_CARVEOUT, percent of the
// maximum) so that kernels of different chunks can share an SM without the SM draining to re-partition L1 / shared memory
void orb_carveout_blur(int pct) {
    cudaFuncSetAttribute(blur_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}

int orb_launch_blur(orb_ctx* c, const Geometry& g, int F, cudaStream_t st) {
    blur_kernel<<<dim3(g.blur_items / 256, F), 256, 0, st>>>(c->d_pyr, c->d_blur, g);
    c->launches++;
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}
