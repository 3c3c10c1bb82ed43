// orb_stereo.cu — Frame::ComputeStereoMatches (reference orb_slam2/src/Frame.cc:502-676) on the device.
//
// One warp per left keypoint: (1) row-band candidate search — a right keypoint is a candidate iff the left
// row lies in [floor(y-2s), ceil(y+2s)] (the vRowIndices table, Frame.cc:519-529), octave within +-1 and
// uR in [uL-maxD, uL]; minimum Hamming distance below TH_HIGH, lowest right index on ties (the table is
// filled in ascending iR order); (2) if < (TH_HIGH+TH_LOW)/2: 11x11 centre-subtracted L1 patch distance
// over 11 horizontal shifts in pyramid level kpL.octave (exact in integers), parabola sub-pixel fit and
// depth in separately-rounded fp32.  A second kernel applies the median cut (Frame.cc:662-675).
#include <string.h>

#include <algorithm>
#include <vector>

#include "orb_internal.cuh"

namespace {

#define TH_HIGH 100
#define TH_LOW 50

struct StereoScales { float scale[ORB_MAX_LEVELS]; float inv_scale[ORB_MAX_LEVELS]; };

__device__ __forceinline__ int dist256s(const uint4* a, const uint4* b) {
    const uint4 a0 = a[0], a1 = a[1], b0 = b[0], b1 = b[1];
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// per right keypoint, once: the row band [minr, maxr] it is listed under (Frame.cc:519-529), its x, octave and index, packed in
// 16 bytes — and the records SORTED INTO BINS of 16 rows by minr (counting sort, one CTA per pair): the reference looks a left
// keypoint's candidates up in vRowIndices[row]; here the left keypoint's warp scans the bins whose minr can reach its row
// ([row - maxband, row], maxband = the tallest band of the geometry) instead of all Nr right keypoints (the first version's
// N x Nr visits were 45 % of the kernel's instructions).  The order inside a bin is arbitrary: the match is a minimum over
// (distance, right index) keys, exactly the reference's "first minimum in ascending iR".
// Batched form: blockIdx.x = stereo pair; n_r (when not NULL) holds the per-pair keypoint counts on the device and the
// per-pair arrays are `cap` entries apart (bin_start: ST_MAXBINS + 1 per pair).
#define ST_BIN_SHIFT 4
#define ST_MAXBINS 640   // image heights up to 10240 rows
__global__ void __launch_bounds__(1024)
stereo_rows_kernel(const orb_kp* __restrict__ kpsR, int Nr, const StereoScales sc, uint4* __restrict__ rows, int* __restrict__ bin_start, int nbins,
                   const int* __restrict__ n_r, int cap) {
    __shared__ int s_cnt[ST_MAXBINS + 1], s_cur[ST_MAXBINS];
    const int pair = blockIdx.x;
    if (n_r) { Nr = min(n_r[pair], cap); kpsR += (size_t)pair * cap; rows += (size_t)pair * cap; }
    bin_start += (size_t)pair * (ST_MAXBINS + 1);
    for (int b = threadIdx.x; b <= nbins; b += blockDim.x) s_cnt[b] = 0;
    __syncthreads();
    auto record = [&](int i, int& bin) {
        const orb_kp kpR = kpsR[i];
        const float r = __fmul_rn(2.0f, sc.scale[kpR.octave]);
        const int maxr = (int)ceilf(__fadd_rn(kpR.y, r)), minr = (int)floorf(__fsub_rn(kpR.y, r));
        bin = min(max(minr, 0) >> ST_BIN_SHIFT, nbins - 1);
        return make_uint4(__float_as_uint(kpR.x), (unsigned)minr, (unsigned)maxr, (unsigned)kpR.octave | ((unsigned)i << 8));
    };
    for (int i = threadIdx.x; i < Nr; i += blockDim.x) { int bin; record(i, bin); atomicAdd(&s_cnt[bin], 1); }
    __syncthreads();
    if (threadIdx.x < 32) {   // exclusive prefix over the bins, one warp
        int carry = 0;
        for (int b0 = 0; b0 < nbins; b0 += 32) {
            const int b = b0 + (int)threadIdx.x;
            const int v = b < nbins ? s_cnt[b] : 0;
            int x = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if ((int)threadIdx.x >= o) x += y; }
            if (b < nbins) { s_cur[b] = carry + x - v; bin_start[b] = carry + x - v; }
            carry += __shfl_sync(0xffffffffu, x, 31);
        }
        if (threadIdx.x == 0) bin_start[nbins] = carry;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < Nr; i += blockDim.x) { int bin; const uint4 rec = record(i, bin); rows[atomicAdd(&s_cur[bin], 1)] = rec; }
}

__global__ void __launch_bounds__(256)
stereo_match_kernel(const uint8_t* __restrict__ pyrL, const uint8_t* __restrict__ pyrR, const orb_kp* __restrict__ kpsL,
                    const uint8_t* __restrict__ descL, int N, const orb_kp* __restrict__ kpsR,
                    const uint8_t* __restrict__ descR, int Nr, float mbf, float mb, const StereoScales sc,
                    float* __restrict__ uRight, float* __restrict__ depth, int* __restrict__ sad,
                    const uint4* __restrict__ rowsR, const int* __restrict__ bin_start, int nbins, int maxband,
                    const __grid_constant__ Geometry g, const int* __restrict__ n_l, const int* __restrict__ n_r, int cap) {
    const int lane = threadIdx.x & 31;
    const int iL = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int pair = blockIdx.y;      // batched form (n_l != NULL): frame `pair` of both pyramid arenas, per-pair arrays `cap` apart
    if (n_l) {
        N = min(n_l[pair], cap); Nr = min(n_r[pair], cap);
        const size_t o = (size_t)pair * cap;
        kpsL += o; descL += o * 32; kpsR += o; descR += o * 32; uRight += o; depth += o; sad += o; rowsR += o;
    }
    bin_start += (size_t)pair * (ST_MAXBINS + 1);
    if (iL >= N) {
        if (n_l && iL < cap && lane == 0) { uRight[iL] = -1.0f; depth[iL] = -1.0f; sad[iL] = -1; }   // unused slots of the pair's row
        return;
    }
    if (lane == 0) { uRight[iL] = -1.0f; depth[iL] = -1.0f; sad[iL] = -1; }
    const orb_kp kpL = kpsL[iL];
    const int levelL = kpL.octave;
    const float vL = kpL.y, uL = kpL.x;
    const int row = (int)vL;
    const int nRows = g.lv[0].h;
    if (row < 0 || row >= nRows) return;
    const float maxD = __fdiv_rn(mbf, mb);   // minZ = mb, maxD = mbf/minZ   (Frame.cc:532-534)
    const float minU = __fsub_rn(uL, maxD), maxU = uL;
    if (maxU < 0) return;
    const uint4* dL = reinterpret_cast<const uint4*>(descL + (size_t)iL * 32);
    unsigned best = 0xFFFFFFFFu;
    // right keypoints whose band can hold `row`: minr in [row - maxband, row] -> the bins of that range
    const int b_lo = min(max(row - maxband, 0) >> ST_BIN_SHIFT, nbins - 1), b_hi = min(row >> ST_BIN_SHIFT, nbins - 1);
    const int p0 = bin_start[b_lo], p1 = bin_start[b_hi + 1];
    (void)Nr;
    for (int p = p0 + lane; p < p1; p += 32) {
        const uint4 q = __ldg(rowsR + p);                // x | minr | maxr | octave + (index << 8)
        if (row < (int)q.y || row > (int)q.z) continue;
        const int oc = (int)(q.w & 0xFFu), iR = (int)(q.w >> 8);
        if (oc < levelL - 1 || oc > levelL + 1) continue;
        const float xR = __uint_as_float(q.x);
        if (xR >= minU && xR <= maxU) {
            const int d = dist256s(dL, reinterpret_cast<const uint4*>(descR + (size_t)iR * 32));
            if (d < TH_HIGH) best = min(best, ((unsigned)d << 16) | (unsigned)iR);
        }
    }
    best = __reduce_min_sync(0xffffffffu, best);
    if (best == 0xFFFFFFFFu) return;
    const int bestDist = (int)(best >> 16), bestIdxR = (int)(best & 0xFFFFu);
    if (bestDist >= (TH_HIGH + TH_LOW) / 2) return;

    // ---- sub-pixel refinement by correlation (Frame.cc:588-659) ----
    const float uR0 = kpsR[bestIdxR].x;
    const float sf = sc.inv_scale[levelL];
    const float scaleduL = roundf(__fmul_rn(kpL.x, sf));
    const float scaledvL = roundf(__fmul_rn(kpL.y, sf));
    const float scaleduR0 = roundf(__fmul_rn(uR0, sf));
    const int w = 5, Lh = 5;
    const LevelGeom& G = g.lv[levelL];
    const float iniu = __fsub_rn(__fadd_rn(scaleduR0, (float)Lh), (float)w);
    const float endu = __fadd_rn(__fadd_rn(__fadd_rn(scaleduR0, (float)Lh), (float)w), 1.0f);
    if (iniu < 0 || endu >= (float)G.w) return;
    const int y0 = (int)(scaledvL - w), xl0 = (int)(scaleduL - w), xr0 = (int)(scaleduR0 - w);
    const uint8_t* PL = pyrL + G.base + (long long)pair * G.frame_stride + G.ioff;  // interior of frame `pair` (0 in the single-pair form)
    const uint8_t* PR = pyrR + G.base + (long long)pair * G.frame_stride + G.ioff;
    const int cL = PL[(long long)(y0 + w) * G.pitch + xl0 + w];
    // each lane owns up to 4 of the 121 patch pixels
    int il[4], py[4], px[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int p = lane + 32 * k;
        py[k] = p / 11; px[k] = p - py[k] * 11;
        il[k] = (p < 121) ? (int)PL[(long long)(y0 + py[k]) * G.pitch + xl0 + px[k]] - cL : 0;
    }
    int bestSad = 0x7FFFFFFF, bestinc = 0;
    int d_prev = 0, d_at_best_m1 = 0, d_at_best = 0, d_at_best_p1 = 0;
    bool want_next = false;
    for (int inc = -Lh; inc <= Lh; ++inc) {
        const int cR = PR[(long long)(y0 + w) * G.pitch + xr0 + inc + w];
        int s = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int p = lane + 32 * k;
            if (p < 121) s += abs(il[k] - ((int)PR[(long long)(y0 + py[k]) * G.pitch + xr0 + inc + px[k]] - cR));
        }
        s = __reduce_add_sync(0xffffffffu, s);
        if (want_next) { d_at_best_p1 = s; want_next = false; }
        if (s < bestSad) { bestSad = s; bestinc = inc; d_at_best_m1 = d_prev; d_at_best = s; want_next = true; }
        d_prev = s;
    }
    if (lane != 0) return;
    sad[iL] = -2 - bestSad;  // provisional: reached refinement but (so far) no match; decoded by the host / median kernel
    if (bestinc == -Lh || bestinc == Lh) return;
    const float dist1 = (float)d_at_best_m1, dist2 = (float)d_at_best, dist3 = (float)d_at_best_p1;
    const float deltaR = __fdiv_rn(__fsub_rn(dist1, dist3),
                                   __fmul_rn(2.0f, __fsub_rn(__fadd_rn(dist1, dist3), __fmul_rn(2.0f, dist2))));
    if (deltaR < -1 || deltaR > 1) return;
    float bestuR = __fmul_rn(sc.scale[levelL], __fadd_rn(__fadd_rn(scaleduR0, (float)bestinc), deltaR));
    float disparity = __fsub_rn(uL, bestuR);
    if (disparity >= 0 && disparity < maxD) {
        if (disparity <= 0) {
            disparity = 0.01f;                       // float disparity = 0.01 (double literal)
            bestuR = (float)((double)uL - 0.01);     // bestuR = uL - 0.01 (evaluated in double)
        }
        depth[iL] = __fdiv_rn(mbf, disparity);
        uRight[iL] = bestuR;
        sad[iL] = bestSad;
    }
}

// median cut (Frame.cc:662-675): thDist = 1.5f*1.4f*median of the kept SADs, drop everything >= thDist
__global__ void __launch_bounds__(1024)
stereo_median_kernel(int N, float* uRight, float* depth, const int* __restrict__ sad, int* nkept_out, const int* __restrict__ n_l, int cap) {
    __shared__ int s_count, s_median, s_removed;
    if (n_l) {   // batched form: one CTA per pair
        const int pair = blockIdx.x;
        N = min(n_l[pair], cap);
        const size_t o = (size_t)pair * cap;
        uRight += o; depth += o; sad += o; nkept_out += pair;
    }
    if (threadIdx.x == 0) { s_count = 0; s_median = -1; s_removed = 0; }
    __syncthreads();
    int local = 0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) local += (sad[i] >= 0);
    atomicAdd(&s_count, local);
    __syncthreads();
    const int count = s_count;
    if (count == 0) { if (threadIdx.x == 0) *nkept_out = 0; return; }
    const int target = count / 2;
    // VALUE of the element of rank `target` among the kept SADs (vDistIdx[size/2].first after the sort, Frame.cc:663-664):
    // bitwise radix select over the SADs staged in shared memory, one block-wide count per bit (the SADs are sums of
    // 121 byte differences, < 2^15, so ~15 rounds of two barriers instead of an O(N^2) rank count)
    extern __shared__ int s_sad[];
    __shared__ int s_max, s_cnt[2];
    if (threadIdx.x == 0) { s_max = 0; s_cnt[0] = s_cnt[1] = 0; }
    __syncthreads();
    int lmax = 0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) { const int v = sad[i]; s_sad[i] = v; lmax = max(lmax, v); }
    atomicMax(&s_max, lmax);
    __syncthreads();
    unsigned prefix = 0;
    int rank = target;
    for (int bit = 31 - __clz(max(s_max, 1)); bit >= 0; --bit) {
        int c0 = 0;   // kept SADs that agree with the prefix above `bit` and have a 0 there
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            const int v = s_sad[i];
            c0 += (v >= 0) && (((unsigned)v >> (bit + 1)) == (prefix >> (bit + 1))) && !(((unsigned)v >> bit) & 1u);
        }
        int* cnt = &s_cnt[bit & 1];
        if (c0) atomicAdd(cnt, c0);
        __syncthreads();
        const int total0 = *cnt;
        if (rank >= total0) { rank -= total0; prefix |= 1u << bit; }
        if (threadIdx.x == 0) s_cnt[(bit & 1) ^ 1] = 0;   // the other counter is idle now: ready for the next round
        __syncthreads();
    }
    if (threadIdx.x == 0) s_median = (int)prefix;
    __syncthreads();
    const float thDist = __fmul_rn(1.5f * 1.4f, (float)s_median);
    int removed = 0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const int si = sad[i];
        if (si >= 0 && !((float)si < thDist)) { uRight[i] = -1.0f; depth[i] = -1.0f; removed++; }
    }
    atomicAdd(&s_removed, removed);
    __syncthreads();
    if (threadIdx.x == 0) *nkept_out = count - s_removed;
}

}  // namespace

// bins of 16 rows over the image height; the tallest row band of the geometry: ceil(y + r) - floor(y - r) <= 2r + 2, r = 2 * scale[octave]
static int stereo_bins(int h) { return ((std::max(h, 1) - 1) >> ST_BIN_SHIFT) + 1; }
static int stereo_maxband(const orb_ctx* c) { return (int)ceilf(4.0f * c->scale[c->nlevels - 1]) + 2; }

extern "C" int orb_stereo_match(orb_ctx* cl, orb_ctx* cr, const orb_kp* kps_l, const uint8_t* desc_l, int nl,
                                const orb_kp* kps_r, const uint8_t* desc_r, int nr, float bf, float b, float* u_right,
                                float* depth, int* nmatches) {
    if (!cl || !cr || nl < 0 || nr < 0 || !nmatches || (nl && (!kps_l || !desc_l || !u_right || !depth)) ||
        (nr && (!kps_r || !desc_r)))
        return ORB_ERR_INVALID;
    *nmatches = 0;
    for (int i = 0; i < nl; ++i) { u_right[i] = -1.0f; depth[i] = -1.0f; }
    if (!cl->have_geom || !cr->have_geom || cl->last_frames < 1 || cr->last_frames < 1) {
        orb_set_error("orb_stereo_match: both contexts must hold an extracted frame");
        return ORB_ERR_INVALID;
    }
    if (cl->device != cr->device || cl->g.w != cr->g.w || cl->g.h != cr->g.h || cl->nlevels != cr->nlevels ||
        cl->max_batch != cr->max_batch) {
        orb_set_error("orb_stereo_match: left/right contexts differ in device, image size, levels or max_batch");
        return ORB_ERR_INVALID;
    }
    if (nl == 0 || nr == 0) return ORB_OK;
    if (nr > 65535) { orb_set_error("orb_stereo_match: more than 65535 right keypoints"); return ORB_ERR_CAPACITY; }
    if (nl > 10000) { orb_set_error("orb_stereo_match: more than 10000 left keypoints"); return ORB_ERR_CAPACITY; }
    for (int i = 0; i < nl; ++i) if (kps_l[i].octave < 0 || kps_l[i].octave >= cl->nlevels) return ORB_ERR_INVALID;
    for (int i = 0; i < nr; ++i) if (kps_r[i].octave < 0 || kps_r[i].octave >= cl->nlevels) return ORB_ERR_INVALID;
    ORB_CUDA(cudaSetDevice(cl->device));
    ORB_CUDA(cudaStreamSynchronize(cr->stream));  // the right pyramid must be complete; work runs on the left stream
    cudaStream_t st = cl->stream;
    // one pinned + one device slab owned by the left context (grow-only): a single H2D, the two kernels, a single D2H
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_kl = take(sizeof(orb_kp) * nl), o_kr = take(sizeof(orb_kp) * nr), o_dl = take((size_t)32 * nl), o_dr = take((size_t)32 * nr);
    const size_t in_bytes = off;
    const size_t o_ur = take(sizeof(float) * nl), o_dep = take(sizeof(float) * nl), o_nk = take(16);
    const size_t io_bytes = off;
    const size_t o_sad = take(sizeof(int) * nl), o_rows = take(sizeof(uint4) * (size_t)nr), o_bins = take(sizeof(int) * (ST_MAXBINS + 1));
    if (cl->d_scratch_cap < off) {
        ORB_CUDA(cudaStreamSynchronize(st));
        cudaFree(cl->d_scratch); cl->d_scratch = nullptr; cl->d_scratch_cap = 0;
        ORB_CUDA(cudaMalloc(&cl->d_scratch, off * 2));
        cl->d_scratch_cap = off * 2;
    }
    if (cl->h_scratch_cap < io_bytes) {
        ORB_CUDA(cudaStreamSynchronize(st));
        cudaFreeHost(cl->h_scratch); cl->h_scratch = nullptr; cl->h_scratch_cap = 0;
        ORB_CUDA(cudaMallocHost(&cl->h_scratch, io_bytes * 2));
        cl->h_scratch_cap = io_bytes * 2;
    }
    uint8_t *H = cl->h_scratch, *D = cl->d_scratch;
    memcpy(H + o_kl, kps_l, sizeof(orb_kp) * nl); memcpy(H + o_kr, kps_r, sizeof(orb_kp) * nr);
    memcpy(H + o_dl, desc_l, (size_t)32 * nl); memcpy(H + o_dr, desc_r, (size_t)32 * nr);
    ORB_CUDA(cudaMemcpyAsync(D, H, in_bytes, cudaMemcpyHostToDevice, st));
    {
        StereoScales sc;
        for (int l = 0; l < ORB_MAX_LEVELS; ++l) { sc.scale[l] = l < cl->nlevels ? cl->scale[l] : 1.f; sc.inv_scale[l] = l < cl->nlevels ? cl->inv_scale[l] : 1.f; }
        const int nbins = stereo_bins(cl->g.h), maxband = stereo_maxband(cl);
        if (nbins > ST_MAXBINS) { orb_set_error("orb_stereo_match: images taller than %d rows are not supported", ST_MAXBINS << ST_BIN_SHIFT); return ORB_ERR_CAPACITY; }
        stereo_rows_kernel<<<1, 1024, 0, st>>>((const orb_kp*)(D + o_kr), nr, sc, (uint4*)(D + o_rows), (int*)(D + o_bins), nbins, nullptr, 0);
        stereo_match_kernel<<<(nl + 7) / 8, 256, 0, st>>>(cl->d_pyr, cr->d_pyr, (const orb_kp*)(D + o_kl), D + o_dl, nl, (const orb_kp*)(D + o_kr),
                                                         D + o_dr, nr, bf, b, sc, (float*)(D + o_ur), (float*)(D + o_dep), (int*)(D + o_sad),
                                                         (const uint4*)(D + o_rows), (const int*)(D + o_bins), nbins, maxband, cl->g, nullptr, nullptr, 0);
        stereo_median_kernel<<<1, 1024, sizeof(int) * (size_t)nl, st>>>(nl, (float*)(D + o_ur), (float*)(D + o_dep), (const int*)(D + o_sad), (int*)(D + o_nk), nullptr, 0);
        cl->launches += 3;
    }
    ORB_CUDA(cudaGetLastError());
    ORB_CUDA(cudaMemcpyAsync(H + o_ur, D + o_ur, io_bytes - o_ur, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaStreamSynchronize(st));
    memcpy(u_right, H + o_ur, sizeof(float) * nl);
    memcpy(depth, H + o_dep, sizeof(float) * nl);
    *nmatches = *(const int*)(H + o_nk);
    return ORB_OK;
}


// Batched, device-resident form: pair p = frame p of both contexts' pyramid arenas (the last orb_extract_batch_device call of
// each), keypoints / descriptors / counts exactly as that call left them on the device.  Asynchronous on the LEFT context's
// stream; the right context's stream is joined with an event.  No host copy, no host synchronisation.
extern "C" int orb_stereo_match_batch_device(orb_ctx* cl, orb_ctx* cr, int npairs, const orb_kp* d_kps_l, const uint8_t* d_desc_l,
                                             const int32_t* d_n_l, const orb_kp* d_kps_r, const uint8_t* d_desc_r, const int32_t* d_n_r,
                                             int cap, float bf, float b, float* d_u_right, float* d_depth, int32_t* d_nmatches) {
    if (!cl || !cr || npairs < 0 || cap <= 0 || !d_kps_l || !d_desc_l || !d_n_l || !d_kps_r || !d_desc_r || !d_n_r || !d_u_right || !d_depth ||
        !d_nmatches)
        return ORB_ERR_INVALID;
    if (npairs == 0) return ORB_OK;
    if (!cl->have_geom || !cr->have_geom || cl->last_frames < npairs || cr->last_frames < npairs) {
        orb_set_error("orb_stereo_match_batch_device: both contexts must hold %d extracted frames", npairs);
        return ORB_ERR_INVALID;
    }
    if (cl->device != cr->device || cl->g.w != cr->g.w || cl->g.h != cr->g.h || cl->nlevels != cr->nlevels || cl->max_batch != cr->max_batch) {
        orb_set_error("orb_stereo_match_batch_device: left/right contexts differ in device, image size, levels or max_batch");
        return ORB_ERR_INVALID;
    }
    if (cap > 10000) { orb_set_error("orb_stereo_match_batch_device: more than 10000 keypoints per frame"); return ORB_ERR_CAPACITY; }
    ORB_CUDA(cudaSetDevice(cl->device));
    cudaStream_t st = cl->stream;
    if (cr->stream != st) {   // the right pyramids / keypoints must be complete before the left stream reads them
        cudaEvent_t ev;
        ORB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        ORB_CUDA(cudaEventRecord(ev, cr->stream));
        ORB_CUDA(cudaStreamWaitEvent(st, ev, 0));
        ORB_CUDA(cudaEventDestroy(ev));
    }
    const size_t per = (size_t)npairs * cap;
    const size_t bins_bytes = ((size_t)npairs * (ST_MAXBINS + 1) * sizeof(int) + 255) & ~(size_t)255;
    const size_t need = per * (sizeof(int) + sizeof(uint4)) + 512 + bins_bytes;
    if (cl->d_scratch_cap < need) {
        ORB_CUDA(cudaStreamSynchronize(st));
        cudaFree(cl->d_scratch); cl->d_scratch = nullptr; cl->d_scratch_cap = 0;
        ORB_CUDA(cudaMalloc(&cl->d_scratch, need));
        cl->d_scratch_cap = need;
    }
    uint4* d_rows = reinterpret_cast<uint4*>(cl->d_scratch);
    int* d_sad = reinterpret_cast<int*>(cl->d_scratch + ((per * sizeof(uint4) + 255) & ~(size_t)255));
    int* d_bins = reinterpret_cast<int*>(cl->d_scratch + ((per * sizeof(uint4) + 255) & ~(size_t)255) + ((per * sizeof(int) + 255) & ~(size_t)255));
    const int nbins = stereo_bins(cl->g.h), maxband = stereo_maxband(cl);
    if (nbins > ST_MAXBINS) { orb_set_error("orb_stereo_match_batch_device: images taller than %d rows are not supported", ST_MAXBINS << ST_BIN_SHIFT); return ORB_ERR_CAPACITY; }
    StereoScales sc;
    for (int l = 0; l < ORB_MAX_LEVELS; ++l) { sc.scale[l] = l < cl->nlevels ? cl->scale[l] : 1.f; sc.inv_scale[l] = l < cl->nlevels ? cl->inv_scale[l] : 1.f; }
    stereo_rows_kernel<<<npairs, 1024, 0, st>>>(d_kps_r, 0, sc, d_rows, d_bins, nbins, d_n_r, cap);
    stereo_match_kernel<<<dim3((cap + 7) / 8, npairs), 256, 0, st>>>(cl->d_pyr, cr->d_pyr, d_kps_l, d_desc_l, 0, d_kps_r, d_desc_r, 0, bf, b, sc, d_u_right,
                                                                    d_depth, d_sad, d_rows, d_bins, nbins, maxband, cl->g, d_n_l, d_n_r, cap);
    stereo_median_kernel<<<npairs, 1024, sizeof(int) * (size_t)cap, st>>>(0, d_u_right, d_depth, d_sad, d_nmatches, d_n_l, cap);
    cl->launches += 3;
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}
