// orb_extract_kernels.cu — sm_100a kernels of the ORB extractor, part 2, and the launch sequence of
// ORBextractor::operator() (reference orb_slam2/src/ORBextractor.cc:1083-1149).
//
// Stages (batched over frames; blockIdx.y = frame):
//   K1  orb_pyramid.cu   ComputePyramid                        ORBextractor.cc:1152-1185
//   K2  orb_fast.cu      per-cell FAST + retry                 ORBextractor.cc:820-863
//   K3  quadtree_kernel  DistributeOctTree                     ORBextractor.cc:561-787       (this file)
//   K5  orb_blur.cu      GaussianBlur 7x7 sigma 2              ORBextractor.cc:1129-1130
//   K4+K6 orient_describe_kernel  IC_Angle + rBRIEF + output   ORBextractor.cc:77-147,1134-1147 (this file)
//
// All arithmetic is integer or separately-rounded IEEE fp32 (compiled with -fmad=false) so the results are
// bit-identical to the CPU reference semantics (OpenCV 4.13.0 primitives).
#include "orb_internal.cuh"

namespace {

__constant__ int c_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

// ======================================================================================================
// K3: DistributeOctTree as block-wide scans.  One CTA per (level, frame).
// Nodes keep stable ids; `list` holds the ids in std::list order.  One iteration expands the nodes of a
// processing order S (phase 1: every multi-point node in list order, ORBextractor.cc:630-689; phase 2: the
// previous iteration's multi-point children sorted by (size, creation index) and walked from the back with
// the size>=N cut, ORBextractor.cc:700-761).  New list = reverse(children in creation order) ++ (old list
// minus expanded parents), exactly what push_front / erase produce.
// ======================================================================================================
#ifndef QT_THREADS
#define QT_THREADS 256        // throughput shape (many frames in flight)
#endif
#define QT_THREADS_LAT 1024   // latency shape (a few frames): 4x the threads for the key passes of the big levels

struct QtNode { short x0, y0, x1, y1; };

template <int NT>
__device__ __forceinline__ int block_exclusive_scan(int* data, int n, int* s_warp /*[NT / 32]*/, int* s_carry) {
    // in-place exclusive scan of data[0..n), returns the total; all threads must call
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) *s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += NT) {
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
        // totals of the warps before this one: one load per lane + a warp reduction (redux.sync) instead of a loop over s_warp
        const int wt = lane < NT / 32 ? s_warp[lane] : 0;
        const int woff = __reduce_add_sync(0xffffffffu, lane < wid ? wt : 0);
        const int carry = *s_carry;
        if (i < n) data[i] = carry + woff + x - v;
        __syncthreads();
        if (threadIdx.x == NT - 1) *s_carry = carry + woff + x;
        __syncthreads();
    }
    return *s_carry;
}

template <int NT>
__global__ void __launch_bounds__(NT)
quadtree_kernel(const unsigned long long* __restrict__ corners, const int* __restrict__ corner_count,
                unsigned short* __restrict__ node_of_key, unsigned long long* __restrict__ kept,
                int* __restrict__ kept_count, int* __restrict__ tie_count, const __grid_constant__ Geometry g, int level0) {
    extern __shared__ __align__(16) uint8_t smem[];
    // longest first: level 0 holds the most corners and the largest quota, so all frames' level-0 CTAs are dispatched before
    // any level-1 CTA (x = frame varies fastest) and the short upper-level CTAs fill the tail of the launch
    const int l = blockIdx.y + level0, f = blockIdx.x;   // level0: first level of this launch (the latency path launches two level ranges)
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
    __shared__ int s_warp[NT / 32], s_carry, s_L, s_nS, s_nExp, s_ncand, s_ncand2, s_nexp_children, s_nextid;

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
            const int nS = block_exclusive_scan<NT>(scan_a, Lcur, s_warp, &s_carry);
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
        block_exclusive_scan<NT>(scan_b, nS, s_warp, &s_carry);
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
        const int T = block_exclusive_scan<NT>(scan_a, nExp, s_warp, &s_carry);  // total children created
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
        const int nkeep = block_exclusive_scan<NT>(scan_a, Lcur, s_warp, &s_carry);
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
#ifndef OD_KPW
#define OD_KPW 4                       // keypoints per warp
#endif
#define OD_KPB (OD_WARPS * OD_KPW)     // keypoints per CTA (== 32: one lane of warp 0 per keypoint in phase 2)
#define OD_ITEMS 297                   // 33 rows x 9 aligned words (31 patch rows + 2 all-zero rows): 11 steps of 27 lanes
#define OD_TAPR 19                     // |tap offset| <= 19 after rotation
#ifndef OD_WPITCH
#define OD_WPITCH 80                    // bytes per row of the staged tap window: four 16-byte chunks hold columns px-19 .. px+19; 80 (not 64) spreads the rows over the banks
#endif
#define OD_WBYTES ((((2 * OD_TAPR + 1) * OD_WPITCH) + 127) / 128 * 128)   // per-warp window, a multiple of 128 bytes (TMA destination alignment)

static_assert(ORB_TAP_BOX_W * ORB_TAP_BOX_H <= OD_WBYTES && ORB_TAP_BOX_W % 16 == 0 && ORB_TAP_BOX_W >= 64, "tap window box must fit the per-warp window");
// ---- TMA (cp.async.bulk.tensor) + mbarrier plumbing for the tap window ----
__device__ __forceinline__ unsigned od_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void od_mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(od_smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void od_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(od_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void od_mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "OD_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra OD_WAIT_DONE;\n"
        "bra OD_WAIT_LOOP;\n"
        "OD_WAIT_DONE:\n"
        "}\n" ::"r"(od_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void od_tma_load_3d(void* dst, const CUtensorMap* map, int x, int y, int z, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(od_smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(od_smem_u32(bar)) : "memory");
}

__device__ __forceinline__ int dp4a_su(unsigned a_signed, unsigned b_unsigned, int c) {
    int d;
    asm("dp4a.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_signed), "r"(b_unsigned), "r"(c));
    return d;
}

// One CTA = 32 consecutive OUTPUT keypoints of one frame (level-major order, ORBextractor.cc:1122-1147):
//   1. every warp: prefix of the per-level kept counts -> (level, index) of its 4 keypoints; patch moments with
//      aligned word loads: lane k takes (row, word) items k, k+32, ... of the 31 x 9-word patch window and two
//      IDP.4A per word against a precomputed table of signed u-weights / circular-mask bytes (indexed by the
//      alignment of the patch), then  m10 += sum u*I,  m01 += v * sum I
//   2. warp 0, one LANE per keypoint: fastAtan2 + the double-precision sincos of pin (iii)
//   3. every warp builds its 4 descriptors (lane = descriptor byte); the 512-point pattern sits in shared memory as
//      float4 [8][32] so that a warp-wide read is conflict-free (a per-lane index into __constant__ serialises)
// KPW = keypoints per warp: 4 for throughput (32 per CTA), 1 for the latency shape of a few frames (8 per CTA, 4x the CTAs and
// a 4x shorter serial chain per warp).
#ifndef OD_MINB
#define OD_MINB 5
#endif
template <int KPW>
__global__ void __launch_bounds__(OD_WARPS * 32, OD_MINB)
orient_describe_kernel(const uint8_t* __restrict__ pyr, const uint8_t* __restrict__ blur,
                       const unsigned long long* __restrict__ kept, const int* __restrict__ kept_count,
                       const uint2* __restrict__ mom_tab, orb_kp* __restrict__ kps_out, uint8_t* __restrict__ desc_out,
                       int cap, int* __restrict__ n_out, const __grid_constant__ Geometry g, const FastTmaps* __restrict__ btm, int f0) {
    __shared__ float4 s_pat[8 * 32];
    __shared__ __align__(128) uint8_t s_win[OD_WARPS][OD_WBYTES];   // per warp: the blurred tap window of the keypoint in work
    __shared__ __align__(8) unsigned long long s_bar[OD_WARPS];    // TMA form: one mbarrier per warp
    const bool tma = btm != nullptr;
    const int wpitch = tma ? ORB_TAP_BOX_W : OD_WPITCH;            // the TMA unit writes the box densely
    unsigned bar_phase = 0;
    __shared__ int s_m[(OD_WARPS * KPW)][2];
    __shared__ float s_ang[(OD_WARPS * KPW)], s_a[(OD_WARPS * KPW)], s_b[(OD_WARPS * KPW)];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = blockIdx.y;
    // ---- per-level kept counts: lane l holds level l, inclusive prefix by shuffles ----
    const int kc = lane < g.nlevels ? kept_count[f * g.nlevels + lane] : 0;
    int cum = kc;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, cum, o);
        if (lane >= o) cum += y;
    }
    const int total = min(__shfl_sync(0xffffffffu, cum, ORB_MAX_LEVELS - 1), cap);
    if (blockIdx.x == 0 && warp == 0) {
        const int tot_all = __shfl_sync(0xffffffffu, cum, ORB_MAX_LEVELS - 1);
        if (lane == 0) n_out[f] = tot_all;
    }
    if (blockIdx.x * (OD_WARPS * KPW) >= total) return;   // uniform per CTA
    // pattern: byte `i` of the descriptor uses points 16 i .. 16 i + 15; s_pat[k * 32 + i] = (x0, y0, x1, y1) of bit k
    // (the table lies behind the moment weights, already in this layout: one coalesced 16-byte load per thread; reading it from
    // the __constant__ pattern cost 128 serialised constant-cache accesses per warp, 15 % of the kernel's stall samples)
    s_pat[threadIdx.x] = __ldg(reinterpret_cast<const float4*>(mom_tab + 4 * OD_ITEMS) + threadIdx.x);
    // ---- output index -> (level, index in level) for this warp's 4 keypoints ----
    const int od_row = lane / 9, od_word = lane - 9 * od_row;   // phase 1: this lane's (row within a 3-row step, word) item
    int lv[KPW], px[KPW], py[KPW], off[KPW], sc[KPW];
#pragma unroll
    for (int q = 0; q < KPW; ++q) {
        const int o = blockIdx.x * (OD_WARPS * KPW) + warp * KPW + q;
        const unsigned below = __ballot_sync(0xffffffffu, lane < g.nlevels && o < cum);   // levels whose prefix exceeds o
        lv[q] = -1; px[q] = py[q] = sc[q] = 0; off[q] = o;
        if (o < total && below) {
            const int l = __ffs(below) - 1;
            const int first = __shfl_sync(0xffffffffu, cum - kc, l);
            const unsigned long long rec = kept[(long long)f * g.total_kp_slots + g.lv[l].kp_base + (o - first)];
            lv[q] = l; sc[q] = corner_score(rec);
            px[q] = corner_x(rec) + ORB_MINB; py[q] = corner_y(rec) + ORB_MINB;   // ORBextractor.cc:881-882
        }
    }
    // ---- the 512 descriptor taps of a keypoint touch ~420 different 32-byte sectors when gathered from L1, but the
    //      (2*19+1)^2 window they live in is only ~80 sectors: the warp copies the window into shared memory with
    //      16-byte cp.async (5 per lane, rows py-19 .. py+19, the 3-4 aligned chunks that hold columns px-19 .. px+19) and
    //      gathers from there.  One window per warp (20 KB per CTA): a larger footprint would shrink the L1 that the
    //      moment / record loads of this kernel live on.  Keypoints lie >= 19 px inside the level (EDGE_THRESHOLD), so
    //      the window never leaves the blurred level's rows; chunk columns past the row end are padding / the next row. ----
    if (tma && lane == 0) od_mbar_init(&s_bar[warp], 1);
    __syncwarp();
    auto stage_window = [&](int lq, int x, int y) {
        const LevelGeom& L = g.lv[lq];
        const int x0 = x - OD_TAPR, xa = x0 & ~15;
        if (tma) {
            // ONE bulk tensor copy per keypoint: box 64 bytes x 39 rows from the 16-byte aligned column xa (the TMA unit faults on an
            // unaligned box start), issued by lane 0, landing on the warp's mbarrier; no LSU wavefronts, no per-lane address math
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the lanes' generic reads of the previous window are done (syncwarp)
                od_mbar_expect_tx(&s_bar[warp], ORB_TAP_BOX_W * ORB_TAP_BOX_H);
                od_tma_load_3d(&s_win[warp][0], &btm->m[lq], xa, y - OD_TAPR, f0 + f, &s_bar[warp]);
            }
            return;
        }
        const int nch = ((x0 + 2 * OD_TAPR) >> 4) - (xa >> 4) + 1;                 // 3 or 4 chunks per row
        const uint8_t* src = blur + L.bbase + (long long)f * L.bframe_stride + (long long)(y - OD_TAPR) * L.bpitch + xa;
        const unsigned dst = (unsigned)__cvta_generic_to_shared(&s_win[warp][0]);
#pragma unroll
        for (int j = 0; j < ((2 * OD_TAPR + 1) * 4 + 31) / 32; ++j) {
            const int item = lane + 32 * j, r = item >> 2, cc = item & 3;
            if (r < 2 * OD_TAPR + 1 && cc < nch)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + r * OD_WPITCH + cc * 16),
                             "l"(src + (long long)r * L.bpitch + cc * 16) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (lv[0] >= 0) stage_window(lv[0], px[0], py[0]);   // flies behind phases 1 and 2
    // ---- phase 1: IC_Angle moments (ORBextractor.cc:77-104) ----
#pragma unroll
    for (int q = 0; q < KPW; ++q) {
        int m10 = 0, m01 = 0;
        if (lv[q] >= 0) {
            const LevelGeom& L = g.lv[lv[q]];
            const int a = (px[q] - ORB_HALF_PATCH) & 3;                       // alignment of the patch's first column
            const uint8_t* p0 = pyr + L.base + (long long)f * L.frame_stride + L.ioff + (py[q] - ORB_HALF_PATCH) * L.pitch +
                                (px[q] - ORB_HALF_PATCH - a);                 // 4-byte aligned
            // item (row, word): 27 lanes take 3 rows x 9 words per step, 11 steps cover rows 0..32 (the table is zero from row 31
            // on and those rows are not loaded); a lane keeps its word column and walks down 3 rows per step, so the loop has no
            // index arithmetic: two loads and two IDP.4A per item (the v factor of m01 sits in the table's second word)
            if (lane < 27) {
                const uint2* tab = mom_tab + a * OD_ITEMS + lane;
                const unsigned pw3 = 3u * ((unsigned)L.pitch >> 2);
                const unsigned* pb = reinterpret_cast<const unsigned*>(p0);   // warp-uniform base + 32-bit word offsets
                unsigned po = od_row * ((unsigned)L.pitch >> 2) + od_word;
#pragma unroll
                for (int j = 0; j < 11; ++j) {
                    if (j < 10 || od_row == 0) {
                        const unsigned pix = __ldg(pb + po);
                        const uint2 w = __ldg(tab + 27 * j);
                        m10 = dp4a_su(w.x, pix, m10);                             // sum u * I
                        m01 = dp4a_su(w.y, pix, m01);                             // sum v * I (weights v inside the circular extent, else 0)
                    }
                    po += pw3;
                }
            }
        }
        m10 = __reduce_add_sync(0xffffffffu, m10);   // redux.sync: one instruction per sum instead of 5 shuffles + 5 adds
        m01 = __reduce_add_sync(0xffffffffu, m01);
        if (lane == 0) { s_m[warp * KPW + q][0] = m01; s_m[warp * KPW + q][1] = m10; }
    }
    __syncthreads();
    // ---- phase 2: angle and steering coefficients, one lane per keypoint ----
    if (warp == 0 && lane < OD_WARPS * KPW) {
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
    for (int q = 0; q < KPW; ++q) {
        if (lv[q] < 0) continue;   // warp-uniform
        const int l = lv[q];
        const LevelGeom& L = g.lv[l];
        const int kq = warp * KPW + q;
        const float a = s_a[kq], b = s_b[kq];
        // taps are addressed with non-negative offsets from the staged window's top-left chunk
        if (tma) { od_mbar_wait(&s_bar[warp], bar_phase); bar_phase ^= 1u; }
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        const uint8_t* b2 = &s_win[warp][0];
        // cvRound (round half to even) of the rotated tap coordinates without a float -> int conversion (F2I runs on the
        // quarter-rate XU pipe): x + 1.5 * 2^23 rounds x to an integer n in the same mode and leaves 0x4B400000 + n in the
        // register; the two biases fold into the window offset (unsigned arithmetic modulo 2^32)
        const float kMagic = 12582912.f;
        const unsigned kBits = 0x4B400000u;
        const unsigned centre = OD_TAPR * wpitch + OD_TAPR + (unsigned)((px[q] - OD_TAPR) & 15) - kBits * (unsigned)wpitch - kBits;
        unsigned val = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float4 pt = s_pat[k * 32 + lane];
            const unsigned r0 = __float_as_uint(__fadd_rn(__fadd_rn(__fmul_rn(pt.x, b), __fmul_rn(pt.y, a)), kMagic));
            const unsigned c0 = __float_as_uint(__fadd_rn(__fsub_rn(__fmul_rn(pt.x, a), __fmul_rn(pt.y, b)), kMagic));
            const unsigned r1 = __float_as_uint(__fadd_rn(__fadd_rn(__fmul_rn(pt.z, b), __fmul_rn(pt.w, a)), kMagic));
            const unsigned c1 = __float_as_uint(__fadd_rn(__fsub_rn(__fmul_rn(pt.z, a), __fmul_rn(pt.w, b)), kMagic));
            const unsigned t0 = b2[centre + r0 * (unsigned)wpitch + c0], t1 = b2[centre + r1 * (unsigned)wpitch + c1];
            val |= (unsigned)(t0 < t1) << k;
        }
        __syncwarp();   // every lane has read the window: the next keypoint's copy may overwrite it
        if (q + 1 < KPW && lv[(q + 1) % KPW] >= 0) stage_window(lv[(q + 1) % KPW], px[(q + 1) % KPW], py[(q + 1) % KPW]);
        // 32-byte descriptor row: gather 4 lanes' bytes into one word, 8 lanes store 8 words (one 32-B sector)
        unsigned w = val;
        w |= __shfl_down_sync(0xffffffffu, val, 1) << 8;
        w |= __shfl_down_sync(0xffffffffu, val, 2) << 16;
        w |= __shfl_down_sync(0xffffffffu, val, 3) << 24;
        const long long o = (long long)f * cap + off[q];
        if ((lane & 3) == 0) reinterpret_cast<unsigned*>(desc_out + o * 32)[lane >> 2] = w;
        {   // the 7 words of the cv::KeyPoint-compatible record, one per lane (every lane holds the warp-uniform fields)
            static_assert(sizeof(orb_kp) == 28, "orb_kp is 7 words");
            const float kx = (l != 0) ? __fmul_rn((float)px[q], L.scale) : (float)px[q];   // ORBextractor.cc:1139-1145
            const float ky = (l != 0) ? __fmul_rn((float)py[q], L.scale) : (float)py[q];
            unsigned wv = __float_as_uint(kx);                        // x
            wv = lane == 1 ? __float_as_uint(ky) : wv;                 // y
            wv = lane == 2 ? __float_as_uint(L.size) : wv;             // size
            wv = lane == 3 ? __float_as_uint(s_ang[kq]) : wv;          // angle
            wv = lane == 4 ? __float_as_uint((float)sc[q]) : wv;       // response
            wv = lane == 5 ? (unsigned)l : wv;                         // octave
            wv = lane == 6 ? 0xFFFFFFFFu : wv;                         // class_id = -1
            if (lane < 7) reinterpret_cast<unsigned*>(kps_out + o)[lane] = wv;
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

// experiment / tuning knob: one shared-memory carve-out for every kernel of the chain (ORB_B200_CARVEOUT, percent of the
// maximum) so that kernels of different chunks can share an SM without the SM draining to re-partition L1 / shared memory
void orb_carveout_extract(int pct) {
    cudaFuncSetAttribute(quadtree_kernel<QT_THREADS>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(quadtree_kernel<QT_THREADS_LAT>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(orient_describe_kernel<OD_KPW>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(orient_describe_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}

int orb_launch_extract(orb_ctx* c, const uint8_t* d_imgs, int pixel_format, int F, int f0, size_t row_stride, size_t frame_stride,
                       orb_kp* d_kps, uint8_t* d_desc, int cap, int* d_n_out, cudaStream_t st) {
    // frames [f0, f0 + F) of the arena: a per-launch copy of the geometry with shifted bases, so the kernels index
    // frames by blockIdx.y alone; d_imgs / d_kps / d_desc / d_n_out already point at the chunk's first frame
    Geometry& g = c->gl;
    g = c->g;
    for (int l = 0; l < g.nlevels; ++l) {
        g.lv[l].base += (long long)f0 * g.lv[l].frame_stride;
        g.lv[l].bbase += (long long)f0 * g.lv[l].bframe_stride;
        g.lv[l].corner_base += (long long)f0 * g.lv[l].corner_cap;
    }
    int* d_cc = c->d_corner_count + (size_t)f0 * g.nlevels;
    int* d_tie = c->d_corner_count + ((size_t)c->max_batch + f0) * g.nlevels;  // second half: tie-at-cut counters
    unsigned long long* d_kept = c->d_kept + (size_t)f0 * g.total_kp_slots;
    int* d_kept_count = c->d_kept_count + (size_t)f0 * g.nlevels;
    cudaEvent_t* ev = nullptr;
    if (c->profile) {
        const int slot = c->prof_head;
        if (c->prof_pending[slot]) { int rc = orb_profile_harvest(c, slot); if (rc != ORB_OK) return rc; }
        ev = c->prof_ev[slot];
        c->prof_frames[slot] = F;
        c->prof_pending[slot] = true;
        c->prof_head = (slot + 1) % ORB_PROF_RING;
    }
    // Two independent chains follow the pyramid interior:  border -> blur  (memory-bound, no shared memory)  and
    // FAST -> quadtree  (ALU-pipe-bound, shared-memory heavy).  With ORB_B200_OVERLAP=1 they run on two streams and
    // meet before the orientation / descriptor kernel; measured gain on B200: none, so the default is aux == st.
    const int which = (st == c->st_c2) ? 1 : 0;
    cudaStream_t aux = (c->overlap && c->st_aux[which]) ? c->st_aux[which] : st;
#define ORB_STAGE_MARK(i, s_) do { if (ev) ORB_CUDA(cudaEventRecord(ev[i], s_)); } while (0)
    ORB_CUDA(cudaMemsetAsync(d_cc, 0, sizeof(int) * (size_t)F * g.nlevels, st));
    ORB_CUDA(cudaMemsetAsync(d_tie, 0, sizeof(int) * (size_t)F * g.nlevels, st));
    // launch of K3 for the levels [l0, l1) on stream s
    auto launch_quadtree = [&](int l0, int l1, cudaStream_t s) -> int {
        if (l1 <= l0) return ORB_OK;
        const size_t smem = (size_t)g.max_node_cap * 80;
        if (smem > 48 * 1024 && !c->qt_attr_set) {   // large nFeatures: opt in to > 48 KB of dynamic shared memory
            ORB_CUDA(cudaFuncSetAttribute(quadtree_kernel<QT_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            ORB_CUDA(cudaFuncSetAttribute(quadtree_kernel<QT_THREADS_LAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            c->qt_attr_set = true;
        }
        if (F >= 8)
            quadtree_kernel<QT_THREADS><<<dim3(F, l1 - l0), QT_THREADS, smem, s>>>(c->d_corners, d_cc, c->d_node_of_key, d_kept,
                                                                                   d_kept_count, d_tie, g, l0);
        else
            quadtree_kernel<QT_THREADS_LAT><<<dim3(F, l1 - l0), QT_THREADS_LAT, smem, s>>>(c->d_corners, d_cc, c->d_node_of_key, d_kept,
                                                                                           d_kept_count, d_tie, g, l0);
        c->launches++;
        return ORB_OK;
    };
    ORB_STAGE_MARK(0, st);
    // Latency shape (a few frames, parallel branches allowed): the upper pyramid levels are one cluster launch of ~20 us
    // that only the FAST / quadtree work of THOSE levels depends on, so FAST + quadtree of the lower levels run next to it
    // on a third stream.
    int tail_first = g.nlevels;
    bool split = false;
    static const bool no_split = getenv("ORB_B200_NO_SPLIT") != nullptr;   // A/B knob
    if (!no_split && aux != st && F < 8 && c->st_aux[1 - which] && pixel_format == ORB_PIX_GRAY8) {
        orb_launch_pyramid(c, g, d_imgs, pixel_format, F, row_stride, frame_stride, st, -1, &tail_first);   // query only (phase -1 launches nothing)
        split = tail_first < g.nlevels;
    }
    if (split) {
        cudaStream_t early = c->st_aux[1 - which];
        if (!c->ev_head) {
            ORB_CUDA(cudaEventCreateWithFlags(&c->ev_head, cudaEventDisableTiming));
            ORB_CUDA(cudaEventCreateWithFlags(&c->ev_early, cudaEventDisableTiming));
        }
        { int rc = orb_launch_pyramid(c, g, d_imgs, pixel_format, F, row_stride, frame_stride, st, 1, nullptr); if (rc != ORB_OK) return rc; }
        ORB_CUDA(cudaEventRecord(c->ev_head, st));
        ORB_CUDA(cudaStreamWaitEvent(early, c->ev_head, 0));
        { int rc = orb_launch_fast(c, g, d_cc, F, f0, early, 0, tail_first); if (rc != ORB_OK) return rc; }
        { int rc = launch_quadtree(0, tail_first, early); if (rc != ORB_OK) return rc; }
        ORB_CUDA(cudaEventRecord(c->ev_early, early));
        { int rc = orb_launch_pyramid(c, g, d_imgs, pixel_format, F, row_stride, frame_stride, st, 2, nullptr); if (rc != ORB_OK) return rc; }
    } else {
        int rc = orb_launch_pyramid(c, g, d_imgs, pixel_format, F, row_stride, frame_stride, st); if (rc != ORB_OK) return rc;   // K1 interior
    }
    ORB_STAGE_MARK(1, st);
    if (aux != st) {
        ORB_CUDA(cudaEventRecord(c->ev_pyr[which], st));
        ORB_CUDA(cudaStreamWaitEvent(aux, c->ev_pyr[which], 0));
    }
    ORB_STAGE_MARK(6, aux);
    ORB_STAGE_MARK(7, aux);   // (the border pass that used to run here is part of the pyramid kernels now)
    { int rc = orb_launch_blur(c, g, F, aux); if (rc != ORB_OK) return rc; }                                        // K5
    ORB_STAGE_MARK(8, aux);
    if (aux != st) ORB_CUDA(cudaEventRecord(c->ev_blur[which], aux));
    ORB_STAGE_MARK(9, st);
    { int rc = orb_launch_fast(c, g, d_cc, F, f0, st, split ? tail_first : 0, g.nlevels); if (rc != ORB_OK) return rc; }   // K2
    ORB_STAGE_MARK(2, st);
    { int rc = launch_quadtree(split ? tail_first : 0, g.nlevels, st); if (rc != ORB_OK) return rc; }                     // K3
    if (split) ORB_CUDA(cudaStreamWaitEvent(st, c->ev_early, 0));
    ORB_STAGE_MARK(3, st);
    if (aux != st) ORB_CUDA(cudaStreamWaitEvent(st, c->ev_blur[which], 0));
    ORB_STAGE_MARK(4, st);
    {   // K4 + K6
        if (F >= 8)
            orient_describe_kernel<OD_KPW><<<dim3((g.total_kp_slots + OD_KPB - 1) / OD_KPB, F), OD_WARPS * 32, 0, st>>>(
                c->d_pyr, c->d_blur, d_kept, d_kept_count, c->d_mom_tab, d_kps, d_desc, cap, d_n_out, g, c->d_btmaps, f0);
        else   // latency shape: one keypoint per warp
            orient_describe_kernel<1><<<dim3((g.total_kp_slots + OD_WARPS - 1) / OD_WARPS, F), OD_WARPS * 32, 0, st>>>(
                c->d_pyr, c->d_blur, d_kept, d_kept_count, c->d_mom_tab, d_kps, d_desc, cap, d_n_out, g, c->d_btmaps, f0);
        c->launches++;
    }
    ORB_STAGE_MARK(5, st);
#undef ORB_STAGE_MARK
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}
