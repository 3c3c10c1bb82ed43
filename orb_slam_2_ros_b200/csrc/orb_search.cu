// orb_search.cu — windowed and brute-force matching with the reference's sequential "already matched" rule.
//
//   orb_search_by_projection : ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th)   ORBmatcher.cc:45-129
//                              ORBmatcher::SearchByProjection(Cur, Last, th, bMono)              ORBmatcher.cc:1330-1472
//                              candidate sets = Frame::GetFeaturesInArea                         Frame.cc:354-412
//                              over Frame::AssignFeaturesToGrid / PosInGrid                      Frame.cc:239-256,415-425
//   orb_match_bruteforce     : SearchByBoW inner loop                                            ORBmatcher.cc:196-252
//
// GPU formulation: everything that is independent per query runs in parallel
// (grid build, window enumeration in the reference's (ix, iy, insertion) order, Hamming distances, per-row
// top-K); the greedy assignment, whose result depends on the order of the queries, is resolved by ONE warp
// that walks the queries in order over those precomputed lists, then applies the rotation-histogram filter.
#include <algorithm>
#include <vector>

#include "orb_internal.cuh"

namespace {

#define GRID_COLS 64   // FRAME_GRID_COLS  (reference Frame.h:38)
#define GRID_ROWS 48   // FRAME_GRID_ROWS  (reference Frame.h:37)
#define HISTO_LENGTH 30
#define GB_MAX_N 8192  // max target keypoints for the in-smem grid sort

__device__ __forceinline__ int dist256(const uint4* a, const uint4* b) {
    const uint4 a0 = a[0], a1 = a[1], b0 = b[0], b1 = b[1];
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// ---- grid build: keys (cell << 16 | index) bitonic-sorted in shared memory => per cell ascending index,
//      which is the insertion order of mGrid[x][y].push_back(i) -----------------------------------------
__global__ void __launch_bounds__(1024)
grid_build_kernel(const orb_kp* __restrict__ kps, int n, int npad, float min_x, float min_y, float inv_w, float inv_h,
                  unsigned* __restrict__ items /*[npad]*/, int* __restrict__ cell_start /*[GRID_COLS*GRID_ROWS+1]*/) {
    extern __shared__ unsigned skey[];
    for (int i = threadIdx.x; i < npad; i += blockDim.x) {
        unsigned key = 0xFFFFFFFFu;
        if (i < n) {
            // PosInGrid (Frame.cc:415-425): round() = half away from zero
            const int px = (int)roundf(__fmul_rn(__fsub_rn(kps[i].x, min_x), inv_w));
            const int py = (int)roundf(__fmul_rn(__fsub_rn(kps[i].y, min_y), inv_h));
            if (px >= 0 && px < GRID_COLS && py >= 0 && py < GRID_ROWS) key = ((unsigned)(px * GRID_ROWS + py) << 16) | (unsigned)i;
        }
        skey[i] = key;
    }
    __syncthreads();
    for (int k = 2; k <= npad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < npad; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned a = skey[i], b = skey[ixj];
                    const bool up = ((i & k) == 0);
                    if ((a > b) == up) { skey[i] = b; skey[ixj] = a; }
                }
            }
            __syncthreads();
        }
    for (int i = threadIdx.x; i < npad; i += blockDim.x) items[i] = skey[i];
    // cell_start[c] = first position with cell >= c
    for (int c = threadIdx.x; c <= GRID_COLS * GRID_ROWS; c += blockDim.x) {
        int lo = 0, hi = npad;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const unsigned key = skey[mid];
            const unsigned cellv = (key == 0xFFFFFFFFu) ? 0xFFFFu : (key >> 16);
            if (cellv < (unsigned)c) lo = mid + 1; else hi = mid;
        }
        cell_start[c] = lo;
    }
}

struct SearchArgs {
    const orb_kp* kps; const uint8_t* desc; const float* u_right; int n;
    int nq; const float *q_u, *q_v, *q_radius; const int *q_min_level, *q_max_level; const uint8_t* q_desc;
    const float *q_ur, *q_er_max, *q_angle; const uint8_t *q_valid, *q_obs;
    float min_x, min_y, inv_w, inv_h;
    const unsigned* items; const int* cell_start;
    int* cand_count; int* cand_base; int* cand_total; int cand_cap;
    int* cand_idx; unsigned short* cand_dist;
};

// ---- one warp per query: GetFeaturesInArea in reference order + distances -----------------------------------
// pass 0 counts, pass 1 (after the warp reserved its segment) writes (index, distance) in candidate order.
__global__ void __launch_bounds__(256)
window_candidates_kernel(const SearchArgs a) {
    const int lane = threadIdx.x & 31;
    const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (qi >= a.nq) return;
    if (lane == 0) { a.cand_count[qi] = 0; a.cand_base[qi] = 0; }
    if (a.q_valid && !a.q_valid[qi]) return;
    const float x = a.q_u[qi], y = a.q_v[qi], r = a.q_radius[qi];
    const int minLevel = a.q_min_level[qi], maxLevel = a.q_max_level[qi];
    // Frame.cc:361-375 (float arithmetic, every op rounded separately)
    const int nMinCellX = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, a.min_x), r), a.inv_w)));
    if (nMinCellX >= GRID_COLS) return;
    const int nMaxCellX = min(GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, a.min_x), r), a.inv_w)));
    if (nMaxCellX < 0) return;
    const int nMinCellY = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, a.min_y), r), a.inv_h)));
    if (nMinCellY >= GRID_ROWS) return;
    const int nMaxCellY = min(GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, a.min_y), r), a.inv_h)));
    if (nMaxCellY < 0) return;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    const uint4* dq = reinterpret_cast<const uint4*>(a.q_desc + (size_t)qi * 32);
    int base = 0, total = 0;
    for (int pass = 0; pass < 2; ++pass) {
        int written = 0;
        for (int ix = nMinCellX; ix <= nMaxCellX; ++ix) {
            // cells (ix, nMinCellY..nMaxCellY) are contiguous in the sorted item array
            const int s = a.cell_start[ix * GRID_ROWS + nMinCellY], e = a.cell_start[ix * GRID_ROWS + nMaxCellY + 1];
            for (int p0 = s; p0 < e; p0 += 32) {
                const int p = p0 + lane;
                bool ok = false;
                int id = -1;
                if (p < e) {
                    id = (int)(a.items[p] & 0xFFFFu);
                    const orb_kp kp = a.kps[id];
                    ok = true;
                    if (bCheckLevels) {
                        if (kp.octave < minLevel) ok = false;
                        if (maxLevel >= 0 && kp.octave > maxLevel) ok = false;
                    }
                    const float dx = __fsub_rn(kp.x, x), dy = __fsub_rn(kp.y, y);
                    if (!(fabsf(dx) < r && fabsf(dy) < r)) ok = false;
                }
                const unsigned m = __ballot_sync(0xffffffffu, ok);
                if (pass == 1 && ok) {
                    const int pos = base + written + __popc(m & ((1u << lane) - 1));
                    int d = dist256(dq, reinterpret_cast<const uint4*>(a.desc + (size_t)id * 32));
                    // the stereo consistency test of the reference loop (ORBmatcher.cc:91-96, 1409-1415) is folded
                    // into the list: a candidate that fails it gets the sentinel distance 0xFFFF
                    if (a.u_right && a.u_right[id] > 0) {
                        const float er = fabsf(__fsub_rn(a.q_ur[qi], a.u_right[id]));
                        if (er > a.q_er_max[qi]) d = 0xFFFF;
                    }
                    a.cand_idx[pos] = id;
                    a.cand_dist[pos] = (unsigned short)d;
                }
                written += __popc(m);
            }
        }
        if (pass == 0) {
            total = written;
            if (total == 0) return;
            if (lane == 0) base = atomicAdd(a.cand_total, total);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base + total > a.cand_cap) {  // arena too small: report, host retries with a larger one
                if (lane == 0) { a.cand_count[qi] = total; a.cand_base[qi] = -1; }
                return;
            }
            if (lane == 0) { a.cand_count[qi] = total; a.cand_base[qi] = base; }
        }
    }
}

__device__ __forceinline__ unsigned warp_min_u32(unsigned v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ void three_maxima(const int* cnt, int& ind1, int& ind2, int& ind3) {  // ORBmatcher.cc:1603-1644
    int max1 = 0, max2 = 0, max3 = 0;
    ind1 = ind2 = ind3 = -1;
    for (int i = 0; i < HISTO_LENGTH; ++i) {
        const int s = cnt[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
    else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { ind3 = -1; }
}

__device__ __forceinline__ int rot_bin(float aq, float at) {  // ORBmatcher.cc:1436-1441 (factor = 1/HISTO_LENGTH)
    float rot = __fsub_rn(aq, at);
    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
    int bin = (int)roundf(__fmul_rn(rot, 1.0f / HISTO_LENGTH));
    if (bin == HISTO_LENGTH) bin = 0;
    return bin;
}

// ---- the order-dependent part: one warp walks the queries in reference order -------------------------------
__global__ void __launch_bounds__(32)
window_resolve_kernel(const SearchArgs a, int mode, int th_dist, float nn_ratio, int check_ori, uint8_t* taken,
                      int* match_of_query, int* target_query, signed char* match_bin, int* nmatches_out) {
    const int lane = threadIdx.x;
    __shared__ int hist[HISTO_LENGTH];
    if (lane < HISTO_LENGTH) hist[lane] = 0;
    for (int i = lane; i < a.n; i += 32) target_query[i] = -1;
    __syncwarp();
    int nmatches = 0;
    for (int qi = 0; qi < a.nq; ++qi) {
        if (lane == 0) { match_of_query[qi] = -1; match_bin[qi] = -1; }
        const int cnt = a.cand_count[qi];
        if (cnt == 0) continue;
        const int base = a.cand_base[qi];
        // best = min over (dist, position) of the candidates that are not taken and passed the stereo test
        unsigned k1 = 0xFFFFFFFFu;
        for (int p = lane; p < cnt; p += 32) {
            const unsigned d = a.cand_dist[base + p];
            if (d < 256u && !taken[a.cand_idx[base + p]]) k1 = min(k1, (d << 16) | (unsigned)p);
        }
        k1 = warp_min_u32(k1);
        if (k1 == 0xFFFFFFFFu) continue;
        const int bestDist = (int)(k1 >> 16), bestPos = (int)(k1 & 0xFFFFu);
        const int bestIdx = a.cand_idx[base + bestPos];
        if (bestDist > th_dist) continue;
        if (mode == ORB_MODE_LOCAL_POINTS) {
            unsigned k2 = 0xFFFFFFFFu;
            for (int p = lane; p < cnt; p += 32) {
                const unsigned d = a.cand_dist[base + p];
                if (p != bestPos && d < 256u && !taken[a.cand_idx[base + p]]) k2 = min(k2, (d << 16) | (unsigned)p);
            }
            k2 = warp_min_u32(k2);
            int bestDist2 = 256, bestLevel2 = -1;
            if (k2 != 0xFFFFFFFFu) { bestDist2 = (int)(k2 >> 16); bestLevel2 = a.kps[a.cand_idx[base + (k2 & 0xFFFFu)]].octave; }
            const int bestLevel = a.kps[bestIdx].octave;
            if (bestLevel == bestLevel2 && (float)bestDist > __fmul_rn(nn_ratio, (float)bestDist2)) continue;  // ORBmatcher.cc:120
        }
        if (lane == 0) {
            match_of_query[qi] = bestIdx;
            target_query[bestIdx] = qi;
            if (!a.q_obs || a.q_obs[qi]) taken[bestIdx] = 1;
            if (mode == ORB_MODE_TRACK_LAST && check_ori) {
                const int bin = rot_bin(a.q_angle[qi], a.kps[bestIdx].angle);
                match_bin[qi] = (signed char)bin;
                hist[bin]++;
            }
        }
        nmatches++;
        __syncwarp();
    }
    __syncwarp();
    if (mode == ORB_MODE_TRACK_LAST && check_ori) {
        int ind1, ind2, ind3;
        three_maxima(hist, ind1, ind2, ind3);
        int removed = 0;
        for (int qi = lane; qi < a.nq; qi += 32) {
            const int bin = match_bin[qi];
            if (bin >= 0 && bin != ind1 && bin != ind2 && bin != ind3) {
                target_query[match_of_query[qi]] = -1;   // ORBmatcher.cc:1462-1466
                removed++;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, o);
        nmatches -= removed;
        __syncwarp();
        // per-query view: a query whose target was nulled loses its match
        for (int qi = lane; qi < a.nq; qi += 32) {
            const int m = match_of_query[qi];
            if (m >= 0 && target_query[m] == -1) match_of_query[qi] = -1;
        }
    }
    if (lane == 0) *nmatches_out = nmatches;
}

// =============================== brute force with mask (SearchByBoW inner loop) ===============================
#define BF_K 4   // unmasked top-K per query kept for the optimistic resolve

// one warp per query row: distances to every target (stored, u16) + sorted top-K packed keys (dist<<16 | j)
__global__ void __launch_bounds__(256)
bf_rows_kernel(const uint8_t* __restrict__ d1, int n1, const uint8_t* __restrict__ d2, int n2,
               unsigned short* __restrict__ D, unsigned* __restrict__ topk) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n1) return;
    const uint4* dq = reinterpret_cast<const uint4*>(d1 + (size_t)i * 32);
    unsigned best[BF_K];
#pragma unroll
    for (int k = 0; k < BF_K; ++k) best[k] = 0xFFFFFFFFu;
    auto insert = [&](unsigned key) {
#pragma unroll
        for (int k = 0; k < BF_K; ++k) {
            const unsigned lo = min(best[k], key);
            key = max(best[k], key);
            best[k] = lo;
        }
    };
    for (int j = lane; j < n2; j += 32) {
        const int d = dist256(dq, reinterpret_cast<const uint4*>(d2 + (size_t)j * 32));
        D[(size_t)i * n2 + j] = (unsigned short)d;
        insert(((unsigned)d << 16) | (unsigned)j);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned other[BF_K];
#pragma unroll
        for (int k = 0; k < BF_K; ++k) other[k] = __shfl_xor_sync(0xffffffffu, best[k], o);
#pragma unroll
        for (int k = 0; k < BF_K; ++k) insert(other[k]);
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < BF_K; ++k) topk[(size_t)i * BF_K + k] = best[k];
    }
}

__global__ void __launch_bounds__(32)
bf_resolve_kernel(const unsigned short* __restrict__ D, const unsigned* __restrict__ topk, int n1, int n2,
                  const float* __restrict__ angle1, const float* __restrict__ angle2, int th_dist, float nn_ratio,
                  int check_ori, int* owner /*[n2]*/, int* match12, signed char* match_bin, int* nmatches_out) {
    const int lane = threadIdx.x;
    __shared__ int hist[HISTO_LENGTH];
    if (lane < HISTO_LENGTH) hist[lane] = 0;
    for (int j = lane; j < n2; j += 32) owner[j] = -1;
    __syncwarp();
    int nmatches = 0;
    for (int i = 0; i < n1; ++i) {
        if (lane == 0) { match12[i] = -1; match_bin[i] = -1; }
        // optimistic: the first two untaken entries of the sorted unmasked top-K are the masked best / second
        unsigned key = 0xFFFFFFFFu;
        bool free_ = false;
        if (lane < BF_K) {
            key = topk[(size_t)i * BF_K + lane];
            free_ = (key != 0xFFFFFFFFu) && (owner[key & 0xFFFFu] < 0);
        }
        const unsigned fm = __ballot_sync(0xffffffffu, free_);
        const unsigned present = __ballot_sync(0xffffffffu, key != 0xFFFFFFFFu);
        unsigned k1, k2;
        if (__popc(fm) >= 2 || __popc(present) < BF_K) {
            // enough free entries, or the list holds ALL targets (n2 < K): exact
            const int p1 = fm ? __ffs(fm) - 1 : -1;
            const unsigned fm2 = fm & (fm - 1);
            const int p2 = fm2 ? __ffs(fm2) - 1 : -1;
            k1 = p1 >= 0 ? __shfl_sync(0xffffffffu, key, p1) : 0xFFFFFFFFu;
            k2 = p2 >= 0 ? __shfl_sync(0xffffffffu, key, p2) : 0xFFFFFFFFu;
        } else {
            // fallback: masked scan of the whole row
            unsigned a1 = 0xFFFFFFFFu, a2 = 0xFFFFFFFFu;
            for (int j = lane; j < n2; j += 32) {
                if (owner[j] >= 0) continue;
                const unsigned k = ((unsigned)D[(size_t)i * n2 + j] << 16) | (unsigned)j;
                a2 = min(a2, max(k, a1));
                a1 = min(a1, k);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned b1 = __shfl_xor_sync(0xffffffffu, a1, o), b2 = __shfl_xor_sync(0xffffffffu, a2, o);
                const unsigned lo = min(a1, b1), hi = max(a1, b1);
                a2 = min(hi, min(a2, b2));
                a1 = lo;
            }
            k1 = a1; k2 = a2;
        }
        if (k1 == 0xFFFFFFFFu) continue;
        const int best1 = (int)(k1 >> 16), bestIdx = (int)(k1 & 0xFFFFu);
        const int best2 = (k2 == 0xFFFFFFFFu) ? 256 : (int)(k2 >> 16);
        if (best1 >= 256) continue;
        if (best1 <= th_dist && (float)best1 < __fmul_rn(nn_ratio, (float)best2)) {   // ORBmatcher.cc:229-231
            if (lane == 0) {
                owner[bestIdx] = i;
                match12[i] = bestIdx;
                if (check_ori) {
                    const int bin = rot_bin(angle1[i], angle2[bestIdx]);
                    match_bin[i] = (signed char)bin;
                    hist[bin]++;
                }
            }
            nmatches++;
        }
        __syncwarp();
    }
    __syncwarp();
    if (check_ori) {
        int ind1, ind2, ind3;
        three_maxima(hist, ind1, ind2, ind3);
        int removed = 0;
        for (int i = lane; i < n1; i += 32) {
            const int bin = match_bin[i];
            if (bin >= 0 && bin != ind1 && bin != ind2 && bin != ind3) { owner[match12[i]] = -1; match12[i] = -1; removed++; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, o);
        nmatches -= removed;
    }
    if (lane == 0) *nmatches_out = nmatches;
}

// small RAII arena for the one-shot host-pointer entry points
struct DevBuf {
    std::vector<void*> ptrs;
    ~DevBuf() { for (void* p : ptrs) cudaFree(p); }
    template <typename T> T* alloc(size_t n) {
        void* p = nullptr;
        if (cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)) != cudaSuccess) return nullptr;
        ptrs.push_back(p);
        return (T*)p;
    }
    template <typename T> T* upload(const T* h, size_t n, cudaStream_t st) {
        if (!h) return nullptr;
        T* d = alloc<T>(n);
        if (d && n) cudaMemcpyAsync(d, h, n * sizeof(T), cudaMemcpyHostToDevice, st);
        return d;
    }
};

}  // namespace

extern "C" {

int orb_search_by_projection(int device, const orb_search_params* prm, const orb_kp* kps_un, const uint8_t* desc,
                             const float* u_right, int n, uint8_t* taken, int nq, const float* q_u, const float* q_v,
                             const float* q_radius, const int32_t* q_min_level, const int32_t* q_max_level,
                             const uint8_t* q_desc, const float* q_ur, const float* q_er_max, const float* q_angle,
                             const uint8_t* q_valid, const uint8_t* q_obs, int32_t* match_of_query, int32_t* target_query,
                             int* nmatches) {
    if (!prm || n < 0 || nq < 0 || !nmatches) return ORB_ERR_INVALID;
    *nmatches = 0;
    if (nq && (!q_u || !q_v || !q_radius || !q_min_level || !q_max_level || !q_desc || !match_of_query)) return ORB_ERR_INVALID;
    if (n && (!kps_un || !desc || !taken)) return ORB_ERR_INVALID;
    if (u_right && (!q_ur || !q_er_max)) return ORB_ERR_INVALID;
    if (prm->mode == ORB_MODE_TRACK_LAST && prm->check_orientation && nq && !q_angle) return ORB_ERR_INVALID;
    for (int i = 0; i < nq; ++i) match_of_query[i] = -1;
    if (target_query) for (int i = 0; i < n; ++i) target_query[i] = -1;
    if (n == 0 || nq == 0) return ORB_OK;
    if (n > GB_MAX_N) { orb_set_error("orb_search_by_projection: more than %d target keypoints", GB_MAX_N); return ORB_ERR_CAPACITY; }
    if (orb_device_count() <= 0) { orb_set_error("no CUDA device visible: liborb_b200 has no CPU fallback"); return ORB_ERR_NO_DEVICE; }
    ORB_CUDA(cudaSetDevice(device));
    cudaStream_t st = nullptr;  // legacy default stream: this one-shot entry point is synchronous anyway
    DevBuf B;
    SearchArgs a;
    memset(&a, 0, sizeof(a));
    a.n = n; a.nq = nq;
    a.kps = B.upload(kps_un, n, st); a.desc = B.upload(desc, (size_t)n * 32, st); a.u_right = B.upload(u_right, n, st);
    a.q_u = B.upload(q_u, nq, st); a.q_v = B.upload(q_v, nq, st); a.q_radius = B.upload(q_radius, nq, st);
    a.q_min_level = B.upload(q_min_level, nq, st); a.q_max_level = B.upload(q_max_level, nq, st);
    a.q_desc = B.upload(q_desc, (size_t)nq * 32, st);
    a.q_ur = B.upload(q_ur, nq, st); a.q_er_max = B.upload(q_er_max, nq, st); a.q_angle = B.upload(q_angle, nq, st);
    a.q_valid = B.upload(q_valid, nq, st); a.q_obs = B.upload(q_obs, nq, st);
    uint8_t* d_taken = B.upload(taken, n, st);
    a.min_x = prm->min_x; a.min_y = prm->min_y;
    a.inv_w = (float)GRID_COLS / (prm->max_x - prm->min_x);   // Frame.cc:162-163
    a.inv_h = (float)GRID_ROWS / (prm->max_y - prm->min_y);
    int npad = 32;
    while (npad < n) npad <<= 1;
    unsigned* d_items = B.alloc<unsigned>(npad);
    int* d_cell_start = B.alloc<int>(GRID_COLS * GRID_ROWS + 1);
    int* d_count = B.alloc<int>(nq); int* d_base = B.alloc<int>(nq); int* d_total = B.alloc<int>(2);
    int* d_moq = B.alloc<int>(nq); int* d_tq = B.alloc<int>(n); signed char* d_bin = B.alloc<signed char>(nq);
    if (!a.kps || !a.desc || !d_taken || !d_items || !d_cell_start || !d_count || !d_base || !d_total || !d_moq || !d_tq || !d_bin) {
        orb_set_error("orb_search_by_projection: device allocation failed"); cudaGetLastError(); return ORB_ERR_CUDA;
    }
    a.items = d_items; a.cell_start = d_cell_start; a.cand_count = d_count; a.cand_base = d_base; a.cand_total = d_total;
    grid_build_kernel<<<1, 1024, npad * sizeof(unsigned), st>>>(a.kps, n, npad, a.min_x, a.min_y, a.inv_w, a.inv_h, d_items, d_cell_start);
    int cand_cap = std::max(nq * 128, 4096);
    for (int attempt = 0; attempt < 2; ++attempt) {
        a.cand_cap = cand_cap;
        a.cand_idx = B.alloc<int>(cand_cap);
        a.cand_dist = B.alloc<unsigned short>(cand_cap);
        if (!a.cand_idx || !a.cand_dist) { orb_set_error("orb_search_by_projection: candidate arena allocation failed"); cudaGetLastError(); return ORB_ERR_CUDA; }
        ORB_CUDA(cudaMemsetAsync(d_total, 0, sizeof(int) * 2, st));
        window_candidates_kernel<<<(nq + 7) / 8, 256, 0, st>>>(a);
        int total = 0;
        ORB_CUDA(cudaMemcpyAsync(&total, d_total, sizeof(int), cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaStreamSynchronize(st));
        if (total <= cand_cap) break;
        if (attempt == 1) { orb_set_error("orb_search_by_projection: candidate arena overflow"); return ORB_ERR_CAPACITY; }
        cand_cap = total;
    }
    window_resolve_kernel<<<1, 32, 0, st>>>(a, prm->mode, prm->th_dist, prm->nn_ratio, prm->check_orientation, d_taken, d_moq, d_tq,
                                           d_bin, d_total + 1);
    ORB_CUDA(cudaGetLastError());
    ORB_CUDA(cudaMemcpyAsync(match_of_query, d_moq, sizeof(int) * nq, cudaMemcpyDeviceToHost, st));
    if (target_query) ORB_CUDA(cudaMemcpyAsync(target_query, d_tq, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaMemcpyAsync(taken, d_taken, n, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaMemcpyAsync(nmatches, d_total + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaStreamSynchronize(st));
    return ORB_OK;
}

int orb_match_bruteforce(int device, const uint8_t* desc1, const float* angle1, int n1, const uint8_t* desc2,
                         const float* angle2, int n2, int th_dist, float nn_ratio, int check_orientation, int32_t* match12,
                         int* nmatches) {
    if (n1 < 0 || n2 < 0 || !nmatches || (n1 && (!desc1 || !match12)) || (n2 && !desc2)) return ORB_ERR_INVALID;
    if (check_orientation && ((n1 && !angle1) || (n2 && !angle2))) return ORB_ERR_INVALID;
    *nmatches = 0;
    for (int i = 0; i < n1; ++i) match12[i] = -1;
    if (n1 == 0 || n2 == 0) return ORB_OK;
    if (n2 > 65535) { orb_set_error("orb_match_bruteforce: more than 65535 targets"); return ORB_ERR_CAPACITY; }
    if (orb_device_count() <= 0) { orb_set_error("no CUDA device visible: liborb_b200 has no CPU fallback"); return ORB_ERR_NO_DEVICE; }
    ORB_CUDA(cudaSetDevice(device));
    cudaStream_t st = nullptr;
    DevBuf B;
    const uint8_t* d1 = B.upload(desc1, (size_t)n1 * 32, st);
    const uint8_t* d2 = B.upload(desc2, (size_t)n2 * 32, st);
    const float* a1 = B.upload(angle1, n1, st);
    const float* a2 = B.upload(angle2, n2, st);
    unsigned short* D = B.alloc<unsigned short>((size_t)n1 * n2);
    unsigned* topk = B.alloc<unsigned>((size_t)n1 * BF_K);
    int* owner = B.alloc<int>(n2); int* m12 = B.alloc<int>(n1); signed char* bin = B.alloc<signed char>(n1); int* d_nm = B.alloc<int>(1);
    if (!d1 || !d2 || !D || !topk || !owner || !m12 || !bin || !d_nm) { orb_set_error("orb_match_bruteforce: device allocation failed"); cudaGetLastError(); return ORB_ERR_CUDA; }
    bf_rows_kernel<<<(n1 + 7) / 8, 256, 0, st>>>(d1, n1, d2, n2, D, topk);
    bf_resolve_kernel<<<1, 32, 0, st>>>(D, topk, n1, n2, a1, a2, th_dist, nn_ratio, check_orientation, owner, m12, bin, d_nm);
    ORB_CUDA(cudaGetLastError());
    ORB_CUDA(cudaMemcpyAsync(match12, m12, sizeof(int) * n1, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaMemcpyAsync(nmatches, d_nm, sizeof(int), cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaStreamSynchronize(st));
    return ORB_OK;
}

}  // extern "C"
