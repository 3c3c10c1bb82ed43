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
#include <string.h>

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
__device__ __forceinline__ void grid_build_body(const orb_kp* __restrict__ kps, int n, int npad, float min_x, float min_y, float inv_w, float inv_h,
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

#define SR_K 8   // unmasked top-K per query kept for the optimistic resolve
#define WC_STAGE 160   // candidates of one window staged in shared memory by window_candidates_body (per warp)
#define SR_CH 1024  // queries staged in shared memory per chunk of the sequential walk
#ifndef SR_THREADS
#define SR_THREADS 1024
#endif
#define BF_NONE 0x7FFFFFFF   // owner of a target nobody has taken

struct SearchArgs {
    const orb_kp* kps; const uint8_t* desc; const float* u_right; int n;
    int nq; const float *q_u, *q_v, *q_radius; const int *q_min_level, *q_max_level; const uint8_t* q_desc;
    const float *q_ur, *q_er_max, *q_angle; const uint8_t *q_valid, *q_obs;
    float min_x, min_y, inv_w, inv_h;
    const unsigned* items; const int* cell_start;
    int* cand_count; int* cand_base; int* cand_total; int cand_cap;
    int* cand_idx; unsigned short* cand_dist;
    unsigned* topk;   // [nq][SR_K] sorted keys (dist << 16 | position in the candidate list), 0xFFFFFFFF = none
};


// the SR_K smallest keys of a warp, given every lane's ascending list best[0..SR_K): SR_K rounds of "warp minimum of the
// list heads, the owning lane pops" (keys are unique: they carry a position).  Lane k returns the k-th smallest.
__device__ __forceinline__ unsigned warp_topk_extract(unsigned (&best)[SR_K], int lane) {
    unsigned mine = 0xFFFFFFFFu;
#pragma unroll
    for (int k = 0; k < SR_K; ++k) {
        const unsigned m = __reduce_min_sync(0xffffffffu, best[0]);   // redux.sync: one instruction instead of 5 shuffles + 5 min
        if (lane == k) mine = m;
        if (best[0] == m && m != 0xFFFFFFFFu) {
#pragma unroll
            for (int i = 0; i + 1 < SR_K; ++i) best[i] = best[i + 1];
            best[SR_K - 1] = 0xFFFFFFFFu;
        }
    }
    return mine;
}

// ---- one warp per query: GetFeaturesInArea in reference order + distances -----------------------------------
// pass 0 counts, pass 1 (after the warp reserved its segment) writes (index, distance) in candidate order.
__device__ __forceinline__ void window_candidates_body(const SearchArgs& a) {
    const int lane = threadIdx.x & 31;
    const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (qi >= a.nq) return;
    if (lane == 0) { a.cand_count[qi] = 0; a.cand_base[qi] = 0; }
    if (lane < SR_K) a.topk[(size_t)qi * SR_K + lane] = 0xFFFFFFFFu;
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
    unsigned best[SR_K];
#pragma unroll
    for (int k = 0; k < SR_K; ++k) best[k] = 0xFFFFFFFFu;
    auto insert = [&](unsigned key) {
#pragma unroll
        for (int k = 0; k < SR_K; ++k) {
            const unsigned lo = min(best[k], key);
            key = max(best[k], key);
            best[k] = lo;
        }
    };
    // One walk over the window's cells: the candidates (index, distance) are staged in a per-warp shared-memory buffer, the
    // arena segment is reserved once their number is known, and the staged entries are copied out.  (The first version walked
    // the cells twice — count, then write — and gathered every keypoint record twice in a latency-bound kernel.)  A window with
    // more than WC_STAGE candidates takes the second walk after all, writing straight to the arena.
    __shared__ unsigned s_stage[8][WC_STAGE];
    unsigned* stage = s_stage[threadIdx.x >> 5];
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
                const int lpos = written + __popc(m & ((1u << lane) - 1));   // position in candidate order
                if (ok && (pass == 1 || lpos < WC_STAGE)) {
                    int d = dist256(dq, reinterpret_cast<const uint4*>(a.desc + (size_t)id * 32));
                    // the stereo consistency test of the reference loop (ORBmatcher.cc:91-96, 1409-1415) is folded
                    // into the list: a candidate that fails it gets the sentinel distance 0xFFFF
                    if (a.u_right && a.u_right[id] > 0) {
                        const float er = fabsf(__fsub_rn(a.q_ur[qi], a.u_right[id]));
                        if (er > a.q_er_max[qi]) d = 0xFFFF;
                    }
                    if (pass == 0) {
                        stage[lpos] = (unsigned)id | ((unsigned)d << 16);
                    } else {
                        a.cand_idx[base + lpos] = id;
                        a.cand_dist[base + lpos] = (unsigned short)d;
                    }
                    if (d < 256) { const unsigned key = ((unsigned)d << 16) | (unsigned)lpos; if (key < best[SR_K - 1]) insert(key); }
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
            if (total <= WC_STAGE) {          // the usual case: copy the staged list out, done
                __syncwarp();
                for (int k = lane; k < total; k += 32) {
                    const unsigned e = stage[k];
                    a.cand_idx[base + k] = (int)(e & 0xFFFFu);
                    a.cand_dist[base + k] = (unsigned short)(e >> 16);
                }
                break;
            }
#pragma unroll
            for (int k = 0; k < SR_K; ++k) best[k] = 0xFFFFFFFFu;   // the second walk rebuilds the list over all candidates
        }
    }
    // warp merge of the per-lane sorted top-K lists
    const unsigned mine = warp_topk_extract(best, lane);
    if (lane < SR_K) a.topk[(size_t)qi * SR_K + lane] = mine;
}

__device__ __forceinline__ unsigned warp_min_u32(unsigned v) { return __reduce_min_sync(0xffffffffu, v); }

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

// ---- the order-dependent part: the queries are walked in reference order -----------------------------------------
// One CTA.  Per chunk of SR_CH queries all threads stage the sorted top-K lists (target index, distance, octave)
// in shared memory; then warp 0 walks the chunk sequentially — every lane executes the same broadcast shared-memory
// reads, lane 0 commits — so one query costs a few dependent shared-memory accesses instead of global round trips.
// `taken` lives in shared memory.  The top-K list answers a query whenever enough of its entries are still free;
// otherwise (list truncated and too many entries taken) warp 0 scans the query's full candidate list.
struct SrStage {
    unsigned short id[SR_CH][SR_K];   // target index, 0xFFFF = none
    uint8_t dist[SR_CH][SR_K];
    uint8_t oct[SR_CH][SR_K];
    uint8_t flags[SR_CH];             // bit 0: list truncated (all K entries valid), bit 1: query blocks its target
};

__device__ __forceinline__ void window_resolve_body(const SearchArgs& a, int mode, int th_dist, float nn_ratio, int check_ori, uint8_t* taken_g,
                                                    int* match_of_query, int* target_query, signed char* match_bin, int* assigned, int* nmatches_out,
                                                    int* overflow, int smem_bytes) {
    extern __shared__ __align__(16) uint8_t sr_smem[];
    SrStage& S = *reinterpret_cast<SrStage*>(sr_smem);
    const bool init = (mode == ORB_MODE_INITIALIZATION);
    // shared-memory layout: INITIALIZATION = stage | taken[n] | vMatchedDistance[n] | vnMatches21[n];
    // the other modes (parallel rounds, below) = own_prev[n] | own_new[n] (int) | taken[n]
    int* own_prev = reinterpret_cast<int*>(sr_smem);
    int* own_new = own_prev + a.n;
    uint8_t* taken = init ? sr_smem + sizeof(SrStage) : reinterpret_cast<uint8_t*>(own_new + a.n);   // [n]
    // SearchForInitialization state (ORBmatcher.cc:416-417): vMatchedDistance and vnMatches21, 16 bit each
    const int npad = (a.n + 3) & ~3;
    unsigned short* vmd = reinterpret_cast<unsigned short*>(taken + npad);      // 0xFFFF = INT_MAX
    unsigned short* own = vmd + npad;                                           // 0xFFFF = -1
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ int hist[HISTO_LENGTH];
    __shared__ int s_nmatches;
    if (*a.cand_total > a.cand_cap) {    // candidate arena too small: the host retries with a larger one
        if (threadIdx.x == 0) *overflow = *a.cand_total;
        return;
    }
    if (threadIdx.x == 0) { *overflow = 0; s_nmatches = 0; }
    if (threadIdx.x < HISTO_LENGTH) hist[threadIdx.x] = 0;
    for (int i = threadIdx.x; i < a.n; i += SR_THREADS) {
        target_query[i] = -1; taken[i] = init ? 0 : taken_g[i];
        if (init) { vmd[i] = 0xFFFFu; own[i] = 0xFFFFu; }
    }
    const int need = (mode == ORB_MODE_TRACK_LAST) ? 1 : 2;
    int nmatches = 0;
    if (!init) {
        // ---- TRACK_LAST / LOCAL_POINTS: the "already matched" rule (ORBmatcher.cc:87-89, 1405-1407) as a parallel
        // fixed-point iteration (see bf_resolve_kernel): a target is invisible to query qi when it was taken before the
        // call or is owned by an earlier BLOCKING query; own[t] = min { qi blocking : choice[qi] == t } is rebuilt every
        // round with atomicMin until no choice changes ----
        __shared__ int s_changed;
        for (int qi = threadIdx.x; qi < a.nq; qi += SR_THREADS) { match_of_query[qi] = -1; match_bin[qi] = -1; }
        for (int i = threadIdx.x; i < a.n; i += SR_THREADS) own_prev[i] = BF_NONE;
        const int nq_up = (a.nq + 31) & ~31;
        for (int round = 0; round <= a.nq; ++round) {
            for (int i = threadIdx.x; i < a.n; i += SR_THREADS) own_new[i] = BF_NONE;
            if (threadIdx.x == 0) s_changed = 0;
            __syncthreads();
            for (int qi = threadIdx.x; qi < nq_up; qi += SR_THREADS) {
                int d1 = 256, d2 = 256, i1 = -1, i2 = -1, nfree = 0, base = 0;
                bool fallback = false;
                if (qi < a.nq) {
                    base = a.cand_base[qi];
                    unsigned key[SR_K];
#pragma unroll
                    for (int k = 0; k < SR_K; ++k) key[k] = a.topk[(size_t)qi * SR_K + k];
#pragma unroll
                    for (int k = 0; k < SR_K; ++k) {
                        if (key[k] == 0xFFFFFFFFu || nfree >= need) break;
                        const int id = a.cand_idx[base + (key[k] & 0xFFFFu)];
                        if (taken[id] != 0 || own_prev[id] < qi) continue;
                        if (nfree == 0) { d1 = (int)(key[k] >> 16); i1 = id; } else { d2 = (int)(key[k] >> 16); i2 = id; }
                        ++nfree;
                    }
                    fallback = nfree < need && key[SR_K - 1] != 0xFFFFFFFFu;   // truncated list, too many entries taken
                }
                unsigned todo = __ballot_sync(0xffffffffu, fallback);
                while (todo) {   // the warp scans the full candidate list of each such query
                    const int src = __ffs((int)todo) - 1;
                    todo &= todo - 1u;
                    const int fq = __shfl_sync(0xffffffffu, qi, src);
                    const int cnt = a.cand_count[fq], fbase = a.cand_base[fq];
                    unsigned a1 = 0xFFFFFFFFu, a2 = 0xFFFFFFFFu;
                    for (int p = lane; p < cnt; p += 32) {
                        const unsigned d = a.cand_dist[fbase + p];
                        if (d >= 256u) continue;
                        const int id = a.cand_idx[fbase + p];
                        if (taken[id] != 0 || own_prev[id] < fq) continue;
                        const unsigned k = (d << 16) | (unsigned)p;
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
                    if (lane == src) {
                        d1 = d2 = 256; i1 = i2 = -1;
                        if (a1 != 0xFFFFFFFFu) { d1 = (int)(a1 >> 16); i1 = a.cand_idx[fbase + (a1 & 0xFFFFu)]; }
                        if (a2 != 0xFFFFFFFFu) { d2 = (int)(a2 >> 16); i2 = a.cand_idx[fbase + (a2 & 0xFFFFu)]; }
                    }
                }
                if (qi < a.nq) {
                    int c = -1;
                    if (i1 >= 0 && d1 <= th_dist) {
                        c = i1;
                        if (mode == ORB_MODE_LOCAL_POINTS) {
                            const int o1 = a.kps[i1].octave, bestLevel2 = (d2 < 256 && i2 >= 0) ? a.kps[i2].octave : -1;
                            if (o1 == bestLevel2 && (float)d1 > __fmul_rn(nn_ratio, (float)d2)) c = -1;   // ORBmatcher.cc:120
                        }
                    }
                    if (c != match_of_query[qi]) { match_of_query[qi] = c; s_changed = 1; }
                    if (c >= 0 && (!a.q_obs || a.q_obs[qi])) atomicMin(own_new + c, qi);   // only such queries block their target
                }
            }
            __syncthreads();
            const int changed = s_changed;
            int* t = own_prev; own_prev = own_new; own_new = t;
            __syncthreads();
            if (!changed) break;
        }
        // final state: taken, target_query (the LAST query that matched the target, as the sequential overwrite leaves it)
        for (int i = threadIdx.x; i < a.n; i += SR_THREADS) if (own_prev[i] != BF_NONE) taken[i] = 1;
        for (int qi = threadIdx.x; qi < a.nq; qi += SR_THREADS) {
            const int m = match_of_query[qi];
            if (m >= 0) { atomicMax(target_query + m, qi); nmatches++; }
        }
        atomicAdd(&s_nmatches, nmatches);
        __syncthreads();
        nmatches = s_nmatches;
        __syncthreads();
    }
    // ---- INITIALIZATION (ORBmatcher.cc:406-521) in parallel rounds.  A target can be taken again by a later query with a
    // strictly smaller distance (:445, 464-471), so the state of a target is its whole list of takers; the distances of the
    // successive takers decrease, i.e. vMatchedDistance seen by query qi = min { d_j : j < qi took the target }.  The
    // sequential result is again the unique solution of choice[qi] = decide(qi, takers before qi): every round rebuilds the
    // per-target taker lists (up to INIT_C entries, d << 16 | query) from the previous choices and re-evaluates all queries;
    // a target with more takers than that, or too many targets for the shared-memory lists, falls back to the walk below.
    bool init_done = false;
    constexpr int INIT_C = 4;
    if (init && a.nq < 65535 && (size_t)a.n * (4 * INIT_C + 4) <= (size_t)smem_bytes) {
        unsigned* lst = reinterpret_cast<unsigned*>(sr_smem);           // [n][INIT_C]
        int* cnt = reinterpret_cast<int*>(lst + (size_t)a.n * INIT_C);  // [n]
        __shared__ int s_chg, s_over;
        __syncthreads();                                                // the set-up loop above wrote arrays these alias
        for (int qi = threadIdx.x; qi < a.nq; qi += SR_THREADS) { assigned[qi] = -1; match_of_query[qi] = -1; match_bin[qi] = -1; }
        const int nq_up = (a.nq + 31) & ~31;
        auto visible = [&](int id, int dist, int qi) -> bool {          // no earlier taker matched the target at least as well
            const int c = min(cnt[id], INIT_C);
            for (int e = 0; e < c; ++e) {
                const unsigned v = lst[id * INIT_C + e];
                if ((int)(v & 0xFFFFu) < qi && (int)(v >> 16) <= dist) return false;
            }
            return true;
        };
        bool over = false;
        for (int round = 0; round <= a.nq; ++round) {
            for (int i = threadIdx.x; i < a.n; i += SR_THREADS) cnt[i] = 0;
            if (threadIdx.x == 0) { s_chg = 0; s_over = 0; }
            __syncthreads();
            for (int qi = threadIdx.x; qi < a.nq; qi += SR_THREADS) {
                const int v = assigned[qi];                              // d << 16 | target, or -1
                if (v >= 0) {
                    const int slot = atomicAdd(&cnt[v & 0xFFFF], 1);
                    if (slot < INIT_C) lst[(v & 0xFFFF) * INIT_C + slot] = ((unsigned)(v >> 16) << 16) | (unsigned)qi;
                    else s_over = 1;
                }
            }
            __syncthreads();
            if (s_over) { over = true; break; }                           // uniform
            for (int qi = threadIdx.x; qi < nq_up; qi += SR_THREADS) {
                int d1 = 256, d2 = 256, i1 = -1, nfree = 0;
                bool fallback = false;
                if (qi < a.nq) {
                    const int base = a.cand_base[qi];
                    unsigned key[SR_K];
#pragma unroll
                    for (int k = 0; k < SR_K; ++k) key[k] = a.topk[(size_t)qi * SR_K + k];
#pragma unroll
                    for (int k = 0; k < SR_K; ++k) {
                        if (key[k] == 0xFFFFFFFFu || nfree >= 2) break;
                        const int id = a.cand_idx[base + (key[k] & 0xFFFFu)], dist = (int)(key[k] >> 16);
                        if (!visible(id, dist, qi)) continue;
                        if (nfree == 0) { d1 = dist; i1 = id; } else d2 = dist;
                        ++nfree;
                    }
                    fallback = nfree < 2 && key[SR_K - 1] != 0xFFFFFFFFu;
                }
                unsigned todo = __ballot_sync(0xffffffffu, fallback);
                while (todo) {   // truncated list with too few visible entries: the warp scans the query's whole candidate list
                    const int src = __ffs((int)todo) - 1;
                    todo &= todo - 1u;
                    const int fq = __shfl_sync(0xffffffffu, qi, src);
                    const int fcnt = a.cand_count[fq], fbase = a.cand_base[fq];
                    unsigned a1 = 0xFFFFFFFFu, a2 = 0xFFFFFFFFu;
                    for (int p = lane; p < fcnt; p += 32) {
                        const unsigned d = a.cand_dist[fbase + p];
                        if (d >= 256u) continue;
                        if (!visible(a.cand_idx[fbase + p], (int)d, fq)) continue;
                        const unsigned k = (d << 16) | (unsigned)p;
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
                    if (lane == src) {
                        d1 = d2 = 256; i1 = -1;
                        if (a1 != 0xFFFFFFFFu) { d1 = (int)(a1 >> 16); i1 = a.cand_idx[fbase + (a1 & 0xFFFFu)]; }
                        if (a2 != 0xFFFFFFFFu) d2 = (int)(a2 >> 16);
                    }
                }
                if (qi < a.nq) {
                    int c = -1;
                    // bestDist <= TH_LOW and bestDist < (float)bestDist2 * mfNNratio, bestDist2 = INT_MAX without a second (ORBmatcher.cc:460-462)
                    if (i1 >= 0 && d1 <= th_dist && !(d2 < 256 && !((float)d1 < __fmul_rn((float)d2, nn_ratio)))) c = (d1 << 16) | i1;
                    if (c != assigned[qi]) { assigned[qi] = c; s_chg = 1; }
                }
            }
            __syncthreads();
            const int changed = s_chg;
            __syncthreads();
            if (!changed) break;
        }
        if (!over) {
            // the lists of the last round ARE the final takers: vnMatches21 = the latest taker, one match per taken target
            for (int i = threadIdx.x; i < a.n; i += SR_THREADS) {
                const int c = min(cnt[i], INIT_C);
                int latest = -1;
                for (int e = 0; e < c; ++e) latest = max(latest, (int)(lst[i * INIT_C + e] & 0xFFFFu));
                target_query[i] = latest;
                nmatches += latest >= 0;
            }
            __syncthreads();
            for (int qi = threadIdx.x; qi < a.nq; qi += SR_THREADS) {
                const int v = assigned[qi];
                const int id = v >= 0 ? (v & 0xFFFF) : -1;
                assigned[qi] = id;                                        // every assignment made, stale ones included (rotation histogram)
                match_of_query[qi] = (id >= 0 && target_query[id] == qi) ? id : -1;
            }
            atomicAdd(&s_nmatches, nmatches);
            __syncthreads();
            nmatches = s_nmatches;
            __syncthreads();
            init_done = true;
        } else {
            __syncthreads();
            nmatches = 0;
        }
    }
    if (init && !init_done) {   // state of the sequential walk (it aliases the lists above)
        for (int i = threadIdx.x; i < a.n; i += SR_THREADS) { target_query[i] = -1; taken[i] = 0; vmd[i] = 0xFFFFu; own[i] = 0xFFFFu; }
    }
    for (int q0 = 0; init && !init_done && q0 < a.nq; q0 += SR_CH) {
        const int nb = min(SR_CH, a.nq - q0);
        __syncthreads();
        // ---- stage the chunk ----
        for (int e = threadIdx.x; e < nb * SR_K; e += SR_THREADS) {
            const int j = e / SR_K, k = e - j * SR_K, qi = q0 + j;
            const unsigned key = a.topk[(size_t)qi * SR_K + k];
            int id = 0xFFFF, oc = 0;
            if (key != 0xFFFFFFFFu) {
                id = a.cand_idx[a.cand_base[qi] + (key & 0xFFFFu)];
                if (mode == ORB_MODE_LOCAL_POINTS) oc = a.kps[id].octave;
            }
            S.id[j][k] = (unsigned short)id; S.dist[j][k] = (uint8_t)(key >> 16); S.oct[j][k] = (uint8_t)oc;
            if (k == SR_K - 1) S.flags[j] = (uint8_t)((key != 0xFFFFFFFFu ? 1 : 0) | ((!a.q_obs || a.q_obs[qi]) ? 2 : 0));
        }
        for (int j = threadIdx.x; j < nb; j += SR_THREADS) { match_of_query[q0 + j] = -1; match_bin[q0 + j] = -1; if (init) assigned[q0 + j] = -1; }
        __syncthreads();
        // ---- sequential walk by warp 0 ----
        if (warp == 0) {
            for (int j = 0; j < nb; ++j) {
                const int qi = q0 + j;
                int d1 = 256, d2 = 256, i1 = -1, o1 = -1, o2 = -1, nfree = 0;
#pragma unroll
                for (int k = 0; k < SR_K; ++k) {
                    const int id = S.id[j][k];
                    if (id == 0xFFFF || nfree >= need) break;
                    // unavailable: already matched (ORBmatcher.cc:87-89, 1405-1407) / matched at least as well (:445)
                    if (init ? (vmd[id] <= S.dist[j][k]) : (taken[id] != 0)) continue;
                    if (nfree == 0) { d1 = S.dist[j][k]; i1 = id; o1 = S.oct[j][k]; } else { d2 = S.dist[j][k]; o2 = S.oct[j][k]; }
                    ++nfree;
                }
                const int flags = S.flags[j];
                if (nfree < need && (flags & 1)) {
                    // the list was truncated and too many of its entries are taken: scan all candidates of the query
                    const int cnt = a.cand_count[qi], base = a.cand_base[qi];
                    unsigned a1 = 0xFFFFFFFFu, a2 = 0xFFFFFFFFu;
                    for (int p = lane; p < cnt; p += 32) {
                        const unsigned d = a.cand_dist[base + p];
                        if (d >= 256u) continue;
                        const int id = a.cand_idx[base + p];
                        if (init ? (vmd[id] <= d) : (taken[id] != 0)) continue;
                        const unsigned k = (d << 16) | (unsigned)p;
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
                    d1 = d2 = 256; i1 = -1;
                    if (a1 != 0xFFFFFFFFu) { d1 = (int)(a1 >> 16); i1 = a.cand_idx[base + (a1 & 0xFFFFu)]; }
                    int i2 = -1;
                    if (a2 != 0xFFFFFFFFu) { d2 = (int)(a2 >> 16); i2 = a.cand_idx[base + (a2 & 0xFFFFu)]; }
                    if (mode == ORB_MODE_LOCAL_POINTS) { o1 = i1 >= 0 ? a.kps[i1].octave : -1; o2 = i2 >= 0 ? a.kps[i2].octave : -1; }
                }
                if (i1 < 0 || d1 > th_dist) continue;
                if (mode == ORB_MODE_LOCAL_POINTS) {
                    const int bestLevel2 = (d2 < 256) ? o2 : -1;
                    if (o1 == bestLevel2 && (float)d1 > __fmul_rn(nn_ratio, (float)d2)) continue;   // ORBmatcher.cc:120
                }
                if (init) {
                    // bestDist < (float)bestDist2 * mfNNratio with bestDist2 = INT_MAX when there is no second (ORBmatcher.cc:462)
                    if (d2 < 256 && !((float)d1 < __fmul_rn((float)d2, nn_ratio))) continue;
                    const int prev = own[i1];
                    if (lane == 0) {
                        if (prev != 0xFFFF) match_of_query[prev] = -1;                              // ORBmatcher.cc:464-468
                        match_of_query[qi] = i1;
                        own[i1] = (unsigned short)qi;
                        vmd[i1] = (unsigned short)d1;
                        assigned[qi] = i1;
                    }
                    nmatches += (prev != 0xFFFF) ? 0 : 1;
                    __syncwarp();
                    continue;
                }
                if (lane == 0) {
                    match_of_query[qi] = i1;
                    target_query[i1] = qi;
                    if (flags & 2) taken[i1] = 1;
                }
                nmatches++;
                __syncwarp();
            }
        }
    }
    if (init && !init_done) {
        if (threadIdx.x == 0) s_nmatches = nmatches;
        __syncthreads();
        nmatches = s_nmatches;
    }
    if (init) {
        if (!init_done) for (int i = threadIdx.x; i < a.n; i += SR_THREADS) target_query[i] = own[i] == 0xFFFFu ? -1 : (int)own[i];   // vnMatches21
        if (check_ori) {
            // rotation histogram of every assignment made, stale ones included (ORBmatcher.cc:473-482, 488-511)
            for (int qi = threadIdx.x; qi < a.nq; qi += SR_THREADS) {
                const int t = assigned[qi];
                if (t >= 0) {
                    const int bin = rot_bin(a.q_angle[qi], a.kps[t].angle);
                    match_bin[qi] = (signed char)bin;
                    atomicAdd(&hist[bin], 1);
                }
            }
            __syncthreads();
            int ind1, ind2, ind3;
            three_maxima(hist, ind1, ind2, ind3);
            __syncthreads();
            if (threadIdx.x == 0) s_nmatches = 0;
            __syncthreads();
            int removed = 0;
            for (int qi = threadIdx.x; qi < a.nq; qi += SR_THREADS) {
                const int bin = match_bin[qi];
                if (bin >= 0 && bin != ind1 && bin != ind2 && bin != ind3 && match_of_query[qi] >= 0) { match_of_query[qi] = -1; removed++; }
            }
            atomicAdd(&s_nmatches, removed);
            __syncthreads();
            nmatches -= s_nmatches;
        }
        if (threadIdx.x == 0) *nmatches_out = nmatches;
        return;
    }
    for (int i = threadIdx.x; i < a.n; i += SR_THREADS) taken_g[i] = taken[i];
    if (mode == ORB_MODE_TRACK_LAST && check_ori) {
        // rotation histogram of every match made (ORBmatcher.cc:1433-1441), after the walk: its global loads are parallel
        for (int qi = threadIdx.x; qi < a.nq; qi += SR_THREADS) {
            const int m = match_of_query[qi];
            if (m >= 0) {
                const int bin = rot_bin(a.q_angle[qi], a.kps[m].angle);
                match_bin[qi] = (signed char)bin;
                atomicAdd(&hist[bin], 1);
            }
        }
        __syncthreads();
        int ind1, ind2, ind3;
        three_maxima(hist, ind1, ind2, ind3);
        __syncthreads();
        if (threadIdx.x == 0) s_nmatches = 0;
        __syncthreads();
        int removed = 0;
        for (int qi = threadIdx.x; qi < a.nq; qi += SR_THREADS) {
            const int bin = match_bin[qi];
            if (bin >= 0 && bin != ind1 && bin != ind2 && bin != ind3) {
                target_query[match_of_query[qi]] = -2;   // ORBmatcher.cc:1462-1466: the reference NULLs the keypoint's map point (also one it held before)
                removed++;
            }
        }
        atomicAdd(&s_nmatches, removed);
        __syncthreads();
        nmatches -= s_nmatches;
        // per-query view: a query whose target was nulled loses its match
        for (int qi = threadIdx.x; qi < a.nq; qi += SR_THREADS) {
            const int m = match_of_query[qi];
            if (m >= 0 && target_query[m] == -2) match_of_query[qi] = -1;
        }
    }
    if (threadIdx.x == 0) *nmatches_out = nmatches;
}

// =============================== brute force with mask (SearchByBoW inner loop) ===============================
#define BF_K SR_K

// one warp per query row (latency shape, one pair): distances to every target (stored, u16) + sorted top-K packed keys (dist<<16 | j)
__device__ __forceinline__ void bf_rows_body(const uint8_t* __restrict__ d1, int n1, const uint8_t* __restrict__ d2, int n2,
                                             unsigned short* __restrict__ D, int dpitch, unsigned* __restrict__ topk,
                                             unsigned track_key /* see bf_rows_tiled_body */) {
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
        D[(size_t)i * dpitch + j] = (unsigned short)d;   // the one-pair form keeps the distance row for the resolve step's fallback scan
        const unsigned key = ((unsigned)d << 16) | (unsigned)j;
        if (key < track_key && key < best[BF_K - 1]) insert(key);   // only distances that can change a decision (track_key <= 256 << 16: a distance of
                                                                     // 256 never becomes best or second best, ORBmatcher.cc:217-226)
    }
    const unsigned mine = warp_topk_extract(best, lane);
    if (lane < BF_K) topk[(size_t)i * BF_K + lane] = mine;
}

// Throughput shape (batches of pairs): one THREAD per query, the query in registers, the targets streamed through shared memory
// in tiles and read with broadcast LDS.128 — the map-wide search kernel's structure (orb_hamming.cu): carry-save popcount tree,
// 4 POPC per compare, key = dist << 16 | j built by multiply-adds on the FMA pipe, the list touched once per group of 4 targets.
// The warp-per-query body above spends 63 warp-instructions per warp-compare (two 16-byte global loads per lane per compare, 8
// POPC), this one about 27.
//   Only distances that can change a decision are listed: a match needs best <= th_dist, and once second-best exceeds
// track = max { d : th_dist >= nn_ratio * d (fp32) } the ratio test (ORBmatcher.cc:229-231) passes for every admissible best, so
// the resolve step never needs the exact value of a distance above `track` (nor its index).  The list of a query therefore
// holds the targets with dist <= track (sorted, at most SR_K; a full list still means "truncated" to the resolve step, which then
// rescans the row) — on real descriptor sets that is a handful of entries, so the insertion path is rare and the common case of
// a group is 2 min + 1 compare.  Exactness: bf_resolve_body reads a missing second entry as 256, and 256 decides like any
// distance above `track`.
#ifndef BFT_THREADS
#define BFT_THREADS 128   // queries per CTA: 1000 queries x 128 pairs = 1024 CTAs of 4 warps (256-thread CTAs leave the SMs with 3 or 4 long CTAs each)
#endif
#define BFT_TILE 128   // targets per shared-memory tile (4 KB)
__device__ __forceinline__ unsigned bft_lop3_xor3(unsigned a, unsigned b, unsigned c) {
    unsigned d; asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}
__device__ __forceinline__ unsigned bft_lop3_maj(unsigned a, unsigned b, unsigned c) {
    unsigned d; asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}
__device__ __forceinline__ unsigned bft_key(const uint4& a, const uint4& b, const unsigned (&q)[8], unsigned j, unsigned m16, unsigned m17,
                                            unsigned m18) {
    const unsigned x0 = a.x ^ q[0], x1 = a.y ^ q[1], x2 = a.z ^ q[2], x3 = a.w ^ q[3];
    const unsigned x4 = b.x ^ q[4], x5 = b.y ^ q[5], x6 = b.z ^ q[6], x7 = b.w ^ q[7];
    const unsigned s1 = bft_lop3_xor3(x0, x1, x2), c1 = bft_lop3_maj(x0, x1, x2);
    const unsigned s2 = bft_lop3_xor3(x3, x4, x5), c2 = bft_lop3_maj(x3, x4, x5);
    const unsigned s3 = bft_lop3_xor3(s1, s2, x6), c3 = bft_lop3_maj(s1, s2, x6);
    const unsigned t1 = bft_lop3_xor3(c1, c2, c3), f1 = bft_lop3_maj(c1, c2, c3);
    unsigned k;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(k) : "r"(__popc(s3)), "r"(m16), "r"(j));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(k) : "r"(__popc(x7)), "r"(m16));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(k) : "r"(__popc(t1)), "r"(m17));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(k) : "r"(__popc(f1)), "r"(m18));
    return k;
}
__device__ __forceinline__ void bf_rows_tiled_body(const uint8_t* __restrict__ d1, int n1, const uint8_t* __restrict__ d2, int n2,
                                                   unsigned* __restrict__ topk, unsigned track_key /* (track + 1) << 16 */, unsigned m16,
                                                   unsigned m17, unsigned m18) {
    __shared__ uint4 tile[2][BFT_TILE * 2];
    const int tid = threadIdx.x;
    const int i = blockIdx.x * BFT_THREADS + tid;
    unsigned q[8];
    {
        uint4 a = make_uint4(0, 0, 0, 0), b = a;
        if (i < n1) { a = __ldg(reinterpret_cast<const uint4*>(d1) + 2 * (size_t)i); b = __ldg(reinterpret_cast<const uint4*>(d1) + 2 * (size_t)i + 1); }
        q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = a.w; q[4] = b.x; q[5] = b.y; q[6] = b.z; q[7] = b.w;
    }
    unsigned best[SR_K];
#pragma unroll
    for (int k = 0; k < SR_K; ++k) best[k] = 0xFFFFFFFFu;
    unsigned lim = track_key;                            // keys >= lim cannot enter the list
    auto insert = [&](unsigned key) {
        if (key >= lim) return;
#pragma unroll
        for (int k = 0; k < SR_K; ++k) {
            const unsigned lo = min(best[k], key);
            key = max(best[k], key);
            best[k] = lo;
        }
        lim = min(track_key, best[SR_K - 1]);
    };
    const int ntiles = (n2 + BFT_TILE - 1) / BFT_TILE;
    auto issue = [&](int t, int buf) {                   // tile t -> 256 chunks of 16 bytes
        const int chunks = min(BFT_TILE, n2 - t * BFT_TILE) * 2;
#pragma unroll
        for (int c = tid; c < BFT_TILE * 2; c += BFT_THREADS) {
            if (c < chunks) {
                const unsigned sa = (unsigned)__cvta_generic_to_shared(&tile[buf][c]);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(reinterpret_cast<const uint4*>(d2) + (size_t)t * BFT_TILE * 2 + c));
            }
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
    if (ntiles > 0) issue(0, 0);
    for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        if (t + 1 < ntiles) { issue(t + 1, buf ^ 1); asm volatile("cp.async.wait_group 1;\n" ::); }
        else asm volatile("cp.async.wait_group 0;\n" ::);
        __syncthreads();
        const int rows = min(BFT_TILE, n2 - t * BFT_TILE);
        const unsigned j0 = (unsigned)(t * BFT_TILE);
        int r = 0;
        for (; r + 4 <= rows; r += 4) {
            unsigned key[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) key[u] = bft_key(tile[buf][2 * (r + u)], tile[buf][2 * (r + u) + 1], q, j0 + r + u, m16, m17, m18);
            if (min(__vimin3_u32(key[0], key[1], key[2]), key[3]) < lim) {
#pragma unroll
                for (int u = 0; u < 4; ++u) insert(key[u]);
            }
        }
        for (; r < rows; ++r) insert(bft_key(tile[buf][2 * r], tile[buf][2 * r + 1], q, j0 + r, m16, m17, m18));
        __syncthreads();
    }
    if (i < n1) {
#pragma unroll
        for (int k = 0; k < SR_K; ++k) topk[(size_t)i * SR_K + k] = best[k];
    }
}
// the largest second-best distance whose exact value can still matter: (float)best < nn_ratio * (float)second (ORBmatcher.cc:229-231,
// separately rounded fp32) holds for every best <= th_dist as soon as second > track
static int bf_track_limit(int th_dist, float nn_ratio) {
    int b = 0;
    while (b < 256 && !((float)th_dist < nn_ratio * (float)b)) ++b;   // smallest second-best that passes for best = th_dist
    return std::min(255, std::max(b - 1, th_dist));
}

// The order-dependent part ("a target taken by an earlier query is invisible to the later ones", ORBmatcher.cc:210) as a
// parallel FIXED-POINT iteration instead of a sequential walk.  The sequential result is the unique solution of
//     choice[i] = decide(i, { t : own[t] < i }),   own[t] = min { i : choice[i] == t }
// (induction over i).  Every round evaluates all queries in parallel against the owners of the previous round and
// rebuilds `own` with atomicMin; when a round changes no choice the recurrence holds for every i, i.e. the state IS the
// sequential result.  Query i is final after round i + 1 at the latest, typical inputs converge in 3-6 rounds of a few
// microseconds (the sequential walk of 1000 queries took 330 us).  The optimistic top-K list answers a query when two of
// its entries are still visible; otherwise the query's warp scans its whole distance row cooperatively.
template <bool SMEM>
__device__ __forceinline__ void bf_resolve_body(const uint8_t* __restrict__ d1, const uint8_t* __restrict__ d2, const unsigned short* __restrict__ D, int dpitch,
                                                const unsigned* __restrict__ topk, int n1, int n2,
                                                const float* __restrict__ angle1, const float* __restrict__ angle2, int th_dist, float nn_ratio,
                                                int check_ori, int* owner /*[n2]*/, int* owner_scratch /*[n2], !SMEM only*/, int* match12,
                                                signed char* match_bin, int* nmatches_out) {
    extern __shared__ __align__(16) uint8_t sr_smem[];
    int* own_prev = SMEM ? reinterpret_cast<int*>(sr_smem) : owner;
    int* own_new = SMEM ? reinterpret_cast<int*>(sr_smem) + n2 : owner_scratch;
    const int lane = threadIdx.x & 31;
    __shared__ int hist[HISTO_LENGTH];
    __shared__ int s_nmatches, s_changed;
    if (threadIdx.x < HISTO_LENGTH) hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_nmatches = 0;
    for (int j = threadIdx.x; j < n2; j += SR_THREADS) own_prev[j] = BF_NONE;
    for (int i = threadIdx.x; i < n1; i += SR_THREADS) { match12[i] = -1; match_bin[i] = -1; }
    const int n1_up = (n1 + 31) & ~31;   // whole warps enter the cooperative fallback
    for (int round = 0; round <= n1; ++round) {
        for (int j = threadIdx.x; j < n2; j += SR_THREADS) own_new[j] = BF_NONE;
        if (threadIdx.x == 0) s_changed = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < n1_up; i += SR_THREADS) {
            int best1 = 256, best2 = 256, bestIdx = -1, nfree = 0;
            bool fallback = false;
            if (i < n1) {
                // optimistic: the first two visible entries of the sorted unmasked top-K are the masked best / second
                unsigned key[SR_K];
#pragma unroll
                for (int k = 0; k < SR_K; ++k) key[k] = topk[(size_t)i * SR_K + k];
#pragma unroll
                for (int k = 0; k < SR_K; ++k) {
                    if (key[k] == 0xFFFFFFFFu || nfree >= 2) break;
                    const int id = (int)(key[k] & 0xFFFFu);
                    if ((SMEM ? own_prev[id] : __ldcg(own_prev + id)) < i) continue;   // taken by an earlier query
                    if (nfree == 0) { best1 = (int)(key[k] >> 16); bestIdx = id; } else best2 = (int)(key[k] >> 16);
                    ++nfree;
                }
                fallback = nfree < 2 && key[SR_K - 1] != 0xFFFFFFFFu;   // list truncated and too many of its entries taken
            }
            // fallback: the lanes of the warp scan the distances of each such query to all targets
            unsigned todo = __ballot_sync(0xffffffffu, fallback);
            while (todo) {
                const int src = __ffs((int)todo) - 1;
                todo &= todo - 1u;
                const int fi = __shfl_sync(0xffffffffu, i, src);
                unsigned a1 = 0xFFFFFFFFu, a2 = 0xFFFFFFFFu;
                const uint4* dq = reinterpret_cast<const uint4*>(d1 + (size_t)fi * 32);
                for (int jj = lane; jj < n2; jj += 32) {
                    if ((SMEM ? own_prev[jj] : __ldcg(own_prev + jj)) < fi) continue;
                    // the batched form keeps no distance matrix (2 MB per pair): its fallback recomputes the row
                    const unsigned dd = D ? (unsigned)D[(size_t)fi * dpitch + jj] : (unsigned)dist256(dq, reinterpret_cast<const uint4*>(d2 + (size_t)jj * 32));
                    if (dd >= 256u) continue;                               // invisible to the reference's strict '<' updates
                    const unsigned k = (dd << 16) | (unsigned)jj;
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
                if (lane == src) {
                    best1 = best2 = 256; bestIdx = -1;
                    if (a1 != 0xFFFFFFFFu) { best1 = (int)(a1 >> 16); bestIdx = (int)(a1 & 0xFFFFu); }
                    if (a2 != 0xFFFFFFFFu) best2 = (int)(a2 >> 16);
                }
            }
            if (i < n1) {
                int c = -1;
                if (bestIdx >= 0 && best1 < 256 && best1 <= th_dist && (float)best1 < __fmul_rn(nn_ratio, (float)best2)) c = bestIdx;   // ORBmatcher.cc:229-231
                if (c != match12[i]) { match12[i] = c; s_changed = 1; }
                if (c >= 0) atomicMin(own_new + c, i);
            }
        }
        __syncthreads();
        const int changed = s_changed;
        int* t = own_prev; own_prev = own_new; own_new = t;
        __syncthreads();
        if (!changed) break;
    }
    // own_prev holds the final owners (every query takes its target: at most one query per target)
    int nmatches = 0;
    for (int i = threadIdx.x; i < n1; i += SR_THREADS) nmatches += match12[i] >= 0;
    for (int j = threadIdx.x; j < n2; j += SR_THREADS) {
        const int o = SMEM ? own_prev[j] : __ldcg(own_prev + j);
        owner_scratch[j] = (o == BF_NONE) ? -1 : o;   // the result array the host reads (see the launch)
    }
    atomicAdd(&s_nmatches, nmatches);
    __syncthreads();
    nmatches = s_nmatches;
    int* owner_out = owner_scratch;
    if (check_ori) {
        for (int i = threadIdx.x; i < n1; i += SR_THREADS) {      // rotation histogram (ORBmatcher.cc:236-247), parallel
            const int m = match12[i];
            if (m >= 0) {
                const int bin = rot_bin(angle1[i], angle2[m]);
                match_bin[i] = (signed char)bin;
                atomicAdd(&hist[bin], 1);
            }
        }
        __syncthreads();
        int ind1, ind2, ind3;
        three_maxima(hist, ind1, ind2, ind3);
        __syncthreads();
        if (threadIdx.x == 0) s_nmatches = 0;
        __syncthreads();
        int removed = 0;
        for (int i = threadIdx.x; i < n1; i += SR_THREADS) {
            const int bin = match_bin[i];
            if (bin >= 0 && bin != ind1 && bin != ind2 && bin != ind3) { owner_out[match12[i]] = -1; match12[i] = -1; removed++; }
        }
        atomicAdd(&s_nmatches, removed);
        __syncthreads();
        nmatches -= s_nmatches;
    }
    if (threadIdx.x == 0) *nmatches_out = nmatches;
}

// ---- launch forms of the bodies above: one pair (arguments in the kernel parameters) and a batch of pairs -------------
__global__ void __launch_bounds__(1024)
grid_build_kernel(const orb_kp* __restrict__ kps, int n, int npad, float min_x, float min_y, float inv_w, float inv_h,
                  unsigned* __restrict__ items, int* __restrict__ cell_start) {
    grid_build_body(kps, n, npad, min_x, min_y, inv_w, inv_h, items, cell_start);
}
__global__ void __launch_bounds__(256) window_candidates_kernel(const SearchArgs a) { window_candidates_body(a); }
__global__ void __launch_bounds__(SR_THREADS, 1)
window_resolve_kernel(const SearchArgs a, int mode, int th_dist, float nn_ratio, int check_ori, uint8_t* taken_g, int* match_of_query,
                      int* target_query, signed char* match_bin, int* assigned, int* nmatches_out, int* overflow, int smem_bytes) {
    window_resolve_body(a, mode, th_dist, nn_ratio, check_ori, taken_g, match_of_query, target_query, match_bin, assigned, nmatches_out, overflow,
                        smem_bytes);
}
__global__ void __launch_bounds__(256)
bf_rows_kernel(const uint8_t* __restrict__ d1, int n1, const uint8_t* __restrict__ d2, int n2, unsigned short* __restrict__ D, int dpitch,
               unsigned* __restrict__ topk, unsigned track_key) {
    bf_rows_body(d1, n1, d2, n2, D, dpitch, topk, track_key);
}
template <bool SMEM>
__global__ void __launch_bounds__(SR_THREADS, 1)
bf_resolve_kernel(const uint8_t* __restrict__ d1, const uint8_t* __restrict__ d2, const unsigned short* __restrict__ D, int dpitch,
                  const unsigned* __restrict__ topk, int n1, int n2,
                  const float* __restrict__ angle1, const float* __restrict__ angle2, int th_dist, float nn_ratio, int check_ori, int* owner,
                  int* owner_scratch, int* match12, signed char* match_bin, int* nmatches_out) {
    bf_resolve_body<SMEM>(d1, d2, D, dpitch, topk, n1, n2, angle1, angle2, th_dist, nn_ratio, check_ori, owner, owner_scratch, match12, match_bin, nmatches_out);
}

// Batched forms (per-pair matching shards like frames): blockIdx.y (one-CTA kernels: blockIdx.x) = pair.  The
// host builds one job per pair with the pair's device pointers; the keypoint / query counts stay on the device (they are
// outputs of the extractor) and are read here.
struct WindowJob {
    SearchArgs a;                       // a.n / a.nq = capacities; the live counts come from n_ptr / nq_ptr
    const int *n_ptr, *nq_ptr;
    uint8_t* taken; int* moq; int* tq; signed char* bin; int* asg; int* scal;   // scal: [0] candidate total, [1] nmatches, [2] overflow
};
__device__ __forceinline__ SearchArgs job_args(const WindowJob& J) {
    SearchArgs a = J.a;
    a.n = min(*J.n_ptr, J.a.n);
    a.nq = min(*J.nq_ptr, J.a.nq);
    return a;
}
__device__ __forceinline__ int pow2_at_least(int n) { int p = 32; while (p < n) p <<= 1; return p; }
__global__ void __launch_bounds__(1024) grid_build_batch_kernel(const WindowJob* __restrict__ jobs) {
    const WindowJob& J = jobs[blockIdx.x];
    const SearchArgs a = job_args(J);
    if (threadIdx.x < 4) J.scal[threadIdx.x] = 0;
    grid_build_body(a.kps, a.n, pow2_at_least(a.n), a.min_x, a.min_y, a.inv_w, a.inv_h, const_cast<unsigned*>(a.items), const_cast<int*>(a.cell_start));
}
__global__ void __launch_bounds__(256) window_candidates_batch_kernel(const WindowJob* __restrict__ jobs) {
    const SearchArgs a = job_args(jobs[blockIdx.y]);
    window_candidates_body(a);
}
__global__ void __launch_bounds__(SR_THREADS, 1)
window_resolve_batch_kernel(const WindowJob* __restrict__ jobs, int mode, int th_dist, float nn_ratio, int check_ori, int smem_bytes) {
    const WindowJob& J = jobs[blockIdx.x];
    const SearchArgs a = job_args(J);
    // slots beyond the live counts read as "no match"
    for (int i = a.nq + threadIdx.x; i < J.a.nq; i += SR_THREADS) J.moq[i] = -1;
    for (int i = a.n + threadIdx.x; i < J.a.n; i += SR_THREADS) J.tq[i] = -1;
    if (a.n == 0 || a.nq == 0) {
        for (int i = threadIdx.x; i < a.nq; i += SR_THREADS) J.moq[i] = -1;
        for (int i = threadIdx.x; i < a.n; i += SR_THREADS) J.tq[i] = -1;
        if (threadIdx.x == 0) { J.scal[1] = 0; J.scal[2] = 0; }
        return;
    }
    window_resolve_body(a, mode, th_dist, nn_ratio, check_ori, J.taken, J.moq, J.tq, J.bin, J.asg, J.scal + 1, J.scal + 2, smem_bytes);
}

__global__ void search_batch_finish_kernel(const WindowJob* __restrict__ jobs, int npairs, int* __restrict__ nmatches) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < npairs) nmatches[p] = jobs[p].scal[2] > 0 ? -1 : jobs[p].scal[1];
}

struct BfJob {
    const uint8_t *d1, *d2; const orb_kp *k1, *k2; const int *n1_ptr, *n2_ptr; int cap1, cap2;
    unsigned* topk; float *a1, *a2; int *owner, *owner2, *m12; signed char* bin; int* nm;
};
__global__ void __launch_bounds__(BFT_THREADS) bf_rows_batch_kernel(const BfJob* __restrict__ jobs, unsigned track_key, unsigned m16, unsigned m17,
                                                                    unsigned m18) {
    const BfJob& J = jobs[blockIdx.y];
    const int n1 = min(*J.n1_ptr, J.cap1), n2 = min(*J.n2_ptr, J.cap2);
    // dense angle arrays for the rotation histogram (the extractor's keypoints are 28-byte records)
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n1) J.a1[t] = J.k1[t].angle;
    for (int j = t; j < n2; j += gridDim.x * blockDim.x) J.a2[j] = J.k2[j].angle;
    if (blockIdx.x * BFT_THREADS >= n1) return;   // uniform per CTA
    bf_rows_tiled_body(J.d1, n1, J.d2, n2, J.topk, track_key, m16, m17, m18);
}
template <bool SMEM>
__global__ void __launch_bounds__(SR_THREADS, 1)
bf_resolve_batch_kernel(const BfJob* __restrict__ jobs, int th_dist, float nn_ratio, int check_ori) {
    const BfJob& J = jobs[blockIdx.x];
    const int n1 = min(*J.n1_ptr, J.cap1), n2 = min(*J.n2_ptr, J.cap2);
    for (int i = n1 + threadIdx.x; i < J.cap1; i += SR_THREADS) J.m12[i] = -1;
    if (n1 == 0 || n2 == 0) {
        for (int i = threadIdx.x; i < n1; i += SR_THREADS) J.m12[i] = -1;
        if (threadIdx.x == 0) *J.nm = 0;
        return;
    }
    bf_resolve_body<SMEM>(J.d1, J.d2, nullptr, 0, J.topk, n1, n2, J.a1, J.a2, th_dist, nn_ratio, check_ori, J.owner, J.owner2, J.m12, J.bin, J.nm);
}

// ---- best / second-best over explicit candidate lists (CSR): one warp per query ------------------------------
// Ties keep the FIRST candidate in list order (strict '<' of ORBmatcher.cc:217-226): key = dist << 23 | position.
__global__ void __launch_bounds__(256)
csr_top2_kernel(const uint8_t* __restrict__ q, int nq, const uint8_t* __restrict__ db, const int* __restrict__ off,
                const int* __restrict__ idx, orb_top2* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= nq) return;
    const uint4* dq = reinterpret_cast<const uint4*>(q + (size_t)i * 32);
    const int c0 = off[i], cnt = off[i + 1] - c0;
    unsigned k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu;
    for (int p = lane; p < cnt; p += 32) {
        const int d = dist256(dq, reinterpret_cast<const uint4*>(db + (size_t)idx[c0 + p] * 32));
        const unsigned k = ((unsigned)d << 23) | (unsigned)p;
        k2 = min(k2, max(k, k1));
        k1 = min(k1, k);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned b1 = __shfl_xor_sync(0xffffffffu, k1, o), b2 = __shfl_xor_sync(0xffffffffu, k2, o);
        const unsigned lo = min(k1, b1), hi = max(k1, b1);
        k2 = min(hi, min(k2, b2));
        k1 = lo;
    }
    if (lane == 0) {
        orb_top2 r;
        r.best_dist = k1 == 0xFFFFFFFFu ? 256 : (int)(k1 >> 23);
        r.best_idx = k1 == 0xFFFFFFFFu ? -1 : idx[c0 + (k1 & 0x7FFFFFu)];
        r.second_dist = k2 == 0xFFFFFFFFu ? 256 : (int)(k2 >> 23);
        r.second_idx = k2 == 0xFFFFFFFFu ? -1 : idx[c0 + (k2 & 0x7FFFFFu)];
        out[i] = r;
    }
}

// =============================== SearchByBoW over two FeatureVectors ===============================================
// (ORBmatcher.cc:160-289 KeyFrame -> Frame, :524-657 KeyFrame -> KeyFrame.)  A keypoint belongs to exactly one
// vocabulary node, so the "target already matched" state never crosses nodes: ONE WARP PER SHARED NODE walks the
// node's query features in order (the sequential part), the 32 lanes score the node's free targets in parallel and
// reduce to best / second with key = dist << 16 | position in the node's list (strict '<': first in list order wins).
__global__ void __launch_bounds__(256)
bow_match_kernel(const uint8_t* __restrict__ d1, const uint8_t* __restrict__ valid1, const int* __restrict__ fv1_node,
                 const int* __restrict__ fv1_start, const int* __restrict__ fv1_feat, int nfv1, const uint8_t* __restrict__ d2,
                 const uint8_t* __restrict__ valid2, const int* __restrict__ fv2_node, const int* __restrict__ fv2_start,
                 const int* __restrict__ fv2_feat, int nfv2, int th_dist, int strict, float nn_ratio, int* match12, int* match21) {
    const int lane = threadIdx.x & 31;
    const int a = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (a >= nfv1) return;
    const int node = fv1_node[a];
    int lo = 0, hi = nfv2;                                   // lower_bound of the node id in the other vector
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (fv2_node[mid] < node) lo = mid + 1; else hi = mid;
    }
    if (lo >= nfv2 || fv2_node[lo] != node) return;
    const int s2 = fv2_start[lo], c2 = fv2_start[lo + 1] - s2;
    for (int i1 = fv1_start[a]; i1 < fv1_start[a + 1]; ++i1) {
        const int idx1 = fv1_feat[i1];
        if (valid1 && !valid1[idx1]) continue;
        const uint4* dq = reinterpret_cast<const uint4*>(d1 + (size_t)idx1 * 32);
        unsigned k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu;
        for (int p = lane; p < c2; p += 32) {
            const int idx2 = fv2_feat[s2 + p];
            if (match21[idx2] >= 0 || (valid2 && !valid2[idx2])) continue;
            const unsigned k = ((unsigned)dist256(dq, reinterpret_cast<const uint4*>(d2 + (size_t)idx2 * 32)) << 16) | (unsigned)p;
            k2 = min(k2, max(k, k1));
            k1 = min(k1, k);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned b1 = __shfl_xor_sync(0xffffffffu, k1, o), b2 = __shfl_xor_sync(0xffffffffu, k2, o);
            const unsigned l = min(k1, b1), h = max(k1, b1);
            k2 = min(h, min(k2, b2));
            k1 = l;
        }
        const int best1 = k1 == 0xFFFFFFFFu ? 256 : (int)(k1 >> 16), best2 = k2 == 0xFFFFFFFFu ? 256 : (int)(k2 >> 16);
        if ((strict ? best1 < th_dist : best1 <= th_dist) && (float)best1 < __fmul_rn(nn_ratio, (float)best2)) {
            const int idx2 = fv2_feat[s2 + (int)(k1 & 0xFFFFu)];
            if (lane == 0) { match12[idx1] = idx2; match21[idx2] = idx1; }
        }
        __syncwarp();   // the commit is visible to every lane's next scan of match21
    }
}

// rotation histogram + three-maxima filter over the finished matches (ORBmatcher.cc:236-247, 268-285): one CTA
__global__ void __launch_bounds__(1024)
bow_rot_kernel(const float* __restrict__ angle1, const float* __restrict__ angle2, int n1, int check_ori, int* match12, int* match21,
               int* nmatches_out) {
    __shared__ int hist[HISTO_LENGTH];
    __shared__ int s_n;
    if (threadIdx.x < HISTO_LENGTH) hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    int cnt = 0;
    for (int i = threadIdx.x; i < n1; i += blockDim.x) {
        const int m = match12[i];
        if (m >= 0) { ++cnt; if (check_ori) atomicAdd(&hist[rot_bin(angle1[i], angle2[m])], 1); }
    }
    __syncthreads();
    if (check_ori) {
        int ind1, ind2, ind3;
        three_maxima(hist, ind1, ind2, ind3);
        for (int i = threadIdx.x; i < n1; i += blockDim.x) {
            const int m = match12[i];
            if (m < 0) continue;
            const int bin = rot_bin(angle1[i], angle2[m]);
            if (bin != ind1 && bin != ind2 && bin != ind3) { if (match21) match21[m] = -1; match12[i] = -1; --cnt; }
        }
    }
    atomicAdd(&s_n, cnt);
    __syncthreads();
    if (threadIdx.x == 0) *nmatches_out = s_n;
}

// =============================== SearchForTriangulation (ORBmatcher.cc:659-825) ===================================
// No loop-carried state (the reference never sets vbMatched2): one warp per shared node, its query keypoints one after
// the other, the lanes gate and score the node's targets.  The reference keeps a candidate when `dist <= bestDist` and the
// epipolar gates pass (OM:741-757), i.e. the minimum distance, LAST in list order on ties: key = dist << 16 | (0xFFFF - pos).
struct TriFrame {
    const orb_kp* kps; const uint8_t* desc; const uint8_t* has_mp; const float* u_right;
    const int *fv_node, *fv_start, *fv_feat; int nfv;
};
struct TriParams { float F[9]; float ex, ey; float scale[ORB_MAX_LEVELS]; float sigma2[ORB_MAX_LEVELS]; int nlevels, only_stereo, th_low; };

__global__ void __launch_bounds__(256)
bow_triangulation_kernel(const TriFrame A, const TriFrame B, const __grid_constant__ TriParams P, int* __restrict__ match12) {
    const int lane = threadIdx.x & 31;
    const int a = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (a >= A.nfv) return;
    const int node = A.fv_node[a];
    int lo = 0, hi = B.nfv;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (B.fv_node[mid] < node) lo = mid + 1; else hi = mid;
    }
    if (lo >= B.nfv || B.fv_node[lo] != node) return;
    const int s2 = B.fv_start[lo], c2 = B.fv_start[lo + 1] - s2;
    for (int i1 = A.fv_start[a]; i1 < A.fv_start[a + 1]; ++i1) {
        const int idx1 = A.fv_feat[i1];
        if (A.has_mp && A.has_mp[idx1]) continue;                                   // OM:703-705
        const bool st1 = A.u_right && A.u_right[idx1] >= 0.0f;
        if (P.only_stereo && !st1) continue;
        const orb_kp kp1 = A.kps[idx1];
        // epipolar line l = x1' F12 (OM:143-145), every product and sum rounded separately
        const float la = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, P.F[0]), __fmul_rn(kp1.y, P.F[3])), P.F[6]);
        const float lb = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, P.F[1]), __fmul_rn(kp1.y, P.F[4])), P.F[7]);
        const float lc = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, P.F[2]), __fmul_rn(kp1.y, P.F[5])), P.F[8]);
        const float den = __fadd_rn(__fmul_rn(la, la), __fmul_rn(lb, lb));
        const uint4* dq = reinterpret_cast<const uint4*>(A.desc + (size_t)idx1 * 32);
        unsigned best = 0xFFFFFFFFu;
        for (int p = lane; p < c2; p += 32) {
            const int idx2 = B.fv_feat[s2 + p];
            if (B.has_mp && B.has_mp[idx2]) continue;                               // OM:726-728
            const bool st2 = B.u_right && B.u_right[idx2] >= 0.0f;
            if (P.only_stereo && !st2) continue;
            const int dist = dist256(dq, reinterpret_cast<const uint4*>(B.desc + (size_t)idx2 * 32));
            if (dist > P.th_low) continue;
            const orb_kp kp2 = B.kps[idx2];
            const int oct = min(max(kp2.octave, 0), ORB_MAX_LEVELS - 1);
            if (!st1 && !st2) {                                                     // too close to the epipole (OM:745-751)
                const float dx = __fsub_rn(P.ex, kp2.x), dy = __fsub_rn(P.ey, kp2.y);
                if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < __fmul_rn(100.0f, P.scale[oct])) continue;
            }
            const float num = __fadd_rn(__fadd_rn(__fmul_rn(la, kp2.x), __fmul_rn(lb, kp2.y)), lc);
            if (den == 0.0f) continue;
            const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
            if (!((double)dsqr < __dmul_rn(3.84, (double)P.sigma2[oct]))) continue;   // OM:156
            best = min(best, ((unsigned)dist << 16) | (unsigned)(0xFFFF - p));
        }
        best = warp_min_u32(best);
        if (lane == 0 && best != 0xFFFFFFFFu) match12[idx1] = B.fv_feat[s2 + (0xFFFF - (int)(best & 0xFFFFu))];
    }
}

// =============================== MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:288-361), batched ============
// One CTA (4 warps) per map point; warp w takes the rows i = w, w+4, ...: the lanes compute the N distances of row i, the
// median vDists[0.5*(N-1)] is found by rank counting (N <= 32: shuffles) or a 257-bin shared-memory histogram; the point's
// answer is the row with the smallest median, first on ties (key = median << 16 | i).
#define DD_WARPS 4
__global__ void __launch_bounds__(DD_WARPS * 32)
distinctive_kernel(const uint8_t* __restrict__ desc, const int* __restrict__ off, int* __restrict__ best) {
    __shared__ int hist[DD_WARPS][264];
    __shared__ unsigned s_key[DD_WARPS];
    const int p = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int o0 = off[p], N = off[p + 1] - o0;
    if (N <= 0) { if (threadIdx.x == 0) best[p] = -1; return; }
    const int k = (N - 1) >> 1;                                  // (size_t)(0.5 * (N - 1))
    unsigned mykey = 0xFFFFFFFFu;
    for (int i = warp; i < N; i += DD_WARPS) {
        const uint4* di = reinterpret_cast<const uint4*>(desc + (size_t)(o0 + i) * 32);
        int median;
        if (N <= 32) {
            const int d = lane < N ? dist256(di, reinterpret_cast<const uint4*>(desc + (size_t)(o0 + lane) * 32)) : 0x7FFF;
            int rank = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int dj = __shfl_sync(0xffffffffu, d, j);
                rank += (dj < d || (dj == d && j < lane)) ? 1 : 0;
            }
            const unsigned who = __ballot_sync(0xffffffffu, rank == k && lane < N);
            median = __shfl_sync(0xffffffffu, d, __ffs(who) - 1);
        } else {
            for (int b = lane; b < 264; b += 32) hist[warp][b] = 0;
            __syncwarp();
            for (int j = lane; j < N; j += 32) atomicAdd(&hist[warp][dist256(di, reinterpret_cast<const uint4*>(desc + (size_t)(o0 + j) * 32))], 1);
            __syncwarp();
            // the k-th smallest = the first bin whose inclusive prefix count exceeds k: 9 bins per lane, warp scan
            int c[9], tot = 0;
#pragma unroll
            for (int b = 0; b < 9; ++b) { c[b] = lane * 9 + b < 257 ? hist[warp][lane * 9 + b] : 0; tot += c[b]; }
            int incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += y;
            }
            int run = incl - tot, mine = 0x7FFF;
#pragma unroll
            for (int b = 0; b < 9; ++b) {
                run += c[b];
                if (run > k && mine == 0x7FFF) mine = lane * 9 + b;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mine = min(mine, __shfl_xor_sync(0xffffffffu, mine, o));
            median = mine;
            __syncwarp();
        }
        mykey = min(mykey, ((unsigned)median << 16) | (unsigned)min(i, 0xFFFF));
    }
    if (lane == 0) s_key[warp] = mykey;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned m = s_key[0];
        for (int w = 1; w < DD_WARPS; ++w) m = min(m, s_key[w]);
        best[p] = (int)(m & 0xFFFFu);
    }
}

// =============================== search half of ORBmatcher::Fuse (ORBmatcher.cc:827-977, 979-1102) ================
// One warp per projected map point.  The sorted item array of grid_build_kernel lists the window's candidates in the
// reference's visiting order (ix, iy, insertion) at ascending positions, so "first minimum" = min(dist << 16 | position).
struct FuseArgs {
    const orb_kp* kps; const uint8_t* desc; const float* u_right; const float* inv_sigma2; int nlevels;
    int nq; const float *q_u, *q_v, *q_ur, *q_radius; const int* q_level; const uint8_t *q_desc, *q_valid;
    float min_x, min_y, inv_w, inv_h;
    const unsigned* items; const int* cell_start;
    int *best_idx, *best_dist;
};

__global__ void __launch_bounds__(256)
fuse_search_kernel(const FuseArgs a) {
    const int lane = threadIdx.x & 31;
    const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (qi >= a.nq) return;
    if (lane == 0) { a.best_idx[qi] = -1; a.best_dist[qi] = 256; }
    if (a.q_valid && !a.q_valid[qi]) return;
    const float x = a.q_u[qi], y = a.q_v[qi], r = a.q_radius[qi];
    const int lvl = a.q_level[qi];
    // KeyFrame::GetFeaturesInArea (KeyFrame.cc:700-739), float arithmetic with every op rounded separately
    const int nMinCellX = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, a.min_x), r), a.inv_w)));
    if (nMinCellX >= GRID_COLS) return;
    const int nMaxCellX = min(GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, a.min_x), r), a.inv_w)));
    if (nMaxCellX < 0) return;
    const int nMinCellY = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, a.min_y), r), a.inv_h)));
    if (nMinCellY >= GRID_ROWS) return;
    const int nMaxCellY = min(GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, a.min_y), r), a.inv_h)));
    if (nMaxCellY < 0) return;
    const uint4* dq = reinterpret_cast<const uint4*>(a.q_desc + (size_t)qi * 32);
    const float ur = a.q_ur ? a.q_ur[qi] : 0.0f;
    unsigned best = 0xFFFFFFFFu;
    for (int ix = nMinCellX; ix <= nMaxCellX; ++ix) {
        const int s = a.cell_start[ix * GRID_ROWS + nMinCellY], e = a.cell_start[ix * GRID_ROWS + nMaxCellY + 1];
        for (int p = s + lane; p < e; p += 32) {
            const int id = (int)(a.items[p] & 0xFFFFu);
            const orb_kp kp = a.kps[id];
            const float dx = __fsub_rn(kp.x, x), dy = __fsub_rn(kp.y, y);
            if (!(fabsf(dx) < r && fabsf(dy) < r)) continue;
            if (kp.octave < lvl - 1 || kp.octave > lvl) continue;                                   // ORBmatcher.cc:909-910
            if (a.inv_sigma2) {
                const float inv = a.inv_sigma2[min(max(kp.octave, 0), a.nlevels - 1)];
                const float ex = __fsub_rn(x, kp.x), ey = __fsub_rn(y, kp.y);
                float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
                const float kpr = a.u_right ? a.u_right[id] : -1.0f;
                if (kpr >= 0.0f) {
                    const float er = __fsub_rn(ur, kpr);
                    e2 = __fadd_rn(e2, __fmul_rn(er, er));
                    if ((double)__fmul_rn(e2, inv) > 7.8) continue;                                  // :914-927
                } else {
                    if ((double)__fmul_rn(e2, inv) > 5.99) continue;                                 // :929-939
                }
            }
            const unsigned d = (unsigned)dist256(dq, reinterpret_cast<const uint4*>(a.desc + (size_t)id * 32));
            best = min(best, (d << 16) | (unsigned)p);
        }
    }
    best = warp_min_u32(best);
    if (lane == 0 && best != 0xFFFFFFFFu) {
        a.best_idx[qi] = (int)(a.items[best & 0xFFFFu] & 0xFFFFu);
        a.best_dist[qi] = (int)(best >> 16);
    }
}

// Per-thread, per-device workspace of the host-pointer entry points: one grow-only device slab, one grow-only pinned
// slab and a private stream.  A call packs all its inputs into the pinned slab, issues ONE H2D copy, the kernels and
// ONE D2H copy, and synchronises once — no cudaMalloc / cudaFree on the call path.
struct Workspace {
    int device = -1;
    cudaStream_t st = nullptr;
    uint8_t *d = nullptr, *h = nullptr;
    size_t d_cap = 0, h_cap = 0;
    std::vector<uint8_t> jobs_on_device;   // batched entry points: the job array currently at the head of the device slab
    ~Workspace() {
        // (the CUDA context may already be gone at thread exit: ignore errors)
        if (device >= 0 && cudaSetDevice(device) == cudaSuccess) { cudaFree(d); cudaFreeHost(h); if (st) cudaStreamDestroy(st); }
    }
    int prepare(int dev, size_t dbytes, size_t hbytes, bool batched = false) {
        ORB_CUDA(cudaSetDevice(dev));
        if (!batched) jobs_on_device.clear();   // the single-pair entry points overwrite the head of the slab
        if (device != dev) {
            if (device >= 0) { cudaSetDevice(device); cudaFree(d); cudaFreeHost(h); if (st) cudaStreamDestroy(st); cudaSetDevice(dev); }
            d = h = nullptr; d_cap = h_cap = 0; st = nullptr; device = dev; jobs_on_device.clear();
            ORB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        }
        if (d_cap < dbytes) {
            ORB_CUDA(cudaStreamSynchronize(st));
            cudaFree(d); d = nullptr; d_cap = 0; jobs_on_device.clear();
            ORB_CUDA(cudaMalloc(&d, dbytes + dbytes / 2));
            d_cap = dbytes + dbytes / 2;
        }
        if (h_cap < hbytes) {
            ORB_CUDA(cudaStreamSynchronize(st));
            cudaFreeHost(h); h = nullptr; h_cap = 0;
            ORB_CUDA(cudaMallocHost(&h, hbytes + hbytes / 2));
            h_cap = hbytes + hbytes / 2;
        }
        return ORB_OK;
    }
};
thread_local Workspace g_ws;
thread_local Workspace g_ws_batch[2];   // batched brute force / batched windowed search: own slabs, so their job arrays stay cached

// Upload a batched call's job array to the head of the device slab unless the identical array is already there: a pipeline that
// calls with the same buffers every step (the normal case) then issues no copy at all — a pageable-source cudaMemcpyAsync stalls
// the host until the stream has drained, which serialises launch and execution.
int upload_jobs(Workspace& W, const void* jobs, size_t bytes, cudaStream_t st) {
    if (W.jobs_on_device.size() == bytes && memcmp(W.jobs_on_device.data(), jobs, bytes) == 0) return ORB_OK;
    W.jobs_on_device.assign((const uint8_t*)jobs, (const uint8_t*)jobs + bytes);
    ORB_CUDA(cudaMemcpyAsync(W.d, jobs, bytes, cudaMemcpyHostToDevice, st));
    return ORB_OK;
}

// carve-up helper: the same offsets are valid in the pinned slab (inputs / outputs) and the device slab
struct Carver {
    size_t off = 0;
    size_t take(size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; }
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
    if (prm->mode != ORB_MODE_TRACK_LAST && prm->mode != ORB_MODE_LOCAL_POINTS && prm->mode != ORB_MODE_INITIALIZATION) return ORB_ERR_INVALID;
    if (prm->mode != ORB_MODE_LOCAL_POINTS && prm->check_orientation && nq && !q_angle) return ORB_ERR_INVALID;
    if (prm->mode == ORB_MODE_INITIALIZATION && nq > 65535) { orb_set_error("SearchForInitialization: more than 65535 queries"); return ORB_ERR_CAPACITY; }
    for (int i = 0; i < nq; ++i) match_of_query[i] = -1;
    if (target_query) for (int i = 0; i < n; ++i) target_query[i] = -1;
    if (n == 0 || nq == 0) return ORB_OK;
    if (n > GB_MAX_N) { orb_set_error("orb_search_by_projection: more than %d target keypoints", GB_MAX_N); return ORB_ERR_CAPACITY; }
    if (orb_device_count() <= 0) { orb_set_error("no CUDA device visible: liborb_b200 has no CPU fallback"); return ORB_ERR_NO_DEVICE; }
    int npad = 32;
    while (npad < n) npad <<= 1;
    int cand_cap = std::max(nq * 128, 4096);
    for (int attempt = 0; attempt < 2; ++attempt) {
        // ---- layout: [inputs | outputs] mirrored in the pinned and device slabs, then device-only scratch ----
        Carver c;
        const size_t o_kps = c.take(sizeof(orb_kp) * n), o_desc = c.take((size_t)32 * n), o_ur = c.take(sizeof(float) * n);
        const size_t o_qu = c.take(4 * (size_t)nq), o_qv = c.take(4 * (size_t)nq), o_qr = c.take(4 * (size_t)nq);
        const size_t o_qmin = c.take(4 * (size_t)nq), o_qmax = c.take(4 * (size_t)nq), o_qdesc = c.take((size_t)32 * nq);
        const size_t o_qur = c.take(4 * (size_t)nq), o_qer = c.take(4 * (size_t)nq), o_qang = c.take(4 * (size_t)nq);
        const size_t o_qvalid = c.take(nq), o_qobs = c.take(nq), o_taken = c.take(n);
        const size_t in_bytes = c.off;
        const size_t o_moq = c.take(4 * (size_t)nq), o_tq = c.take(4 * (size_t)n), o_scal = c.take(16);
        const size_t io_bytes = c.off;   // outputs: [o_taken .. io_bytes) is copied back (taken + results)
        const size_t o_items = c.take(4 * (size_t)npad), o_cells = c.take(4 * (GRID_COLS * GRID_ROWS + 1));
        const size_t o_cnt = c.take(4 * (size_t)nq), o_base = c.take(4 * (size_t)nq), o_bin = c.take(nq);
        const size_t o_topk = c.take(4 * (size_t)nq * SR_K), o_asg = c.take(4 * (size_t)nq);
        const size_t o_cidx = c.take(4 * (size_t)cand_cap), o_cdist = c.take(2 * (size_t)cand_cap);
        Workspace& W = g_ws;
        int rc = W.prepare(device, c.off, io_bytes);
        if (rc != ORB_OK) return rc;
        cudaStream_t st = W.st;
        uint8_t *H = W.h, *Dv = W.d;
        auto put = [&](size_t off, const void* src, size_t bytes) { if (src) memcpy(H + off, src, bytes); };
        put(o_kps, kps_un, sizeof(orb_kp) * n); put(o_desc, desc, (size_t)32 * n); put(o_ur, u_right, sizeof(float) * n);
        put(o_qu, q_u, 4 * (size_t)nq); put(o_qv, q_v, 4 * (size_t)nq); put(o_qr, q_radius, 4 * (size_t)nq);
        put(o_qmin, q_min_level, 4 * (size_t)nq); put(o_qmax, q_max_level, 4 * (size_t)nq); put(o_qdesc, q_desc, (size_t)32 * nq);
        put(o_qur, q_ur, 4 * (size_t)nq); put(o_qer, q_er_max, 4 * (size_t)nq); put(o_qang, q_angle, 4 * (size_t)nq);
        put(o_qvalid, q_valid, nq); put(o_qobs, q_obs, nq); put(o_taken, taken, n);
        ORB_CUDA(cudaMemcpyAsync(Dv, H, in_bytes, cudaMemcpyHostToDevice, st));
        SearchArgs a;
        memset(&a, 0, sizeof(a));
        a.n = n; a.nq = nq;
        a.kps = (const orb_kp*)(Dv + o_kps); a.desc = Dv + o_desc; a.u_right = u_right ? (const float*)(Dv + o_ur) : nullptr;
        a.q_u = (const float*)(Dv + o_qu); a.q_v = (const float*)(Dv + o_qv); a.q_radius = (const float*)(Dv + o_qr);
        a.q_min_level = (const int*)(Dv + o_qmin); a.q_max_level = (const int*)(Dv + o_qmax); a.q_desc = Dv + o_qdesc;
        a.q_ur = q_ur ? (const float*)(Dv + o_qur) : nullptr; a.q_er_max = q_er_max ? (const float*)(Dv + o_qer) : nullptr;
        a.q_angle = q_angle ? (const float*)(Dv + o_qang) : nullptr;
        a.q_valid = q_valid ? Dv + o_qvalid : nullptr; a.q_obs = q_obs ? Dv + o_qobs : nullptr;
        a.min_x = prm->min_x; a.min_y = prm->min_y;
        a.inv_w = (float)GRID_COLS / (prm->max_x - prm->min_x);   // Frame.cc:162-163
        a.inv_h = (float)GRID_ROWS / (prm->max_y - prm->min_y);
        int* d_scal = (int*)(Dv + o_scal);                        // [0] candidate total, [1] nmatches, [2] overflow
        a.items = (const unsigned*)(Dv + o_items); a.cell_start = (const int*)(Dv + o_cells);
        a.cand_count = (int*)(Dv + o_cnt); a.cand_base = (int*)(Dv + o_base); a.cand_total = d_scal;
        a.cand_cap = cand_cap; a.cand_idx = (int*)(Dv + o_cidx); a.cand_dist = (unsigned short*)(Dv + o_cdist);
        a.topk = (unsigned*)(Dv + o_topk);
        ORB_CUDA(cudaMemsetAsync(d_scal, 0, 16, st));
        grid_build_kernel<<<1, 1024, npad * sizeof(unsigned), st>>>(a.kps, n, npad, a.min_x, a.min_y, a.inv_w, a.inv_h,
                                                                    (unsigned*)(Dv + o_items), (int*)(Dv + o_cells));
        window_candidates_kernel<<<(nq + 7) / 8, 256, 0, st>>>(a);
        // INITIALIZATION: stage + taken[n] + vMatchedDistance[n] + vnMatches21[n]; other modes: two int owner arrays + taken[n]
        const size_t rsmem_max = std::max(sizeof(SrStage) + GB_MAX_N * 5, (size_t)9 * GB_MAX_N + 16);
        size_t rsmem = std::max(sizeof(SrStage) + (size_t)((n + 3) & ~3) * 5, (size_t)9 * ((n + 3) & ~3) + 16);
        if (prm->mode == ORB_MODE_INITIALIZATION)   // taker lists of the parallel rounds: 20 bytes per target when they fit
            rsmem = std::max(rsmem, std::min((size_t)20 * ((n + 3) & ~3), rsmem_max));
        static thread_local int attr_dev = -1;
        if (attr_dev != device) {
            ORB_CUDA(cudaFuncSetAttribute(window_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(sizeof(SrStage) + GB_MAX_N * 5, (size_t)9 * GB_MAX_N + 16)));
            attr_dev = device;
        }
        window_resolve_kernel<<<1, SR_THREADS, rsmem, st>>>(a, prm->mode, prm->th_dist, prm->nn_ratio, prm->check_orientation,
                                                            Dv + o_taken, (int*)(Dv + o_moq), (int*)(Dv + o_tq),
                                                            (signed char*)(Dv + o_bin), (int*)(Dv + o_asg), d_scal + 1, d_scal + 2, (int)rsmem);
        ORB_CUDA(cudaGetLastError());
        ORB_CUDA(cudaMemcpyAsync(H + o_taken, Dv + o_taken, io_bytes - o_taken, cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaStreamSynchronize(st));
        const int* scal = (const int*)(H + o_scal);
        if (scal[2] > 0) {                                        // candidate arena overflow: once more with the exact size
            if (attempt == 1) { orb_set_error("orb_search_by_projection: candidate arena overflow"); return ORB_ERR_CAPACITY; }
            cand_cap = scal[2];
            continue;
        }
        memcpy(taken, H + o_taken, n);
        memcpy(match_of_query, H + o_moq, 4 * (size_t)nq);
        if (target_query) memcpy(target_query, H + o_tq, 4 * (size_t)n);
        *nmatches = scal[1];
        return ORB_OK;
    }
    return ORB_ERR_CAPACITY;
}

int orb_hamming_top2_csr(int device, const uint8_t* q, int nq, const uint8_t* db, int64_t ndb, const int32_t* cand_off,
                         const int32_t* cand_idx, orb_top2* out) {
    if (nq < 0 || ndb < 0 || (nq && (!q || !cand_off || !out))) return ORB_ERR_INVALID;
    if (nq == 0) return ORB_OK;
    const int64_t total = cand_off[nq];
    if (cand_off[0] != 0 || total < 0 || (total && (!cand_idx || !db))) return ORB_ERR_INVALID;
    for (int i = 0; i < nq; ++i) {
        if (cand_off[i + 1] < cand_off[i]) return ORB_ERR_INVALID;
        if (cand_off[i + 1] - cand_off[i] >= (1 << 23)) { orb_set_error("orb_hamming_top2_csr: more than 8M candidates for one query"); return ORB_ERR_CAPACITY; }
    }
    for (int64_t c = 0; c < total; ++c) if (cand_idx[c] < 0 || cand_idx[c] >= ndb) { orb_set_error("orb_hamming_top2_csr: candidate index out of range"); return ORB_ERR_INVALID; }
    if (orb_device_count() <= 0) { orb_set_error("no CUDA device visible: liborb_b200 has no CPU fallback"); return ORB_ERR_NO_DEVICE; }
    Carver c;
    const size_t o_q = c.take((size_t)32 * nq), o_db = c.take((size_t)32 * ndb), o_off = c.take(4 * (size_t)(nq + 1)), o_idx = c.take(4 * (size_t)total);
    const size_t in_bytes = c.off;
    const size_t o_out = c.take(sizeof(orb_top2) * (size_t)nq);
    Workspace& W = g_ws;
    int rc = W.prepare(device, c.off, c.off);
    if (rc != ORB_OK) return rc;
    uint8_t *H = W.h, *Dv = W.d;
    memcpy(H + o_q, q, (size_t)32 * nq);
    if (ndb) memcpy(H + o_db, db, (size_t)32 * ndb);
    memcpy(H + o_off, cand_off, 4 * (size_t)(nq + 1));
    if (total) memcpy(H + o_idx, cand_idx, 4 * (size_t)total);
    ORB_CUDA(cudaMemcpyAsync(Dv, H, in_bytes, cudaMemcpyHostToDevice, W.st));
    csr_top2_kernel<<<(nq + 7) / 8, 256, 0, W.st>>>(Dv + o_q, nq, Dv + o_db, (const int*)(Dv + o_off), (const int*)(Dv + o_idx), (orb_top2*)(Dv + o_out));
    ORB_CUDA(cudaGetLastError());
    ORB_CUDA(cudaMemcpyAsync(H + o_out, Dv + o_out, sizeof(orb_top2) * (size_t)nq, cudaMemcpyDeviceToHost, W.st));
    ORB_CUDA(cudaStreamSynchronize(W.st));
    memcpy(out, H + o_out, sizeof(orb_top2) * (size_t)nq);
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
    Carver c;
    const size_t o_d1 = c.take((size_t)32 * n1), o_d2 = c.take((size_t)32 * n2), o_a1 = c.take(4 * (size_t)n1), o_a2 = c.take(4 * (size_t)n2);
    const size_t in_bytes = c.off;
    const size_t o_m12 = c.take(4 * (size_t)n1), o_nm = c.take(16);
    const size_t io_bytes = c.off;
    const int dpitch = (n2 + 7) & ~7;   // 16-byte aligned rows of the u16 distance matrix
    const size_t o_D = c.take(2 * (size_t)n1 * dpitch), o_topk = c.take(4 * (size_t)n1 * BF_K), o_owner = c.take(4 * (size_t)n2), o_owner2 = c.take(4 * (size_t)n2), o_bin = c.take(n1);
    Workspace& W = g_ws;
    int rc = W.prepare(device, c.off, io_bytes);
    if (rc != ORB_OK) return rc;
    cudaStream_t st = W.st;
    uint8_t *H = W.h, *Dv = W.d;
    memcpy(H + o_d1, desc1, (size_t)32 * n1); memcpy(H + o_d2, desc2, (size_t)32 * n2);
    if (angle1) memcpy(H + o_a1, angle1, 4 * (size_t)n1);
    if (angle2) memcpy(H + o_a2, angle2, 4 * (size_t)n2);
    ORB_CUDA(cudaMemcpyAsync(Dv, H, in_bytes, cudaMemcpyHostToDevice, st));
    bf_rows_kernel<<<(n1 + 7) / 8, 256, 0, st>>>(Dv + o_d1, n1, Dv + o_d2, n2, (unsigned short*)(Dv + o_D), dpitch, (unsigned*)(Dv + o_topk),
                                                 (unsigned)(bf_track_limit(th_dist, nn_ratio) + 1) << 16);   // (unpruned lists: 101 instead of 89 us per call)
    {
        const size_t own_smem = 8 * (size_t)n2;   // two owner arrays in shared memory when they fit, else in the workspace
        static bool attr_set[64] = {};   // per device
        if (!attr_set[device & 63]) {
            ORB_CUDA(cudaFuncSetAttribute(bf_resolve_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr_set[device & 63] = true;
        }
        if (own_smem <= 200 * 1024)
            bf_resolve_kernel<true><<<1, SR_THREADS, own_smem, st>>>(Dv + o_d1, Dv + o_d2, (const unsigned short*)(Dv + o_D), dpitch, (const unsigned*)(Dv + o_topk), n1, n2,
                                                                     (const float*)(Dv + o_a1), (const float*)(Dv + o_a2), th_dist, nn_ratio, check_orientation,
                                                                     (int*)(Dv + o_owner), (int*)(Dv + o_owner2), (int*)(Dv + o_m12),
                                                                     (signed char*)(Dv + o_bin), (int*)(Dv + o_nm));
        else
            bf_resolve_kernel<false><<<1, SR_THREADS, 0, st>>>(Dv + o_d1, Dv + o_d2, (const unsigned short*)(Dv + o_D), dpitch, (const unsigned*)(Dv + o_topk), n1, n2,
                                                               (const float*)(Dv + o_a1), (const float*)(Dv + o_a2), th_dist, nn_ratio, check_orientation,
                                                               (int*)(Dv + o_owner), (int*)(Dv + o_owner2), (int*)(Dv + o_m12),
                                                               (signed char*)(Dv + o_bin), (int*)(Dv + o_nm));
    }
    ORB_CUDA(cudaGetLastError());
    ORB_CUDA(cudaMemcpyAsync(H + o_m12, Dv + o_m12, io_bytes - o_m12, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaStreamSynchronize(st));
    memcpy(match12, H + o_m12, 4 * (size_t)n1);
    *nmatches = *(const int*)(H + o_nm);
    return ORB_OK;
}

int orb_search_by_bow(int device, const uint8_t* desc1, const float* angle1, const uint8_t* valid1, int n1, const int32_t* fv1_node,
                      const int32_t* fv1_start, const int32_t* fv1_feat, int nfv1, const uint8_t* desc2, const float* angle2,
                      const uint8_t* valid2, int n2, const int32_t* fv2_node, const int32_t* fv2_start, const int32_t* fv2_feat, int nfv2,
                      int th_dist, int strict, float nn_ratio, int check_orientation, int32_t* match12, int32_t* match21, int* nmatches) {
    if (n1 < 0 || n2 < 0 || nfv1 < 0 || nfv2 < 0 || !nmatches || (n1 && (!desc1 || !match12)) || (n2 && (!desc2 || !match21))) return ORB_ERR_INVALID;
    if ((nfv1 && (!fv1_node || !fv1_start || !fv1_feat)) || (nfv2 && (!fv2_node || !fv2_start || !fv2_feat))) return ORB_ERR_INVALID;
    if (check_orientation && ((n1 && !angle1) || (n2 && !angle2))) return ORB_ERR_INVALID;
    *nmatches = 0;
    for (int i = 0; i < n1; ++i) match12[i] = -1;
    for (int j = 0; j < n2; ++j) match21[j] = -1;
    if (n1 == 0 || n2 == 0 || nfv1 == 0 || nfv2 == 0) return ORB_OK;
    const int nf1 = fv1_start[nfv1], nf2 = fv2_start[nfv2];
    if (nf1 < 0 || nf1 > n1 || nf2 < 0 || nf2 > n2) { orb_set_error("orb_search_by_bow: feature vector larger than the keypoint set"); return ORB_ERR_INVALID; }
    for (int i = 0; i < nf1; ++i) if (fv1_feat[i] < 0 || fv1_feat[i] >= n1) { orb_set_error("orb_search_by_bow: fv1_feat out of range"); return ORB_ERR_INVALID; }
    for (int i = 0; i < nf2; ++i) if (fv2_feat[i] < 0 || fv2_feat[i] >= n2) { orb_set_error("orb_search_by_bow: fv2_feat out of range"); return ORB_ERR_INVALID; }
    for (int j = 0; j < nfv2; ++j) if (fv2_start[j + 1] - fv2_start[j] > 65535) { orb_set_error("orb_search_by_bow: more than 65535 features in one node"); return ORB_ERR_CAPACITY; }
    if (orb_device_count() <= 0) { orb_set_error("no CUDA device visible: liborb_b200 has no CPU fallback"); return ORB_ERR_NO_DEVICE; }
    Carver c;
    const size_t o_d1 = c.take((size_t)32 * n1), o_d2 = c.take((size_t)32 * n2), o_a1 = c.take(4 * (size_t)n1), o_a2 = c.take(4 * (size_t)n2);
    const size_t o_v1 = c.take(n1), o_v2 = c.take(n2);
    const size_t o_n1 = c.take(4 * (size_t)nfv1), o_s1 = c.take(4 * ((size_t)nfv1 + 1)), o_f1 = c.take(4 * (size_t)std::max(nf1, 1));
    const size_t o_n2 = c.take(4 * (size_t)nfv2), o_s2 = c.take(4 * ((size_t)nfv2 + 1)), o_f2 = c.take(4 * (size_t)std::max(nf2, 1));
    const size_t o_m12 = c.take(4 * (size_t)n1), o_m21 = c.take(4 * (size_t)n2);   // initialised to -1 on the host: part of the upload
    const size_t in_bytes = c.off;
    const size_t o_nm = c.take(16);
    const size_t io_bytes = c.off;
    Workspace& W = g_ws;
    int rc = W.prepare(device, io_bytes, io_bytes);
    if (rc != ORB_OK) return rc;
    cudaStream_t st = W.st;
    uint8_t *H = W.h, *Dv = W.d;
    memcpy(H + o_d1, desc1, (size_t)32 * n1); memcpy(H + o_d2, desc2, (size_t)32 * n2);
    if (angle1) memcpy(H + o_a1, angle1, 4 * (size_t)n1);
    if (angle2) memcpy(H + o_a2, angle2, 4 * (size_t)n2);
    if (valid1) memcpy(H + o_v1, valid1, n1);
    if (valid2) memcpy(H + o_v2, valid2, n2);
    memcpy(H + o_n1, fv1_node, 4 * (size_t)nfv1); memcpy(H + o_s1, fv1_start, 4 * ((size_t)nfv1 + 1)); memcpy(H + o_f1, fv1_feat, 4 * (size_t)nf1);
    memcpy(H + o_n2, fv2_node, 4 * (size_t)nfv2); memcpy(H + o_s2, fv2_start, 4 * ((size_t)nfv2 + 1)); memcpy(H + o_f2, fv2_feat, 4 * (size_t)nf2);
    memset(H + o_m12, 0xFF, 4 * (size_t)n1); memset(H + o_m21, 0xFF, 4 * (size_t)n2);
    ORB_CUDA(cudaMemcpyAsync(Dv, H, in_bytes, cudaMemcpyHostToDevice, st));
    bow_match_kernel<<<(nfv1 + 7) / 8, 256, 0, st>>>(Dv + o_d1, valid1 ? Dv + o_v1 : nullptr, (const int*)(Dv + o_n1), (const int*)(Dv + o_s1),
                                                     (const int*)(Dv + o_f1), nfv1, Dv + o_d2, valid2 ? Dv + o_v2 : nullptr, (const int*)(Dv + o_n2),
                                                     (const int*)(Dv + o_s2), (const int*)(Dv + o_f2), nfv2, th_dist, strict, nn_ratio,
                                                     (int*)(Dv + o_m12), (int*)(Dv + o_m21));
    bow_rot_kernel<<<1, 1024, 0, st>>>((const float*)(Dv + o_a1), (const float*)(Dv + o_a2), n1, check_orientation, (int*)(Dv + o_m12),
                                       (int*)(Dv + o_m21), (int*)(Dv + o_nm));
    ORB_CUDA(cudaGetLastError());
    ORB_CUDA(cudaMemcpyAsync(H + o_m12, Dv + o_m12, io_bytes - o_m12, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaStreamSynchronize(st));
    memcpy(match12, H + o_m12, 4 * (size_t)n1);
    memcpy(match21, H + o_m21, 4 * (size_t)n2);
    *nmatches = *(const int*)(H + o_nm);
    return ORB_OK;
}

int orb_search_for_triangulation(int device, const orb_kp* kps1, const uint8_t* desc1, const uint8_t* has_mp1, const float* u_right1, int n1,
                                 const int32_t* fv1_node, const int32_t* fv1_start, const int32_t* fv1_feat, int nfv1, const orb_kp* kps2,
                                 const uint8_t* desc2, const uint8_t* has_mp2, const float* u_right2, int n2, const int32_t* fv2_node,
                                 const int32_t* fv2_start, const int32_t* fv2_feat, int nfv2, const float* F12, float ex, float ey,
                                 const float* scale_factors, const float* level_sigma2, int nlevels, int only_stereo, int check_orientation,
                                 int32_t* match12, int* nmatches) {
    if (n1 < 0 || n2 < 0 || nfv1 < 0 || nfv2 < 0 || !nmatches || !F12 || !scale_factors || !level_sigma2) return ORB_ERR_INVALID;
    if ((n1 && (!kps1 || !desc1 || !match12)) || (n2 && (!kps2 || !desc2))) return ORB_ERR_INVALID;
    if ((nfv1 && (!fv1_node || !fv1_start || !fv1_feat)) || (nfv2 && (!fv2_node || !fv2_start || !fv2_feat))) return ORB_ERR_INVALID;
    if (nlevels < 1 || nlevels > ORB_MAX_LEVELS) { orb_set_error("orb_search_for_triangulation: 1..%d levels", ORB_MAX_LEVELS); return ORB_ERR_INVALID; }
    *nmatches = 0;
    for (int i = 0; i < n1; ++i) match12[i] = -1;
    if (n1 == 0 || n2 == 0 || nfv1 == 0 || nfv2 == 0) return ORB_OK;
    const int nf1 = fv1_start[nfv1], nf2 = fv2_start[nfv2];
    if (nf1 < 0 || nf1 > n1 || nf2 < 0 || nf2 > n2) { orb_set_error("orb_search_for_triangulation: feature vector larger than the keypoint set"); return ORB_ERR_INVALID; }
    for (int i = 0; i < nf1; ++i) if (fv1_feat[i] < 0 || fv1_feat[i] >= n1) { orb_set_error("orb_search_for_triangulation: fv1_feat out of range"); return ORB_ERR_INVALID; }
    for (int i = 0; i < nf2; ++i) if (fv2_feat[i] < 0 || fv2_feat[i] >= n2) { orb_set_error("orb_search_for_triangulation: fv2_feat out of range"); return ORB_ERR_INVALID; }
    for (int j = 0; j < nfv2; ++j) if (fv2_start[j + 1] - fv2_start[j] > 65535) { orb_set_error("orb_search_for_triangulation: more than 65535 features in one node"); return ORB_ERR_CAPACITY; }
    for (int j = 0; j < n2; ++j) if (kps2[j].octave < 0 || kps2[j].octave >= nlevels) { orb_set_error("orb_search_for_triangulation: octave out of range"); return ORB_ERR_INVALID; }
    if (orb_device_count() <= 0) { orb_set_error("no CUDA device visible: liborb_b200 has no CPU fallback"); return ORB_ERR_NO_DEVICE; }
    Carver c;
    const size_t o_k1 = c.take(sizeof(orb_kp) * (size_t)n1), o_k2 = c.take(sizeof(orb_kp) * (size_t)n2);
    const size_t o_d1 = c.take((size_t)32 * n1), o_d2 = c.take((size_t)32 * n2);
    const size_t o_h1 = c.take(n1), o_h2 = c.take(n2), o_u1 = c.take(4 * (size_t)n1), o_u2 = c.take(4 * (size_t)n2);
    const size_t o_a1 = c.take(4 * (size_t)n1), o_a2 = c.take(4 * (size_t)n2);     // dense angles for the rotation filter
    const size_t o_n1 = c.take(4 * (size_t)nfv1), o_s1 = c.take(4 * ((size_t)nfv1 + 1)), o_f1 = c.take(4 * (size_t)std::max(nf1, 1));
    const size_t o_n2 = c.take(4 * (size_t)nfv2), o_s2 = c.take(4 * ((size_t)nfv2 + 1)), o_f2 = c.take(4 * (size_t)std::max(nf2, 1));
    const size_t o_m12 = c.take(4 * (size_t)n1);
    const size_t in_bytes = c.off;
    const size_t o_nm = c.take(16);
    const size_t io_bytes = c.off;
    Workspace& W = g_ws;
    int rc = W.prepare(device, io_bytes, io_bytes);
    if (rc != ORB_OK) return rc;
    cudaStream_t st = W.st;
    uint8_t *H = W.h, *Dv = W.d;
    memcpy(H + o_k1, kps1, sizeof(orb_kp) * (size_t)n1); memcpy(H + o_k2, kps2, sizeof(orb_kp) * (size_t)n2);
    memcpy(H + o_d1, desc1, (size_t)32 * n1); memcpy(H + o_d2, desc2, (size_t)32 * n2);
    if (has_mp1) memcpy(H + o_h1, has_mp1, n1);
    if (has_mp2) memcpy(H + o_h2, has_mp2, n2);
    if (u_right1) memcpy(H + o_u1, u_right1, 4 * (size_t)n1);
    if (u_right2) memcpy(H + o_u2, u_right2, 4 * (size_t)n2);
    for (int i = 0; i < n1; ++i) reinterpret_cast<float*>(H + o_a1)[i] = kps1[i].angle;
    for (int j = 0; j < n2; ++j) reinterpret_cast<float*>(H + o_a2)[j] = kps2[j].angle;
    memcpy(H + o_n1, fv1_node, 4 * (size_t)nfv1); memcpy(H + o_s1, fv1_start, 4 * ((size_t)nfv1 + 1)); memcpy(H + o_f1, fv1_feat, 4 * (size_t)nf1);
    memcpy(H + o_n2, fv2_node, 4 * (size_t)nfv2); memcpy(H + o_s2, fv2_start, 4 * ((size_t)nfv2 + 1)); memcpy(H + o_f2, fv2_feat, 4 * (size_t)nf2);
    memset(H + o_m12, 0xFF, 4 * (size_t)n1);
    ORB_CUDA(cudaMemcpyAsync(Dv, H, in_bytes, cudaMemcpyHostToDevice, st));
    TriFrame A{(const orb_kp*)(Dv + o_k1), Dv + o_d1, has_mp1 ? Dv + o_h1 : nullptr, u_right1 ? (const float*)(Dv + o_u1) : nullptr,
               (const int*)(Dv + o_n1), (const int*)(Dv + o_s1), (const int*)(Dv + o_f1), nfv1};
    TriFrame B{(const orb_kp*)(Dv + o_k2), Dv + o_d2, has_mp2 ? Dv + o_h2 : nullptr, u_right2 ? (const float*)(Dv + o_u2) : nullptr,
               (const int*)(Dv + o_n2), (const int*)(Dv + o_s2), (const int*)(Dv + o_f2), nfv2};
    TriParams P;
    memset(&P, 0, sizeof(P));
    memcpy(P.F, F12, sizeof(P.F));
    P.ex = ex; P.ey = ey; P.nlevels = nlevels; P.only_stereo = only_stereo; P.th_low = 50;   // ORBmatcher::TH_LOW
    for (int l = 0; l < nlevels; ++l) { P.scale[l] = scale_factors[l]; P.sigma2[l] = level_sigma2[l]; }
    bow_triangulation_kernel<<<(nfv1 + 7) / 8, 256, 0, st>>>(A, B, P, (int*)(Dv + o_m12));
    bow_rot_kernel<<<1, 1024, 0, st>>>((const float*)(Dv + o_a1), (const float*)(Dv + o_a2), n1, check_orientation, (int*)(Dv + o_m12), nullptr,
                                       (int*)(Dv + o_nm));
    ORB_CUDA(cudaGetLastError());
    ORB_CUDA(cudaMemcpyAsync(H + o_m12, Dv + o_m12, io_bytes - o_m12, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaStreamSynchronize(st));
    memcpy(match12, H + o_m12, 4 * (size_t)n1);
    *nmatches = *(const int*)(H + o_nm);
    return ORB_OK;
}

int orb_search_by_sim3(int device, const orb_kp* kps1_un, const uint8_t* desc1, int n1, const float* bounds1, const orb_kp* kps2_un,
                       const uint8_t* desc2, int n2, const float* bounds2, const float* q12_u, const float* q12_v, const float* q12_radius,
                       const int32_t* q12_level, const uint8_t* q12_desc, const uint8_t* q12_valid, const float* q21_u, const float* q21_v,
                       const float* q21_radius, const int32_t* q21_level, const uint8_t* q21_desc, const uint8_t* q21_valid, int th_dist,
                       int32_t* match12, int* nfound) {
    if (n1 < 0 || n2 < 0 || !nfound || !bounds1 || !bounds2) return ORB_ERR_INVALID;
    if (n1 && (!kps1_un || !desc1 || !q12_u || !q12_v || !q12_radius || !q12_level || !q12_desc || !match12)) return ORB_ERR_INVALID;
    if (n2 && (!kps2_un || !desc2 || !q21_u || !q21_v || !q21_radius || !q21_level || !q21_desc)) return ORB_ERR_INVALID;
    *nfound = 0;
    for (int i = 0; i < n1; ++i) match12[i] = -1;
    if (n1 == 0 || n2 == 0) return ORB_OK;
    // Each direction (ORBmatcher.cc:1150-1213, 1216-1279) is a windowed best-only search without any "already matched"
    // state: the TRACK_LAST walk with no target taken and no query blocking its target (q_obs = 0), octave band
    // [pred - 1, pred], no rotation histogram.
    std::vector<int32_t> m1(n1, -1), m2(n2, -1), lo, hi;
    std::vector<uint8_t> taken, zeros;
    auto one_way = [&](const orb_kp* kps, const uint8_t* desc, int n, const float* b, int nq, const float* u, const float* v, const float* r,
                       const int32_t* lvl, const uint8_t* qd, const uint8_t* valid, int32_t* out) -> int {
        orb_search_params prm;
        prm.mode = ORB_MODE_TRACK_LAST; prm.th_dist = th_dist; prm.nn_ratio = 1.0f; prm.check_orientation = 0;
        prm.min_x = b[0]; prm.min_y = b[1]; prm.max_x = b[2]; prm.max_y = b[3];
        lo.resize(nq); hi.resize(nq);
        for (int i = 0; i < nq; ++i) { lo[i] = lvl[i] - 1; hi[i] = lvl[i]; }
        taken.assign(n, 0); zeros.assign(nq, 0);
        int nm = 0;
        return orb_search_by_projection(device, &prm, kps, desc, nullptr, n, taken.data(), nq, u, v, r, lo.data(), hi.data(), qd, nullptr, nullptr,
                                        nullptr, valid, zeros.data(), out, nullptr, &nm);
    };
    int rc = one_way(kps2_un, desc2, n2, bounds2, n1, q12_u, q12_v, q12_radius, q12_level, q12_desc, q12_valid, m1.data());
    if (rc != ORB_OK) return rc;
    rc = one_way(kps1_un, desc1, n1, bounds1, n2, q21_u, q21_v, q21_radius, q21_level, q21_desc, q21_valid, m2.data());
    if (rc != ORB_OK) return rc;
    int found = 0;
    for (int i1 = 0; i1 < n1; ++i1) {                      // agreement check, ORBmatcher.cc:1282-1299
        const int idx2 = m1[i1];
        if (idx2 >= 0 && m2[idx2] == i1) { match12[i1] = idx2; ++found; }
    }
    *nfound = found;
    return ORB_OK;
}

int orb_distinctive_descriptors(int device, const uint8_t* desc32, const int32_t* off, int npoints, int32_t* best_idx, uint8_t* best_desc32) {
    if (npoints < 0 || !off || (npoints && !best_idx)) return ORB_ERR_INVALID;
    if (npoints == 0) return ORB_OK;
    if (off[0] != 0) { orb_set_error("orb_distinctive_descriptors: off[0] must be 0"); return ORB_ERR_INVALID; }
    for (int p = 0; p < npoints; ++p) {
        if (off[p + 1] < off[p]) { orb_set_error("orb_distinctive_descriptors: off must be non-decreasing"); return ORB_ERR_INVALID; }
        if (off[p + 1] - off[p] > 65535) { orb_set_error("orb_distinctive_descriptors: more than 65535 observations of one point"); return ORB_ERR_CAPACITY; }
    }
    const int total = off[npoints];
    if (total && !desc32) return ORB_ERR_INVALID;
    if (orb_device_count() <= 0) { orb_set_error("no CUDA device visible: liborb_b200 has no CPU fallback"); return ORB_ERR_NO_DEVICE; }
    Carver c;
    const size_t o_d = c.take((size_t)32 * std::max(total, 1)), o_off = c.take(4 * ((size_t)npoints + 1));
    const size_t in_bytes = c.off;
    const size_t o_best = c.take(4 * (size_t)npoints);
    Workspace& W = g_ws;
    int rc = W.prepare(device, c.off, c.off);
    if (rc != ORB_OK) return rc;
    if (total) memcpy(W.h + o_d, desc32, (size_t)32 * total);
    memcpy(W.h + o_off, off, 4 * ((size_t)npoints + 1));
    ORB_CUDA(cudaMemcpyAsync(W.d, W.h, in_bytes, cudaMemcpyHostToDevice, W.st));
    distinctive_kernel<<<npoints, DD_WARPS * 32, 0, W.st>>>(W.d + o_d, (const int*)(W.d + o_off), (int*)(W.d + o_best));
    ORB_CUDA(cudaGetLastError());
    ORB_CUDA(cudaMemcpyAsync(W.h + o_best, W.d + o_best, 4 * (size_t)npoints, cudaMemcpyDeviceToHost, W.st));
    ORB_CUDA(cudaStreamSynchronize(W.st));
    memcpy(best_idx, W.h + o_best, 4 * (size_t)npoints);
    if (best_desc32)
        for (int p = 0; p < npoints; ++p)
            if (best_idx[p] >= 0) memcpy(best_desc32 + (size_t)p * 32, desc32 + (size_t)(off[p] + best_idx[p]) * 32, 32);
    return ORB_OK;
}

int orb_fuse_search(int device, const orb_kp* kps_un, const uint8_t* desc, const float* u_right, int n, const float* bounds,
                    const float* inv_level_sigma2, int nlevels, int nq, const float* q_u, const float* q_v, const float* q_ur,
                    const float* q_radius, const int32_t* q_level, const uint8_t* q_desc, const uint8_t* q_valid, int32_t* best_idx,
                    int32_t* best_dist) {
    if (n < 0 || nq < 0 || !bounds || (nq && (!q_u || !q_v || !q_radius || !q_level || !q_desc || !best_idx || !best_dist))) return ORB_ERR_INVALID;
    if (n && (!kps_un || !desc)) return ORB_ERR_INVALID;
    if (inv_level_sigma2 && (nlevels < 1 || nlevels > ORB_MAX_LEVELS)) { orb_set_error("orb_fuse_search: 1..%d levels", ORB_MAX_LEVELS); return ORB_ERR_INVALID; }
    if (inv_level_sigma2 && u_right && !q_ur) return ORB_ERR_INVALID;
    for (int i = 0; i < nq; ++i) { best_idx[i] = -1; best_dist[i] = 256; }
    if (n == 0 || nq == 0) return ORB_OK;
    if (n > GB_MAX_N) { orb_set_error("orb_fuse_search: more than %d keypoints", GB_MAX_N); return ORB_ERR_CAPACITY; }
    if (orb_device_count() <= 0) { orb_set_error("no CUDA device visible: liborb_b200 has no CPU fallback"); return ORB_ERR_NO_DEVICE; }
    int npad = 32;
    while (npad < n) npad <<= 1;
    Carver c;
    const size_t o_kps = c.take(sizeof(orb_kp) * (size_t)n), o_desc = c.take((size_t)32 * n), o_ur = c.take(4 * (size_t)n), o_sig = c.take(4 * ORB_MAX_LEVELS);
    const size_t o_qu = c.take(4 * (size_t)nq), o_qv = c.take(4 * (size_t)nq), o_qur = c.take(4 * (size_t)nq), o_qr = c.take(4 * (size_t)nq);
    const size_t o_ql = c.take(4 * (size_t)nq), o_qd = c.take((size_t)32 * nq), o_qval = c.take(nq);
    const size_t in_bytes = c.off;
    const size_t o_bi = c.take(4 * (size_t)nq), o_bd = c.take(4 * (size_t)nq);
    const size_t io_bytes = c.off;
    const size_t o_items = c.take(4 * (size_t)npad), o_cells = c.take(4 * (GRID_COLS * GRID_ROWS + 1));
    Workspace& W = g_ws;
    int rc = W.prepare(device, c.off, io_bytes);
    if (rc != ORB_OK) return rc;
    cudaStream_t st = W.st;
    uint8_t *H = W.h, *Dv = W.d;
    auto put = [&](size_t off, const void* src, size_t bytes) { if (src) memcpy(H + off, src, bytes); };
    put(o_kps, kps_un, sizeof(orb_kp) * (size_t)n); put(o_desc, desc, (size_t)32 * n); put(o_ur, u_right, 4 * (size_t)n);
    put(o_sig, inv_level_sigma2, 4 * (size_t)(inv_level_sigma2 ? nlevels : 0));
    put(o_qu, q_u, 4 * (size_t)nq); put(o_qv, q_v, 4 * (size_t)nq); put(o_qur, q_ur, 4 * (size_t)nq); put(o_qr, q_radius, 4 * (size_t)nq);
    put(o_ql, q_level, 4 * (size_t)nq); put(o_qd, q_desc, (size_t)32 * nq); put(o_qval, q_valid, nq);
    ORB_CUDA(cudaMemcpyAsync(Dv, H, in_bytes, cudaMemcpyHostToDevice, st));
    FuseArgs a;
    memset(&a, 0, sizeof(a));
    a.kps = (const orb_kp*)(Dv + o_kps); a.desc = Dv + o_desc; a.u_right = u_right ? (const float*)(Dv + o_ur) : nullptr;
    a.inv_sigma2 = inv_level_sigma2 ? (const float*)(Dv + o_sig) : nullptr; a.nlevels = nlevels;
    a.nq = nq; a.q_u = (const float*)(Dv + o_qu); a.q_v = (const float*)(Dv + o_qv); a.q_ur = q_ur ? (const float*)(Dv + o_qur) : nullptr;
    a.q_radius = (const float*)(Dv + o_qr); a.q_level = (const int*)(Dv + o_ql); a.q_desc = Dv + o_qd; a.q_valid = q_valid ? Dv + o_qval : nullptr;
    a.min_x = bounds[0]; a.min_y = bounds[1];
    a.inv_w = (float)GRID_COLS / (bounds[2] - bounds[0]);   // Frame.cc:162-163 / KeyFrame.cc:43-44
    a.inv_h = (float)GRID_ROWS / (bounds[3] - bounds[1]);
    a.items = (const unsigned*)(Dv + o_items); a.cell_start = (const int*)(Dv + o_cells);
    a.best_idx = (int*)(Dv + o_bi); a.best_dist = (int*)(Dv + o_bd);
    grid_build_kernel<<<1, 1024, npad * sizeof(unsigned), st>>>(a.kps, n, npad, a.min_x, a.min_y, a.inv_w, a.inv_h, (unsigned*)(Dv + o_items),
                                                                (int*)(Dv + o_cells));
    fuse_search_kernel<<<(nq + 7) / 8, 256, 0, st>>>(a);
    ORB_CUDA(cudaGetLastError());
    ORB_CUDA(cudaMemcpyAsync(H + o_bi, Dv + o_bi, io_bytes - o_bi, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaStreamSynchronize(st));
    memcpy(best_idx, H + o_bi, 4 * (size_t)nq);
    memcpy(best_dist, H + o_bd, 4 * (size_t)nq);
    return ORB_OK;
}


/* ---- batched, device-resident pair matching (BASELINE config 2) ------------------------------------ */
int orb_match_bruteforce_batch_device(int device, int npairs, const orb_kp* d_kps1, const uint8_t* d_desc1, const int32_t* d_n1, int cap1,
                                      const orb_kp* d_kps2, const uint8_t* d_desc2, const int32_t* d_n2, int cap2, int th_dist, float nn_ratio,
                                      int check_orientation, int32_t* d_match12, int32_t* d_nmatches, void* cuda_stream) {
    if (npairs < 0 || cap1 <= 0 || cap2 <= 0 || !d_kps1 || !d_desc1 || !d_n1 || !d_kps2 || !d_desc2 || !d_n2 || !d_match12 || !d_nmatches)
        return ORB_ERR_INVALID;
    if (npairs == 0) return ORB_OK;
    if (cap2 > 65535) { orb_set_error("orb_match_bruteforce_batch_device: more than 65535 targets"); return ORB_ERR_CAPACITY; }
    if (orb_device_count() <= 0) { orb_set_error("no CUDA device visible: liborb_b200 has no CPU fallback"); return ORB_ERR_NO_DEVICE; }
    Carver c;   // per-pair scratch
    const size_t o_topk = c.take(4 * (size_t)cap1 * BF_K), o_a1 = c.take(4 * (size_t)cap1), o_a2 = c.take(4 * (size_t)cap2);
    const size_t o_owner = c.take(4 * (size_t)cap2), o_owner2 = c.take(4 * (size_t)cap2), o_bin = c.take(cap1);
    const size_t per_pair = c.off;
    const size_t jobs_bytes = ((size_t)npairs * sizeof(BfJob) + 255) & ~(size_t)255;
    Workspace& W = g_ws_batch[0];
    int rc = W.prepare(device, jobs_bytes + per_pair * (size_t)npairs, 256, true);
    if (rc != ORB_OK) return rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    std::vector<BfJob> jobs(npairs);
    for (int p = 0; p < npairs; ++p) {
        uint8_t* S = W.d + jobs_bytes + per_pair * (size_t)p;
        BfJob& J = jobs[p];
        J.d1 = d_desc1 + (size_t)p * cap1 * 32; J.d2 = d_desc2 + (size_t)p * cap2 * 32;
        J.k1 = d_kps1 + (size_t)p * cap1; J.k2 = d_kps2 + (size_t)p * cap2;
        J.n1_ptr = d_n1 + p; J.n2_ptr = d_n2 + p; J.cap1 = cap1; J.cap2 = cap2;
        J.topk = (unsigned*)(S + o_topk); J.a1 = (float*)(S + o_a1); J.a2 = (float*)(S + o_a2);
        J.owner = (int*)(S + o_owner); J.owner2 = (int*)(S + o_owner2); J.m12 = d_match12 + (size_t)p * cap1; J.bin = (signed char*)(S + o_bin);
        J.nm = d_nmatches + p;
    }
    rc = upload_jobs(W, jobs.data(), (size_t)npairs * sizeof(BfJob), st);
    if (rc != ORB_OK) return rc;
    const BfJob* d_jobs = (const BfJob*)W.d;
    const unsigned track_key = (unsigned)(bf_track_limit(th_dist, nn_ratio) + 1) << 16;
    bf_rows_batch_kernel<<<dim3((cap1 + BFT_THREADS - 1) / BFT_THREADS, npairs), BFT_THREADS, 0, st>>>(d_jobs, track_key, 1u << 16, 2u << 16, 4u << 16);
    {
        const size_t own_smem = 8 * (size_t)cap2;   // the two owner arrays of a pair in shared memory when they fit, else in the workspace
        static bool attr_set[64] = {};   // per device
        if (!attr_set[device & 63]) {
            ORB_CUDA(cudaFuncSetAttribute(bf_resolve_batch_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr_set[device & 63] = true;
        }
        if (own_smem <= 200 * 1024) bf_resolve_batch_kernel<true><<<npairs, SR_THREADS, own_smem, st>>>(d_jobs, th_dist, nn_ratio, check_orientation);
        else bf_resolve_batch_kernel<false><<<npairs, SR_THREADS, 0, st>>>(d_jobs, th_dist, nn_ratio, check_orientation);
    }
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

int orb_search_by_projection_batch_device(int device, const orb_search_params* prm, int npairs, const orb_search_batch* b, void* cuda_stream) {
    if (!prm || !b || npairs < 0 || b->cap_n <= 0 || b->cap_q <= 0) return ORB_ERR_INVALID;
    if (!b->d_kps_un || !b->d_desc || !b->d_n || !b->d_taken || !b->d_nq || !b->d_q_u || !b->d_q_v || !b->d_q_radius || !b->d_q_min_level ||
        !b->d_q_max_level || !b->d_q_desc || !b->d_match_of_query || !b->d_target_query || !b->d_nmatches)
        return ORB_ERR_INVALID;
    if (b->d_u_right && (!b->d_q_ur || !b->d_q_er_max)) return ORB_ERR_INVALID;
    if (prm->mode != ORB_MODE_TRACK_LAST && prm->mode != ORB_MODE_LOCAL_POINTS) {
        orb_set_error("orb_search_by_projection_batch_device: modes TRACK_LAST and LOCAL_POINTS only");
        return ORB_ERR_INVALID;
    }
    if (prm->mode == ORB_MODE_TRACK_LAST && prm->check_orientation && !b->d_q_angle) return ORB_ERR_INVALID;
    if (npairs == 0) return ORB_OK;
    const int n = b->cap_n, nq = b->cap_q;
    if (n > GB_MAX_N) { orb_set_error("orb_search_by_projection_batch_device: more than %d target keypoints", GB_MAX_N); return ORB_ERR_CAPACITY; }
    if (orb_device_count() <= 0) { orb_set_error("no CUDA device visible: liborb_b200 has no CPU fallback"); return ORB_ERR_NO_DEVICE; }
    int npad = 32;
    while (npad < n) npad <<= 1;
    const int cand_cap = std::max(nq * 160, 4096);
    Carver c;   // per-pair scratch
    const size_t o_items = c.take(4 * (size_t)npad), o_cells = c.take(4 * (GRID_COLS * GRID_ROWS + 1));
    const size_t o_cnt = c.take(4 * (size_t)nq), o_base = c.take(4 * (size_t)nq), o_bin = c.take(nq);
    const size_t o_topk = c.take(4 * (size_t)nq * SR_K), o_asg = c.take(4 * (size_t)nq), o_scal = c.take(16);
    const size_t o_cidx = c.take(4 * (size_t)cand_cap), o_cdist = c.take(2 * (size_t)cand_cap);
    const size_t per_pair = c.off;
    const size_t jobs_bytes = ((size_t)npairs * sizeof(WindowJob) + 255) & ~(size_t)255;
    Workspace& W = g_ws_batch[1];
    int rc = W.prepare(device, jobs_bytes + per_pair * (size_t)npairs, 256, true);
    if (rc != ORB_OK) return rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    std::vector<WindowJob> jobs(npairs);
    for (int p = 0; p < npairs; ++p) {
        uint8_t* S = W.d + jobs_bytes + per_pair * (size_t)p;
        WindowJob& J = jobs[p];
        memset(&J, 0, sizeof(J));
        SearchArgs& a = J.a;
        const size_t on = (size_t)p * n, oq = (size_t)p * nq;
        a.n = n; a.nq = nq;
        a.kps = b->d_kps_un + on; a.desc = b->d_desc + on * 32; a.u_right = b->d_u_right ? b->d_u_right + on : nullptr;
        a.q_u = b->d_q_u + oq; a.q_v = b->d_q_v + oq; a.q_radius = b->d_q_radius + oq;
        a.q_min_level = b->d_q_min_level + oq; a.q_max_level = b->d_q_max_level + oq; a.q_desc = b->d_q_desc + oq * 32;
        a.q_ur = b->d_q_ur ? b->d_q_ur + oq : nullptr; a.q_er_max = b->d_q_er_max ? b->d_q_er_max + oq : nullptr;
        a.q_angle = b->d_q_angle ? b->d_q_angle + oq : nullptr;
        a.q_valid = b->d_q_valid ? b->d_q_valid + oq : nullptr; a.q_obs = b->d_q_obs ? b->d_q_obs + oq : nullptr;
        a.min_x = prm->min_x; a.min_y = prm->min_y;
        a.inv_w = (float)GRID_COLS / (prm->max_x - prm->min_x);   // Frame.cc:162-163
        a.inv_h = (float)GRID_ROWS / (prm->max_y - prm->min_y);
        int* scal = (int*)(S + o_scal);
        a.items = (const unsigned*)(S + o_items); a.cell_start = (const int*)(S + o_cells);
        a.cand_count = (int*)(S + o_cnt); a.cand_base = (int*)(S + o_base); a.cand_total = scal;
        a.cand_cap = cand_cap; a.cand_idx = (int*)(S + o_cidx); a.cand_dist = (unsigned short*)(S + o_cdist);
        a.topk = (unsigned*)(S + o_topk);
        J.n_ptr = b->d_n + p; J.nq_ptr = b->d_nq + p;
        J.taken = b->d_taken + on; J.moq = b->d_match_of_query + oq; J.tq = b->d_target_query + on;
        J.bin = (signed char*)(S + o_bin); J.asg = (int*)(S + o_asg); J.scal = scal;
    }
    rc = upload_jobs(W, jobs.data(), (size_t)npairs * sizeof(WindowJob), st);
    if (rc != ORB_OK) return rc;
    const WindowJob* d_jobs = (const WindowJob*)W.d;
    const size_t rsmem = std::max(sizeof(SrStage) + (size_t)((n + 3) & ~3) * 5, (size_t)9 * ((n + 3) & ~3) + 16);
    static thread_local int attr_dev = -1;
    if (attr_dev != device) {
        ORB_CUDA(cudaFuncSetAttribute(window_resolve_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)std::max(sizeof(SrStage) + GB_MAX_N * 5, (size_t)9 * GB_MAX_N + 16)));
        attr_dev = device;
    }
    grid_build_batch_kernel<<<npairs, 1024, npad * sizeof(unsigned), st>>>(d_jobs);
    window_candidates_batch_kernel<<<dim3((nq + 7) / 8, npairs), 256, 0, st>>>(d_jobs);
    window_resolve_batch_kernel<<<npairs, SR_THREADS, rsmem, st>>>(d_jobs, prm->mode, prm->th_dist, prm->nn_ratio, prm->check_orientation, (int)rsmem);
    ORB_CUDA(cudaGetLastError());
    // per-pair status: nmatches = scal[1]; a pair whose candidate arena overflowed reports -1 (re-run it through orb_search_by_projection)
    search_batch_finish_kernel<<<(npairs + 255) / 256, 256, 0, st>>>(d_jobs, npairs, b->d_nmatches);
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}
}  // extern "C"
