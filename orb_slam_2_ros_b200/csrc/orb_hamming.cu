// orb_hamming.cu — 256-bit Hamming best / second-best search (sm_100a).
//
// Replaces the inner loops built on ORBmatcher::DescriptorDistance
// (reference orb_slam2/src/ORBmatcher.cc:1649-1665; best/second update rule :217-226).
//
// Kernel shape: every thread OWNS HT_QPT query descriptors in registers; the database streams through
// shared memory in cp.async double-buffered tiles and every row is read with two broadcast LDS.128, so a
// compare costs 8 LOP3 + 8 POPC + a 3-input add tree, and the per-thread running (best, second) needs no
// cross-lane traffic at all.  best/second are kept as packed keys  key = dist << 23 | local_row  so that
//   k2 = min(k2, max(key, k1));  k1 = min(k1, key)
// implements the reference's strict '<' / 'else if <' update including "lowest index wins ties" and
// "second may equal best".  The database is split over blockIdx.x; a merge kernel combines the splits.
// This path is INT/popc-pipe bound: no tensor cores, HBM traffic is negligible.
#include <algorithm>
#include <vector>

#include "orb_internal.cuh"

namespace {

#define HT_THREADS 256
#ifndef HT_TILE
#define HT_TILE 256          // database rows per shared-memory stage (8 KB)
#endif
#ifndef HAMMING_CTAS_PER_SM
#define HAMMING_CTAS_PER_SM 4
#endif
#define HT_IDX_BITS 23       // local row index bits in a packed key (split length <= 8M rows)

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ unsigned lop3_xor3(unsigned a, unsigned b, unsigned c) {
    unsigned d; asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}
#ifndef HAMMING_VARIANT
#define HAMMING_VARIANT 0
#endif
#ifndef HAMMING_GROUP
#define HAMMING_GROUP 4      // rows per top-2 update group (1 = update after every row)
#endif
#ifndef HAMMING_QPT
#define HAMMING_QPT 2        // queries per thread when there are enough queries
#endif
__device__ __forceinline__ unsigned lop3_maj(unsigned a, unsigned b, unsigned c) {
    unsigned d; asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}

// popcount of the 256-bit XOR.  POPC issues at a quarter of the LOP3 rate on sm_100 (measured: orb_bench_issue_rate),
// so the 8 words first go through a carry-save adder tree (Harley-Seal: 4 CSAs = 8 LOP3) that leaves 4 words of
// weight 1, 1, 2, 4: 4 POPC instead of 8, the rest on the ALU / FMA pipes.  Exact, like the reference's bit-hack
// (ORBmatcher.cc:1649-1665).
__device__ __forceinline__ int hamming256(const uint4& a, const uint4& b, const unsigned (&q)[8]) {
    const unsigned x0 = a.x ^ q[0], x1 = a.y ^ q[1], x2 = a.z ^ q[2], x3 = a.w ^ q[3];
    const unsigned x4 = b.x ^ q[4], x5 = b.y ^ q[5], x6 = b.z ^ q[6], x7 = b.w ^ q[7];
    const unsigned s1 = lop3_xor3(x0, x1, x2), c1 = lop3_maj(x0, x1, x2);
    const unsigned s2 = lop3_xor3(x3, x4, x5), c2 = lop3_maj(x3, x4, x5);
    const unsigned s3 = lop3_xor3(s1, s2, x6), c3 = lop3_maj(s1, s2, x6);
#if HAMMING_VARIANT == 1
    // 3 CSAs + 5 POPC: two LOP3 fewer on the ALU pipe, one POPC more on the XU pipe
    return __popc(s3) + __popc(x7) + 2 * (__popc(c1) + __popc(c2) + __popc(c3));
#else
    const unsigned t1 = lop3_xor3(c1, c2, c3), f1 = lop3_maj(c1, c2, c3);
    return __popc(s3) + __popc(x7) + 2 * __popc(t1) + 4 * __popc(f1);
#endif
}

// the packed key  dist << 23 | row  straight from the four partial popcounts, as a chain of four multiply-adds with multipliers the
// compiler cannot turn into shifts (kernel argument m23 = 1 << 23): IMAD runs on the FMA pipe, which this kernel leaves idle,
// while shift / LEA / OR forms would take ALU-pipe slots, the pipe that binds it
__device__ __forceinline__ unsigned hamming256_key(const uint4& a, const uint4& b, const unsigned (&q)[8], unsigned row, unsigned m23,
                                                   unsigned m24, unsigned m25) {
    const unsigned x0 = a.x ^ q[0], x1 = a.y ^ q[1], x2 = a.z ^ q[2], x3 = a.w ^ q[3];
    const unsigned x4 = b.x ^ q[4], x5 = b.y ^ q[5], x6 = b.z ^ q[6], x7 = b.w ^ q[7];
    const unsigned s1 = lop3_xor3(x0, x1, x2), c1 = lop3_maj(x0, x1, x2);
    const unsigned s2 = lop3_xor3(x3, x4, x5), c2 = lop3_maj(x3, x4, x5);
    const unsigned s3 = lop3_xor3(s1, s2, x6), c3 = lop3_maj(s1, s2, x6);
    unsigned k;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(k) : "r"(__popc(s3)), "r"(m23), "r"(row));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(k) : "r"(__popc(x7)), "r"(m23));
#if HAMMING_VARIANT == 1
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(k) : "r"(__popc(c1)), "r"(m24));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(k) : "r"(__popc(c2)), "r"(m24));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(k) : "r"(__popc(c3)), "r"(m24));
#else
    const unsigned t1 = lop3_xor3(c1, c2, c3), f1 = lop3_maj(c1, c2, c3);
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(k) : "r"(__popc(t1)), "r"(m24));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(k) : "r"(__popc(f1)), "r"(m25));
#endif
    return k;
}

// the plain form (8 POPC), kept for the issue-rate microbenchmark that defines the popc roofline
__device__ __forceinline__ int hamming256_popc8(const uint4& a, const uint4& b, const unsigned (&q)[8]) {
    const int s0 = __popc(a.x ^ q[0]) + __popc(a.y ^ q[1]) + __popc(a.z ^ q[2]);
    const int s1 = __popc(a.w ^ q[3]) + __popc(b.x ^ q[4]) + __popc(b.y ^ q[5]);
    return s0 + s1 + __popc(b.z ^ q[6]) + __popc(b.w ^ q[7]);
}

#ifdef HAMMING_MINB
#define HT_BOUNDS __launch_bounds__(HT_THREADS, HAMMING_MINB)
#else
#define HT_BOUNDS __launch_bounds__(HT_THREADS)
#endif
template <int QPT>
__global__ void HT_BOUNDS
hamming_top2_kernel(const uint4* __restrict__ q, int nq, const uint4* __restrict__ db, long long ndb,
                    long long rows_per_split, uint2* __restrict__ partial, int nq_pad, unsigned m23, unsigned m24, unsigned m25) {
    __shared__ uint4 sdb[2][HT_TILE * 2];
    const int tid = threadIdx.x;
    const long long r0 = (long long)blockIdx.x * rows_per_split;
    const long long r1 = min(ndb, r0 + rows_per_split);
    const int qbase = blockIdx.y * (HT_THREADS * QPT);

    unsigned qw[QPT][8];
    unsigned k1[QPT], k2[QPT];
#pragma unroll
    for (int j = 0; j < QPT; ++j) {
        const int qi = qbase + j * HT_THREADS + tid;
        uint4 a = make_uint4(0, 0, 0, 0), b = a;
        if (qi < nq) { a = __ldg(q + 2 * (long long)qi); b = __ldg(q + 2 * (long long)qi + 1); }
        qw[j][0] = a.x; qw[j][1] = a.y; qw[j][2] = a.z; qw[j][3] = a.w;
        qw[j][4] = b.x; qw[j][5] = b.y; qw[j][6] = b.z; qw[j][7] = b.w;
        k1[j] = 0xFFFFFFFFu; k2[j] = 0xFFFFFFFFu;
    }
    const long long nrows = r1 - r0;
    const int ntiles = (int)((nrows + HT_TILE - 1) / HT_TILE);
    auto issue = [&](int t, int buf) {
        // tile t: rows [r0 + t*HT_TILE, ...) -> 2 x 16-byte chunks per thread
        const long long row0 = r0 + (long long)t * HT_TILE;
        const long long chunks = min((long long)HT_TILE, r1 - row0) * 2;
#pragma unroll
        for (int k = 0; k < 2 * HT_TILE / HT_THREADS; ++k) {
            const int cidx = tid + k * HT_THREADS;
            if (cidx < chunks) cp_async16(&sdb[buf][cidx], db + row0 * 2 + cidx);
        }
        cp_async_commit();
    };
    if (ntiles > 0) issue(0, 0);
    for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        if (t + 1 < ntiles) { issue(t + 1, buf ^ 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        const int rows = (int)min((long long)HT_TILE, nrows - (long long)t * HT_TILE);
        const unsigned idx0 = (unsigned)(t * HT_TILE);
#if HAMMING_GROUP > 1
        // rows in groups of HAMMING_GROUP: the 3-op top-2 update (ALU pipe, the binding one) runs only when the smallest key
        // of the group beats the thread's current second best — after the first few thousand rows that is rare, and the
        // common case costs 2 min3/min + 1 compare per group instead of 3 min/max per row.  Same result: keys that do not beat
        // k2 change neither k1 nor k2, and the update order inside a group is the row order.
        int r = 0;
        for (; r + HAMMING_GROUP <= rows; r += HAMMING_GROUP) {
            unsigned key[QPT][HAMMING_GROUP];
#pragma unroll
            for (int u = 0; u < HAMMING_GROUP; ++u) {
                const uint4 a = sdb[buf][2 * (r + u)], b = sdb[buf][2 * (r + u) + 1];
#pragma unroll
                for (int j = 0; j < QPT; ++j)
                    key[j][u] = hamming256_key(a, b, qw[j], idx0 + r + u, m23, m24, m25);
            }
#pragma unroll
            for (int j = 0; j < QPT; ++j) {
                unsigned m = key[j][0];
#pragma unroll
                for (int u = 1; u + 1 < HAMMING_GROUP; u += 2) m = __vimin3_u32(m, key[j][u], key[j][u + 1]);
                if ((HAMMING_GROUP & 1) == 0) m = min(m, key[j][HAMMING_GROUP - 1]);
                if (m < k2[j]) {
#pragma unroll
                    for (int u = 0; u < HAMMING_GROUP; ++u) {
                        k2[j] = min(k2[j], max(key[j][u], k1[j]));
                        k1[j] = min(k1[j], key[j][u]);
                    }
                }
            }
        }
        for (; r < rows; ++r) {
#else
#pragma unroll 4
        for (int r = 0; r < rows; ++r) {
#endif
            const uint4 a = sdb[buf][2 * r], b = sdb[buf][2 * r + 1];
#pragma unroll
            for (int j = 0; j < QPT; ++j) {
                const unsigned key = ((unsigned)hamming256(a, b, qw[j]) << HT_IDX_BITS) | (idx0 + r);
                k2[j] = min(k2[j], max(key, k1[j]));
                k1[j] = min(k1[j], key);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < QPT; ++j) {
        const int qi = qbase + j * HT_THREADS + tid;
        if (qi < nq) partial[(long long)blockIdx.x * nq_pad + qi] = make_uint2(k1[j], k2[j]);
    }
}

// lexicographic (dist, global row) top-2 over the splits of one query
__global__ void hamming_merge_kernel(const uint2* __restrict__ partial, int nsplit, int nq, int nq_pad,
                                     long long rows_per_split, long long index_base, orb_top2* __restrict__ out) {
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    int d1 = 256, d2 = 256;
    long long i1 = -1, i2 = -1;
    for (int s = 0; s < nsplit; ++s) {
        const uint2 p = partial[(long long)s * nq_pad + qi];
        const unsigned ks[2] = {p.x, p.y};
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (ks[k] == 0xFFFFFFFFu) continue;
            const int d = (int)(ks[k] >> HT_IDX_BITS);
            const long long g = (long long)s * rows_per_split + (ks[k] & ((1u << HT_IDX_BITS) - 1));
            // splits are visited in ascending row order, so '<' keeps the lowest row on ties
            if (d < d1) { d2 = d1; i2 = i1; d1 = d; i1 = g; }
            else if (d < d2) { d2 = d; i2 = g; }
        }
    }
    orb_top2 o;
    o.best_dist = d1; o.second_dist = d2;
    o.best_idx = i1 < 0 ? -1 : index_base + i1;
    o.second_idx = i2 < 0 ? -1 : index_base + i2;
    out[qi] = o;
}

// exact merge of per-shard results on the device (after the all-gather): parts[nparts][nq] hold global indices;
// lexicographic (dist, index) top-2 of the union, shards in any order
__global__ void top2_merge_kernel(const orb_top2* __restrict__ parts, int nparts, int nq, orb_top2* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    int d1 = 256, d2 = 256;
    long long i1 = 0x7FFFFFFFFFFFFFFFll, i2 = 0x7FFFFFFFFFFFFFFFll;   // max = absent
    for (int s = 0; s < nparts; ++s) {
        const orb_top2 p = parts[(size_t)s * nq + i];
        const int ds[2] = {p.best_dist, p.second_dist};
        const long long is[2] = {p.best_idx, p.second_idx};
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (is[k] < 0) continue;
            if (ds[k] < d1 || (ds[k] == d1 && is[k] < i1)) { d2 = d1; i2 = i1; d1 = ds[k]; i1 = is[k]; }
            else if (ds[k] < d2 || (ds[k] == d2 && is[k] < i2)) { d2 = ds[k]; i2 = is[k]; }
        }
    }
    orb_top2 o;
    o.best_dist = d1; o.second_dist = d2;
    o.best_idx = i1 == 0x7FFFFFFFFFFFFFFFll ? -1 : i1;
    o.second_idx = i2 == 0x7FFFFFFFFFFFFFFFll ? -1 : i2;
    out[i] = o;
}

// ---- register-only issue-rate microbenchmarks: the denominators of the popc roofline -------------------
__global__ void popc_peak_kernel(unsigned* out, int iters, unsigned seed) {
    unsigned a0 = seed + threadIdx.x, a1 = a0 * 3u, a2 = a0 * 5u, a3 = a0 * 7u;
    unsigned a4 = a0 * 11u, a5 = a0 * 13u, a6 = a0 * 17u, a7 = a0 * 19u;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {  // 64 dependent-free-ish POPCs per iteration (8 chains)
            a0 = __popc(a0) + 0x55u; a1 = __popc(a1) + 0x33u; a2 = __popc(a2) + 0x0Fu; a3 = __popc(a3) + 0x71u;
            a4 = __popc(a4) + 0x5Au; a5 = __popc(a5) + 0x3Cu; a6 = __popc(a6) + 0x69u; a7 = __popc(a7) + 0x17u;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}
// the compare itself with all operands in registers: CSA = false: 8 xor + 8 popc + adds + the 3-op packed top-2 update per row
// (SURVEY's popc8 form); CSA = true: the search kernel's own mix — carry-save tree + 4 popc, key by multiply-adds, top-2 update
// per group of HAMMING_GROUP rows
template <bool CSA>
__global__ void compare_peak_kernel(unsigned* out, int iters, unsigned seed, unsigned m23, unsigned m24, unsigned m25) {
    unsigned q[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) q[k] = seed * (k + 1) + threadIdx.x * 2654435761u;
    uint4 a = make_uint4(seed, seed ^ 0x1234567u, seed * 3u, seed * 7u), b = make_uint4(~seed, seed * 5u, seed * 9u, seed + 77u);
    unsigned k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu;
    constexpr int G = (CSA && HAMMING_GROUP > 1) ? HAMMING_GROUP : 1;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u0 = 0; u0 < 8; u0 += G) {
            unsigned key[G];
#pragma unroll
            for (int u = 0; u < G; ++u) {
                // every word of the row changes every iteration (8 extra IADD on the ALU pipe), so no popc is hoisted
                a.x += 0x9E3779B9u; a.y += a.x; a.z += a.y; a.w += a.z; b.x += a.w; b.y += b.x; b.z += b.y; b.w += b.z;
                key[u] = CSA ? hamming256_key(a, b, q, (unsigned)(i * 8 + u0 + u), m23, m24, m25)
                             : (((unsigned)hamming256_popc8(a, b, q) << HT_IDX_BITS) | (unsigned)(i * 8 + u0 + u));
            }
            unsigned m = key[0];
#pragma unroll
            for (int u = 1; u < G; ++u) m = min(m, key[u]);
            if (G == 1 || m < k2) {
#pragma unroll
                for (int u = 0; u < G; ++u) {
                    k2 = min(k2, max(key[u], k1));
                    k1 = min(k1, key[u]);
                }
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = k1 ^ k2;
}

struct HammingPlan { int qpt, grid_y, nsplit, nq_pad; long long rows_per_split; };

HammingPlan make_plan(int nq, long long ndb, int sms) {
    HammingPlan p;
    p.qpt = (nq > 148 * HT_THREADS / 2) ? 2 : 1;   // enough queries to fill the chip with 2 per thread?
    if (nq >= 1024) p.qpt = 2;
    if (nq >= 1024 && HAMMING_QPT == 4) p.qpt = 4;
    const int per_cta = HT_THREADS * p.qpt;
    p.grid_y = (nq + per_cta - 1) / per_cta;
    const int target_ctas = sms * HAMMING_CTAS_PER_SM;
    long long nsplit = std::max<long long>(1, target_ctas / p.grid_y);
    nsplit = std::min<long long>(nsplit, std::max<long long>(1, (ndb + 63) / 64));   // >= 64 rows per split
    long long rps = (ndb + nsplit - 1) / nsplit;
    rps = std::max<long long>(rps, 1);
    const long long max_rps = (1ll << HT_IDX_BITS);
    if (rps > max_rps) rps = max_rps;
    rps = (rps + 1) / 2 * 2;   // keep tiles 32-byte aligned pairs
    p.rows_per_split = rps;
    p.nsplit = (int)std::max<long long>(1, (ndb + rps - 1) / rps);
    p.nq_pad = (nq + 31) / 32 * 32;
    return p;
}

}  // namespace

struct orb_db {
    int device = 0, sms = 148;
    long long cap = 0, n = 0, index_base = 0;
    uint8_t* d_rows = nullptr;
    cudaStream_t stream = nullptr; bool own_stream = false;
    uint2* d_partial = nullptr; size_t partial_elems = 0;
    uint8_t* d_q = nullptr; orb_top2* d_out = nullptr; int q_cap = 0;
    long long launches = 0;
    // sharded form (orb_db_create_sharded): one shard per rank, per-shard results all-gathered over NCCL and merged on the device
    void* comm = nullptr; int rank = 0, world = 1;
    orb_top2* d_part = nullptr; orb_top2* d_all = nullptr; int shard_q_cap = 0;
    // CUDA-event timers of {search kernel, merge kernel}: ring of event triples, harvested lazily
    bool profile = false;
    cudaEvent_t prof_ev[ORB_PROF_RING][3] = {};
    bool prof_pending[ORB_PROF_RING] = {};
    int prof_head = 0;
    double prof_ms[2] = {0, 0};
    long long prof_calls = 0;
};

static int db_harvest(orb_db* db, int slot) {
    cudaEvent_t* ev = db->prof_ev[slot];
    ORB_CUDA(cudaEventSynchronize(ev[2]));
    for (int s = 0; s < 2; ++s) {
        float ms = 0.f;
        ORB_CUDA(cudaEventElapsedTime(&ms, ev[s], ev[s + 1]));
        db->prof_ms[s] += ms;
    }
    db->prof_calls++;
    db->prof_pending[slot] = false;
    return ORB_OK;
}

static int db_launch(orb_db* db, const uint8_t* d_q, int nq, orb_top2* d_out) {
    if (nq == 0) return ORB_OK;
    const HammingPlan p = make_plan(nq, db->n, db->sms);
    const size_t need = (size_t)p.nsplit * p.nq_pad;
    if (db->partial_elems < need) {
        ORB_CUDA(cudaStreamSynchronize(db->stream));
        cudaFree(db->d_partial); db->d_partial = nullptr;
        ORB_CUDA(cudaMalloc(&db->d_partial, need * sizeof(uint2)));
        db->partial_elems = need;
    }
    cudaEvent_t* ev = nullptr;
    if (db->profile) {
        const int slot = db->prof_head;
        if (db->prof_pending[slot]) { int rc = db_harvest(db, slot); if (rc != ORB_OK) return rc; }
        ev = db->prof_ev[slot];
        db->prof_pending[slot] = true;
        db->prof_head = (slot + 1) % ORB_PROF_RING;
        ORB_CUDA(cudaEventRecord(ev[0], db->stream));
    }
    if (db->n > 0) {
        dim3 grid(p.nsplit, p.grid_y);
        if (p.qpt == 4)
            hamming_top2_kernel<4><<<grid, HT_THREADS, 0, db->stream>>>((const uint4*)d_q, nq, (const uint4*)db->d_rows, db->n,
                                                                       p.rows_per_split, db->d_partial, p.nq_pad, 1u << HT_IDX_BITS, 2u << HT_IDX_BITS, 4u << HT_IDX_BITS);
        else if (p.qpt == 2)
            hamming_top2_kernel<2><<<grid, HT_THREADS, 0, db->stream>>>((const uint4*)d_q, nq, (const uint4*)db->d_rows, db->n,
                                                                       p.rows_per_split, db->d_partial, p.nq_pad, 1u << HT_IDX_BITS, 2u << HT_IDX_BITS, 4u << HT_IDX_BITS);
        else
            hamming_top2_kernel<1><<<grid, HT_THREADS, 0, db->stream>>>((const uint4*)d_q, nq, (const uint4*)db->d_rows, db->n,
                                                                       p.rows_per_split, db->d_partial, p.nq_pad, 1u << HT_IDX_BITS, 2u << HT_IDX_BITS, 4u << HT_IDX_BITS);
        db->launches++;
    }
    if (ev) ORB_CUDA(cudaEventRecord(ev[1], db->stream));
    hamming_merge_kernel<<<(nq + 127) / 128, 128, 0, db->stream>>>(db->d_partial, db->n > 0 ? p.nsplit : 0, nq, p.nq_pad,
                                                                  p.rows_per_split, db->index_base, d_out);
    db->launches++;
    if (ev) ORB_CUDA(cudaEventRecord(ev[2], db->stream));
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

// ---- NCCL, bound at run time ------------------------------------------------------------------------------------------
// liborb_b200.so must load on hosts without NCCL (single-GPU use, the CPU-only build check), so the five entry points the
// sharded database needs are resolved with dlopen / dlsym on first use: "libnccl.so.2" — the copy already mapped into the
// process (e.g. the one PyTorch ships) if there is one, else the system's.  The ABI of these calls is stable across NCCL 2.x.
#include <dlfcn.h>
namespace {
struct NcclId { char internal[128]; };   // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
struct NcclApi {
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int /*ncclDataType_t*/, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
NcclApi* nccl_api() {
    static NcclApi api = [] {
        NcclApi a;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
        if (!h) return a;
        a.GetUniqueId = (int (*)(NcclId*))dlsym(h, "ncclGetUniqueId");
        a.CommInitRank = (int (*)(void**, int, NcclId, int))dlsym(h, "ncclCommInitRank");
        a.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
        a.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(h, "ncclAllGather");
        a.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.GetErrorString;
        return a;
    }();
    return &api;
}
#define ORB_NCCL(call)                                                                                        \
    do {                                                                                                      \
        const int _r = (call);                                                                                \
        if (_r != 0) { orb_set_error("NCCL: %s (%s)", nccl_api()->GetErrorString(_r), #call); return ORB_ERR_CUDA; } \
    } while (0)
}  // namespace

extern "C" {

int orb_db_create(orb_db** out, int device, int64_t capacity_rows, int64_t index_base) {
    if (!out || capacity_rows < 0) return ORB_ERR_INVALID;
    *out = nullptr;
    if (orb_device_count() <= 0) { orb_set_error("no CUDA device visible: liborb_b200 has no CPU fallback"); return ORB_ERR_NO_DEVICE; }
    ORB_CUDA(cudaSetDevice(device));
    orb_db* db = new orb_db;
    db->device = device; db->cap = capacity_rows; db->index_base = index_base;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) db->sms = prop.multiProcessorCount;
    if (cudaMalloc(&db->d_rows, (size_t)std::max<int64_t>(capacity_rows, 1) * 32) != cudaSuccess) {
        orb_set_error("orb_db_create: cudaMalloc of %lld rows failed", (long long)capacity_rows);
        cudaGetLastError(); delete db; return ORB_ERR_CUDA;
    }
    if (cudaStreamCreateWithFlags(&db->stream, cudaStreamNonBlocking) != cudaSuccess) { cudaFree(db->d_rows); delete db; return ORB_ERR_CUDA; }
    db->own_stream = true;
    *out = db;
    return ORB_OK;
}

void orb_db_destroy(orb_db* db) {
    if (!db) return;
    cudaSetDevice(db->device);
    if (db->stream) cudaStreamSynchronize(db->stream);
    cudaFree(db->d_rows); cudaFree(db->d_partial); cudaFree(db->d_q); cudaFree(db->d_out); cudaFree(db->d_part); cudaFree(db->d_all);
    if (db->comm && nccl_api()->ok) nccl_api()->CommDestroy(db->comm);
    if (db->prof_ev[0][0])
        for (int r = 0; r < ORB_PROF_RING; ++r)
            for (int s = 0; s < 3; ++s) cudaEventDestroy(db->prof_ev[r][s]);
    if (db->own_stream && db->stream) cudaStreamDestroy(db->stream);
    delete db;
}

int orb_db_add(orb_db* db, const uint8_t* desc, int64_t nrows) {
    if (!db || (!desc && nrows) || nrows < 0) return ORB_ERR_INVALID;
    if (db->n + nrows > db->cap) { orb_set_error("orb_db_add: capacity %lld exceeded", db->cap); return ORB_ERR_CAPACITY; }
    ORB_CUDA(cudaSetDevice(db->device));
    ORB_CUDA(cudaMemcpyAsync(db->d_rows + db->n * 32, desc, (size_t)nrows * 32, cudaMemcpyHostToDevice, db->stream));
    ORB_CUDA(cudaStreamSynchronize(db->stream));
    db->n += nrows;
    return ORB_OK;
}

int orb_db_add_device(orb_db* db, const uint8_t* d_desc, int64_t nrows) {
    if (!db || (!d_desc && nrows) || nrows < 0) return ORB_ERR_INVALID;
    if (db->n + nrows > db->cap) { orb_set_error("orb_db_add_device: capacity %lld exceeded", db->cap); return ORB_ERR_CAPACITY; }
    ORB_CUDA(cudaSetDevice(db->device));
    ORB_CUDA(cudaMemcpyAsync(db->d_rows + db->n * 32, d_desc, (size_t)nrows * 32, cudaMemcpyDeviceToDevice, db->stream));
    db->n += nrows;
    return ORB_OK;
}

int orb_db_profile_enable(orb_db* db, int enable) {
    if (!db) return ORB_ERR_INVALID;
    ORB_CUDA(cudaSetDevice(db->device));
    if (enable && !db->prof_ev[0][0])
        for (int r = 0; r < ORB_PROF_RING; ++r)
            for (int s = 0; s < 3; ++s) ORB_CUDA(cudaEventCreate(&db->prof_ev[r][s]));
    db->profile = enable != 0;
    return ORB_OK;
}

int orb_db_profile_read(orb_db* db, double* search_ms, double* merge_ms, int64_t* calls, int reset) {
    if (!db) return ORB_ERR_INVALID;
    if (db->prof_ev[0][0]) {
        ORB_CUDA(cudaSetDevice(db->device));
        for (int r = 0; r < ORB_PROF_RING; ++r)
            if (db->prof_pending[r]) { int rc = db_harvest(db, r); if (rc != ORB_OK) return rc; }
    }
    if (search_ms) *search_ms = db->prof_ms[0];
    if (merge_ms) *merge_ms = db->prof_ms[1];
    if (calls) *calls = db->prof_calls;
    if (reset) { db->prof_ms[0] = db->prof_ms[1] = 0; db->prof_calls = 0; }
    return ORB_OK;
}

int64_t orb_db_size(orb_db* db) { return db ? db->n : 0; }
int64_t orb_db_launch_count(orb_db* db) { return db ? db->launches : 0; }

int orb_db_set_stream(orb_db* db, void* s) {
    if (!db) return ORB_ERR_INVALID;
    ORB_CUDA(cudaSetDevice(db->device));
    ORB_CUDA(cudaStreamSynchronize(db->stream));
    if (db->own_stream) { cudaStreamDestroy(db->stream); db->own_stream = false; }
    db->stream = (cudaStream_t)s;
    if (!db->stream) { ORB_CUDA(cudaStreamCreateWithFlags(&db->stream, cudaStreamNonBlocking)); db->own_stream = true; }
    return ORB_OK;
}

int orb_db_query_top2_device(orb_db* db, const uint8_t* d_q, int nq, orb_top2* d_out) {
    if (!db || nq < 0 || (nq && (!d_q || !d_out))) return ORB_ERR_INVALID;
    ORB_CUDA(cudaSetDevice(db->device));
    return db_launch(db, d_q, nq, d_out);
}

int orb_db_query_top2(orb_db* db, const uint8_t* q, int nq, orb_top2* out) {
    if (!db || nq < 0 || (nq && (!q || !out))) return ORB_ERR_INVALID;
    if (nq == 0) return ORB_OK;
    ORB_CUDA(cudaSetDevice(db->device));
    if (db->q_cap < nq) {
        ORB_CUDA(cudaStreamSynchronize(db->stream));
        cudaFree(db->d_q); cudaFree(db->d_out); db->d_q = nullptr; db->d_out = nullptr;
        ORB_CUDA(cudaMalloc(&db->d_q, (size_t)nq * 32));
        ORB_CUDA(cudaMalloc(&db->d_out, (size_t)nq * sizeof(orb_top2)));
        db->q_cap = nq;
    }
    ORB_CUDA(cudaMemcpyAsync(db->d_q, q, (size_t)nq * 32, cudaMemcpyHostToDevice, db->stream));
    int rc = db_launch(db, db->d_q, nq, db->d_out);
    if (rc != ORB_OK) return rc;
    ORB_CUDA(cudaMemcpyAsync(out, db->d_out, (size_t)nq * sizeof(orb_top2), cudaMemcpyDeviceToHost, db->stream));
    ORB_CUDA(cudaStreamSynchronize(db->stream));
    return ORB_OK;
}

/* ---- sharded database: one shard per rank / GPU, exact global top-2 (BASELINE config 5) ---------- */
int orb_shard_unique_id(uint8_t* id128) {
    if (!id128) return ORB_ERR_INVALID;
    if (!nccl_api()->ok) { orb_set_error("NCCL (libnccl.so.2) not found: the sharded database needs it"); return ORB_ERR_INVALID; }
    NcclId id;
    ORB_NCCL(nccl_api()->GetUniqueId(&id));
    memcpy(id128, id.internal, 128);
    return ORB_OK;
}

int orb_db_create_sharded(orb_db** out, int device, int64_t capacity_rows, int64_t index_base, int rank, int world, const uint8_t* id128) {
    if (!out || world < 1 || rank < 0 || rank >= world || (world > 1 && !id128)) return ORB_ERR_INVALID;
    int rc = orb_db_create(out, device, capacity_rows, index_base);
    if (rc != ORB_OK) return rc;
    orb_db* db = *out;
    db->rank = rank; db->world = world;
    if (world > 1) {
        if (!nccl_api()->ok) { orb_db_destroy(db); *out = nullptr; orb_set_error("NCCL (libnccl.so.2) not found: the sharded database needs it"); return ORB_ERR_INVALID; }
        NcclId id;
        memcpy(id.internal, id128, 128);
        const int r = nccl_api()->CommInitRank(&db->comm, world, id, rank);   // collective: every rank calls it
        if (r != 0) { orb_set_error("NCCL: %s (ncclCommInitRank)", nccl_api()->GetErrorString(r)); orb_db_destroy(db); *out = nullptr; return ORB_ERR_CUDA; }
    }
    return ORB_OK;
}

/* every rank passes the same nq queries; d_out on every rank receives the exact global result.  Asynchronous on the shard's stream:
 * per-shard search -> ncclAllGather of the nq x 24-byte results -> device merge. */
int orb_db_query_top2_sharded(orb_db* db, const uint8_t* d_q, int nq, orb_top2* d_out) {
    if (!db || nq < 0 || (nq && (!d_q || !d_out))) return ORB_ERR_INVALID;
    if (nq == 0) return ORB_OK;
    ORB_CUDA(cudaSetDevice(db->device));
    if (db->world == 1) return db_launch(db, d_q, nq, d_out);
    if (db->shard_q_cap < nq) {
        ORB_CUDA(cudaStreamSynchronize(db->stream));
        cudaFree(db->d_part); cudaFree(db->d_all); db->d_part = nullptr; db->d_all = nullptr; db->shard_q_cap = 0;
        ORB_CUDA(cudaMalloc(&db->d_part, (size_t)nq * sizeof(orb_top2)));
        ORB_CUDA(cudaMalloc(&db->d_all, (size_t)nq * sizeof(orb_top2) * db->world));
        db->shard_q_cap = nq;
    }
    int rc = db_launch(db, d_q, nq, db->d_part);
    if (rc != ORB_OK) return rc;
    ORB_NCCL(nccl_api()->AllGather(db->d_part, db->d_all, (size_t)nq * sizeof(orb_top2), /*ncclChar*/ 0, db->comm, db->stream));
    top2_merge_kernel<<<(nq + 127) / 128, 128, 0, db->stream>>>(db->d_all, db->world, nq, d_out);
    db->launches++;
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

/* host-pointer convenience form of the same (blocking) */
int orb_db_query_top2_sharded_host(orb_db* db, const uint8_t* q, int nq, orb_top2* out) {
    if (!db || nq < 0 || (nq && (!q || !out))) return ORB_ERR_INVALID;
    if (nq == 0) return ORB_OK;
    ORB_CUDA(cudaSetDevice(db->device));
    if (db->q_cap < nq) {
        ORB_CUDA(cudaStreamSynchronize(db->stream));
        cudaFree(db->d_q); cudaFree(db->d_out); db->d_q = nullptr; db->d_out = nullptr;
        ORB_CUDA(cudaMalloc(&db->d_q, (size_t)nq * 32));
        ORB_CUDA(cudaMalloc(&db->d_out, (size_t)nq * sizeof(orb_top2)));
        db->q_cap = nq;
    }
    ORB_CUDA(cudaMemcpyAsync(db->d_q, q, (size_t)nq * 32, cudaMemcpyHostToDevice, db->stream));
    int rc = orb_db_query_top2_sharded(db, db->d_q, nq, db->d_out);
    if (rc != ORB_OK) return rc;
    ORB_CUDA(cudaMemcpyAsync(out, db->d_out, (size_t)nq * sizeof(orb_top2), cudaMemcpyDeviceToHost, db->stream));
    ORB_CUDA(cudaStreamSynchronize(db->stream));
    return ORB_OK;
}

int orb_hamming_top2(int device, const uint8_t* q, int nq, const uint8_t* dbrows, int64_t ndb, orb_top2* out) {
    if (nq < 0 || ndb < 0 || (nq && (!q || !out)) || (ndb && !dbrows)) return ORB_ERR_INVALID;
    // one grow-only database per calling thread and device, reused across calls (no cudaMalloc / cudaFree on the call path)
    struct Cache { orb_db* db = nullptr; int device = -1; ~Cache() { /* the CUDA context may be gone at thread exit: leak, do not touch it */ } };
    static thread_local Cache cache;
    if (cache.db && (cache.device != device || cache.db->cap < ndb)) { orb_db_destroy(cache.db); cache.db = nullptr; }
    if (!cache.db) {
        const int rc = orb_db_create(&cache.db, device, std::max<int64_t>(ndb + ndb / 2, 1024), 0);
        if (rc != ORB_OK) { cache.db = nullptr; return rc; }
        cache.device = device;
    }
    cache.db->n = 0;
    int rc = orb_db_add(cache.db, dbrows, ndb);
    if (rc == ORB_OK) rc = orb_db_query_top2(cache.db, q, nq, out);
    return rc;
}

int orb_top2_merge_device(int device, const orb_top2* d_parts, int nparts, int nq, orb_top2* d_out, void* cuda_stream) {
    if (!d_parts || !d_out || nparts <= 0 || nq < 0) return ORB_ERR_INVALID;
    if (nq == 0) return ORB_OK;
    if (orb_device_count() <= 0) { orb_set_error("no CUDA device visible"); return ORB_ERR_NO_DEVICE; }
    ORB_CUDA(cudaSetDevice(device));
    top2_merge_kernel<<<(nq + 127) / 128, 128, 0, (cudaStream_t)cuda_stream>>>(d_parts, nparts, nq, d_out);
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

int orb_top2_merge(const orb_top2* parts, int nparts, int nq, orb_top2* out) {
    if (!parts || !out || nparts <= 0 || nq < 0) return ORB_ERR_INVALID;
    for (int i = 0; i < nq; ++i) {
        // lexicographic (dist, global index) top-2 of the union; shards may arrive in any order
        int d1 = 256, d2 = 256;
        int64_t i1 = INT64_MAX, i2 = INT64_MAX;   // INT64_MAX = absent
        auto less = [](int da, int64_t ia, int db_, int64_t ib) { return da < db_ || (da == db_ && ia < ib); };
        for (int s = 0; s < nparts; ++s) {
            const orb_top2& p = parts[(size_t)s * nq + i];
            const int ds[2] = {p.best_dist, p.second_dist};
            const int64_t is[2] = {p.best_idx, p.second_idx};
            for (int k = 0; k < 2; ++k) {
                if (is[k] < 0) continue;
                if (less(ds[k], is[k], d1, i1)) { d2 = d1; i2 = i1; d1 = ds[k]; i1 = is[k]; }
                else if (less(ds[k], is[k], d2, i2)) { d2 = ds[k]; i2 = is[k]; }
            }
        }
        if (i1 == INT64_MAX) i1 = -1;
        if (i2 == INT64_MAX) i2 = -1;
        out[i].best_dist = d1; out[i].second_dist = d2; out[i].best_idx = i1; out[i].second_idx = i2;
    }
    return ORB_OK;
}

/* issue-rate microbenchmarks (not in the public header; bench.py binds them through ctypes):
 * kind 0: POPC only -> *gops = 1e9 popc/s ; kind 1: register-only 256-bit compare with 8 POPC + top-2 update ->
 * 1e9 compares/s ; kind 2: the same with the carry-save (4 POPC) popcount the search kernel uses */
int orb_bench_issue_rate(int device, int kind, int iters, double* gops) {
    if (!gops || iters <= 0) return ORB_ERR_INVALID;
    if (orb_device_count() <= 0) { orb_set_error("no CUDA device visible"); return ORB_ERR_NO_DEVICE; }
    ORB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    ORB_CUDA(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    unsigned* d_out = nullptr;
    ORB_CUDA(cudaMalloc(&d_out, sizeof(unsigned) * blocks * threads));
    cudaEvent_t e0, e1;
    ORB_CUDA(cudaEventCreate(&e0)); ORB_CUDA(cudaEventCreate(&e1));
    float best_ms = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        ORB_CUDA(cudaEventRecord(e0));
        if (kind == 0) popc_peak_kernel<<<blocks, threads>>>(d_out, iters, 12345u + rep);
        else if (kind == 1) compare_peak_kernel<false><<<blocks, threads>>>(d_out, iters, 12345u + rep, 1u << HT_IDX_BITS, 2u << HT_IDX_BITS, 4u << HT_IDX_BITS);
        else compare_peak_kernel<true><<<blocks, threads>>>(d_out, iters, 12345u + rep, 1u << HT_IDX_BITS, 2u << HT_IDX_BITS, 4u << HT_IDX_BITS);
        ORB_CUDA(cudaEventRecord(e1));
        ORB_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        ORB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0) best_ms = std::min(best_ms, ms);
    }
    const double ops = (double)blocks * threads * iters * (kind == 0 ? 64.0 : 8.0);
    *gops = ops / (best_ms * 1e-3) / 1e9;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_out);
    return ORB_OK;
}

}  // extern "C"
