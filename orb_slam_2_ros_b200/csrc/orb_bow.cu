// orb_bow.cu — bag-of-words transform of ORB descriptors: ORBVocabulary::transform
// = DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>::transform (reference
// orb_slam2/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1140-1272, FORB.cpp:81-101, BowVector.cpp:33-87,
// FeatureVector.cpp:29-44; callers Frame::ComputeBoW Frame.cc:428-435, KeyFrame::ComputeBoW KeyFrame.cc:68-77).
//
// Vocabulary layout in HBM ("slots"): every non-root node sits in exactly one children list; the lists are laid out
// back to back in node-id order of the parents, children in the reference's push_back order (TemplatedVocabulary.h:1411).
//   desc[slot][32]  rec[slot] = {first child slot, #children, node id, word id}   weight[slot] (double)
// so one level of the descent is ONE round of independent loads (the k children's 32-byte descriptors + their 16-byte
// records, 352 B per child group) followed by a half-warp arg-min.  The ORBvoc-size tree (1.1 M nodes, 56 MB) stays
// resident in the 126 MB L2.
//
// bow_descend_kernel : 16 lanes per descriptor (2 per warp); lane c scores child c (c+16, ... for k > 16) with POPC over
//                      two uint4 loads; key = dist << 8 | child index, min over the half warp = the reference's
//                      "first child with the smallest distance" (strict <, :1251-1262).  Per feature: word id, weight,
//                      node id at level L - levelsup.
// bow_assemble_kernel: one CTA per frame.  BowVector = sort (word id, feature index) keys in shared memory (bitonic),
//                      one thread per word adds the weights in feature order (the order of BowVector::addWeight's
//                      double additions), one thread accumulates the norm in word order (BowVector::normalize), all
//                      threads divide.  FeatureVector = sort (node id, feature index) keys, segment heads.
//                      Double arithmetic uses explicit __dadd_rn / __dmul_rn / IEEE division: bit-exact vs the host.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "orb_internal.cuh"

struct orb_voc {
    int device = 0;
    int k = 0, L = 0, scoring = 0, weighting = 0, n_nodes = 0, n_words = 0, n_slots = 0;
    int4 root = {0, 0, 0, 0};
    uint8_t* d_desc = nullptr;
    int4* d_rec = nullptr;
    double* d_weight = nullptr;
};

namespace {

#define BOW_LANES 16
#define BOW_ASM_THREADS 1024
#define BOW_MAX_FRAME 8192   // features per frame of the vector assembly (shared-memory sort)

__global__ void __launch_bounds__(256)
bow_descend_kernel(const uint8_t* __restrict__ desc, int n, const uint8_t* __restrict__ vdesc, const int4* __restrict__ vrec,
                   const double* __restrict__ vweight, const int4 root, const int nid_level, int32_t* __restrict__ word_id,
                   double* __restrict__ weight, int32_t* __restrict__ node_id) {
    const int lane = threadIdx.x & (BOW_LANES - 1);
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) / BOW_LANES;
    if (i >= n) return;   // whole 16-lane groups leave together (n is a count of groups)
    const unsigned gmask = 0xFFFFu << (threadIdx.x & 16);
    const uint4 f0 = __ldg(reinterpret_cast<const uint4*>(desc + (size_t)i * 32));
    const uint4 f1 = __ldg(reinterpret_cast<const uint4*>(desc + (size_t)i * 32) + 1);
    int4 cur = root;
    int cur_slot = -1, level = 0, nid = 0;
    bool nid_set = nid_level <= 0;                       // TemplatedVocabulary.h:1240: root
    do {
        ++level;
        unsigned best = 0xFFFFFFFFu;
        int4 brec = make_int4(0, 0, 0, 0);
        for (int c = lane; c < cur.y; c += BOW_LANES) {
            const int slot = cur.x + c;
            const uint4 d0 = __ldg(reinterpret_cast<const uint4*>(vdesc + (size_t)slot * 32));
            const uint4 d1 = __ldg(reinterpret_cast<const uint4*>(vdesc + (size_t)slot * 32) + 1);
            const int4 r = __ldg(vrec + slot);
            const int dist = __popc(f0.x ^ d0.x) + __popc(f0.y ^ d0.y) + __popc(f0.z ^ d0.z) + __popc(f0.w ^ d0.w) +
                             __popc(f1.x ^ d1.x) + __popc(f1.y ^ d1.y) + __popc(f1.z ^ d1.z) + __popc(f1.w ^ d1.w);
            const unsigned key = ((unsigned)dist << 8) | (unsigned)c;
            if (key < best) { best = key; brec = r; }
        }
        unsigned m = best;
#pragma unroll
        for (int o = BOW_LANES / 2; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(gmask, m, o, BOW_LANES));
        const int c = (int)(m & 0xFFu), src = c & (BOW_LANES - 1);   // the lane that scored child c
        const bool mine = (best == m);
        // the winner's record: only the owning lane holds it for key m
        int4 w;
        w.x = __shfl_sync(gmask, mine ? brec.x : 0, src, BOW_LANES);
        w.y = __shfl_sync(gmask, mine ? brec.y : 0, src, BOW_LANES);
        w.z = __shfl_sync(gmask, mine ? brec.z : 0, src, BOW_LANES);
        w.w = __shfl_sync(gmask, mine ? brec.w : 0, src, BOW_LANES);
        cur_slot = cur.x + c;
        cur = w;
        if (level == nid_level) { nid = cur.z; nid_set = true; }   // :1264-1265
    } while (cur.y > 0);                                           // !isLeaf()
    if (!nid_set) nid = cur.z;                                     // pin (iv): leaf above level L - levelsup
    if (lane == 0) {
        word_id[i] = cur.w;
        weight[i] = __ldg(vweight + cur_slot);
        node_id[i] = nid;
    }
}

__device__ __forceinline__ void bitonic_sort_u64(unsigned long long* a, int N) {
    for (int k = 2; k <= N; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (N >> 1); t += blockDim.x) {
                const int i = 2 * t - (t & (j - 1));
                const unsigned long long x = a[i], y = a[i + j];
                const bool up = (i & k) == 0;
                if ((x > y) == up) { a[i] = y; a[i + j] = x; }
            }
            __syncthreads();
        }
    }
}

// exclusive block scan of one int per thread (blockDim = 1024)
__device__ __forceinline__ int block_excl_scan(int v, int* warp_tot, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = warp_tot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += y;
        }
        warp_tot[lane] = w;   // inclusive
    }
    __syncthreads();
    total = warp_tot[31];
    const int base = warp ? warp_tot[warp - 1] : 0;
    __syncthreads();
    return base + incl - v;
}

// norm_kind: 0 = L1, 1 = L2, -1 = none; tf: 1 = TF / TF_IDF (addWeight), 0 = IDF / BINARY (addIfNotExist)
__global__ void __launch_bounds__(BOW_ASM_THREADS)
bow_assemble_kernel(const int32_t* __restrict__ desc_off, const int32_t* __restrict__ word_id, const double* __restrict__ weight,
                    const int32_t* __restrict__ node_id, const int N, const int tf, const int norm_kind,
                    int32_t* __restrict__ bow_n, int32_t* __restrict__ bow_word, double* __restrict__ bow_value,
                    int32_t* __restrict__ fv_n, int32_t* __restrict__ fv_node, int32_t* __restrict__ fv_start,
                    int32_t* __restrict__ fv_feat) {
    extern __shared__ __align__(16) unsigned char bow_smem[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(bow_smem);       // [N]
    double* vals = reinterpret_cast<double*>(bow_smem + (size_t)N * 8);                // [N]
    __shared__ int s_wtot[32];
    __shared__ double s_norm;
    const int f = blockIdx.x;
    const int o0 = desc_off[f], n = desc_off[f + 1] - o0;
    const int per = N / BOW_ASM_THREADS > 0 ? N / BOW_ASM_THREADS : 1;                 // consecutive entries per thread
    // ---------------- BowVector ----------------
    for (int p = threadIdx.x; p < N; p += blockDim.x) {
        unsigned long long key = ~0ull;
        if (p < n && weight[o0 + p] > 0.0) key = ((unsigned long long)(unsigned)word_id[o0 + p] << 32) | (unsigned)p;   // w > 0: not stopped (:1170)
        keys[p] = key;
    }
    __syncthreads();
    bitonic_sort_u64(keys, N);
    // heads: one thread per word adds the weights of its features in feature order (BowVector::addWeight, BowVector.cpp:33-45)
    int cnt = 0;
    const int p0 = threadIdx.x * per;
    for (int p = p0; p < p0 + per && p < N; ++p) {
        const unsigned long long key = keys[p];
        if (key == ~0ull) break;
        const unsigned w = (unsigned)(key >> 32);
        if (p > 0 && (unsigned)(keys[p - 1] >> 32) == w) continue;
        double v = weight[o0 + (int)(unsigned)key];
        if (tf)
            for (int q = p + 1; q < N && (unsigned)(keys[q] >> 32) == w; ++q) v = __dadd_rn(v, weight[o0 + (int)(unsigned)keys[q]]);
        vals[p] = v;
        ++cnt;
    }
    int nb;
    int out = block_excl_scan(cnt, s_wtot, nb);
    // compact (word, value) to the front; values are staged in registers first (vals is both source and destination)
    {
        double vloc[BOW_MAX_FRAME / BOW_ASM_THREADS];
        unsigned wloc[BOW_MAX_FRAME / BOW_ASM_THREADS];
        int c = 0;
        for (int p = p0; p < p0 + per && p < N; ++p) {
            const unsigned long long key = keys[p];
            if (key == ~0ull) break;
            const unsigned w = (unsigned)(key >> 32);
            if (p > 0 && (unsigned)(keys[p - 1] >> 32) == w) continue;
            vloc[c] = vals[p]; wloc[c] = w; ++c;
        }
        __syncthreads();
        for (int j = 0; j < c; ++j) { vals[out + j] = vloc[j]; bow_word[o0 + out + j] = (int32_t)wloc[j]; }
    }
    __syncthreads();
    if (tf && norm_kind < 0 && nb > 0) {          // TemplatedVocabulary.h:1177-1183: unnecessary when normalising
        const double nd = (double)nb;
        for (int j = threadIdx.x; j < nb; j += blockDim.x) vals[j] = vals[j] / nd;
        __syncthreads();
    }
    if (norm_kind >= 0) {                          // BowVector::normalize, BowVector.cpp:63-87: sequential, word order
        if (threadIdx.x == 0) {
            double nrm = 0.0;
            if (norm_kind == 0) { for (int j = 0; j < nb; ++j) nrm = __dadd_rn(nrm, fabs(vals[j])); }
            else { for (int j = 0; j < nb; ++j) nrm = __dadd_rn(nrm, __dmul_rn(vals[j], vals[j])); nrm = sqrt(nrm); }
            s_norm = nrm;
        }
        __syncthreads();
        const double nrm = s_norm;
        if (nrm > 0.0) for (int j = threadIdx.x; j < nb; j += blockDim.x) vals[j] = vals[j] / nrm;
        __syncthreads();
    }
    for (int j = threadIdx.x; j < nb; j += blockDim.x) bow_value[o0 + j] = vals[j];
    if (threadIdx.x == 0) bow_n[f] = nb;
    __syncthreads();
    // ---------------- FeatureVector (FeatureVector::addFeature, FeatureVector.cpp:29-44) ----------------
    for (int p = threadIdx.x; p < N; p += blockDim.x) {
        unsigned long long key = ~0ull;
        if (p < n && weight[o0 + p] > 0.0) key = ((unsigned long long)(unsigned)node_id[o0 + p] << 32) | (unsigned)p;
        keys[p] = key;
    }
    __syncthreads();
    bitonic_sort_u64(keys, N);
    cnt = 0;
    int nvalid_local = 0;
    for (int p = p0; p < p0 + per && p < N; ++p) {
        const unsigned long long key = keys[p];
        if (key == ~0ull) break;
        ++nvalid_local;
        fv_feat[o0 + p] = (int32_t)(unsigned)key;
        if (p == 0 || (unsigned)(keys[p - 1] >> 32) != (unsigned)(key >> 32)) ++cnt;
    }
    int nf;
    out = block_excl_scan(cnt, s_wtot, nf);
    for (int p = p0; p < p0 + per && p < N; ++p) {
        const unsigned long long key = keys[p];
        if (key == ~0ull) break;
        if (p == 0 || (unsigned)(keys[p - 1] >> 32) != (unsigned)(key >> 32)) {
            fv_node[o0 + out] = (int32_t)(unsigned)(key >> 32);
            fv_start[o0 + f + out] = p;
            ++out;
        }
    }
    int nvalid;
    block_excl_scan(nvalid_local, s_wtot, nvalid);
    if (threadIdx.x == 0) { fv_n[f] = nf; fv_start[o0 + f + nf] = nvalid; }
}

// per-thread workspace of the host-pointer entry points (same scheme as orb_search.cu): grow-only slabs + a stream
struct Workspace {
    int device = -1;
    cudaStream_t st = nullptr;
    uint8_t *d = nullptr, *h = nullptr;
    size_t d_cap = 0, h_cap = 0;
    ~Workspace() {
        if (device >= 0 && cudaSetDevice(device) == cudaSuccess) { cudaFree(d); cudaFreeHost(h); if (st) cudaStreamDestroy(st); }
    }
    int prepare(int dev, size_t dbytes, size_t hbytes) {
        ORB_CUDA(cudaSetDevice(dev));
        if (device != dev) {
            if (device >= 0) { cudaSetDevice(device); cudaFree(d); cudaFreeHost(h); if (st) cudaStreamDestroy(st); cudaSetDevice(dev); }
            d = h = nullptr; d_cap = h_cap = 0; st = nullptr; device = dev;
            ORB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        }
        if (d_cap < dbytes) {
            ORB_CUDA(cudaStreamSynchronize(st));
            cudaFree(d); d = nullptr; d_cap = 0;
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
thread_local Workspace g_bow_ws;

struct Carver {
    size_t off = 0;
    size_t take(size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; }
};

bool must_normalize(int scoring, int& norm) {   // ScoringObject.h:73-90
    norm = scoring == 1 ? 1 : 0;
    return scoring != 5;
}

int launch_descend(const orb_voc* v, const uint8_t* d_desc, int n, int levelsup, int32_t* d_word, double* d_weight, int32_t* d_node,
                   cudaStream_t st) {
    if (n <= 0) return ORB_OK;
    const long long threads = (long long)n * BOW_LANES;
    bow_descend_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(d_desc, n, v->d_desc, v->d_rec, v->d_weight, v->root,
                                                                          v->L - levelsup, d_word, d_weight, d_node);
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

int launch_assemble(const orb_voc* v, const int32_t* d_desc_off, int nframes, int max_frame, const int32_t* d_word, const double* d_weight,
                    const int32_t* d_node, int32_t* d_bow_n, int32_t* d_bow_word, double* d_bow_value, int32_t* d_fv_n, int32_t* d_fv_node,
                    int32_t* d_fv_start, int32_t* d_fv_feat, cudaStream_t st) {
    if (nframes <= 0) return ORB_OK;
    int N = BOW_ASM_THREADS;
    while (N < max_frame) N <<= 1;
    if (N > BOW_MAX_FRAME) { orb_set_error("orb_bow_transform: more than %d descriptors in one frame", BOW_MAX_FRAME); return ORB_ERR_CAPACITY; }
    const size_t smem = (size_t)N * 16;
    static thread_local int attr_dev = -1;
    if (attr_dev != v->device) {
        ORB_CUDA(cudaFuncSetAttribute(bow_assemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BOW_MAX_FRAME * 16));
        attr_dev = v->device;
    }
    int norm;
    const bool must = must_normalize(v->scoring, norm);
    const int tf = (v->weighting == 0 || v->weighting == 1) ? 1 : 0;
    bow_assemble_kernel<<<nframes, BOW_ASM_THREADS, smem, st>>>(d_desc_off, d_word, d_weight, d_node, N, tf, must ? norm : -1, d_bow_n,
                                                                 d_bow_word, d_bow_value, d_fv_n, d_fv_node, d_fv_start, d_fv_feat);
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

}  // namespace

extern "C" {

int orb_voc_create(orb_voc** voc, int device, int k, int L, int scoring, int weighting, int n_nodes, const int32_t* parent,
                   const uint8_t* is_leaf, const uint8_t* desc32, const double* weight) {
    if (!voc) return ORB_ERR_INVALID;
    *voc = nullptr;
    if (n_nodes < 1 || (n_nodes > 1 && (!parent || !is_leaf || !desc32 || !weight))) return ORB_ERR_INVALID;
    if (L < 0 || scoring < 0 || scoring > 5 || weighting < 0 || weighting > 3) { orb_set_error("orb_voc_create: bad L / scoring / weighting"); return ORB_ERR_INVALID; }
    for (int i = 1; i < n_nodes; ++i)
        if (parent[i] < 0 || parent[i] >= i) { orb_set_error("orb_voc_create: node %d has parent %d (parents must precede their children)", i, parent[i]); return ORB_ERR_INVALID; }
    if (orb_device_count() <= 0) { orb_set_error("no CUDA device visible: liborb_b200 has no CPU fallback"); return ORB_ERR_NO_DEVICE; }
    // children lists in push_back order (TemplatedVocabulary.h:1411) -> slots
    std::vector<int> nchild(n_nodes, 0), first(n_nodes + 1, 0), fill(n_nodes, 0), slot_of(n_nodes, -1), word_of(n_nodes, 0);
    for (int i = 1; i < n_nodes; ++i) nchild[parent[i]]++;
    for (int i = 0; i < n_nodes; ++i) first[i + 1] = first[i] + nchild[i];
    int n_words = 0;
    for (int i = 1; i < n_nodes; ++i) {
        slot_of[i] = first[parent[i]] + fill[parent[i]]++;
        if (is_leaf[i]) word_of[i] = n_words++;     // word ids in node order (:1427-1433); Node() default word_id = 0
    }
    for (int i = 0; i < n_nodes; ++i)
        if (nchild[i] > 255) { orb_set_error("orb_voc_create: node %d has %d children (max 255)", i, nchild[i]); return ORB_ERR_INVALID; }
    const int n_slots = n_nodes - 1;
    std::vector<uint8_t> sdesc((size_t)std::max(n_slots, 1) * 32);
    std::vector<int4> srec(std::max(n_slots, 1));
    std::vector<double> sw(std::max(n_slots, 1));
    for (int i = 1; i < n_nodes; ++i) {
        const int s = slot_of[i];
        memcpy(&sdesc[(size_t)s * 32], desc32 + (size_t)i * 32, 32);
        srec[s] = make_int4(first[i], nchild[i], i, word_of[i]);
        sw[s] = weight[i];
    }
    orb_voc* v = new orb_voc;
    v->device = device; v->k = k; v->L = L; v->scoring = scoring; v->weighting = weighting;
    v->n_nodes = n_nodes; v->n_words = n_words; v->n_slots = n_slots;
    v->root = make_int4(first[0], nchild[0], 0, 0);
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaMalloc(&v->d_desc, sdesc.size());
    if (e == cudaSuccess) e = cudaMalloc(&v->d_rec, srec.size() * sizeof(int4));
    if (e == cudaSuccess) e = cudaMalloc(&v->d_weight, sw.size() * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpy(v->d_desc, sdesc.data(), sdesc.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(v->d_rec, srec.data(), srec.size() * sizeof(int4), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(v->d_weight, sw.data(), sw.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        orb_set_error("orb_voc_create: %s", cudaGetErrorString(e));
        cudaFree(v->d_desc); cudaFree(v->d_rec); cudaFree(v->d_weight);
        delete v;
        return ORB_ERR_CUDA;
    }
    *voc = v;
    return ORB_OK;
}

// ORBVocabulary::loadFromTextFile (TemplatedVocabulary.h:1351-1441): `k L scoring weighting`, then one line per node
// `parent isLeaf d0 .. d31 weight`; node ids in file order.  Blank lines are skipped (DESIGN.md pin (v)).
int orb_voc_load_text(orb_voc** voc, int device, const char* path) {
    if (!voc || !path) return ORB_ERR_INVALID;
    *voc = nullptr;
    FILE* f = fopen(path, "r");
    if (!f) { orb_set_error("orb_voc_load_text: cannot open %s", path); return ORB_ERR_INVALID; }
    std::vector<char> line(1 << 16);
    auto next_long = [](char*& p, long& out) -> bool {
        char* e;
        out = strtol(p, &e, 10);
        if (e == p) return false;
        p = e;
        return true;
    };
    long k = 0, L = 0, n1 = 0, n2 = 0;
    bool ok = fgets(line.data(), (int)line.size(), f) != nullptr;
    if (ok) {
        char* p = line.data();
        ok = next_long(p, k) && next_long(p, L) && next_long(p, n1) && next_long(p, n2);
    }
    if (!ok || k < 0 || k > 20 || L < 1 || L > 10 || n1 < 0 || n1 > 5 || n2 < 0 || n2 > 3) {   // :1379-1383
        fclose(f);
        orb_set_error("Vocabulary loading failure: This is not a correct text file!");
        return ORB_ERR_INVALID;
    }
    std::vector<int32_t> parent(1, 0);
    std::vector<uint8_t> leaf(1, 0), desc(32, 0);
    std::vector<double> weight(1, 0.0);
    while (fgets(line.data(), (int)line.size(), f)) {
        char* p = line.data();
        long pid, isleaf;
        if (!next_long(p, pid)) continue;
        if (pid < 0 || pid >= (long)parent.size()) { fclose(f); orb_set_error("orb_voc_load_text: node %zu has parent %ld", parent.size(), pid); return ORB_ERR_INVALID; }
        if (!next_long(p, isleaf)) isleaf = 0;
        uint8_t d[32] = {0};
        for (int i = 0; i < 32; ++i) {
            long b;
            if (next_long(p, b)) d[i] = (unsigned char)b;
        }
        char* e;
        double w = strtod(p, &e);
        if (e == p) w = 0.0;
        parent.push_back((int32_t)pid); leaf.push_back(isleaf > 0 ? 1 : 0); desc.insert(desc.end(), d, d + 32); weight.push_back(w);
    }
    fclose(f);
    return orb_voc_create(voc, device, (int)k, (int)L, (int)n1, (int)n2, (int)parent.size(), parent.data(), leaf.data(), desc.data(), weight.data());
}

void orb_voc_destroy(orb_voc* v) {
    if (!v) return;
    if (cudaSetDevice(v->device) == cudaSuccess) { cudaFree(v->d_desc); cudaFree(v->d_rec); cudaFree(v->d_weight); }
    delete v;
}

int orb_voc_info(orb_voc* v, int* k, int* L, int* n_nodes, int* n_words, int* scoring, int* weighting) {
    if (!v) return ORB_ERR_INVALID;
    if (k) *k = v->k;
    if (L) *L = v->L;
    if (n_nodes) *n_nodes = v->n_nodes;
    if (n_words) *n_words = v->n_words;
    if (scoring) *scoring = v->scoring;
    if (weighting) *weighting = v->weighting;
    return ORB_OK;
}

int orb_bow_transform_features_device(orb_voc* v, const uint8_t* d_desc32, int n, int levelsup, int32_t* d_word_id, double* d_weight,
                                      int32_t* d_node_id, void* cuda_stream) {
    if (!v || n < 0 || (n && (!d_desc32 || !d_word_id || !d_weight || !d_node_id))) return ORB_ERR_INVALID;
    if (v->root.y == 0) { orb_set_error("orb_bow_transform_features: empty vocabulary"); return ORB_ERR_INVALID; }
    ORB_CUDA(cudaSetDevice(v->device));
    return launch_descend(v, d_desc32, n, levelsup, d_word_id, d_weight, d_node_id, (cudaStream_t)cuda_stream);
}

int orb_bow_transform_features(orb_voc* v, const uint8_t* desc32, int n, int levelsup, int32_t* word_id, double* weight, int32_t* node_id) {
    if (!v || n < 0 || (n && (!desc32 || !word_id || !weight || !node_id))) return ORB_ERR_INVALID;
    if (n == 0) return ORB_OK;
    if (v->root.y == 0) { orb_set_error("orb_bow_transform_features: empty vocabulary"); return ORB_ERR_INVALID; }
    Carver c;
    const size_t o_desc = c.take((size_t)n * 32), in_bytes = c.off;
    const size_t o_word = c.take((size_t)n * 4), o_node = c.take((size_t)n * 4), o_w = c.take((size_t)n * 8);
    Workspace& W = g_bow_ws;
    int rc = W.prepare(v->device, c.off, c.off);
    if (rc != ORB_OK) return rc;
    memcpy(W.h + o_desc, desc32, (size_t)n * 32);
    ORB_CUDA(cudaMemcpyAsync(W.d, W.h, in_bytes, cudaMemcpyHostToDevice, W.st));
    rc = launch_descend(v, W.d + o_desc, n, levelsup, (int32_t*)(W.d + o_word), (double*)(W.d + o_w), (int32_t*)(W.d + o_node), W.st);
    if (rc != ORB_OK) return rc;
    ORB_CUDA(cudaMemcpyAsync(W.h + in_bytes, W.d + in_bytes, c.off - in_bytes, cudaMemcpyDeviceToHost, W.st));
    ORB_CUDA(cudaStreamSynchronize(W.st));
    memcpy(word_id, W.h + o_word, (size_t)n * 4);
    memcpy(node_id, W.h + o_node, (size_t)n * 4);
    memcpy(weight, W.h + o_w, (size_t)n * 8);
    return ORB_OK;
}

// all pointers on the device; scratch = 16 bytes per descriptor (word id, node id, weight); asynchronous on cuda_stream
int orb_bow_transform_device(orb_voc* v, const uint8_t* d_desc32, const int32_t* d_desc_off, int nframes, int n_total, int max_frame,
                             int levelsup, void* d_scratch, int32_t* d_bow_n, int32_t* d_bow_word, double* d_bow_value, int32_t* d_fv_n,
                             int32_t* d_fv_node, int32_t* d_fv_start, int32_t* d_fv_feat, void* cuda_stream) {
    if (!v || nframes < 0 || n_total < 0 || !d_desc_off || !d_bow_n || !d_fv_n || !d_fv_start) return ORB_ERR_INVALID;
    if (n_total && (!d_desc32 || !d_scratch || !d_bow_word || !d_bow_value || !d_fv_node || !d_fv_feat)) return ORB_ERR_INVALID;
    ORB_CUDA(cudaSetDevice(v->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    double* d_w = (double*)d_scratch;
    int32_t* d_word = (int32_t*)((uint8_t*)d_scratch + (size_t)n_total * 8);
    int32_t* d_node = d_word + n_total;
    if (v->root.y == 0) {   // empty(): both vectors stay empty (TemplatedVocabulary.h:1147-1150)
        ORB_CUDA(cudaMemsetAsync(d_bow_n, 0, (size_t)nframes * 4, st));
        ORB_CUDA(cudaMemsetAsync(d_fv_n, 0, (size_t)nframes * 4, st));
        ORB_CUDA(cudaMemsetAsync(d_fv_start, 0, ((size_t)n_total + nframes) * 4, st));
        return ORB_OK;
    }
    int rc = launch_descend(v, d_desc32, n_total, levelsup, d_word, d_w, d_node, st);
    if (rc != ORB_OK) return rc;
    return launch_assemble(v, d_desc_off, nframes, max_frame, d_word, d_w, d_node, d_bow_n, d_bow_word, d_bow_value, d_fv_n, d_fv_node,
                           d_fv_start, d_fv_feat, st);
}

int orb_bow_transform(orb_voc* v, const uint8_t* desc32, const int32_t* desc_off, int nframes, int levelsup, int32_t* bow_n,
                      int32_t* bow_word, double* bow_value, int32_t* fv_n, int32_t* fv_node, int32_t* fv_start, int32_t* fv_feat) {
    if (!v || nframes < 0 || !desc_off || (nframes && (!bow_n || !fv_n || !fv_start))) return ORB_ERR_INVALID;
    if (nframes == 0) return ORB_OK;
    int max_frame = 0;
    if (desc_off[0] != 0) { orb_set_error("orb_bow_transform: desc_off[0] must be 0"); return ORB_ERR_INVALID; }
    for (int f = 0; f < nframes; ++f) {
        if (desc_off[f + 1] < desc_off[f]) { orb_set_error("orb_bow_transform: desc_off must be non-decreasing"); return ORB_ERR_INVALID; }
        max_frame = std::max(max_frame, desc_off[f + 1] - desc_off[f]);
    }
    const int n = desc_off[nframes];
    if (n && (!desc32 || !bow_word || !bow_value || !fv_node || !fv_feat)) return ORB_ERR_INVALID;
    if (max_frame > BOW_MAX_FRAME) { orb_set_error("orb_bow_transform: more than %d descriptors in one frame", BOW_MAX_FRAME); return ORB_ERR_CAPACITY; }
    if (orb_device_count() <= 0) { orb_set_error("no CUDA device visible: liborb_b200 has no CPU fallback"); return ORB_ERR_NO_DEVICE; }
    Carver c;
    const size_t o_desc = c.take((size_t)n * 32), o_off = c.take(((size_t)nframes + 1) * 4), in_bytes = c.off;
    const size_t o_bn = c.take((size_t)nframes * 4), o_fn = c.take((size_t)nframes * 4);
    const size_t o_bw = c.take((size_t)n * 4), o_bv = c.take((size_t)n * 8), o_fnode = c.take((size_t)n * 4);
    const size_t o_fs = c.take(((size_t)n + nframes) * 4), o_ff = c.take((size_t)n * 4), io_bytes = c.off;
    const size_t o_scr = c.take((size_t)n * 16 + 16);
    Workspace& W = g_bow_ws;
    int rc = W.prepare(v->device, c.off, io_bytes);
    if (rc != ORB_OK) return rc;
    if (n) memcpy(W.h + o_desc, desc32, (size_t)n * 32);
    memcpy(W.h + o_off, desc_off, ((size_t)nframes + 1) * 4);
    ORB_CUDA(cudaMemcpyAsync(W.d, W.h, in_bytes, cudaMemcpyHostToDevice, W.st));
    rc = orb_bow_transform_device(v, W.d + o_desc, (const int32_t*)(W.d + o_off), nframes, n, max_frame, levelsup, W.d + o_scr,
                                  (int32_t*)(W.d + o_bn), (int32_t*)(W.d + o_bw), (double*)(W.d + o_bv), (int32_t*)(W.d + o_fn),
                                  (int32_t*)(W.d + o_fnode), (int32_t*)(W.d + o_fs), (int32_t*)(W.d + o_ff), W.st);
    if (rc != ORB_OK) return rc;
    ORB_CUDA(cudaMemcpyAsync(W.h + in_bytes, W.d + in_bytes, io_bytes - in_bytes, cudaMemcpyDeviceToHost, W.st));
    ORB_CUDA(cudaStreamSynchronize(W.st));
    memcpy(bow_n, W.h + o_bn, (size_t)nframes * 4);
    memcpy(fv_n, W.h + o_fn, (size_t)nframes * 4);
    memcpy(fv_start, W.h + o_fs, ((size_t)n + nframes) * 4);
    if (n) {
        memcpy(bow_word, W.h + o_bw, (size_t)n * 4); memcpy(bow_value, W.h + o_bv, (size_t)n * 8);
        memcpy(fv_node, W.h + o_fnode, (size_t)n * 4); memcpy(fv_feat, W.h + o_ff, (size_t)n * 4);
    }
    return ORB_OK;
}

}  // extern "C"
