// orb_oracle_bow.cpp — CPU ORACLE (test infrastructure only, see orb_oracle.h) for the bag-of-words transform of
// ORB descriptors: ORBVocabulary = DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>
// (reference orb_slam2/include/ORBVocabulary.h:31; orb_slam2/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h, FORB.cpp,
// BowVector.cpp, FeatureVector.cpp, ScoringObject.cpp).  Callers in the reference: Frame::ComputeBoW (Frame.cc:428-435),
// KeyFrame::ComputeBoW (KeyFrame.cc:68-77), both with levelsup = 4.
//
// The restatement keeps the reference's containers (std::vector children lists, std::map BowVector / FeatureVector) and
// its loop order, so tie-breaking (first child with the minimum distance, TemplatedVocabulary.h:1251-1262) and the
// order of the double-precision additions (BowVector::addWeight in feature order, BowVector::normalize in word order)
// are the reference's.  Pinned to the reference's own DBoW2 sources compiled in oracle/_ref (tests/test_ref_pin.py::test_reference_dbow2_equals_oracle); the vocabulary file ORBvoc.txt is absent
// from the checkout, so tests use synthetic trees written in the reference's text format.
//
// Stated pin (iv): when the descent reaches a leaf ABOVE level L - levelsup, the reference leaves *nid unwritten
// (TemplatedVocabulary.h:1264-1265 never fires; the caller's `NodeId nid` is uninitialised, :1163-1168).  Oracle and
// product return the leaf's own node id in that case.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "orb_oracle.h"

namespace {

struct Node {  // TemplatedVocabulary.h:96-131
    int id = 0;
    double weight = 0;
    std::vector<int> children;
    int parent = 0;
    uint8_t desc[32] = {0};
    int word_id = 0;
    bool isLeaf() const { return children.empty(); }
};

struct Voc {
    int k = 0, L = 0, scoring = 0, weighting = 0;
    std::vector<Node> nodes;
    std::vector<int> words;  // word id -> node id
};

int forb_distance(const uint8_t* a, const uint8_t* b) {  // FORB.cpp:81-101
    const int32_t* pa = reinterpret_cast<const int32_t*>(a);
    const int32_t* pb = reinterpret_cast<const int32_t*>(b);
    int dist = 0;
    for (int i = 0; i < 8; i++, pa++, pb++) {
        unsigned int v = *pa ^ *pb;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

// TemplatedVocabulary.h:1231-1272
void transform_one(const Voc& v, const uint8_t* feature, int& word_id, double& weight, int* nid, int levelsup) {
    const int nid_level = v.L - levelsup;
    bool nid_set = false;
    if (nid_level <= 0 && nid) { *nid = 0; nid_set = true; }
    int final_id = 0, current_level = 0;
    do {
        ++current_level;
        const std::vector<int>& nodes = v.nodes[final_id].children;
        final_id = nodes[0];
        double best_d = forb_distance(feature, v.nodes[final_id].desc);
        for (size_t c = 1; c < nodes.size(); ++c) {
            const int id = nodes[c];
            const double d = forb_distance(feature, v.nodes[id].desc);
            if (d < best_d) { best_d = d; final_id = id; }
        }
        if (nid && current_level == nid_level) { *nid = final_id; nid_set = true; }
    } while (!v.nodes[final_id].isLeaf());
    if (nid && !nid_set) *nid = final_id;  // pin (iv)
    word_id = v.nodes[final_id].word_id;
    weight = v.nodes[final_id].weight;
}

bool must_normalize(int scoring, int& norm) {  // ScoringObject.h:73-90
    norm = (scoring == 1) ? 1 : 0;             // L2_NORM -> L2, everything else L1
    return scoring != 5;                       // DOT_PRODUCT does not normalise
}

}  // namespace

extern "C" {

void* orc_voc_create(int k, int L, int scoring, int weighting, int n_nodes, const int32_t* parent, const uint8_t* is_leaf,
                     const uint8_t* desc32, const double* weight) {
    // the in-memory form of loadFromTextFile (TemplatedVocabulary.h:1389-1436): node ids in file order, node 0 = root
    Voc* v = new Voc;
    v->k = k; v->L = L; v->scoring = scoring; v->weighting = weighting;
    v->nodes.resize(n_nodes);
    for (int nid = 1; nid < n_nodes; ++nid) {
        Node& n = v->nodes[nid];
        n.id = nid;
        n.parent = parent[nid];
        v->nodes[parent[nid]].children.push_back(nid);
        memcpy(n.desc, desc32 + (size_t)nid * 32, 32);
        n.weight = weight[nid];
        if (is_leaf[nid]) { n.word_id = (int)v->words.size(); v->words.push_back(nid); }
    }
    return v;
}

// TemplatedVocabulary.h:1351-1441.  The reference parses with std::stringstream operator>>; this restatement reads the
// same whitespace-separated decimal tokens with strtol / strtod (C stdio: the oracle library carries a statically linked
// libstdc++ whose iostreams must not be mixed with the host process's).  Stated pin (v): the reference's
// `while(!f.eof())` loop turns the empty string after the last newline into a phantom extra child of the root with an
// uninitialised descriptor (:1396-1420 with every extraction failing); blank lines are skipped here.
void* orc_voc_load_text(const char* path) {
    FILE* f = fopen(path, "r");
    if (!f) return nullptr;
    Voc* v = new Voc;
    std::vector<char> line(1 << 16);
    auto next_long = [](char*& p, long& out) -> bool {
        char* e;
        out = strtol(p, &e, 10);
        if (e == p) return false;
        p = e;
        return true;
    };
    bool ok = fgets(line.data(), (int)line.size(), f) != nullptr;
    long k = 0, L = 0, n1 = 0, n2 = 0;
    if (ok) {
        char* p = line.data();
        ok = next_long(p, k) && next_long(p, L) && next_long(p, n1) && next_long(p, n2);
    }
    if (!ok || k < 0 || k > 20 || L < 1 || L > 10 || n1 < 0 || n1 > 5 || n2 < 0 || n2 > 3) { delete v; fclose(f); return nullptr; }
    v->k = (int)k; v->L = (int)L; v->scoring = (int)n1; v->weighting = (int)n2;
    v->nodes.resize(1);
    while (fgets(line.data(), (int)line.size(), f)) {
        char* p = line.data();
        long pid, leaf;
        if (!next_long(p, pid)) continue;  // blank line: pin (v)
        const int nid = (int)v->nodes.size();
        if (pid < 0 || pid >= nid) { delete v; fclose(f); return nullptr; }  // the reference would index out of range
        v->nodes.resize(v->nodes.size() + 1);
        v->nodes[nid].id = nid;
        v->nodes[nid].parent = (int)pid;
        v->nodes[pid].children.push_back(nid);
        if (!next_long(p, leaf)) leaf = 0;
        for (int iD = 0; iD < 32; iD++) {  // FORB::fromString, FORB.cpp:120-136
            long n;
            if (next_long(p, n)) v->nodes[nid].desc[iD] = (unsigned char)n;
        }
        char* e;
        const double w = strtod(p, &e);
        if (e != p) v->nodes[nid].weight = w;
        if (leaf > 0) { v->nodes[nid].word_id = (int)v->words.size(); v->words.push_back(nid); }
    }
    fclose(f);
    return v;
}

void orc_voc_destroy(void* voc) { delete static_cast<Voc*>(voc); }

void orc_voc_info(void* voc, int32_t* k, int32_t* L, int32_t* n_nodes, int32_t* n_words, int32_t* scoring, int32_t* weighting) {
    Voc* v = static_cast<Voc*>(voc);
    *k = v->k; *L = v->L; *n_nodes = (int)v->nodes.size(); *n_words = (int)v->words.size(); *scoring = v->scoring; *weighting = v->weighting;
}

void orc_voc_export(void* voc, int32_t* parent, uint8_t* is_leaf, uint8_t* desc32, double* weight) {
    Voc* v = static_cast<Voc*>(voc);
    std::vector<char> leaf(v->nodes.size(), 0);
    for (int nid : v->words) leaf[nid] = 1;
    for (size_t i = 0; i < v->nodes.size(); ++i) {
        parent[i] = v->nodes[i].parent; is_leaf[i] = (uint8_t)leaf[i];
        memcpy(desc32 + i * 32, v->nodes[i].desc, 32); weight[i] = v->nodes[i].weight;
    }
}

// per feature: transform(feature, id, weight, &nid, levelsup)
void orc_bow_transform_features(void* voc, const uint8_t* desc32, int n, int levelsup, int32_t* word_id, double* weight, int32_t* node_id) {
    Voc* v = static_cast<Voc*>(voc);
    for (int i = 0; i < n; ++i) {
        int w, nid;
        double wt;
        transform_one(*v, desc32 + (size_t)i * 32, w, wt, &nid, levelsup);
        word_id[i] = w; weight[i] = wt; node_id[i] = nid;
    }
}

// transform(features, BowVector, FeatureVector, levelsup): TemplatedVocabulary.h:1140-1218.  Flat outputs:
// BowVector = (bow_word[j], bow_value[j]) j < *n_bow ascending word id; FeatureVector = fv_node[j] ascending, features of
// node j = fv_feat[fv_start[j] .. fv_start[j+1]).
void orc_bow_transform(void* voc, const uint8_t* desc32, int n, int levelsup, int32_t* n_bow, int32_t* bow_word, double* bow_value,
                       int32_t* n_fv, int32_t* fv_node, int32_t* fv_start, int32_t* fv_feat) {
    Voc* vv = static_cast<Voc*>(voc);
    std::map<unsigned int, double> v;
    std::map<unsigned int, std::vector<unsigned int>> fv;
    *n_bow = 0; *n_fv = 0;
    if (vv->nodes.size() <= 1) { fv_start[0] = 0; return; }
    int norm;
    const bool must = must_normalize(vv->scoring, norm);
    const bool tf = (vv->weighting == 1 || vv->weighting == 0);  // TF || TF_IDF
    for (int i = 0; i < n; ++i) {
        int id, nid;
        double w;
        transform_one(*vv, desc32 + (size_t)i * 32, id, w, &nid, levelsup);
        if (w > 0) {
            auto vit = v.lower_bound((unsigned)id);
            if (vit != v.end() && !(v.key_comp()((unsigned)id, vit->first))) {
                if (tf) vit->second += w;              // BowVector::addWeight; addIfNotExist leaves it
            } else {
                v.insert(vit, std::make_pair((unsigned)id, w));
            }
            fv[(unsigned)nid].push_back((unsigned)i);  // FeatureVector::addFeature
        }
    }
    if (tf && !v.empty() && !must) {
        const double nd = (double)v.size();
        for (auto& e : v) e.second /= nd;
    }
    if (must) {  // BowVector::normalize, BowVector.cpp:63-87
        double nrm = 0.0;
        if (norm == 0) { for (auto& e : v) nrm += fabs(e.second); }
        else { for (auto& e : v) nrm += e.second * e.second; nrm = sqrt(nrm); }
        if (nrm > 0.0) for (auto& e : v) e.second /= nrm;
    }
    int j = 0;
    for (auto& e : v) { bow_word[j] = (int32_t)e.first; bow_value[j] = e.second; ++j; }
    *n_bow = j;
    j = 0;
    int pos = 0;
    for (auto& e : fv) {
        fv_node[j] = (int32_t)e.first; fv_start[j] = pos;
        for (unsigned int fi : e.second) fv_feat[pos++] = (int32_t)fi;
        ++j;
    }
    fv_start[j] = pos;
    *n_fv = j;
}

// L1Scoring::score (ScoringObject.cpp:23-66) on flat sorted vectors
double orc_bow_score_l1(const int32_t* w1, const double* v1, int n1, const int32_t* w2, const double* v2, int n2) {
    int i = 0, j = 0;
    double score = 0;
    while (i < n1 && j < n2) {
        if (w1[i] == w2[j]) { score += fabs(v1[i] - v2[j]) - fabs(v1[i]) - fabs(v2[j]); ++i; ++j; }
        else if (w1[i] < w2[j]) { while (i < n1 && w1[i] < w2[j]) ++i; }
        else { while (j < n2 && w2[j] < w1[i]) ++j; }
    }
    return -score / 2.0;
}

}  // extern "C"
