// ORBVocabulary.cc — see ORBVocabulary.h.  The tree descent and the vector assembly run on the GPU (csrc/orb_bow.cu);
// this file only converts between the reference's std::map containers and the flat arrays of the C ABI.
#include "ORBVocabulary.h"

#include <cmath>
#include <cstdio>
#include <cstring>

#include "../../include/orb_b200.h"

namespace ORB_SLAM2 {

ORBVocabulary::~ORBVocabulary() { orb_voc_destroy(voc_); }

bool ORBVocabulary::loadFromTextFile(const std::string& filename) {
    orb_voc_destroy(voc_);
    voc_ = nullptr;
    if (orb_voc_load_text(&voc_, device_, filename.c_str()) != ORB_OK) {
        fprintf(stderr, "ORBVocabulary::loadFromTextFile: %s\n", orb_last_error());   // the reference prints and returns false
        return false;
    }
    orb_voc_info(voc_, nullptr, nullptr, &nNodes_, &nWords_, nullptr, nullptr);
    return true;
}

bool ORBVocabulary::create(int k, int L, int scoring, int weighting, int nNodes, const int32_t* parent, const uint8_t* isLeaf,
                           const uint8_t* desc32, const double* weight) {
    orb_voc_destroy(voc_);
    voc_ = nullptr;
    if (orb_voc_create(&voc_, device_, k, L, scoring, weighting, nNodes, parent, isLeaf, desc32, weight) != ORB_OK) {
        fprintf(stderr, "ORBVocabulary::create: %s\n", orb_last_error());
        return false;
    }
    orb_voc_info(voc_, nullptr, nullptr, &nNodes_, &nWords_, nullptr, nullptr);
    return true;
}

bool ORBVocabulary::empty() const { return !voc_ || nWords_ == 0; }
unsigned int ORBVocabulary::size() const { return (unsigned int)nWords_; }

void ORBVocabulary::transform(const uint8_t* desc32, int n, DBoW2::BowVector& v, DBoW2::FeatureVector& fv, int levelsup) const {
    v.clear();
    fv.clear();
    if (empty() || n <= 0) return;                                     // TemplatedVocabulary.h:1147-1150
    const int32_t off[2] = {0, n};
    int32_t nb = 0, nf = 0;
    std::vector<int32_t> bw(n), fnode(n), fstart(n + 1), ffeat(n);
    std::vector<double> bv(n);
    if (orb_bow_transform(voc_, desc32, off, 1, levelsup, &nb, bw.data(), bv.data(), &nf, fnode.data(), fstart.data(), ffeat.data()) != ORB_OK) {
        fprintf(stderr, "ORBVocabulary::transform: %s\n", orb_last_error());
        return;
    }
    for (int j = 0; j < nb; ++j) v.insert(v.end(), std::make_pair((DBoW2::WordId)bw[j], bv[j]));       // already in key order
    for (int j = 0; j < nf; ++j) {
        std::vector<unsigned int>& dst = fv.insert(fv.end(), std::make_pair((DBoW2::NodeId)fnode[j], std::vector<unsigned int>()))->second;
        dst.assign(ffeat.begin() + fstart[j], ffeat.begin() + fstart[j + 1]);
    }
}

void ORBVocabulary::transform(const std::vector<cv::Mat>& features, DBoW2::BowVector& v, DBoW2::FeatureVector& fv, int levelsup) const {
    std::vector<uint8_t> flat(features.size() * 32);
    for (size_t i = 0; i < features.size(); ++i) memcpy(&flat[i * 32], features[i].ptr(0), 32);
    transform(flat.data(), (int)features.size(), v, fv, levelsup);
}

double ORBVocabulary::score(const DBoW2::BowVector& v1, const DBoW2::BowVector& v2) const {
    DBoW2::BowVector::const_iterator a = v1.begin(), b = v2.begin();
    double score = 0;
    while (a != v1.end() && b != v2.end()) {
        if (a->first == b->first) { score += fabs(a->second - b->second) - fabs(a->second) - fabs(b->second); ++a; ++b; }
        else if (a->first < b->first) a = v1.lower_bound(b->first);
        else b = v2.lower_bound(a->first);
    }
    return -score / 2.0;
}

}  // namespace ORB_SLAM2
