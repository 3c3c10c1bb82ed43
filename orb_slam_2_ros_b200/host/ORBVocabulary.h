// ORBVocabulary.h — ORB_SLAM2::ORBVocabulary (reference orb_slam2/include/ORBVocabulary.h:31
// = DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB>) over the C ABI.  Same method names and
// argument meaning as the reference for the calls ORB-SLAM2 makes: loadFromTextFile (System.cc), transform(features,
// BowVector&, FeatureVector&, levelsup) (Frame.cc:428-435, KeyFrame.cc:68-77), score (KeyFrameDatabase.cc), size / empty.
// DBoW2::BowVector / FeatureVector keep their reference definitions (std::map subclasses, BowVector.h:56, FeatureVector.h:21).
#ifndef ORBVOCABULARY_H
#define ORBVOCABULARY_H

#include <map>
#include <string>
#include <vector>

#include "cv_compat.h"

struct orb_voc;

// The header this file replaces pulls in DBoW2/TemplatedVocabulary.h, whose line 36 is a global `using namespace std;` — and the
// rest of ORB-SLAM2 has come to depend on it (Frame.h:91 `vector<size_t>`, Frame.cc:79 `thread`, ORBmatcher.h:70 `pair<>`, ...).
// A drop-in replacement has to keep that side effect or those files stop compiling.
using namespace std;

// DBoW2::BowVector / FeatureVector: the reference's own headers when they are on the include path (a real ORB-SLAM2 tree),
// else the two std::map subclasses they are (BowVector.h:56, FeatureVector.h:21)
#if __has_include("Thirdparty/DBoW2/DBoW2/BowVector.h")
#include "Thirdparty/DBoW2/DBoW2/BowVector.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"
#else
namespace DBoW2 {
typedef unsigned int WordId;
typedef double WordValue;
typedef unsigned int NodeId;
class BowVector : public std::map<WordId, WordValue> {};
class FeatureVector : public std::map<NodeId, std::vector<unsigned int> > {};
}  // namespace DBoW2
#endif

namespace ORB_SLAM2 {

class ORBVocabulary {
public:
    explicit ORBVocabulary(int device = 0) : device_(device) {}
    ~ORBVocabulary();
    ORBVocabulary(const ORBVocabulary&) = delete;
    ORBVocabulary& operator=(const ORBVocabulary&) = delete;

    bool loadFromTextFile(const std::string& filename);                     // TemplatedVocabulary.h:1351-1441
    // in-memory form of the same file (node 0 = root, parents precede children)
    bool create(int k, int L, int scoring, int weighting, int nNodes, const int32_t* parent, const uint8_t* isLeaf,
                const uint8_t* desc32, const double* weight);
    bool empty() const;                                                      // TemplatedVocabulary.h:  m_words.empty()
    unsigned int size() const;                                               // number of words
    // TemplatedVocabulary.h:1140-1218.  features: 1 x 32 CV_8U rows (Converter::toDescriptorVector)
    void transform(const std::vector<cv::Mat>& features, DBoW2::BowVector& v, DBoW2::FeatureVector& fv, int levelsup) const;
    // the same on the descriptor matrix itself (n x 32, row stride = 32)
    void transform(const uint8_t* desc32, int n, DBoW2::BowVector& v, DBoW2::FeatureVector& fv, int levelsup) const;
    // L1Scoring::score (ScoringObject.cpp:23-66) — host arithmetic on two short sorted vectors, as in the reference
    double score(const DBoW2::BowVector& a, const DBoW2::BowVector& b) const;
    orb_voc* handle() const { return voc_; }

private:
    orb_voc* voc_ = nullptr;
    int device_;
    int nWords_ = 0, nNodes_ = 0;
};

}  // namespace ORB_SLAM2
#endif
