// host_check.cpp — drives the C++ host shims (ORBextractor / ORBmatcherArrays / ComputeStereoMatches) the way Frame.cc does
// and dumps the results for tests/test_gpu_host_shim.py, which compares them with the oracle.
//   host_check <w> <h> <nfeatures> <nlevels> <left.raw> <out.bin> [<right.raw> <bf> <b>]
//   host_check bow <voc.txt> <w> <h> <a.raw> <b.raw> <out.bin>   (Frame::ComputeBoW on both frames + SearchByBoW)
// out.bin: int32 n | n x KeyPoint(28 B) | n x 32 B | int32 level0 ROI step | w*h level-0 ROI bytes
//          [ | int32 nR | nR x KeyPoint | nR x 32 B | int32 nmatches | n x float mvuRight | n x float mvDepth ]
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "ORBextractor.h"
#include "ORBVocabulary.h"
#include "ORBmatcherArrays.h"
#include <cstring>

using namespace ORB_SLAM2;

static cv::Mat read_raw(const char* path, int w, int h) {
    cv::Mat m(h, w, CV_8U);
    FILE* f = fopen(path, "rb");
    if (!f || fread(m.data, 1, (size_t)w * h, f) != (size_t)w * h) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
    fclose(f);
    return m;
}

// Frame::ComputeBoW (Frame.cc:428-435) on two extracted frames, then ORBmatcherArrays::SearchByBoW(KeyFrame, Frame) with every
// keypoint of the first frame holding a map point.  Dump: per frame int32 n | n x 32 B | int32 nb | nb x (u32 word, f64 value)
// | int32 nf | per node (u32 node, int32 count, count x u32) ; then int32 nmatches | n1 x int32 match12 | n2 x int32 match21.
static int bow_main(char** argv) {
    ORBVocabulary voc;
    if (!voc.loadFromTextFile(argv[2])) return 3;
    const int w = atoi(argv[3]), h = atoi(argv[4]);
    cv::Mat im[2] = {read_raw(argv[5], w, h), read_raw(argv[6], w, h)};
    ORBextractor ex(1000, 1.2f, 8, 20, 7);
    std::vector<cv::KeyPoint> k[2];
    cv::Mat d[2];
    DBoW2::BowVector bv[2];
    DBoW2::FeatureVector fv[2];
    FILE* o = fopen(argv[7], "wb");
    if (!o) return 2;
    for (int f = 0; f < 2; ++f) {
        ex(im[f], cv::Mat(), k[f], d[f]);
        std::vector<cv::Mat> rows;                                     // Converter::toDescriptorVector
        for (int i = 0; i < d[f].rows; ++i) rows.push_back(d[f].row(i));
        voc.transform(rows, bv[f], fv[f], 4);
        int n = (int)k[f].size(), nb = (int)bv[f].size(), nf = (int)fv[f].size();
        fwrite(&n, 4, 1, o);
        if (n) fwrite(d[f].data, 32, n, o);
        fwrite(&nb, 4, 1, o);
        for (DBoW2::BowVector::const_iterator it = bv[f].begin(); it != bv[f].end(); ++it) { fwrite(&it->first, 4, 1, o); fwrite(&it->second, 8, 1, o); }
        fwrite(&nf, 4, 1, o);
        for (DBoW2::FeatureVector::const_iterator it = fv[f].begin(); it != fv[f].end(); ++it) {
            int c = (int)it->second.size();
            fwrite(&it->first, 4, 1, o); fwrite(&c, 4, 1, o); fwrite(it->second.data(), 4, c, o);
        }
    }
    std::vector<float> a[2];
    for (int f = 0; f < 2; ++f) for (size_t i = 0; i < k[f].size(); ++i) a[f].push_back(k[f][i].angle);
    std::vector<int32_t> m12, m21;
    ORBmatcherArrays matcher(0.7f, true);
    int nm = matcher.SearchByBoW(d[0].data, a[0].data(), nullptr, (int)k[0].size(), fv[0], d[1].data, a[1].data(), nullptr, (int)k[1].size(), fv[1],
                                 false, m12, m21);
    fwrite(&nm, 4, 1, o);
    fwrite(m12.data(), 4, m12.size(), o);
    fwrite(m21.data(), 4, m21.size(), o);
    fclose(o);
    printf("host_check bow ok: %d / %d keypoints, %zu / %zu words, %d matches, self score %.3f\n", (int)k[0].size(), (int)k[1].size(), bv[0].size(),
           bv[1].size(), nm, voc.score(bv[0], bv[0]));
    return 0;
}

int main(int argc, char** argv) {
    if (argc == 8 && !strcmp(argv[1], "bow")) return bow_main(argv);
    if (argc != 7 && argc != 10) { fprintf(stderr, "usage: host_check w h nfeatures nlevels left.raw out.bin [right.raw bf b]\n"); return 2; }
    const int w = atoi(argv[1]), h = atoi(argv[2]), nf = atoi(argv[3]), nl = atoi(argv[4]);
    const bool stereo = argc == 10;
    cv::Mat imL = read_raw(argv[5], w, h), imR;
    ORBextractor exL(nf, 1.2f, nl, 20, 7), exR(nf, 1.2f, nl, 20, 7);
    std::vector<cv::KeyPoint> kL, kR;
    cv::Mat dL, dR;
    if (stereo) {
        imR = read_raw(argv[7], w, h);
        // Frame.cc:79-82: the two extractors run on two threads
        std::thread tl([&] { exL(imL, cv::Mat(), kL, dL); });
        std::thread tr([&] { exR(imR, cv::Mat(), kR, dR); });
        tl.join(); tr.join();
    } else {
        exL(imL, cv::Mat(), kL, dL);
    }
    FILE* o = fopen(argv[6], "wb");
    if (!o) return 2;
    int n = (int)kL.size();
    fwrite(&n, 4, 1, o);
    fwrite(kL.data(), sizeof(cv::KeyPoint), n, o);
    if (n) fwrite(dL.data, 32, n, o);
    const cv::Mat& p0 = exL.mvImagePyramid[0];
    int step = (int)p0.step;
    fwrite(&step, 4, 1, o);
    for (int y = 0; y < p0.rows; ++y) fwrite(p0.ptr(y), 1, p0.cols, o);
    if (stereo) {
        int nR = (int)kR.size();
        fwrite(&nR, 4, 1, o);
        fwrite(kR.data(), sizeof(cv::KeyPoint), nR, o);
        if (nR) fwrite(dR.data, 32, nR, o);
        std::vector<float> uR, depth;
        int nm = ComputeStereoMatches(exL, exR, kL, dL, kR, dR, (float)atof(argv[8]), (float)atof(argv[9]), uR, depth);
        fwrite(&nm, 4, 1, o);
        fwrite(uR.data(), 4, n, o);
        fwrite(depth.data(), 4, n, o);
    }
    fclose(o);
    printf("host_check ok: %d keypoints%s\n", n, stereo ? " (stereo)" : "");
    return 0;
}
