// host_check.cpp — drives the C++ host shims (ORBextractor / ORBmatcher / ComputeStereoMatches) the way Frame.cc does
// and dumps the results for tests/test_gpu_host_shim.py, which compares them with the oracle.
//   host_check <w> <h> <nfeatures> <nlevels> <left.raw> <out.bin> [<right.raw> <bf> <b>]
// out.bin: int32 n | n x KeyPoint(28 B) | n x 32 B | int32 level0 ROI step | w*h level-0 ROI bytes
//          [ | int32 nR | nR x KeyPoint | nR x 32 B | int32 nmatches | n x float mvuRight | n x float mvDepth ]
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "ORBextractor.h"
#include "ORBmatcher.h"

using namespace ORB_SLAM2;

static cv::Mat read_raw(const char* path, int w, int h) {
    cv::Mat m(h, w, CV_8U);
    FILE* f = fopen(path, "rb");
    if (!f || fread(m.data, 1, (size_t)w * h, f) != (size_t)w * h) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
    fclose(f);
    return m;
}

int main(int argc, char** argv) {
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
