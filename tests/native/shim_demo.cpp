// Exercises the drop-in C++ headers the way the reference's own call sites do:
//   Frame::Frame            (src/Frame.cpp:25)          (*mpORBExtractor)(img, cv::Mat(), keyPoints, descriptors)
//   Map::loadLocalGraph     (src/Map.cpp:891-1053)      addCamPara / addVertexSE2 / addEdgeSE2 / addVertexSBAXYZ / addEdgeSE2XYZ
//   LocalMapper::localBA    (src/LocalMapper.cpp:239-260) solver stack, setForceStopFlag, initializeOptimization(0), optimize(N)
//   Map::optimizeLocalGraph (src/Map.cpp:754-783)       estimateVertexSE2 / estimateVertexSBAXYZ
// usage: shim_demo <in.bin> <out.bin>   (binary layouts written/read by tests/test_cpp_shim.py)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "se2lam/ORBextractor.h"
#include "se2lam/optimizer.h"

using namespace se2lam;

template <class T> static void rd(FILE* f, T* p, size_t n) { if (fread(p, sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }
template <class T> static void wr(FILE* f, const T* p, size_t n) { fwrite(p, sizeof(T), n, f); }

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    FILE* fi = fopen(argv[1], "rb"); FILE* fo = fopen(argv[2], "wb");
    if (!fi || !fo) return 2;
    // ---------------------------------------------------------------- ORB
    int w, h;
    rd(fi, &w, 1); rd(fi, &h, 1);
    std::vector<unsigned char> pix((size_t)w * h);
    rd(fi, pix.data(), pix.size());
    cv::Mat img(h, w, CV_8UC1, pix.data());
    ORBextractor* mpORBExtractor = new ORBextractor(1000, 1.2f, 8);            // Track.cpp:34
    std::vector<cv::KeyPoint> keyPoints;
    cv::Mat descriptors;
    (*mpORBExtractor)(img, cv::Mat(), keyPoints, descriptors);                  // Frame.cpp:25
    int N = (int)keyPoints.size();
    wr(fo, &N, 1);
    wr(fo, keyPoints.data(), keyPoints.size());
    for (int i = 0; i < N; ++i) wr(fo, descriptors.ptr<unsigned char>(i), 32);
    int levels = mpORBExtractor->GetLevels(); float sf = mpORBExtractor->GetScaleFactor();
    wr(fo, &levels, 1); wr(fo, &sf, 1);
    // empty image -> outputs untouched
    std::vector<cv::KeyPoint> kp2(3); cv::Mat d2;
    (*mpORBExtractor)(cv::Mat(), cv::Mat(), kp2, d2);
    int untouched = (int)kp2.size();
    wr(fo, &untouched, 1);
    delete mpORBExtractor;
    // ---------------------------------------------------------------- local BA
    int P, L, E, O, iters;
    rd(fi, &P, 1); rd(fi, &L, 1); rd(fi, &E, 1); rd(fi, &O, 1); rd(fi, &iters, 1);
    std::vector<double> poses(3 * P), points(3 * L), uv(2 * E), info(3 * E), om(3 * O), oinf(6 * O), cam(3), Tbc(12);
    std::vector<unsigned char> fixed(P);
    std::vector<int> ep(E), el(E), oi(O), oj(O);
    double delta;
    rd(fi, poses.data(), poses.size()); rd(fi, fixed.data(), fixed.size()); rd(fi, points.data(), points.size());
    rd(fi, ep.data(), ep.size()); rd(fi, el.data(), el.size()); rd(fi, uv.data(), uv.size()); rd(fi, info.data(), info.size());
    rd(fi, oi.data(), oi.size()); rd(fi, oj.data(), oj.size()); rd(fi, om.data(), om.size()); rd(fi, oinf.data(), oinf.size());
    rd(fi, cam.data(), 3); rd(fi, Tbc.data(), 12); rd(fi, &delta, 1);

    SlamOptimizer optimizer;                                                    // LocalMapper.cpp:238-246
    SlamLinearSolver* linearSolver = new SlamLinearSolver();
    SlamBlockSolver* blockSolver = new SlamBlockSolver(linearSolver);
    SlamAlgorithm* solver = new SlamAlgorithm(blockSolver);
    optimizer.setAlgorithm(solver);
    optimizer.setVerbose(false);
    bool mbAbortBA = false;
    optimizer.setForceStopFlag(&mbAbortBA);

    cv::Mat Kcam(3, 3, CV_32FC1);                                               // Config::Kcam
    Kcam.at<float>(0, 0) = (float)cam[0]; Kcam.at<float>(0, 2) = (float)cam[1]; Kcam.at<float>(1, 1) = (float)cam[0]; Kcam.at<float>(1, 2) = (float)cam[2];
    Kcam.at<float>(2, 2) = 1.f;
    CamPara* campr = addCamPara(optimizer, Kcam, 0);                            // Map.cpp:897, the reference's own call shape
    for (int i = 0; i < P; ++i)                                                 // Map.cpp:918-932
        addVertexSE2(optimizer, g2o::SE2(poses[3 * i], poses[3 * i + 1], poses[3 * i + 2]), i, fixed[i] != 0);
    for (int o = 0; o < O; ++o) {                                               // Map.cpp:935-956
        g2o::Matrix3D inf;
        inf(0, 0) = oinf[6 * o]; inf(0, 1) = inf(1, 0) = oinf[6 * o + 1]; inf(0, 2) = inf(2, 0) = oinf[6 * o + 2];
        inf(1, 1) = oinf[6 * o + 3]; inf(1, 2) = inf(2, 1) = oinf[6 * o + 4]; inf(2, 2) = oinf[6 * o + 5];
        addEdgeSE2(optimizer, g2o::makeVector3D(om[3 * o], om[3 * o + 1], om[3 * o + 2]), oi[o], oj[o], inf);
    }
    const int maxKFid = P + 1;                                                  // Map.cpp:972
    g2o::Matrix3D Rbc; g2o::Vector3D tbc;
    for (int i = 0; i < 9; ++i) Rbc.d[i] = Tbc[i];
    for (int i = 0; i < 3; ++i) tbc[i] = Tbc[9 + i];
    const g2o::SE3Quat bTc(Rbc, tbc);
    for (int j = 0; j < L; ++j)                                                 // Map.cpp:985-988
        addVertexSBAXYZ(optimizer, g2o::makeVector3D(points[3 * j], points[3 * j + 1], points[3 * j + 2]), maxKFid + j);
    for (int e = 0; e < E; ++e) {                                               // Map.cpp:1048-1049
        g2o::Matrix2D inf; inf(0, 0) = info[3 * e]; inf(0, 1) = inf(1, 0) = info[3 * e + 1]; inf(1, 1) = info[3 * e + 2];
        addEdgeSE2XYZ(optimizer, g2o::makeVector2D(uv[2 * e], uv[2 * e + 1]), ep[e], maxKFid + el[e], campr, bTc, inf, delta);
    }
    optimizer.initializeOptimization(0);                                        // LocalMapper.cpp:259
    int done = optimizer.optimize(iters);                                       // LocalMapper.cpp:260
    wr(fo, &done, 1);
    for (int i = 0; i < P; ++i) { g2o::Vector3D vp = estimateVertexSE2(optimizer, i).toVector(); wr(fo, vp.d, 3); }   // Map.cpp:768
    for (int j = 0; j < L; ++j) { g2o::Vector3D p = estimateVertexSBAXYZ(optimizer, j + maxKFid); wr(fo, p.d, 3); }   // Map.cpp:777
    // per-edge outlier test after the BA, as LocalMapper::removeOutlierChi2 does it (LocalMapper.cpp:187-213): computeError() + chi2()
    double chi2_sum = 0; int n_edges = 0;
    for (g2o::Edge* e : optimizer.edges()) { e->computeError(); chi2_sum += e->chi2(); ++n_edges; }
    wr(fo, &chi2_sum, 1); wr(fo, &n_edges, 1);
    // an aborted BA performs no iteration (setForceStopFlag, Track.cpp:372)
    mbAbortBA = true;
    int aborted = optimizer.optimize(iters);
    wr(fo, &aborted, 1);
    fclose(fi); fclose(fo);
    return 0;
}
