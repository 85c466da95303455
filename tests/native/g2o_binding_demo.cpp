// Drives include/se2lam/g2o_gpu_levenberg.h through the g2o API exactly like the reference with the SlamAlgorithm typedef
// swapped (optimizer.h:32): Map::loadLocalGraph-style graph construction (src/Map.cpp:891-1053, src/optimizer.cpp:17-62,
// 207-215, 316-324), LocalMapper::localBA (src/LocalMapper.cpp:239-260), Map::optimizeLocalGraph read-back - against the
// g2o mock of tests/native/mock_g2o (g2o itself is absent from this container). Same binary layouts as shim_demo.cpp.
// usage: g2o_binding_demo <in.bin> <out.bin>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "se2lam/g2o_gpu_levenberg.h"

typedef g2o::Solver SlamBlockSolver;                          // the block / linear solver stack is unused on the GPU path
typedef se2gpu::G2oGpuLevenberg SlamAlgorithm;                // <- the one-line change of optimizer.h:32
typedef g2o::SparseOptimizer SlamOptimizer;
typedef g2o::CameraParameters CamPara;

template <class T> static void rd(FILE* f, T* p, size_t n) { if (n && fread(p, sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }
template <class T> static void wr(FILE* f, const T* p, size_t n) { fwrite(p, sizeof(T), n, f); }

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    FILE* fi = fopen(argv[1], "rb"); FILE* fo = fopen(argv[2], "wb");
    if (!fi || !fo) return 2;
    int w, h;
    rd(fi, &w, 1); rd(fi, &h, 1);
    std::vector<unsigned char> pix((size_t)w * h);
    rd(fi, pix.data(), pix.size());                            // (the ORB part of the shared input file is skipped)
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
    SlamAlgorithm* solver = new SlamAlgorithm(new SlamBlockSolver());
    g2o::Matrix3D Rbc; g2o::Vector3D tbc;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Rbc(r, c) = Tbc[3 * r + c];
    for (int i = 0; i < 3; ++i) tbc[i] = Tbc[9 + i];
    const g2o::SE3Quat bTc(Rbc, tbc);
    solver->setExtrinsic(bTc);                                                  // Config::bTc
    optimizer.setAlgorithm(solver);
    optimizer.setVerbose(false);
    bool mbAbortBA = false;
    optimizer.setForceStopFlag(&mbAbortBA);

    g2o::Vector2D pp; pp[0] = cam[1]; pp[1] = cam[2];
    CamPara* campr = new CamPara(cam[0], pp, 0.);                               // addCamPara, optimizer.cpp:207-215
    campr->setId(0);
    optimizer.addParameter(campr);
    for (int i = 0; i < P; ++i) {                                               // addVertexSE2, optimizer.cpp:34-43
        g2o::VertexSE2* v = new g2o::VertexSE2;
        v->setId(i); v->setEstimate(g2o::SE2(poses[3 * i], poses[3 * i + 1], poses[3 * i + 2])); v->setFixed(fixed[i] != 0);
        optimizer.addVertex(v);
    }
    for (int o = 0; o < O; ++o) {                                               // addEdgeSE2, optimizer.cpp:52-62
        g2o::PreEdgeSE2* e = new g2o::PreEdgeSE2;
        e->vertices()[0] = optimizer.vertex(oi[o]); e->vertices()[1] = optimizer.vertex(oj[o]);
        g2o::Vector3D m; for (int k = 0; k < 3; ++k) m[k] = om[3 * o + k];
        g2o::Matrix3D inf;
        inf(0, 0) = oinf[6 * o]; inf(0, 1) = inf(1, 0) = oinf[6 * o + 1]; inf(0, 2) = inf(2, 0) = oinf[6 * o + 2];
        inf(1, 1) = oinf[6 * o + 3]; inf(1, 2) = inf(2, 1) = oinf[6 * o + 4]; inf(2, 2) = oinf[6 * o + 5];
        e->setMeasurement(m); e->setInformation(inf);
        optimizer.addEdge(e);
    }
    const int maxKFid = P + 1;
    for (int j = 0; j < L; ++j) {                                               // addVertexSBAXYZ, optimizer.cpp:316-324
        g2o::VertexSBAPointXYZ* v = new g2o::VertexSBAPointXYZ();
        g2o::Vector3D p; for (int k = 0; k < 3; ++k) p[k] = points[3 * j + k];
        v->setEstimate(p); v->setId(maxKFid + j); v->setMarginalized(true); v->setFixed(false);
        optimizer.addVertex(v);
    }
    for (int e = 0; e < E; ++e) {                                               // addEdgeSE2XYZ, optimizer.cpp:17-32
        g2o::EdgeSE2XYZ* ed = new g2o::EdgeSE2XYZ;
        ed->vertices()[0] = optimizer.vertex(ep[e]); ed->vertices()[1] = optimizer.vertex(maxKFid + el[e]);
        ed->setCameraParameter(campr); ed->setExtParameter(bTc);
        g2o::Vector2D m; m[0] = uv[2 * e]; m[1] = uv[2 * e + 1];
        g2o::Matrix2D inf; inf(0, 0) = info[3 * e]; inf(0, 1) = inf(1, 0) = info[3 * e + 1]; inf(1, 1) = info[3 * e + 2];
        ed->setMeasurement(m); ed->setInformation(inf);
        g2o::RobustKernelHuber* rk = new g2o::RobustKernelHuber; rk->setDelta(delta);
        ed->setRobustKernel(rk);
        optimizer.addEdge(ed);
    }
    optimizer.initializeOptimization(0);                                        // LocalMapper.cpp:259
    int done = optimizer.optimize(iters);                                       // LocalMapper.cpp:260: g2o calls solve(0..iters-1)
    int on_gpu = solver->onGpu() ? 1 : 0;
    wr(fo, &done, 1); wr(fo, &on_gpu, 1);
    for (int i = 0; i < P; ++i) { g2o::Vector3D vp = static_cast<g2o::VertexSE2*>(optimizer.vertex(i))->estimate().toVector(); wr(fo, vp.d, 3); }
    for (int j = 0; j < L; ++j) { g2o::Vector3D p = static_cast<g2o::VertexSBAPointXYZ*>(optimizer.vertex(j + maxKFid))->estimate(); wr(fo, p.d, 3); }
    double lam = solver->lastStats().lambda; int trials = solver->levenbergIteration();
    wr(fo, &lam, 1); wr(fo, &trials, 1);
    fclose(fi); fclose(fo);
    return 0;
}
