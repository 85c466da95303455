// Drop-in subset of se2lam's include/se2lam/optimizer.h for the SE(2)-XYZ local-BA path: the solver typedefs
// (reference optimizer.h:30-34) and the graph helpers Map::loadLocalGraph / optimizeLocalGraph call
// (optimizer.h:77-110, 140-141; bodies at src/optimizer.cpp:17-62, 199-215, 316-324, 549-554), with identical
// names, argument order and ownership (the optimizer owns what is added to it). The arithmetic runs in
// libse2gpu.so through the C ABI; see g2o_compat.h for what is covered when real g2o is absent.
// The SE(3) helpers of the reference header (pose-graph / PnP paths) are out of scope.
#ifndef OPTIMIZER_H
#define OPTIMIZER_H

#include "g2o_compat.h"
#include "EdgeSE2XYZ.h"

#if defined(__has_include)
#if __has_include(<opencv2/core/core.hpp>)
#include <opencv2/core/core.hpp>
#ifndef SE2LAM_HAVE_OPENCV
#define SE2LAM_HAVE_OPENCV 1
#endif
#endif
#endif
#ifndef SE2LAM_HAVE_OPENCV
#include "cv_compat.h"
#endif

namespace se2lam{

typedef g2o::BlockSolverGpu SlamBlockSolver;
typedef g2o::LinearSolverGpu SlamLinearSolver;
typedef g2o::OptimizationAlgorithmGpuLM SlamAlgorithm;
typedef g2o::SparseOptimizer SlamOptimizer;
typedef g2o::CameraParameters CamPara;

inline void
initOptimizer(SlamOptimizer &opt, bool verbose=false){                         // optimizer.cpp:199-205
    SlamLinearSolver* linearSolver = new SlamLinearSolver();
    SlamBlockSolver* blockSolver = new SlamBlockSolver(linearSolver);
    SlamAlgorithm* solver = new SlamAlgorithm(blockSolver);
    opt.setAlgorithm(solver);
    opt.setVerbose(verbose);
}

// K as a raw 3x3 row-major float array (Config::Kcam.ptr<float>()); the cv::Mat overload below is the reference's signature
inline CamPara*
addCamPara(SlamOptimizer &opt, const float* K, int id){                        // optimizer.cpp:207-215
    g2o::Vector2D principal_point = g2o::makeVector2D(K[2], K[5]);
    CamPara* campr = new CamPara(K[0], principal_point, 0.);
    campr->setId(id);
    opt.addParameter(campr);
    return campr;
}

// the reference's own signature (optimizer.cpp:207-215): K is Config::Kcam, a 3x3 CV_32F cv::Mat
inline CamPara*
addCamPara(SlamOptimizer &opt, const cv::Mat& K, int id){
    const float Kf[9] = {K.at<float>(0,0), K.at<float>(0,1), K.at<float>(0,2), K.at<float>(1,0), K.at<float>(1,1), K.at<float>(1,2),
                         K.at<float>(2,0), K.at<float>(2,1), K.at<float>(2,2)};
    return addCamPara(opt, Kf, id);
}

inline g2o::VertexSE2*
addVertexSE2(SlamOptimizer &opt, const g2o::SE2& pose, int id, bool fixed = false){   // optimizer.cpp:34-43
    g2o::VertexSE2* v = new g2o::VertexSE2;
    v->setId(id);
    v->setEstimate(pose);
    v->setFixed(fixed);
    opt.addVertex(v);
    return v;
}

inline g2o::SE2
estimateVertexSE2(SlamOptimizer &opt, int id){                                 // optimizer.cpp:45-50
    g2o::VertexSE2* v = static_cast<g2o::VertexSE2*>(opt.vertex(id));
    return v->estimate();
}

inline g2o::PreEdgeSE2*
addEdgeSE2(SlamOptimizer &opt, const g2o::Vector3D& meas, int id0, int id1, const g2o::Matrix3D& info){   // optimizer.cpp:52-62
    g2o::PreEdgeSE2* e = new g2o::PreEdgeSE2;
    e->vertices()[0] = opt.vertex(id0);
    e->vertices()[1] = opt.vertex(id1);
    e->setMeasurement(meas);
    e->setInformation(info);
    opt.addEdge(e);
    return e;
}

inline void
addVertexSBAXYZ(SlamOptimizer &opt, const g2o::Vector3D &xyz, int id, bool marginal=true, bool fixed=false){   // optimizer.cpp:316-324
    g2o::VertexSBAPointXYZ* v = new g2o::VertexSBAPointXYZ();
    v->setEstimate(xyz);
    v->setId(id);
    v->setMarginalized(marginal);
    v->setFixed(fixed);
    opt.addVertex(v);
}

inline g2o::Vector3D
estimateVertexSBAXYZ(SlamOptimizer &opt, int id){                              // optimizer.cpp:549-554
    g2o::VertexSBAPointXYZ* v = static_cast<g2o::VertexSBAPointXYZ*>(opt.vertex(id));
    return v->estimate();
}

inline g2o::EdgeSE2XYZ*
addEdgeSE2XYZ(SlamOptimizer &opt, const g2o::Vector2D& meas, int id0, int id1,
              CamPara* campara, const g2o::SE3Quat &_Tbc, const g2o::Matrix2D &info, double thHuber){   // optimizer.cpp:17-32
    g2o::EdgeSE2XYZ* e = new g2o::EdgeSE2XYZ;
    e->vertices()[0] = opt.vertex(id0);
    e->vertices()[1] = opt.vertex(id1);
    e->setCameraParameter(campara);
    e->setExtParameter(_Tbc);
    e->setMeasurement(meas);
    e->setInformation(info);
    g2o::RobustKernelHuber* rk = new g2o::RobustKernelHuber;
    rk->setDelta(thHuber);
    e->setRobustKernel(rk);
    opt.addEdge(e);
    return e;
}

}// namespace se2lam

#endif // OPTIMIZER_H
