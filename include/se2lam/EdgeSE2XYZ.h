// include/se2lam/EdgeSE2XYZ.h of the drop-in header set (reference include/se2lam/EdgeSE2XYZ.h:32-102: g2o::EdgeSE2XYZ, the
// SE(2) pose - XYZ landmark reprojection edge, and g2o::PreEdgeSE2, the SE(2) odometry prior).
//
// * With g2o installed the project's OWN header is the one to use - the edge classes stay g2o::BaseBinaryEdge subclasses so that
//   graph construction, removeOutlierChi2 and the SE(3) paths keep working; the GPU takes over at the algorithm level
//   (g2o_gpu_levenberg.h). This file then only forwards to it (it shadows it because this include directory is put first
//   for ORBextractor.h / ORBmatcher.h).
// * Without g2o (this build container) the classes come from g2o_compat.h: same names, setters and per-edge queries
//   (setCameraParameter, setExtParameter, setMeasurement, setInformation, setRobustKernel, computeError, chi2, error,
//   information, level / setLevel), recorded for the C ABI. computeError / chi2 evaluate ONE edge on the host the way
//   LocalMapper::removeOutlierChi2 asks for them after a BA (src/LocalMapper.cpp:187-213); the optimisation itself never
//   calls them - linearisation, robust weighting and accumulation of all edges run in libse2gpu.so.
#ifndef EDGE_SE2_XYZ_H
#define EDGE_SE2_XYZ_H

#if defined(__has_include)
#if __has_include(<g2o/core/base_binary_edge.h>)
#define SE2LAM_HAVE_G2O 1
#endif
#endif

#ifdef SE2LAM_HAVE_G2O
#include_next <se2lam/EdgeSE2XYZ.h>
#else
#include "g2o_compat.h"
#endif

#endif
