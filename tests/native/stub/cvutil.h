// TEST STAND-IN for the reference's include/se2lam/cvutil.h (cvu::camprjc / cvu::se3map, cvutil.h:30-34). The test
// feeds map points whose position IS (u, v, 1) with an identity camera and pose, so the projection is exact.
#pragma once
#include "se2lam/cv_compat.h"
namespace cvu {
inline cv::Point3f se3map(const cv::Mat&, const cv::Point3f& p) { return p; }
inline cv::Point2f camprjc(const cv::Mat&, const cv::Point3f& p) { return cv::Point2f(p.x / p.z, p.y / p.z); }
}
