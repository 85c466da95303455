// TEST STAND-IN for the reference's include/se2lam/Frame.h: only the members se2lam::ORBmatcher touches
// (Frame.h:48-71). The real project keeps its own Frame.h; the matcher shim never replaces it.
#pragma once
#include <memory>
#include <vector>
#include "se2lam/cv_compat.h"
namespace se2lam {
class Frame {
public:
    static float minXUn, minYUn, maxXUn, maxYUn;
    static float mfGridElementWidthInv, mfGridElementHeightInv;
    std::vector<cv::KeyPoint> keyPoints;
    std::vector<cv::KeyPoint> keyPointsUn;
    cv::Mat descriptors;
    int N = 0;
    cv::Mat Tcw;
    bool inImgBound(cv::Point2f pt) { return pt.x >= minXUn && pt.x <= maxXUn && pt.y >= minYUn && pt.y <= maxYUn; }
};
typedef std::shared_ptr<Frame> PtrFrame;
}
