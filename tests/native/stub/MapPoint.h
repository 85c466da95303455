// TEST STAND-IN for the reference's include/se2lam/MapPoint.h (members used by ORBmatcher: MapPoint.h:43-57).
#pragma once
#include <memory>
#include "se2lam/cv_compat.h"
namespace se2lam {
class KeyFrame;
class MapPoint {
public:
    bool mNull = false, mGoodPrl = true;
    cv::Point3f mPos;
    cv::Mat mMainDescriptor;
    int mMainOctave = 0;
    bool isGoodPrl() { return mGoodPrl; }
    bool isNull() { return mNull; }
    cv::Point3f getPos() { return mPos; }
};
typedef std::shared_ptr<MapPoint> PtrMapPoint;
}
