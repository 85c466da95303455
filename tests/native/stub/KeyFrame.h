// TEST STAND-IN for the reference's include/se2lam/KeyFrame.h (members used by ORBmatcher: KeyFrame.h:62-125) and for
// DBoW2::FeatureVector (Thirdparty/DBoW2/DBoW2/FeatureVector.h: a std::map<NodeId, std::vector<unsigned int> >).
#pragma once
#include <map>
#include <memory>
#include <vector>
#include "Frame.h"
#include "MapPoint.h"
namespace DBoW2 { typedef std::map<unsigned int, std::vector<unsigned int> > FeatureVector; }
namespace se2lam {
class KeyFrame : public Frame {
public:
    bool mNull = false;
    std::vector<unsigned char> mObserved;            // hasObservation(idx)
    std::vector<PtrMapPoint> mvpMapPoints;
    DBoW2::FeatureVector mFeatVec;
    bool isNull() { return mNull; }
    bool hasObservation(const PtrMapPoint&) { return false; }
    bool hasObservation(int idx) { return idx < (int)mObserved.size() && mObserved[idx]; }
    DBoW2::FeatureVector GetFeatureVector() { return mFeatVec; }
    std::vector<PtrMapPoint> GetMapPointMatches() { return mvpMapPoints; }
};
typedef std::shared_ptr<KeyFrame> PtrKeyFrame;
}
