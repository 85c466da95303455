// Drop-in replacement of se2lam's include/se2lam/ORBmatcher.h (reference lines 42-77): same class, same public
// signatures and members, so Track.cpp:131-132, LocalMapper.cpp:117-118, Localizer.cpp:216-217 / 411-413,
// GlobalMapper.cpp:274-276 and MapPoint.cpp:259 compile unchanged. Like the reference header it includes the
// project's own MapPoint.h / KeyFrame.h / Frame.h (those classes are NOT replaced); the methods flatten the object graph
// to arrays and forward to the C ABI (se2gpu_match_by_window / _match_by_projection / _search_by_bow of libse2gpu.so),
// which keeps one device context per GPU for the life of the process - the reference constructs an ORBmatcher on the stack
// for every call, so nothing is allocated per object. There is no CPU matching path. Header-only: link with -lse2gpu.
#ifndef ORBMATCHER_H
#define ORBMATCHER_H

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#if defined(__has_include)
#if __has_include(<opencv2/core/core.hpp>)
#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>
#ifndef SE2LAM_HAVE_OPENCV
#define SE2LAM_HAVE_OPENCV 1
#endif
#endif
#endif
#ifndef SE2LAM_HAVE_OPENCV
#include "cv_compat.h"
#endif

#include "MapPoint.h"
#include "KeyFrame.h"
#include "Frame.h"
#include "Config.h"      // Config::Kcam   (MatchByProjection, ORBmatcher.cpp:397)
#include "cvutil.h"      // cvu::camprjc / cvu::se3map

#include "../se2gpu.h"

namespace se2lam
{

class ORBmatcher
{
public:

    ORBmatcher(float nnratio=0.6, bool checkOri=true) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}

    // Computes the Hamming distance between two ORB descriptors (ORBmatcher.cpp:110-126). Eight popcounts: evaluated
    // inline - the batched device version is se2gpu_hamming_distance, the matchers below never call this per pair.
    static int DescriptorDistance(const cv::Mat &a, const cv::Mat &b)
    {
        const unsigned char* pa = a.ptr<unsigned char>(0);
        const unsigned char* pb = b.ptr<unsigned char>(0);
        int dist = 0;
        for (int i = 0; i < 8; i++) {
            unsigned int va, vb;
            std::memcpy(&va, pa + 4 * i, 4); std::memcpy(&vb, pb + 4 * i, 4);
            dist += __builtin_popcount(va ^ vb);
        }
        return dist;
    }

    // Search matches between MapPoints in a KeyFrame and ORB in a Frame, constrained to the same vocabulary node.
    int SearchByBoW(PtrKeyFrame pKF1, PtrKeyFrame pKF2,
                    std::map<int, int> &mapIdxMatches12, bool bIfMPOnly = true)
    {
        mapIdxMatches12.clear();
        if (pKF1 == NULL || pKF1->isNull() || pKF2 == NULL || pKF2->isNull())
            return 0;                                                        // :133-135
        BowFlat f1, f2;
        flatten(pKF1, f1); flatten(pKF2, f2);
        std::vector<int> m12(f1.kf.n > 0 ? f1.kf.n : 1, -1);
        const int n = se2gpu_search_by_bow(&f1.kf, &f2.kf, bIfMPOnly ? 1 : 0, mfNNratio, mbCheckOrientation ? 1 : 0, m12.data(), 0);
        if (n < 0) die();
        for (int i = 0; i < f1.kf.n; ++i) if (m12[i] >= 0) mapIdxMatches12[i] = m12[i];
        return n;
    }

    void ComputeThreeMaxima(std::vector<int>* histo, const int L, int &ind1, int &ind2, int &ind3)   // :64-105
    {
        int max1 = 0, max2 = 0, max3 = 0;
        ind1 = ind2 = ind3 = -1;
        for (int i = 0; i < L; i++) {
            const int s = (int)histo[i].size();
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
    }

    static const int TH_LOW = 75;
    static const int TH_HIGH = 100;
    static const int HISTO_LENGTH = 30;

    float mfNNratio;
    bool mbCheckOrientation;

    int MatchByWindow(const Frame& frame1, Frame& frame2,
                           std::vector<cv::Point2f>& vbPrevMatched, const int winSize,
                           std::vector<int>& vnMatches12, const int levelOffset = 1,
                           const int minLevel = 0, const int maxLevel = 8)
    {
        vnMatches12 = std::vector<int>(frame1.N, -1);                        // :282
        if (frame1.N == 0) return 0;
        std::vector<unsigned char> pack1, pack2;
        const se2gpu_grid_params grid = {Frame::minXUn, Frame::minYUn, Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv};
        const int n = se2gpu_match_by_window(kps(frame1.keyPointsUn), rows(frame1.descriptors, frame1.N, pack1), frame1.N,
                                             kps(frame2.keyPointsUn), rows(frame2.descriptors, frame2.N, pack2), frame2.N,
                                             reinterpret_cast<float*>(vbPrevMatched.data()), grid, winSize, levelOffset, minLevel, maxLevel,
                                             mfNNratio, vnMatches12.data(), 0);
        if (n < 0) die();
        return n;
    }

    int MatchByProjection(PtrKeyFrame& pNewKF, std::vector<PtrMapPoint>& localMPs, const int winSize, const int levelOffset,
                          std::vector<int>& vMatchesIdxMP)
    {
        const int N = pNewKF->N, M = (int)localMPs.size();
        vMatchesIdxMP = std::vector<int>(N, -1);                             // :387
        if (N == 0 || M == 0) return 0;
        std::vector<unsigned char> observed(N), valid(M, 0), mpdesc((size_t)M * 32, 0), pack;
        std::vector<float> uv(2 * (size_t)M, 0.f);
        std::vector<int> octave(M, 0);
        for (int k = 0; k < N; ++k) observed[k] = pNewKF->hasObservation(k) ? 1 : 0;            // :421
        for (int i = 0; i < M; ++i) {                                        // :390-400
            PtrMapPoint pMP = localMPs[i];
            if (pMP->isNull() || !pMP->isGoodPrl()) continue;
            if (pNewKF->hasObservation(pMP)) continue;
            cv::Point2f predictUV = cvu::camprjc(Config::Kcam, cvu::se3map(pNewKF->Tcw, pMP->getPos()));
            if (!pNewKF->inImgBound(predictUV)) continue;
            valid[i] = 1; uv[2 * i] = predictUV.x; uv[2 * i + 1] = predictUV.y; octave[i] = pMP->mMainOctave;
            std::memcpy(&mpdesc[(size_t)i * 32], pMP->mMainDescriptor.ptr<unsigned char>(0), 32);
        }
        const se2gpu_grid_params grid = {Frame::minXUn, Frame::minYUn, Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv};
        const int n = se2gpu_match_by_projection(kps(pNewKF->keyPointsUn), rows(pNewKF->descriptors, N, pack), N, observed.data(), valid.data(),
                                                 uv.data(), M, octave.data(), mpdesc.data(), grid, winSize, levelOffset, mfNNratio,
                                                 vMatchesIdxMP.data(), 0);
        if (n < 0) die();
        return n;
    }

    float RadiusByViewingCos(const float &viewCos)                           // :54-60
    {
        if(viewCos>0.998)
            return 2.5;
        else
            return 4.0;
    }

private:
    static const se2gpu_keypoint* kps(const std::vector<cv::KeyPoint>& v)
    {
        static_assert(sizeof(cv::KeyPoint) == sizeof(se2gpu_keypoint), "cv::KeyPoint layout");
        return reinterpret_cast<const se2gpu_keypoint*>(v.data());
    }
    // N x 32 CV_8U descriptor rows as one contiguous block (copied only when the Mat is not continuous)
    static const unsigned char* rows(const cv::Mat& d, int n, std::vector<unsigned char>& pack)
    {
        if (n == 0) return NULL;
        if (d.isContinuous()) return d.ptr<unsigned char>(0);
        pack.resize((size_t)n * 32);
        for (int i = 0; i < n; ++i) std::memcpy(&pack[(size_t)i * 32], d.ptr<unsigned char>(i), 32);
        return pack.data();
    }
    struct BowFlat {
        se2gpu_bow_kf kf;
        std::vector<float> angle; std::vector<unsigned char> has_mp, pack; std::vector<int> node, ptr, feat;
    };
    // DBoW2::FeatureVector (std::map<NodeId, std::vector<unsigned int> >, ascending node ids) -> CSR; :138-146
    static void flatten(PtrKeyFrame& pKF, BowFlat& f)
    {
        const int n = pKF->N;
        f.angle.resize(n); f.has_mp.resize(n);
        std::vector<PtrMapPoint> vpMP = pKF->GetMapPointMatches();
        for (int i = 0; i < n; ++i) {
            f.angle[i] = pKF->keyPointsUn[i].angle;
            f.has_mp[i] = (i < (int)vpMP.size() && vpMP[i] && !vpMP[i]->isNull()) ? 1 : 0;
        }
        DBoW2::FeatureVector fv = pKF->GetFeatureVector();
        f.ptr.push_back(0);
        for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it) {
            f.node.push_back((int)it->first);
            for (size_t k = 0; k < it->second.size(); ++k) f.feat.push_back((int)it->second[k]);
            f.ptr.push_back((int)f.feat.size());
        }
        f.kf.angle = f.angle.data(); f.kf.desc = rows(pKF->descriptors, n, f.pack); f.kf.has_mp = f.has_mp.data(); f.kf.n = n;
        f.kf.node = f.node.data(); f.kf.n_node = (int)f.node.size(); f.kf.ptr = f.ptr.data(); f.kf.feat = f.feat.data();
    }
    static void die()
    {
        std::fprintf(stderr, "se2lam::ORBmatcher (GPU): %s\n", se2gpu_last_error());
        std::abort();                                                        // no CPU fallback by design
    }
};

}// namespace se2lam


#endif // ORBMATCHER_H
