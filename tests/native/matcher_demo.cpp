// Exercises the drop-in include/se2lam/ORBmatcher.h the way the reference's call sites do:
//   Track::mTrack                (src/Track.cpp:129-132)          ORBmatcher matcher(0.9); matcher.MatchByWindow(mRefFrame, mFrame, mPrevMatched, 20, mMatchIdx)
//   LocalMapper::findCorrespd    (src/LocalMapper.cpp:117-118)    ORBmatcher matcher; matcher.MatchByProjection(mNewKF, localMPs, 15, 2, vMatchedIdxMPs)
//   GlobalMapper::DetectLoopClose(src/GlobalMapper.cpp:274-276)   ORBmatcher matcher; matcher.SearchByBoW(pKF, pKFLoop, mapMatches, bIfMPOnly)
// against stand-ins of the project's own Frame / KeyFrame / MapPoint headers (tests/native/stub).
// usage: matcher_demo <in.bin> <out.bin>   (binary layouts written/read by tests/test_cpp_shim.py)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "se2lam/ORBmatcher.h"

using namespace se2lam;

float Frame::minXUn = 0, Frame::minYUn = 0, Frame::maxXUn = 640, Frame::maxYUn = 480;
float Frame::mfGridElementWidthInv = 64.f / 640.f, Frame::mfGridElementHeightInv = 48.f / 480.f;
cv::Mat Config::Kcam;

template <class T> static void rd(FILE* f, T* p, size_t n) { if (n && fread(p, sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }
template <class T> static void wr(FILE* f, const T* p, size_t n) { fwrite(p, sizeof(T), n, f); }

static void read_frame(FILE* fi, Frame& fr) {
    int n; rd(fi, &n, 1);
    fr.N = n; fr.keyPointsUn.resize(n); fr.descriptors.create(n, 32, CV_8U);
    rd(fi, fr.keyPointsUn.data(), n);
    rd(fi, fr.descriptors.ptr<unsigned char>(0), (size_t)n * 32);
    fr.keyPoints = fr.keyPointsUn;
}

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    FILE* fi = fopen(argv[1], "rb"); FILE* fo = fopen(argv[2], "wb");
    if (!fi || !fo) return 2;
    // ---------------------------------------------------------------- MatchByWindow (Track.cpp:113-132)
    {
        Frame mRefFrame, mFrame;
        read_frame(fi, mRefFrame); read_frame(fi, mFrame);
        std::vector<cv::Point2f> mPrevMatched(mRefFrame.N);
        for (int i = 0; i < mRefFrame.N; ++i) mPrevMatched[i] = mRefFrame.keyPointsUn[i].pt;     // Track.cpp:113-116
        std::vector<int> mMatchIdx;
        ORBmatcher matcher(0.9);
        int nMatched = matcher.MatchByWindow(mRefFrame, mFrame, mPrevMatched, 20, mMatchIdx);
        int n = (int)mMatchIdx.size();
        wr(fo, &nMatched, 1); wr(fo, &n, 1); wr(fo, mMatchIdx.data(), n); wr(fo, mPrevMatched.data(), n);
    }
    // ---------------------------------------------------------------- MatchByProjection (LocalMapper.cpp:104-118)
    {
        PtrKeyFrame mNewKF(new KeyFrame);
        read_frame(fi, *mNewKF);
        mNewKF->mObserved.resize(mNewKF->N);
        rd(fi, mNewKF->mObserved.data(), mNewKF->N);
        int M; rd(fi, &M, 1);
        std::vector<unsigned char> valid(M), desc((size_t)M * 32);
        std::vector<float> uv(2 * (size_t)M);
        std::vector<int> octave(M);
        rd(fi, valid.data(), M); rd(fi, uv.data(), uv.size()); rd(fi, octave.data(), M); rd(fi, desc.data(), desc.size());
        std::vector<PtrMapPoint> localMPs(M);
        for (int i = 0; i < M; ++i) {
            localMPs[i].reset(new MapPoint);
            localMPs[i]->mGoodPrl = valid[i] != 0;
            localMPs[i]->mPos = cv::Point3f(uv[2 * i], uv[2 * i + 1], 1.f);
            localMPs[i]->mMainOctave = octave[i];
            localMPs[i]->mMainDescriptor.create(1, 32, CV_8U);
            memcpy(localMPs[i]->mMainDescriptor.ptr<unsigned char>(0), &desc[(size_t)i * 32], 32);
        }
        std::vector<int> vMatchedIdxMPs;
        ORBmatcher matcher;
        int n = matcher.MatchByProjection(mNewKF, localMPs, 15, 2, vMatchedIdxMPs);
        int sz = (int)vMatchedIdxMPs.size();
        wr(fo, &n, 1); wr(fo, &sz, 1); wr(fo, vMatchedIdxMPs.data(), sz);
    }
    // ---------------------------------------------------------------- SearchByBoW (GlobalMapper.cpp:268-276)
    {
        PtrKeyFrame kf[2];
        for (int k = 0; k < 2; ++k) {
            kf[k].reset(new KeyFrame);
            read_frame(fi, *kf[k]);
            std::vector<unsigned char> has(kf[k]->N);
            rd(fi, has.data(), has.size());
            kf[k]->mvpMapPoints.resize(kf[k]->N);
            for (int i = 0; i < kf[k]->N; ++i) if (has[i]) kf[k]->mvpMapPoints[i].reset(new MapPoint);
            int nn; rd(fi, &nn, 1);
            std::vector<int> node(nn), ptr(nn + 1);
            rd(fi, node.data(), nn); rd(fi, ptr.data(), nn + 1);
            std::vector<int> feat(ptr[nn]);
            rd(fi, feat.data(), feat.size());
            for (int a = 0; a < nn; ++a)
                for (int q = ptr[a]; q < ptr[a + 1]; ++q) kf[k]->mFeatVec[(unsigned)node[a]].push_back((unsigned)feat[q]);
        }
        std::map<int, int> mapMatches;
        ORBmatcher matcher;
        int n = matcher.SearchByBoW(kf[0], kf[1], mapMatches, true);
        int sz = (int)mapMatches.size();
        wr(fo, &n, 1); wr(fo, &sz, 1);
        for (std::map<int, int>::iterator it = mapMatches.begin(); it != mapMatches.end(); ++it) { wr(fo, &it->first, 1); wr(fo, &it->second, 1); }
        int dd = ORBmatcher::DescriptorDistance(kf[0]->descriptors, kf[1]->descriptors);      // row 0 vs row 0
        wr(fo, &dd, 1);
    }
    fclose(fi); fclose(fo);
    return 0;
}
