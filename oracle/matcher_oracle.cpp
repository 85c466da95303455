// ORB matcher ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of se2lam::ORBmatcher (reference src/ORBmatcher.cpp, whole file) and of the
// keypoint grid it searches (src/Frame.cpp:64-77 grid fill, :209-286 PosInGrid / GetFeaturesInArea;
// include/se2lam/Frame.h:26-27 FRAME_GRID_ROWS=48, FRAME_GRID_COLS=64), with the reference's object
// graph (Frame / KeyFrame / MapPoint) flattened to plain arrays.  All arithmetic is in the repo
// (integer Hamming, float window tests), so this oracle is a direct restatement; it is pinned by
// the brute-force checks in tests/test_matcher_oracle.py.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may load this.
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

const int TH_HIGH = 100, TH_LOW = 75, HISTO_LENGTH = 30;  // ORBmatcher.cpp:45-47
const int GRID_ROWS = 48, GRID_COLS = 64;                  // Frame.h:26-27

struct KeyPoint { float x, y, size, angle, response; int octave, class_id; };

// ORBmatcher::DescriptorDistance, ORBmatcher.cpp:110-126 (SWAR popcount over 8 x 32 bit)
inline int descriptor_distance(const uint8_t* a, const uint8_t* b) {
    const int32_t* pa = (const int32_t*)a;
    const int32_t* pb = (const int32_t*)b;
    int dist = 0;
    for (int i = 0; i < 8; i++, pa++, pb++) {
        unsigned int v = *pa ^ *pb;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

struct Grid {
    std::vector<int> cell[GRID_COLS][GRID_ROWS];
    float minX, minY, invW, invH;
    const KeyPoint* kp;
    // Frame.cpp:64-77 + PosInGrid :209-220
    void build(const KeyPoint* kps, int n, float minX_, float minY_, float invW_, float invH_) {
        kp = kps; minX = minX_; minY = minY_; invW = invW_; invH = invH_;
        for (int i = 0; i < n; ++i) {
            int posX = (int)roundf((kps[i].x - minX) * invW);
            int posY = (int)roundf((kps[i].y - minY) * invH);
            if (posX < 0 || posX >= GRID_COLS || posY < 0 || posY >= GRID_ROWS) continue;
            cell[posX][posY].push_back(i);
        }
    }
    // Frame::GetFeaturesInArea, Frame.cpp:222-286
    void query(float x, float y, float r, int minLevel, int maxLevel, std::vector<int>& out) const {
        out.clear();
        int nMinCellX = (int)floorf((x - minX - r) * invW);
        nMinCellX = nMinCellX > 0 ? nMinCellX : 0;
        if (nMinCellX >= GRID_COLS) return;
        int nMaxCellX = (int)ceilf((x - minX + r) * invW);
        nMaxCellX = nMaxCellX < GRID_COLS - 1 ? nMaxCellX : GRID_COLS - 1;
        if (nMaxCellX < 0) return;
        int nMinCellY = (int)floorf((y - minY - r) * invH);
        nMinCellY = nMinCellY > 0 ? nMinCellY : 0;
        if (nMinCellY >= GRID_ROWS) return;
        int nMaxCellY = (int)ceilf((y - minY + r) * invH);
        nMaxCellY = nMaxCellY < GRID_ROWS - 1 ? nMaxCellY : GRID_ROWS - 1;
        if (nMaxCellY < 0) return;
        bool bCheckLevels = true, bSameLevel = false;
        if (minLevel == -1 && maxLevel == -1) bCheckLevels = false;
        else if (minLevel == maxLevel) bSameLevel = true;
        for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
            for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
                const std::vector<int>& vCell = cell[ix][iy];
                for (size_t j = 0; j < vCell.size(); j++) {
                    const KeyPoint& k = kp[vCell[j]];
                    if (bCheckLevels && !bSameLevel) { if (k.octave < minLevel || k.octave > maxLevel) continue; }
                    else if (bSameLevel) { if (k.octave != minLevel) continue; }
                    if (std::fabs(k.x - x) > r || std::fabs(k.y - y) > r) continue;
                    out.push_back(vCell[j]);
                }
            }
    }
};

// ORBmatcher::ComputeThreeMaxima, ORBmatcher.cpp:64-105
void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = (int)histo[i].size();
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

}  // namespace

extern "C" {

int matcher_oracle_distance(const uint8_t* a, const uint8_t* b) { return descriptor_distance(a, b); }

// ORBmatcher::MatchByWindow, ORBmatcher.cpp:278-381.  prev [n1 x 2] is vbPrevMatched (in/out).
int matcher_oracle_match_by_window(const void* kp1_, const uint8_t* d1, int n1, const void* kp2_, const uint8_t* d2, int n2,
                                   float* prev, float minX, float minY, float invW, float invH, float winSize,
                                   int levelOffset, int minLevel, int maxLevel, int checkOri /*unused by reference*/,
                                   float nnratio, int* matches12) {
    const KeyPoint* kp1 = (const KeyPoint*)kp1_;
    const KeyPoint* kp2 = (const KeyPoint*)kp2_;
    Grid* g = new Grid;
    g->build(kp2, n2, minX, minY, invW, invH);
    int nmatches = 0;
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = (float)HISTO_LENGTH / 360.0f;
    std::vector<int> vMatchesDistance(n2, INT_MAX), vnMatches21(n2, -1), cand;
    for (int i1 = 0; i1 < n1; i1++) {
        int level1 = kp1[i1].octave;
        if (level1 > maxLevel || level1 < minLevel) continue;
        int minLevel2 = level1 - levelOffset > 0 ? level1 - levelOffset : 0;
        g->query(prev[2 * i1], prev[2 * i1 + 1], winSize, minLevel2, level1 + levelOffset, cand);
        if (cand.empty()) continue;
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (int i2 : cand) {
            int dist = descriptor_distance(d1 + 32 * (size_t)i1, d2 + 32 * (size_t)i2);
            if (vMatchesDistance[i2] <= dist) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist <= TH_LOW) {
            if (bestDist < (float)bestDist2 * nnratio) {
                if (vnMatches21[bestIdx2] >= 0) { matches12[vnMatches21[bestIdx2]] = -1; nmatches--; }
                matches12[i1] = bestIdx2;
                vnMatches21[bestIdx2] = i1;
                vMatchesDistance[bestIdx2] = bestDist;
                nmatches++;
                float rot = kp1[i1].angle - kp2[bestIdx2].angle;
                if (rot < 0.0) rot += 360.f;
                int bin = (int)roundf(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                rotHist[bin].push_back(i1);
            }
        }
    }
    {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0; j < rotHist[i].size(); j++) {
                int idx1 = rotHist[i][j];
                if (matches12[idx1] >= 0) { matches12[idx1] = -1; nmatches--; }
            }
        }
    }
    for (int i1 = 0; i1 < n1; i1++)
        if (matches12[i1] >= 0) { prev[2 * i1] = kp2[matches12[i1]].x; prev[2 * i1 + 1] = kp2[matches12[i1]].y; }
    delete g;
    return nmatches;
}

// ORBmatcher::MatchByProjection, ORBmatcher.cpp:383-454, flattened:
//   mp_valid[i]  = !isNull && isGoodPrl && !pNewKF->hasObservation(pMP) && inImgBound(predictUV)
//   mp_uv[i]     = predictUV (caller: cvu::camprjc(Kcam, cvu::se3map(Tcw, pos)))
//   mp_octave[i] = pMP->mMainOctave ; mp_desc = pMP->mMainDescriptor
//   kf_observed[idx] = pNewKF->hasObservation(idx)
int matcher_oracle_match_by_projection(const void* kfkp_, const uint8_t* kfdesc, int n, const uint8_t* kf_observed,
                                       const uint8_t* mp_valid, const float* mp_uv, int nmp, const int* mp_octave,
                                       const uint8_t* mp_desc, int winSize, float minX, float minY, float invW, float invH,
                                       int levelOffset, float nnratio, int* matchesIdxMP) {
    const KeyPoint* kp = (const KeyPoint*)kfkp_;
    Grid* g = new Grid;
    g->build(kp, n, minX, minY, invW, invH);
    int nmatches = 0;
    for (int i = 0; i < n; ++i) matchesIdxMP[i] = -1;
    std::vector<int> vMatchesDistance(n, INT_MAX), cand;
    for (int i = 0; i < nmp; i++) {
        if (!mp_valid[i]) continue;
        const int predictLevel = mp_octave[i];
        const int levelWinSize = predictLevel * winSize;
        const int minLevel = predictLevel > levelOffset ? predictLevel - levelOffset : 0;
        g->query(mp_uv[2 * i], mp_uv[2 * i + 1], (float)levelWinSize, minLevel, predictLevel + levelOffset, cand);
        if (cand.empty()) continue;
        int bestDist = INT_MAX, bestLevel = -1, bestDist2 = INT_MAX, bestLevel2 = -1, bestIdx = -1;
        for (int idx : cand) {
            if (kf_observed[idx]) continue;
            const int dist = descriptor_distance(mp_desc + 32 * (size_t)i, kfdesc + 32 * (size_t)idx);
            if (vMatchesDistance[idx] <= dist) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = kp[idx].octave; bestIdx = idx; }
            else if (dist < bestDist2) { bestLevel2 = kp[idx].octave; bestDist2 = dist; }
        }
        if (bestDist <= TH_HIGH) {
            if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
            if (matchesIdxMP[bestIdx] >= 0) { matchesIdxMP[bestIdx] = -1; nmatches--; }
            matchesIdxMP[bestIdx] = i;
            vMatchesDistance[bestIdx] = bestDist;
            nmatches++;
        }
    }
    delete g;
    return nmatches;
}

// ORBmatcher::SearchByBoW, ORBmatcher.cpp:128-276, with each DBoW2::FeatureVector flattened to
// ascending node ids + CSR feature lists.  has_mp{1,2}[idx] = (pMP && !pMP->isNull()).
int matcher_oracle_search_by_bow(const float* angle1, const uint8_t* d1, const uint8_t* has_mp1, int n1, const int* node1,
                                 int nnode1, const int* ptr1, const int* feat1, const float* angle2, const uint8_t* d2,
                                 const uint8_t* has_mp2, int n2, const int* node2, int nnode2, const int* ptr2,
                                 const int* feat2, int mpOnly, float nnratio, int checkOri, int* matches12) {
    struct Pack { const float* angle2; const uint8_t* d2; const uint8_t* has_mp2; const int* node2; long nnode2; const int* ptr2; const int* feat2; long n2; };
    const Pack pk{angle2, d2, has_mp2, node2, nnode2, ptr2, feat2, n2};
    const Pack* P2 = &pk;
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    std::vector<uint8_t> vbMatched2(P2->n2, 0);
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = (float)HISTO_LENGTH / 360.0f;
    int nmatches = 0;
    int a = 0, b = 0;
    while (a < nnode1 && b < P2->nnode2) {
        if (node1[a] == P2->node2[b]) {
            for (int i1 = ptr1[a]; i1 < ptr1[a + 1]; i1++) {
                int idx1 = feat1[i1];
                if (mpOnly && !has_mp1[idx1]) continue;
                int bestDist1 = INT_MAX, bestIdx2 = -1, bestDist2 = INT_MAX;
                for (int i2 = P2->ptr2[b]; i2 < P2->ptr2[b + 1]; i2++) {
                    int idx2 = P2->feat2[i2];
                    if (mpOnly && !P2->has_mp2[idx2]) continue;
                    if (vbMatched2[idx2]) continue;
                    int dist = descriptor_distance(d1 + 32 * (size_t)idx1, P2->d2 + 32 * (size_t)idx2);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = idx2; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (bestDist1 < TH_LOW) {
                    if ((float)bestDist1 < nnratio * (float)bestDist2) {
                        matches12[idx1] = bestIdx2;
                        vbMatched2[bestIdx2] = 1;
                        if (checkOri) {
                            float rot = angle1[idx1] - P2->angle2[bestIdx2];
                            if (rot < 0.0) rot += 360.0f;
                            int bin = (int)roundf(rot * factor);
                            if (bin == HISTO_LENGTH) bin = 0;
                            rotHist[bin].push_back(idx1);
                        }
                        nmatches++;
                    }
                }
            }
            a++; b++;
        } else if (node1[a] < P2->node2[b]) {
            while (a < nnode1 && node1[a] < P2->node2[b]) a++;      // lower_bound
        } else {
            while (b < P2->nnode2 && P2->node2[b] < node1[a]) b++;
        }
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0; j < rotHist[i].size(); j++) { matches12[rotHist[i][j]] = -1; nmatches--; }
        }
    }
    return nmatches;
}

}  // extern "C"
