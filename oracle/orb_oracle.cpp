// ORB front-end ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of se2lam::ORBextractor::operator() (reference src/ORBextractor.cpp:727-788) and of
// the OpenCV primitives it calls (resize INTER_LINEAR, copyMakeBorder REFLECT_101, FAST-9-16 with
// non-max suppression, KeyPointsFilter::retainBest, fastAtan2, GaussianBlur 7x7 sigma 2, cvRound).
// OpenCV is an un-vendored dependency of the reference (README.MD:27); its arithmetic is restated
// here from its published algorithm and pinned bit-exactly against the cv2 4.13.0 wheel by
// oracle/pin_orb_against_cv2.py (fixtures under tests/golden/).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this. The product path (se2lam_b200/csrc) never links or calls it.
//
// Pins (documented in DESIGN.md):
//  * blur       = float32 separable 7-tap (getGaussianKernel(7,2,CV_32F)); row pass sequential
//                 left-to-right with fused multiply-add, column pass symmetric-paired with fused
//                 multiply-add, round-half-even to u8.  This is bit-for-bit what cv2.sepFilter2D
//                 computes on an AVX2/FMA3 host (OpenCV's scalar RowFilter/SymmColumnFilter loops
//                 are compiled with contraction in the AVX2 dispatch unit) and is the path
//                 GaussianBlur takes for the reference's non-isolated sub-matrix call when no
//                 IPP/HAL intercepts it.
//  * retainBest = std::nth_element(begin, begin+n-1, end, response-greater) + std::partition of
//                 ties + truncation, with THIS toolchain's libstdc++ (GCC 13).
//  * steering   = a=(float)cos((double)angle_rad), b=(float)sin((double)angle_rad).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

const int PATCH_SIZE = 31;       // ORBextractor.cpp:81
const int HALF_PATCH_SIZE = 15;  // :82
const int EDGE_THRESHOLD = 16;   // :83

static const int bit_pattern_31[256 * 4] = {
#include "../se2lam_b200/csrc/orb_pattern_31.inc"
};

inline int cv_round_f(float v) { return (int)lrintf(v); }   // cvRound: round-half-even
inline int cv_round_d(double v) { return (int)lrint(v); }

struct KeyPoint {  // cv::KeyPoint, 28 bytes
    float x, y, size, angle, response;
    int octave, class_id;
};

struct Plane {  // one pyramid level with its 16 px border; roi() points at the level origin
    int w = 0, h = 0, pitch = 0;
    std::vector<uint8_t> buf;
    void alloc(int w_, int h_) {
        w = w_; h = h_; pitch = w + 2 * EDGE_THRESHOLD;
        buf.assign((size_t)pitch * (h + 2 * EDGE_THRESHOLD), 0);
    }
    uint8_t* roi() { return buf.data() + (size_t)EDGE_THRESHOLD * pitch + EDGE_THRESHOLD; }
    const uint8_t* roi() const { return buf.data() + (size_t)EDGE_THRESHOLD * pitch + EDGE_THRESHOLD; }
};

inline int reflect101(int p, int len) {  // cv::borderInterpolate(BORDER_REFLECT_101)
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * (len - 1) - p;
    }
    return p;
}

// copyMakeBorder(..., 16,16,16,16, BORDER_REFLECT_101) of the ROI into its own parent plane
// (ORBextractor.cpp:815-816, 823-824).
void fill_border(Plane& p) {
    uint8_t* r = p.roi();
    const int B = EDGE_THRESHOLD;
    for (int y = -B; y < p.h + B; ++y) {
        int sy = reflect101(y, p.h);
        for (int x = -B; x < p.w + B; ++x) {
            if (y >= 0 && y < p.h && x >= 0 && x < p.w) continue;
            int sx = reflect101(x, p.w);
            r[(ptrdiff_t)y * p.pitch + x] = r[(ptrdiff_t)sy * p.pitch + sx];
        }
    }
}

// cv::resize(src, dst, dsize, 0, 0, INTER_LINEAR) for CV_8UC1: 11-bit fixed-point bilinear
// [upstream OpenCV imgproc/resize.cpp: resizeGeneric_ + HResizeLinear + VResizeLinear].
void resize_linear_u8(const uint8_t* src, int sw, int sh, int spitch, uint8_t* dst, int dw, int dh,
                      int dpitch) {
    const int COEF_BITS = 11;
    const float COEF_SCALE = (float)(1 << COEF_BITS);
    double inv_scale_x = (double)dw / sw, inv_scale_y = (double)dh / sh;
    double scale_x = 1. / inv_scale_x, scale_y = 1. / inv_scale_y;
    std::vector<int> xofs(dw), yofs(dh);
    std::vector<short> ialpha(dw * 2), ibeta(dh * 2);
    for (int dx = 0; dx < dw; ++dx) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = (int)floorf(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        float c0 = 1.f - fx, c1 = fx;
        int a0 = cv_round_f(c0 * COEF_SCALE), a1 = cv_round_f(c1 * COEF_SCALE);
        ialpha[dx * 2] = (short)std::min(std::max(a0, -32768), 32767);
        ialpha[dx * 2 + 1] = (short)std::min(std::max(a1, -32768), 32767);
    }
    for (int dy = 0; dy < dh; ++dy) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = (int)floorf(fy);
        fy -= sy;
        yofs[dy] = sy;
        float c0 = 1.f - fy, c1 = fy;
        ibeta[dy * 2] = (short)cv_round_f(c0 * COEF_SCALE);
        ibeta[dy * 2 + 1] = (short)cv_round_f(c1 * COEF_SCALE);
    }
    std::vector<int> row0(dw), row1(dw);
    for (int dy = 0; dy < dh; ++dy) {
        int sy0 = std::min(std::max(yofs[dy], 0), sh - 1);
        int sy1 = std::min(std::max(yofs[dy] + 1, 0), sh - 1);
        const uint8_t* S0 = src + (size_t)sy0 * spitch;
        const uint8_t* S1 = src + (size_t)sy1 * spitch;
        for (int dx = 0; dx < dw; ++dx) {
            int sx = xofs[dx];
            int sx1 = std::min(sx + 1, sw - 1);  // alpha1 == 0 whenever sx+1 is out of range
            row0[dx] = S0[sx] * ialpha[dx * 2] + S0[sx1] * ialpha[dx * 2 + 1];
            row1[dx] = S1[sx] * ialpha[dx * 2] + S1[sx1] * ialpha[dx * 2 + 1];
        }
        int b0 = ibeta[dy * 2], b1 = ibeta[dy * 2 + 1];
        uint8_t* D = dst + (size_t)dy * dpitch;
        for (int dx = 0; dx < dw; ++dx) {
            int v = (((b0 * (row0[dx] >> 4)) >> 16) + ((b1 * (row1[dx] >> 4)) >> 16) + 2) >> 2;
            D[dx] = (uint8_t)std::min(std::max(v, 0), 255);
        }
    }
}

// cv::FAST(img, kps, threshold, nonmaxSuppression=true), TYPE_9_16, on a cell ROI
// [upstream OpenCV features2d/fast.cpp FAST_t<16> + fast_score.cpp cornerScore<16>].
const int ring_dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
const int ring_dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

inline int corner_score16(const int* d /*25 entries: v - ring[k], wrapped*/, int threshold) {
    int a0 = threshold;
    for (int k = 0; k < 16; k += 2) {
        int a = std::min(d[k + 1], d[k + 2]);
        a = std::min(a, d[k + 3]);
        if (a <= a0) continue;
        a = std::min(a, d[k + 4]); a = std::min(a, d[k + 5]); a = std::min(a, d[k + 6]);
        a = std::min(a, d[k + 7]); a = std::min(a, d[k + 8]);
        a0 = std::max(a0, std::min(a, d[k]));
        a0 = std::max(a0, std::min(a, d[k + 9]));
    }
    int b0 = -a0;
    for (int k = 0; k < 16; k += 2) {
        int b = std::max(d[k + 1], d[k + 2]);
        b = std::max(b, d[k + 3]); b = std::max(b, d[k + 4]); b = std::max(b, d[k + 5]);
        if (b >= b0) continue;
        b = std::max(b, d[k + 6]); b = std::max(b, d[k + 7]); b = std::max(b, d[k + 8]);
        b0 = std::min(b0, std::max(b, d[k]));
        b0 = std::min(b0, std::max(b, d[k + 9]));
    }
    return -b0 - 1;
}

void fast9_16_nms(const uint8_t* img, int cols, int rows, int pitch, int threshold,
                  std::vector<KeyPoint>& out) {
    out.clear();
    if (cols < 7 || rows < 7) return;
    std::vector<uint8_t> score((size_t)cols * rows, 0);
    for (int i = 3; i < rows - 3; ++i) {
        const uint8_t* ptr = img + (size_t)i * pitch;
        for (int j = 3; j < cols - 3; ++j) {
            int v = ptr[j];
            int d[25];
            for (int k = 0; k < 16; ++k) d[k] = v - ptr[j + ring_dy[k] * pitch + ring_dx[k]];
            for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
            // corner test: 9 contiguous ring pixels all darker than v-t (d>t) or all brighter (d<-t)
            bool corner = false;
            int run_dark = 0, run_bright = 0;
            for (int k = 0; k < 25; ++k) {
                run_dark = (d[k] > threshold) ? run_dark + 1 : 0;
                run_bright = (d[k] < -threshold) ? run_bright + 1 : 0;
                if (run_dark > 8 || run_bright > 8) { corner = true; break; }
            }
            if (corner) score[(size_t)i * cols + j] = (uint8_t)corner_score16(d, threshold);
        }
    }
    for (int i = 3; i < rows - 3; ++i)
        for (int j = 3; j < cols - 3; ++j) {
            int s = score[(size_t)i * cols + j];
            if (!s) continue;
            const uint8_t* p = &score[(size_t)(i - 1) * cols + j];
            const uint8_t* c = &score[(size_t)i * cols + j];
            const uint8_t* n = &score[(size_t)(i + 1) * cols + j];
            if (s > p[-1] && s > p[0] && s > p[1] && s > c[-1] && s > c[1] && s > n[-1] && s > n[0] &&
                s > n[1]) {
                KeyPoint kp{(float)j, (float)i, 7.f, -1.f, (float)s, 0, -1};
                out.push_back(kp);
            }
        }
}

// cv::KeyPointsFilter::retainBest [upstream OpenCV features2d/keypoint.cpp]
struct ResponseGreater {
    bool operator()(const KeyPoint& a, const KeyPoint& b) const { return a.response > b.response; }
};
void retain_best(std::vector<KeyPoint>& kps, int n) {
    if (n >= 0 && kps.size() > (size_t)n) {
        if (n == 0) { kps.clear(); return; }
        std::nth_element(kps.begin(), kps.begin() + n - 1, kps.end(), ResponseGreater());
        float ambiguous = kps[n - 1].response;
        auto new_end = std::partition(kps.begin() + n, kps.end(),
                                      [ambiguous](const KeyPoint& k) { return k.response >= ambiguous; });
        kps.resize(new_end - kps.begin());
    }
}

// cv::fastAtan2 (degrees) [upstream OpenCV core/mathfuncs_core: atan_f32 polynomial]
float fast_atan2_deg(float y, float x) {
    const float p1 = 0.9997878412794807f * (float)(180 / M_PI);
    const float p3 = -0.3258083974640975f * (float)(180 / M_PI);
    const float p5 = 0.1555786518463281f * (float)(180 / M_PI);
    const float p7 = -0.04432655554792128f * (float)(180 / M_PI);
    float ax = std::fabs(x), ay = std::fabs(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

struct Extractor {
    int nfeatures, nlevels, fastTh;
    double scaleFactor;
    std::vector<float> mvScaleFactor, mvInvScaleFactor;
    std::vector<int> mnFeaturesPerLevel, umax;
    std::vector<Plane> pyr;       // un-blurred (mvImagePyramid after ComputePyramid)
    std::vector<Plane> blurred;   // after the in-place GaussianBlur of :769
    float gk[7];

    // ORBextractor.cpp:463-520
    Extractor(int nf, float sf, int nl, int ft) : nfeatures(nf), nlevels(nl), fastTh(ft), scaleFactor(sf) {
        mvScaleFactor.resize(nlevels);
        mvScaleFactor[0] = 1;
        for (int i = 1; i < nlevels; i++) mvScaleFactor[i] = (float)(mvScaleFactor[i - 1] * scaleFactor);
        float invScaleFactor = (float)(1.0f / scaleFactor);
        mvInvScaleFactor.resize(nlevels);
        mvInvScaleFactor[0] = 1;
        for (int i = 1; i < nlevels; i++) mvInvScaleFactor[i] = mvInvScaleFactor[i - 1] * invScaleFactor;
        mnFeaturesPerLevel.resize(nlevels);
        float factor = (float)(1.0 / scaleFactor);
        float nDesired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
        int sum = 0;
        for (int level = 0; level < nlevels - 1; level++) {
            mnFeaturesPerLevel[level] = cv_round_f(nDesired);
            sum += mnFeaturesPerLevel[level];
            nDesired *= factor;
        }
        mnFeaturesPerLevel[nlevels - 1] = std::max(nfeatures - sum, 0);
        umax.resize(HALF_PATCH_SIZE + 1);
        int v, v0, vmax = (int)floorf(HALF_PATCH_SIZE * sqrtf(2.f) / 2 + 1);
        int vmin = (int)ceilf(HALF_PATCH_SIZE * sqrtf(2.f) / 2);
        const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
        for (v = 0; v <= vmax; ++v) umax[v] = cv_round_d(sqrt(hp2 - v * v));
        for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
            while (umax[v0] == umax[v0 + 1]) ++v0;
            umax[v] = v0;
            ++v0;
        }
        // cv::getGaussianKernel(7, 2, CV_32F): exp in double, normalise in double, store float
        double g[7], s = 0;
        for (int i = 0; i < 7; ++i) { double x = i - 3; g[i] = std::exp(-0.5 * x * x / 4.0); s += g[i]; }
        // OpenCV computes t = exp(scale2X*x*x) with scale2X = -0.5/(sigma*sigma), sums, then kernel[i] = (float)(t[i]*(1/sum))
        double inv = 1. / s;
        for (int i = 0; i < 7; ++i) gk[i] = (float)(g[i] * inv);
    }

    // ORBextractor.cpp:790-831
    void computePyramid(const uint8_t* img, int w, int h, int stride) {
        pyr.assign(nlevels, Plane());
        for (int level = 0; level < nlevels; ++level) {
            float scale = mvInvScaleFactor[level];
            int sw = cv_round_f((float)w * scale), sh = cv_round_f((float)h * scale);
            pyr[level].alloc(sw, sh);
            if (level != 0) {
                resize_linear_u8(pyr[level - 1].roi(), pyr[level - 1].w, pyr[level - 1].h, pyr[level - 1].pitch,
                                 pyr[level].roi(), sw, sh, pyr[level].pitch);
            } else {
                for (int y = 0; y < h; ++y) memcpy(pyr[0].roi() + (size_t)y * pyr[0].pitch, img + (size_t)y * stride, w);
            }
            fill_border(pyr[level]);
        }
    }

    // ORBextractor.cpp:130-157
    float icAngle(const Plane& p, float ptx, float pty) const {
        int m_01 = 0, m_10 = 0;
        const int step = p.pitch;
        const uint8_t* center = p.roi() + (ptrdiff_t)cv_round_f(pty) * step + cv_round_f(ptx);
        for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
        for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
            int v_sum = 0, d = umax[v];
            for (int u = -d; u <= d; ++u) {
                int val_plus = center[u + v * step], val_minus = center[u - v * step];
                v_sum += (val_plus - val_minus);
                m_10 += u * (val_plus + val_minus);
            }
            m_01 += v * v_sum;
        }
        return fast_atan2_deg((float)m_01, (float)m_10);
    }

    // ORBextractor.cpp:531-716
    void computeKeyPoints(std::vector<std::vector<KeyPoint>>& all, std::vector<int>* dbg_cell_counts) {
        all.assign(nlevels, {});
        float imageRatio = (float)pyr[0].w / pyr[0].h;
        for (int level = 0; level < nlevels; ++level) {
            const int nDesiredFeatures = mnFeaturesPerLevel[level];
            const int levelCols = (int)sqrtf((float)nDesiredFeatures / (5 * imageRatio));
            const int levelRows = (int)(imageRatio * levelCols);
            const int minBorderX = EDGE_THRESHOLD, minBorderY = minBorderX;
            const int maxBorderX = pyr[level].w - EDGE_THRESHOLD;
            const int maxBorderY = pyr[level].h - EDGE_THRESHOLD;
            const int W = maxBorderX - minBorderX, H = maxBorderY - minBorderY;
            const int cellW = (int)ceilf((float)W / levelCols);
            const int cellH = (int)ceilf((float)H / levelRows);
            const int nCells = levelRows * levelCols;
            const int nfeaturesCell = (int)ceilf((float)nDesiredFeatures / nCells);
            std::vector<std::vector<std::vector<KeyPoint>>> cellKeyPoints(levelRows, std::vector<std::vector<KeyPoint>>(levelCols));
            std::vector<std::vector<int>> nToRetain(levelRows, std::vector<int>(levelCols, 0));
            std::vector<std::vector<int>> nTotal(levelRows, std::vector<int>(levelCols, 0));
            std::vector<std::vector<bool>> bNoMore(levelRows, std::vector<bool>(levelCols, false));
            std::vector<int> iniXCol(levelCols), iniYRow(levelRows);
            int nNoMore = 0, nToDistribute = 0;
            float hY = cellH + 6;
            const Plane& P = pyr[level];
            for (int i = 0; i < levelRows; i++) {
                const float iniY = minBorderY + i * cellH - 3;
                iniYRow[i] = iniY;
                if (i == levelRows - 1) {
                    hY = maxBorderY + 3 - iniY;
                    if (hY <= 0) continue;
                }
                float hX = cellW + 6;
                for (int j = 0; j < levelCols; j++) {
                    float iniX;
                    if (i == 0) { iniX = minBorderX + j * cellW - 3; iniXCol[j] = iniX; }
                    else iniX = iniXCol[j];
                    if (j == levelCols - 1) {
                        hX = maxBorderX + 3 - iniX;
                        if (hX <= 0) continue;
                    }
                    int r0 = (int)iniY, r1 = (int)(iniY + hY), c0 = (int)iniX, c1 = (int)(iniX + hX);
                    const uint8_t* cell = P.roi() + (ptrdiff_t)r0 * P.pitch + c0;
                    std::vector<KeyPoint>& kc = cellKeyPoints[i][j];
                    fast9_16_nms(cell, c1 - c0, r1 - r0, P.pitch, fastTh, kc);
                    if (kc.size() <= 3) {
                        kc.clear();
                        fast9_16_nms(cell, c1 - c0, r1 - r0, P.pitch, 7, kc);
                    }
                    const int nKeys = (int)kc.size();
                    nTotal[i][j] = nKeys;
                    if (dbg_cell_counts) dbg_cell_counts->push_back(nKeys);
                    if (nKeys > nfeaturesCell) { nToRetain[i][j] = nfeaturesCell; bNoMore[i][j] = false; }
                    else { nToRetain[i][j] = nKeys; nToDistribute += nfeaturesCell - nKeys; bNoMore[i][j] = true; nNoMore++; }
                }
            }
            while (nToDistribute > 0 && nNoMore < nCells) {
                int nNewFeaturesCell = (int)(nfeaturesCell + ceilf((float)nToDistribute / (nCells - nNoMore)));
                nToDistribute = 0;
                for (int i = 0; i < levelRows; i++)
                    for (int j = 0; j < levelCols; j++)
                        if (!bNoMore[i][j]) {
                            if (nTotal[i][j] > nNewFeaturesCell) { nToRetain[i][j] = nNewFeaturesCell; bNoMore[i][j] = false; }
                            else { nToRetain[i][j] = nTotal[i][j]; nToDistribute += nNewFeaturesCell - nTotal[i][j]; bNoMore[i][j] = true; nNoMore++; }
                        }
            }
            std::vector<KeyPoint>& keypoints = all[level];
            const int scaledPatchSize = (int)(PATCH_SIZE * mvScaleFactor[level]);
            for (int i = 0; i < levelRows; i++)
                for (int j = 0; j < levelCols; j++) {
                    std::vector<KeyPoint>& keysCell = cellKeyPoints[i][j];
                    retain_best(keysCell, nToRetain[i][j]);
                    if ((int)keysCell.size() > nToRetain[i][j]) keysCell.resize(nToRetain[i][j]);
                    for (size_t k = 0; k < keysCell.size(); k++) {
                        keysCell[k].x += iniXCol[j];
                        keysCell[k].y += iniYRow[i];
                        keysCell[k].octave = level;
                        keysCell[k].size = (float)scaledPatchSize;
                        keypoints.push_back(keysCell[k]);
                    }
                }
            if ((int)keypoints.size() > nDesiredFeatures) {
                retain_best(keypoints, nDesiredFeatures);
                keypoints.resize(nDesiredFeatures);
            }
        }
        for (int level = 0; level < nlevels; ++level)
            for (auto& kp : all[level]) kp.angle = icAngle(pyr[level], kp.x, kp.y);
    }

    // GaussianBlur(level, level, Size(7,7), 2, 2, BORDER_REFLECT_101) on the (non-isolated) ROI,
    // ORBextractor.cpp:767-769. Neighbours outside the ROI are the real border-ring pixels.
    void blurLevel(int level) {
        const Plane& s = pyr[level];
        Plane& d = blurred[level];
        d = s;
        std::vector<float> tmp((size_t)(s.h + 6) * s.w);
        const uint8_t* r = s.roi();
        for (int y = -3; y < s.h + 3; ++y) {
            const uint8_t* row = r + (ptrdiff_t)y * s.pitch;
            float* t = &tmp[(size_t)(y + 3) * s.w];
            for (int x = 0; x < s.w; ++x) {
                const uint8_t* S = row + x - 3;
                float f = gk[0] * (float)S[0];
                for (int k = 1; k < 7; ++k) f = fmaf(gk[k], (float)S[k], f);
                t[x] = f;
            }
        }
        uint8_t* o = d.roi();
        for (int y = 0; y < s.h; ++y)
            for (int x = 0; x < s.w; ++x) {
                const float* c = &tmp[(size_t)(y + 3) * s.w + x];
                float f = gk[3] * c[0];
                for (int k = 1; k <= 3; ++k) f = fmaf(gk[3 + k], c[(ptrdiff_t)k * s.w] + c[-(ptrdiff_t)k * s.w], f);
                int iv = cv_round_f(f);
                o[(ptrdiff_t)y * d.pitch + x] = (uint8_t)std::min(std::max(iv, 0), 255);
            }
    }

    // ORBextractor.cpp:160-200
    void descriptor(const Plane& img, const KeyPoint& kpt, uint8_t* desc) const {
        const float factorPI = (float)(M_PI / 180.f);
        float angle = (float)kpt.angle * factorPI;
        float a = (float)cos((double)angle), b = (float)sin((double)angle);
        const int step = img.pitch;
        const uint8_t* center = img.roi() + (ptrdiff_t)cv_round_f(kpt.y) * step + cv_round_f(kpt.x);
        const int* pat = bit_pattern_31;
        for (int i = 0; i < 32; ++i, pat += 32) {
            int val = 0;
            for (int k = 0; k < 8; ++k) {
                int x0 = pat[4 * k], y0 = pat[4 * k + 1], x1 = pat[4 * k + 2], y1 = pat[4 * k + 3];
                volatile float r0a = x0 * b, r0b = y0 * a, c0a = x0 * a, c0b = y0 * b;  // no contraction
                volatile float r1a = x1 * b, r1b = y1 * a, c1a = x1 * a, c1b = y1 * b;
                int t0 = center[cv_round_f(r0a + r0b) * step + cv_round_f(c0a - c0b)];
                int t1 = center[cv_round_f(r1a + r1b) * step + cv_round_f(c1a - c1b)];
                val |= (t0 < t1) << k;
            }
            desc[i] = (uint8_t)val;
        }
    }

    // ORBextractor.cpp:727-788. Returns number of keypoints; kps/desc sized for nfeatures.
    int extract(const uint8_t* img, int w, int h, int stride, KeyPoint* kps_out, uint8_t* desc_out,
                std::vector<int>* dbg_cell_counts = nullptr) {
        if (!img || w <= 0 || h <= 0) return 0;
        computePyramid(img, w, h, stride);
        std::vector<std::vector<KeyPoint>> all;
        computeKeyPoints(all, dbg_cell_counts);
        blurred.assign(nlevels, Plane());
        int offset = 0;
        for (int level = 0; level < nlevels; ++level) {
            std::vector<KeyPoint>& kps = all[level];
            if (kps.empty()) continue;
            blurLevel(level);
            for (size_t i = 0; i < kps.size(); ++i) descriptor(blurred[level], kps[i], desc_out + (size_t)(offset + i) * 32);
            if (level != 0) {
                float scale = mvScaleFactor[level];
                for (auto& kp : kps) { kp.x *= scale; kp.y *= scale; }
            }
            memcpy(kps_out + offset, kps.data(), kps.size() * sizeof(KeyPoint));
            offset += (int)kps.size();
        }
        return offset;
    }
};

}  // namespace

extern "C" {

void* orb_oracle_create(int nfeatures, float scaleFactor, int nlevels, int fastTh) {
    return new Extractor(nfeatures, scaleFactor, nlevels, fastTh);
}
void orb_oracle_destroy(void* h) { delete (Extractor*)h; }

// kps: n x 28 bytes (cv::KeyPoint layout), desc: n x 32 bytes; both sized for >= nfeatures entries.
int orb_oracle_extract(void* h, const uint8_t* img, int w, int h_, int stride, void* kps, uint8_t* desc) {
    return ((Extractor*)h)->extract(img, w, h_, stride, (KeyPoint*)kps, desc);
}

int orb_oracle_level_dims(void* h, int level, int* w, int* hh, int* pitch) {
    Extractor* e = (Extractor*)h;
    if (level < 0 || level >= (int)e->pyr.size()) return -1;
    *w = e->pyr[level].w; *hh = e->pyr[level].h; *pitch = e->pyr[level].pitch;
    return 0;
}
// copies the bordered plane ((h+32) x pitch) of the last extract; blurred=1 -> after GaussianBlur
int orb_oracle_get_level(void* h, int level, int blurred, uint8_t* out) {
    Extractor* e = (Extractor*)h;
    const std::vector<Plane>& v = blurred ? e->blurred : e->pyr;
    if (level < 0 || level >= (int)v.size() || v[level].buf.empty()) return -1;
    memcpy(out, v[level].buf.data(), v[level].buf.size());
    return 0;
}
void orb_oracle_tables(void* h, int* features_per_level, float* scale, float* inv_scale, int* umax16, float* gk7) {
    Extractor* e = (Extractor*)h;
    for (int i = 0; i < e->nlevels; ++i) {
        features_per_level[i] = e->mnFeaturesPerLevel[i];
        scale[i] = e->mvScaleFactor[i];
        inv_scale[i] = e->mvInvScaleFactor[i];
    }
    for (int i = 0; i < 16; ++i) umax16[i] = e->umax[i];
    for (int i = 0; i < 7; ++i) gk7[i] = e->gk[i];
}

// primitive entry points so each restated OpenCV primitive can be pinned against cv2 on its own
void orb_oracle_resize(const uint8_t* src, int sw, int sh, int spitch, uint8_t* dst, int dw, int dh, int dpitch) {
    resize_linear_u8(src, sw, sh, spitch, dst, dw, dh, dpitch);
}
// returns count; out: (x, y, response) triples as float
int orb_oracle_fast(const uint8_t* img, int cols, int rows, int pitch, int threshold, float* out, int max_out) {
    std::vector<KeyPoint> k;
    fast9_16_nms(img, cols, rows, pitch, threshold, k);
    int n = std::min((int)k.size(), max_out);
    for (int i = 0; i < n; ++i) { out[3 * i] = k[i].x; out[3 * i + 1] = k[i].y; out[3 * i + 2] = k[i].response; }
    return (int)k.size();
}
float orb_oracle_fast_atan2(float y, float x) { return fast_atan2_deg(y, x); }
// retainBest on (response, id) pairs: returns the surviving ids in their final order
int orb_oracle_retain_best(const float* responses, int n, int n_points, int* ids_out) {
    std::vector<KeyPoint> k(n);
    for (int i = 0; i < n; ++i) { k[i] = KeyPoint{0, 0, 0, 0, responses[i], 0, i}; }
    retain_best(k, n_points);
    for (size_t i = 0; i < k.size(); ++i) ids_out[i] = k[i].class_id;
    return (int)k.size();
}
// std::nth_element permutation only (for pinning the device introselect port)
void orb_oracle_nth_element(const float* responses, int n, int nth, int* ids_out) {
    std::vector<KeyPoint> k(n);
    for (int i = 0; i < n; ++i) { k[i] = KeyPoint{0, 0, 0, 0, responses[i], 0, i}; }
    std::nth_element(k.begin(), k.begin() + nth, k.end(), ResponseGreater());
    for (int i = 0; i < n; ++i) ids_out[i] = k[i].class_id;
}

}  // extern "C"
