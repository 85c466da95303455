// CPU ORACLE — TEST INFRASTRUCTURE ONLY (never linked, imported or executed by the product path).
//
// cv::undistort(im, img, Config::Kcam, Config::Dcam) as the reference calls it right before the ORB extractor
// (reference src/Frame.cpp:22). The arithmetic lives in OpenCV [upstream, imgproc/undistort]:
//   * the frame is processed in stripes of max(1, 4096 / cols) rows; for the stripe starting at row y the map is
//     initUndistortRectifyMap(A, dist, I, Ar) with Ar = A except Ar(1,2) = cy - y;
//   * ir = Ar^-1 by LU with partial pivoting (double); per row i of the stripe the ray starts at
//     (i*ir1 + ir2, i*ir4 + ir5, i*ir7 + ir8) and is advanced by (ir0, ir3, ir6) per column (accumulated, not recomputed);
//   * x = X/W, y = Y/W, r2 = x^2 + y^2, kr = (1 + ((k3 r2 + k2) r2 + k1) r2) / (1 + ((k6 r2 + k5) r2 + k4) r2),
//     xd = x kr + p1 2xy + p2 (r2 + 2x^2) + s1 r2 + s2 r2^2, yd likewise, u = fx xd + cx, v = fy yd + cy (A, not Ar);
//   * CV_16SC2 map: iu = cvRound(u*32), iv = cvRound(v*32); integer part iu >> 5, fraction index (iv & 31)*32 + (iu & 31);
//   * remap INTER_LINEAR, BORDER_CONSTANT(0): weights (32-a)(32-b)*32 etc. (exact, sum 32768), (sum + 16384) >> 15.
// PINNED: oracle/pin_undistort_against_cv2.py checks this file bit for bit against cv2.undistort (OpenCV 4.13) and writes
// tests/golden/undistort_golden.npz.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

// cv::Mat::inv(DECOMP_LU) for a 3x3 double matrix: hal::LU64f with the identity as right-hand side
bool lu_inv3(const double Ain[9], double out[9]) {
    double A[9], b[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    std::memcpy(A, Ain, sizeof A);
    const int m = 3;
    for (int i = 0; i < m; ++i) {
        int k = i;
        for (int j = i + 1; j < m; ++j) if (std::fabs(A[j * 3 + i]) > std::fabs(A[k * 3 + i])) k = j;
        if (std::fabs(A[k * 3 + i]) < 2.220446049250313e-16 * 100) return false;
        if (k != i) for (int j = 0; j < 3; ++j) { std::swap(A[i * 3 + j], A[k * 3 + j]); std::swap(b[i * 3 + j], b[k * 3 + j]); }
        const double d = -1 / A[i * 3 + i];
        for (int j = i + 1; j < m; ++j) {
            const double alpha = A[j * 3 + i] * d;
            for (int kk = i + 1; kk < m; ++kk) A[j * 3 + kk] += alpha * A[i * 3 + kk];
            for (int kk = 0; kk < 3; ++kk) b[j * 3 + kk] += alpha * b[i * 3 + kk];
        }
    }
    for (int i = m - 1; i >= 0; --i)
        for (int j = 0; j < 3; ++j) {
            double s = b[i * 3 + j];
            for (int kk = i + 1; kk < m; ++kk) s -= A[i * 3 + kk] * b[kk * 3 + j];
            b[i * 3 + j] = s / A[i * 3 + i];
        }
    std::memcpy(out, b, sizeof b);
    return true;
}

}  // namespace

extern "C" {

// map for a w x h frame: m1 [h*w*2] (integer source coordinates x, y), m2 [h*w] (5+5 bit fractions). K row-major 3x3
// float32, dist nd float32 coefficients (k1 k2 p1 p2 [k3 [k4 k5 k6 [s1 s2 s3 s4]]]). Returns 0, or -1 if K is singular.
int undistort_oracle_map(const float* K, const float* dist, int nd, int w, int h, int16_t* m1, uint16_t* m2) {
    double A[9], D[14] = {0};
    for (int i = 0; i < 9; ++i) A[i] = (double)K[i];
    for (int i = 0; i < nd && i < 14; ++i) D[i] = (double)dist[i];
    const double k1 = D[0], k2 = D[1], p1 = D[2], p2 = D[3], k3 = D[4], k4 = D[5], k5 = D[6], k6 = D[7], s1 = D[8], s2 = D[9], s3 = D[10], s4 = D[11];
    const double fx = A[0], fy = A[4], u0 = A[2], v0 = A[5];
    const int stripe0 = std::min(std::max(1, (1 << 12) / std::max(w, 1)), h);
    for (int y0 = 0; y0 < h; y0 += stripe0) {
        const int rows = std::min(stripe0, h - y0);
        double Ar[9], ir[9];
        std::memcpy(Ar, A, sizeof Ar);
        Ar[5] = v0 - y0;
        if (!lu_inv3(Ar, ir)) return -1;
        for (int i = 0; i < rows; ++i) {
            double _x = i * ir[1] + ir[2], _y = i * ir[4] + ir[5], _w = i * ir[7] + ir[8];
            int16_t* o1 = m1 + (size_t)(y0 + i) * w * 2;
            uint16_t* o2 = m2 + (size_t)(y0 + i) * w;
            for (int j = 0; j < w; ++j, _x += ir[0], _y += ir[3], _w += ir[6]) {
                const double ww = 1. / _w, x = _x * ww, y = _y * ww;
                const double x2 = x * x, y2 = y * y, r2 = x2 + y2, _2xy = 2 * x * y;
                const double kr = (1 + ((k3 * r2 + k2) * r2 + k1) * r2) / (1 + ((k6 * r2 + k5) * r2 + k4) * r2);
                const double xd = (x * kr + p1 * _2xy + p2 * (r2 + 2 * x2) + s1 * r2 + s2 * r2 * r2);
                const double yd = (y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy + s3 * r2 + s4 * r2 * r2);
                const double u = fx * xd + u0, v = fy * yd + v0;
                const double su = u * 32, sv = v * 32;
                const int iu = su >= 2147483647.0 ? 2147483647 : su <= -2147483648.0 ? (int)-2147483648LL : (int)std::lrint(su);
                const int iv = sv >= 2147483647.0 ? 2147483647 : sv <= -2147483648.0 ? (int)-2147483648LL : (int)std::lrint(sv);
                o1[2 * j] = (int16_t)(iu >> 5); o1[2 * j + 1] = (int16_t)(iv >> 5);
                o2[j] = (uint16_t)((iv & 31) * 32 + (iu & 31));
            }
        }
    }
    return 0;
}

// cv::remap(src, dst, m1, m2, INTER_LINEAR, BORDER_CONSTANT, 0) for 8-bit single-channel images
void undistort_oracle_remap(const uint8_t* src, int w, int h, int stride, const int16_t* m1, const uint16_t* m2, uint8_t* dst) {
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const size_t idx = (size_t)y * w + x;
            const int sx = m1[2 * idx], sy = m1[2 * idx + 1];
            const int a = m2[idx] & 31, b = m2[idx] >> 5;
            const int w00 = (32 - a) * (32 - b) * 32, w01 = a * (32 - b) * 32, w10 = (32 - a) * b * 32, w11 = a * b * 32;
            auto tap = [&](int yy, int xx) -> int { return (xx >= 0 && xx < w && yy >= 0 && yy < h) ? src[(size_t)yy * stride + xx] : 0; };
            const int val = tap(sy, sx) * w00 + tap(sy, sx + 1) * w01 + tap(sy + 1, sx) * w10 + tap(sy + 1, sx + 1) * w11;
            const int r = (val + (1 << 14)) >> 15;
            dst[idx] = (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
        }
}

int undistort_oracle(const uint8_t* src, int w, int h, int stride, const float* K, const float* dist, int nd, uint8_t* dst) {
    std::vector<int16_t> m1((size_t)w * h * 2);
    std::vector<uint16_t> m2((size_t)w * h);
    if (undistort_oracle_map(K, dist, nd, w, h, m1.data(), m2.data()) != 0) return -1;
    undistort_oracle_remap(src, w, h, stride, m1.data(), m2.data(), dst);
    return 0;
}

}  // extern "C"
