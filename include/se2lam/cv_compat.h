// Minimal stand-ins for the OpenCV types that appear in se2lam's ORBextractor / ORBmatcher signatures.
// Used ONLY when real OpenCV headers are not installed (this build container has none), so that the
// header shims in this directory can be compiled and exercised by tests/native/. With OpenCV present
// the shims include <opencv2/...> instead and this file is not used.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

namespace cv {

enum { CV_8U_ = 0 };
#ifndef CV_8U
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_32FC1 5
#endif

struct Point2f { float x = 0, y = 0; Point2f() {} Point2f(float x_, float y_) : x(x_), y(y_) {} };
struct Point3f { float x = 0, y = 0, z = 0; Point3f() {} Point3f(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {} };

struct KeyPoint {   // same 28-byte layout as cv::KeyPoint
    Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1;
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

class Mat {
public:
    int rows = 0, cols = 0;
    size_t step = 0;
    uint8_t* data = nullptr;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, void* ext, size_t step_ = 0) : rows(r), cols(c), step(step_ ? step_ : (size_t)c * esz(type)), data((uint8_t*)ext), type_(type) {}
    void create(int r, int c, int type) {
        if (r == rows && c == cols && type == type_ && own_) return;
        own_ = std::shared_ptr<uint8_t>(new uint8_t[(size_t)r * c * esz(type) + 1](), std::default_delete<uint8_t[]>());
        rows = r; cols = c; type_ = type; step = (size_t)c * esz(type); data = own_.get();
    }
    template <class T> T& at(int r, int c) { return ((T*)(data + (size_t)r * step))[c]; }
    template <class T> const T& at(int r, int c) const { return ((const T*)(data + (size_t)r * step))[c]; }
    void release() { own_.reset(); data = nullptr; rows = cols = 0; step = 0; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return type_; }
    bool isContinuous() const { return step == (size_t)cols * esz(type_); }
    template <class T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step); }
    template <class T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }
    Mat getMat() const { return *this; }
private:
    static size_t esz(int type) { return type == CV_32F ? 4 : 1; }
    std::shared_ptr<uint8_t> own_;
    int type_ = CV_8U;
};

typedef const Mat& InputArray;
typedef Mat& OutputArray;

}  // namespace cv
