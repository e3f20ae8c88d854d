// Minimal stand-in for the slice of OpenCV's core API that the cvsteer class surface touches (cv::Mat_<T>, cv::Point,
// cv::Size).  Used ONLY when <opencv2/core/core.hpp> is not installed (this build image has the Python cv2 wheel but no
// OpenCV C++ headers).  With real OpenCV present, include/cvsteer/SteerableFilters.h includes the real header and this
// file is not used.  Semantics kept: reference-counted shallow copies, row-major, `step` in bytes, (row, col) and
// cv::Point(x, y) element access, create() reallocating only on size change, converting constructor between element
// types without scaling (what cv::Mat_<float>(const cv::Mat&) does for the 8-bit images both reference callers pass).
#ifndef CVSTEER_B200_CV_SHIM_H_
#define CVSTEER_B200_CV_SHIM_H_

#include <cstddef>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>

namespace cv {

struct Point {
    int x, y;
    Point() : x(0), y(0) {}
    Point(int x_, int y_) : x(x_), y(y_) {}
};

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
    bool operator==(const Size& o) const { return width == o.width && height == o.height; }
};

class Exception : public std::runtime_error {
public:
    explicit Exception(const std::string& m) : std::runtime_error(m) {}
};

template <typename T>
class Mat_ {
public:
    int rows, cols;
    size_t step;  // bytes per row
    T* data;

    Mat_() : rows(0), cols(0), step(0), data(nullptr) {}
    Mat_(int r, int c) : Mat_() { create(r, c); }
    Mat_(Size s) : Mat_() { create(s.height, s.width); }
    // borrow external memory (like cv::Mat(rows, cols, type, data, step)); no ownership
    Mat_(int r, int c, T* ext, size_t step_bytes = 0) : rows(r), cols(c), step(step_bytes ? step_bytes : sizeof(T) * c), data(ext) {}
    // converting constructor: element-wise static_cast, no scaling (cv::Mat_<float>(const cv::Mat&) == convertTo)
    template <typename U>
    Mat_(const Mat_<U>& o) : Mat_()
    {
        create(o.rows, o.cols);
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) (*this)(r, c) = static_cast<T>(o(r, c));
    }

    void create(int r, int c)
    {
        if (r == rows && c == cols && data) return;
        rows = r, cols = c, step = sizeof(T) * (size_t)c;
        buf_.reset(new T[(size_t)r * c], std::default_delete<T[]>());
        data = buf_.get();
    }
    void create(Size s) { create(s.height, s.width); }
    bool empty() const { return !data || rows == 0 || cols == 0; }
    Size size() const { return Size(cols, rows); }
    size_t total() const { return (size_t)rows * cols; }
    T* ptr(int r = 0) { return reinterpret_cast<T*>(reinterpret_cast<char*>(data) + step * r); }
    const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(reinterpret_cast<const char*>(data) + step * r); }
    T& operator()(int r, int c) { return ptr(r)[c]; }
    const T& operator()(int r, int c) const { return ptr(r)[c]; }
    T& operator()(const Point& p) { return ptr(p.y)[p.x]; }
    const T& operator()(const Point& p) const { return ptr(p.y)[p.x]; }
    T& operator()(int i) { return data[i]; }  // 1-row vectors (tap kernels)
    const T& operator()(int i) const { return data[i]; }
    Mat_ clone() const
    {
        Mat_ m(rows, cols);
        for (int r = 0; r < rows; ++r) std::memcpy(m.ptr(r), ptr(r), sizeof(T) * cols);
        return m;
    }

private:
    std::shared_ptr<T> buf_;
};

typedef Mat_<float> Mat1f;
typedef Mat_<unsigned char> Mat1b;

}  // namespace cv
#endif
