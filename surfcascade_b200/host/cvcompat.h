// The few OpenCV value types the reference's class signatures use (cv::Size, cv::Rect, cv::Mat).
// With -DSC_HAVE_OPENCV the real <opencv2/opencv.hpp> is used instead, so the kept classes compile
// unchanged inside an OpenCV application.
#ifndef SC_CVCOMPAT_H
#define SC_CVCOMPAT_H

#ifdef SC_HAVE_OPENCV
#include <opencv2/opencv.hpp>
#else
#include <cstddef>
#include <cstring>
#include <memory>

typedef unsigned char uchar;

namespace cv {

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};

struct Rect {
    int x, y, width, height;
    Rect() : x(0), y(0), width(0), height(0) {}
    Rect(int x_, int y_, int w_, int h_) : x(x_), y(y_), width(w_), height(h_) {}
    int area() const { return width * height; }
};

enum { CV_8UC1_COMPAT = 0 };
#ifndef CV_8UC1
#define CV_8UC1 0
#endif

// 8-bit single-channel image: either a view on caller memory or an owning buffer.
class Mat {
public:
    int rows, cols;
    uchar* data;
    size_t step;  // bytes per row

    Mat() : rows(0), cols(0), data(0), step(0) {}
    Mat(int r, int c, int /*type*/, void* ext, size_t step_ = 0) : rows(r), cols(c), data((uchar*)ext), step(step_ ? step_ : (size_t)c) {}
    Mat(int r, int c, int /*type*/) : rows(r), cols(c), data(0), step((size_t)c) {
        own_.reset(new uchar[(size_t)r * c], std::default_delete<uchar[]>());
        data = own_.get();
    }
    uchar* ptr(int y = 0) { return data + (size_t)y * step; }
    const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
    Size size() const { return Size(cols, rows); }
    bool empty() const { return !data || rows <= 0 || cols <= 0; }

private:
    std::shared_ptr<uchar> own_;
};

}  // namespace cv
#endif

#endif
