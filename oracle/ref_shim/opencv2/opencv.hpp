// TEST INFRASTRUCTURE ONLY (oracle build) -- never linked into the product library.
//
// Minimal stand-in for the slice of OpenCV 3.0.0 that the reference's detection/training path
// touches, so that the UNMODIFIED reference sources under /root/reference compile on a box that
// has no OpenCV C++ SDK.  Nothing here is copied from
// OpenCV or from the reference: every function is a restatement of published OpenCV behaviour,
// written for this repo, and cross-checked against the cv2 4.13 wheel in tests/test_ref_shim.py.
//
// Reference call sites served (paths relative to /root/reference/ObjDetector):
//   cv::Mat header ctor / ptr / size   FeatureExtractors/DenseSURFFeatureExtractor.cpp:74,222  ObjDetector.cpp:174
//   cv::integral(u8 -> CV_32FC1)       FeatureExtractors/DenseSURFFeatureExtractor.cpp:75
//   cv::merge                          FeatureExtractors/DenseSURFFeatureExtractor.cpp:80
//   cv::imread(.., IMREAD_GRAYSCALE)   FeatureExtractors/DenseSURFFeatureExtractor.cpp:44,109,134
//   cv::groupRectangles(5-arg)         ObjDetector.cpp:225
#ifndef SC_ORACLE_OPENCV_STANDIN_HPP
#define SC_ORACLE_OPENCV_STANDIN_HPP

#include <pmmintrin.h>
#include <algorithm>
#include <cassert>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

typedef unsigned char uchar;

#define CV_8UC1 0
#define CV_32FC1 5
#define CV_8UC3 16
#define CV_MAKE_F32(cn) (5 + (((cn) - 1) << 3))

namespace cv {

template <typename T>
struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;

template <typename T>
struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w_, T h_) : x(x_), y(y_), width(w_), height(h_) {}
    T area() const { return width * height; }
};
typedef Rect_<int> Rect;

template <typename T, int N>
struct Vec {
    T val[N];
};

enum ImreadModes { IMREAD_GRAYSCALE = 0 };

// Dense row-major matrix.  Owning buffers are 64-byte aligned because the reference reads
// __m128 members straight out of the merged integral (DenseSURFFeatureExtractor.cpp:355-356).
class Mat {
public:
    int rows, cols;
    uchar* data;

    Mat() : rows(0), cols(0), data(0), type_(0) {}
    Mat(int r, int c, int type, void* ext) : rows(r), cols(c), data((uchar*)ext), type_(type) {}
    Mat(int r, int c, int type) : rows(0), cols(0), data(0), type_(0) { create(r, c, type); }

    void create(int r, int c, int type) {
        rows = r; cols = c; type_ = type;
        size_t bytes = (size_t)r * c * elemSize();
        void* p = 0;
        if (posix_memalign(&p, 64, bytes ? bytes : 64) != 0) p = 0;
        own_.reset((uchar*)p, free);
        data = own_.get();
    }
    int type() const { return type_; }
    int channels() const { return (type_ >> 3) + 1; }
    size_t elemSize() const { return (size_t)channels() * ((type_ & 7) == 5 ? 4 : 1); }
    size_t rowBytes() const { return (size_t)cols * elemSize(); }
    uchar* ptr(int y = 0) { return data + (size_t)y * rowBytes(); }
    const uchar* ptr(int y = 0) const { return data + (size_t)y * rowBytes(); }
    Size size() const { return Size(cols, rows); }
    bool empty() const { return data == 0 || rows == 0 || cols == 0; }

private:
    int type_;
    std::shared_ptr<uchar> own_;
};

// Binary PGM (P5, maxval <= 255) reader: the only image format the oracle harness writes.
inline Mat imread(const std::string& path, int /*flags*/) {
    Mat m;
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return m;
    char magic[3] = {0, 0, 0};
    int w = 0, h = 0, maxv = 0;
    if (fscanf(f, "%2s", magic) == 1 && strcmp(magic, "P5") == 0) {
        int vals[3], got = 0;
        while (got < 3) {
            int c = fgetc(f);
            if (c == EOF) break;
            if (c == '#') { while (c != '\n' && c != EOF) c = fgetc(f); continue; }
            if (c == ' ' || c == '\t' || c == '\n' || c == '\r') continue;
            ungetc(c, f);
            if (fscanf(f, "%d", &vals[got]) != 1) break;
            got++;
        }
        if (got == 3) {
            w = vals[0]; h = vals[1]; maxv = vals[2];
            fgetc(f);  // the single whitespace byte before the raster
            if (w > 0 && h > 0 && maxv > 0 && maxv <= 255) {
                m.create(h, w, CV_8UC1);
                if (fread(m.data, 1, (size_t)w * h, f) != (size_t)w * h) m = Mat();
            }
        }
    }
    fclose(f);
    return m;
}

// OpenCV's portable integral_<uchar,float>: a float running sum along the row (exact: it stays
// below 2^24 for any row shorter than 65793 pixels) added to the float value directly above,
// row after row.  Output is (rows+1) x (cols+1) with a zero first row and column.
inline void integral(const Mat& src, Mat& sum, int /*sdepth*/) {
    const int h = src.rows, w = src.cols;
    sum.create(h + 1, w + 1, CV_32FC1);
    float* s = (float*)sum.data;
    const size_t pitch = (size_t)w + 1;
    for (int x = 0; x <= w; x++) s[x] = 0.f;
    for (int y = 0; y < h; y++) {
        const uchar* row = src.ptr(y);
        const float* up = s + (size_t)y * pitch;
        float* cur = s + (size_t)(y + 1) * pitch;
        float run = 0.f;
        cur[0] = 0.f;
        for (int x = 0; x < w; x++) {
            run += (float)row[x];
            cur[x + 1] = up[x + 1] + run;
        }
    }
}

// Interleave n equally sized single-channel float planes into one n-channel matrix.
inline void merge(const std::vector<Mat>& planes, Mat& dst) {
    const int n = (int)planes.size();
    const int h = planes[0].rows, w = planes[0].cols;
    dst.create(h, w, CV_MAKE_F32(n));
    float* d = (float*)dst.data;
    for (int c = 0; c < n; c++) {
        const float* p = (const float*)planes[c].data;
        for (size_t i = 0, e = (size_t)h * w; i < e; i++) d[i * n + c] = p[i];
    }
}

namespace standin_detail {
inline int round_half_even(double v) { return (int)lrint(v); }

inline bool similar_rects(const Rect& a, const Rect& b, double eps) {
    double delta = eps * (std::min(a.width, b.width) + std::min(a.height, b.height)) * 0.5;
    return std::abs(a.x - b.x) <= delta && std::abs(a.y - b.y) <= delta &&
           std::abs(a.x + a.width - b.x - b.width) <= delta &&
           std::abs(a.y + a.height - b.y - b.height) <= delta;
}

// Connected components of the "similar" relation; class ids in order of first member.
inline int components(const std::vector<Rect>& r, std::vector<int>& label, double eps) {
    const int n = (int)r.size();
    std::vector<int> parent(n);
    for (int i = 0; i < n; i++) parent[i] = i;
    struct F {
        static int find(std::vector<int>& p, int i) {
            while (p[i] != i) { p[i] = p[p[i]]; i = p[i]; }
            return i;
        }
    };
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++)
            if (similar_rects(r[i], r[j], eps)) {
                int a = F::find(parent, i), b = F::find(parent, j);
                if (a != b) parent[std::max(a, b)] = std::min(a, b);
            }
    label.assign(n, -1);
    std::vector<int> id(n, -1);
    int k = 0;
    for (int i = 0; i < n; i++) {
        int root = F::find(parent, i);
        if (id[root] < 0) id[root] = k++;
        label[i] = id[root];
    }
    return k;
}
}  // namespace standin_detail

// The (rects, rejectLevels, levelWeights, groupThreshold, eps) overload, SURVEY.md Appendix A.6.
inline void groupRectangles(std::vector<Rect>& rects, std::vector<int>& levels,
                            std::vector<double>& levelWeights, int groupThreshold, double eps = 0.2) {
    using namespace standin_detail;
    if (groupThreshold <= 0 || rects.empty()) return;
    std::vector<int> label;
    const int k = components(rects, label, eps);
    std::vector<Rect> acc(k);
    std::vector<int> members(k, 0), bestLevel(k, 0);
    std::vector<double> bestWeight(k, DBL_MIN);
    const int n = (int)rects.size();
    for (int i = 0; i < n; i++) {
        Rect& a = acc[label[i]];
        a.x += rects[i].x; a.y += rects[i].y; a.width += rects[i].width; a.height += rects[i].height;
        members[label[i]]++;
    }
    const bool haveWeights = !levels.empty() && !levelWeights.empty();
    if (haveWeights)
        for (int i = 0; i < n; i++) {
            int c = label[i];
            if (levels[i] > bestLevel[c]) { bestLevel[c] = levels[i]; bestWeight[c] = levelWeights[i]; }
            else if (levels[i] == bestLevel[c] && levelWeights[i] > bestWeight[c]) bestWeight[c] = levelWeights[i];
        }
    for (int c = 0; c < k; c++) {
        float inv = 1.f / members[c];
        acc[c] = Rect(round_half_even(acc[c].x * inv), round_half_even(acc[c].y * inv),
                      round_half_even(acc[c].width * inv), round_half_even(acc[c].height * inv));
    }
    rects.clear(); levels.clear(); levelWeights.clear();
    for (int i = 0; i < k; i++) {
        if (members[i] <= groupThreshold) continue;
        const Rect& a = acc[i];
        int j = 0;
        for (; j < k; j++) {
            if (j == i || members[j] <= groupThreshold) continue;
            const Rect& b = acc[j];
            int dx = round_half_even(b.width * eps), dy = round_half_even(b.height * eps);
            if (a.x >= b.x - dx && a.y >= b.y - dy && a.x + a.width <= b.x + b.width + dx &&
                a.y + a.height <= b.y + b.height + dy && (members[j] > std::max(3, members[i]) || members[i] < 3))
                break;
        }
        if (j == k) {
            rects.push_back(a);
            levels.push_back(haveWeights ? bestLevel[i] : members[i]);
            levelWeights.push_back(bestWeight[i]);
        }
    }
}

}  // namespace cv

#endif
