// DenseSURFFeatureExtractor over the C-ABI: the integral image stays in GPU memory.
#include "FeatureExtractors/DenseSURFFeatureExtractor.h"

#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "sc_host.h"

namespace {
// The reference's methods cannot fail (they index host memory unchecked); a GPU error is reported as an exception, never abort().
[[noreturn]] void die(sc_handle* h, const char* where) {
    throw std::runtime_error(std::string("DenseSURFFeatureExtractor::") + where + ": " + sc_last_error(h));
}
sc_rect to_sc(const Rect& r) { return sc_rect{r.x, r.y, r.width, r.height}; }
}  // namespace

DenseSURFFeatureExtractor::DenseSURFFeatureExtractor() : handle_(nullptr), owns_handle_(true) {
    if (sc_create(0, &handle_) != SC_OK) {
        handle_ = nullptr;
        throw std::runtime_error("DenseSURFFeatureExtractor: no usable CUDA device; this library has no CPU path");
    }
}

DenseSURFFeatureExtractor::DenseSURFFeatureExtractor(sc_handle* h) : handle_(h), owns_handle_(false) {}

DenseSURFFeatureExtractor::~DenseSURFFeatureExtractor() {
    if (owns_handle_) sc_destroy(handle_);
}

// reference: DenseSURFFeatureExtractor.cpp:65-87
void DenseSURFFeatureExtractor::IntegralImage(Mat img) {
    if (sc_integral(handle_, img.data, img.cols, img.rows, (int)img.step, nullptr) != SC_OK) die(handle_, "IntegralImage");
}

// reference: :351-358
float DenseSURFFeatureExtractor::sum(const Rect& win) {
    const sc_rect r = to_sc(win);
    float s = 0.f;
    if (sc_window_sum(handle_, &r, 1, &s) != SC_OK) die(handle_, "sum");
    return s;
}

// reference: :49-63
void DenseSURFFeatureExtractor::ExtractPatches(std::vector<Rect>& patches) {
    std::vector<sc_rect> pool;
    sc_host::pool_patches(size.width, size.height, &pool);
    for (const sc_rect& p : pool) patches.push_back(Rect(p.x, p.y, p.w, p.h));
}

// reference: :379-415
void DenseSURFFeatureExtractor::CalcFeature(const Rect& patch, std::vector<float>& feature) {
    const sc_rect r = to_sc(patch);
    feature.resize(dim);
    if (sc_features(handle_, &r, 1, feature.data()) != SC_OK) die(handle_, "CalcFeature");
}

// reference: :89-101 -- one GPU call per list instead of one per patch
void DenseSURFFeatureExtractor::ExtractFeatures(const std::vector<Rect>& patches, std::vector<std::vector<float>>& features_win) {
    std::vector<sc_rect> r;
    for (const Rect& p : patches) r.push_back(to_sc(p));
    std::vector<float> flat(r.size() * dim);
    if (sc_features(handle_, r.data(), (int)r.size(), flat.data()) != SC_OK) die(handle_, "ExtractFeatures");
    features_win.resize(r.size());
    for (size_t i = 0; i < r.size(); i++) features_win[i].assign(flat.begin() + i * dim, flat.begin() + (i + 1) * dim);
}

void DenseSURFFeatureExtractor::ExtractFeatures(const std::vector<std::vector<Rect>>& patches,
                                                std::vector<std::vector<std::vector<float>>>& features_win) {
    features_win.resize(patches.size());
    for (size_t i = 0; i < patches.size(); i++) ExtractFeatures(patches[i], features_win[i]);
}

// reference: :459-508
void DenseSURFFeatureExtractor::ProjectPatches(const Rect win2, const std::vector<Rect>& patches1, std::vector<Rect>& patches2) {
    patches2.resize(patches1.size());
    for (size_t i = 0; i < patches1.size(); i++) {
        const sc_rect p = sc_host::project_patch(size.width, win2.width, to_sc(patches1[i]));
        patches2[i] = Rect(p.x + win2.x, p.y + win2.y, p.w, p.h);
    }
}

void DenseSURFFeatureExtractor::ProjectPatches(const Rect win2, const std::vector<std::vector<Rect>>& patches1,
                                               std::vector<std::vector<Rect>>& patches2) {
    patches2.resize(patches1.size());
    for (size_t i = 0; i < patches1.size(); i++) ProjectPatches(win2, patches1[i], patches2[i]);
}
