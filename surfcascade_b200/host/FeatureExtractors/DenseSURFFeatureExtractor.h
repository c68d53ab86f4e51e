// Dense SURF feature extractor (detection side).  Mirrors the reference's DenseSURFFeatureExtractor
// (FeatureExtractors/DenseSURFFeatureExtractor.h:27-61): IntegralImage, sum, ExtractPatches, CalcFeature,
// ExtractFeatures x2, ProjectPatches x2, public size / dim.  The channel + integral image lives in GPU memory
// behind the C-ABI (include/surfcascade.h); sum / CalcFeature / ExtractFeatures read it there.  The file-list /
// imread training helpers (LoadFileList, ExtractNextImageFeatures, FillNegSamples, :27-47,103-195) are out of scope.
#ifndef DENSESURFFEATUREEXTRACTOR_H
#define DENSESURFFEATUREEXTRACTOR_H

#include <string>
#include <vector>

#include "cvcompat.h"

struct sc_handle;

using cv::Mat;
using cv::Rect;
using cv::Size;

class DenseSURFFeatureExtractor
{
    static const int n_cells = 4;
    static const int n_bins = 8;
    sc_handle* handle_;
    bool owns_handle_;

public:
    Size size;
    static const int dim = n_bins * n_cells;

    DenseSURFFeatureExtractor();                       // creates a handle on CUDA device 0
    explicit DenseSURFFeatureExtractor(sc_handle* h);  // shares a caller-owned handle
    ~DenseSURFFeatureExtractor();
    DenseSURFFeatureExtractor(const DenseSURFFeatureExtractor&) = delete;
    DenseSURFFeatureExtractor& operator=(const DenseSURFFeatureExtractor&) = delete;

    sc_handle* handle() const { return handle_; }
    void IntegralImage(Mat img);
    float sum(const Rect& win);
    void ExtractPatches(std::vector<Rect>& patches);
    void CalcFeature(const Rect& patch, std::vector<float>& feature);
    void ExtractFeatures(const std::vector<Rect>& patches, std::vector<std::vector<float>>& features_win);
    void ExtractFeatures(const std::vector<std::vector<Rect>>& patches, std::vector<std::vector<std::vector<float>>>& features_win);
    void ProjectPatches(const Rect win2, const std::vector<std::vector<Rect>>& patches1, std::vector<std::vector<Rect>>& patches2);
    void ProjectPatches(const Rect win2, const std::vector<Rect>& patches1, std::vector<Rect>& patches2);
};

#endif
