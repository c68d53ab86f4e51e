// Test program for the kept C++ class surface (INTEGRATION.md "Level 2"): the reference's detect branch written against
// the classes exactly as a reference maintainer has it -- Model::Load, DenseSURFFeatureExtractor::ExtractPatches,
// CascadeClassifier::GetFittedPatchIndexes, IntegralImage, sum, ProjectPatches, CalcFeature, StageClassifier::Predict2 and
// ->theta (reference: ObjDetector.cpp:107-143 set-up, :174-219 scan) -- with every arithmetic step served by the GPU library
// through those methods.  Prints "x y l score" of every raw detection (17 significant digits) and a counter line; pytest
// (tests/test_gpu_cli_and_classes.py) compares the output with the oracle.  One blocking GPU round trip per call: this is a
// parity surface, not the fast path (that is sc_detect).
//
//     class_detect model.cfg image.pgm base [x y l]...
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "CascadeClassifier/CascadeClassifier.h"
#include "FeatureExtractors/DenseSURFFeatureExtractor.h"
#include "Model.h"

static bool read_pgm(const char* path, Mat* out) {
    std::ifstream f(path, std::ios::binary);
    std::string magic;
    int w = 0, h = 0, mx = 0;
    f >> magic >> w >> h >> mx;
    if (!f.good() || magic != "P5" || mx != 255 || w < 1 || h < 1) return false;
    f.get();
    Mat m(h, w, CV_8UC1);
    f.read((char*)m.data, (std::streamsize)w * h);
    if (!f.good()) return false;
    *out = m;
    return true;
}

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: class_detect model.cfg image.pgm base\n"); return 2; }
    const int base = atoi(argv[3]);
    Model model(argv[1]);
    CascadeClassifier cascade;
    if (model.Load(cascade) != EXIT_SUCCESS) { fprintf(stderr, "class_detect: Model::Load failed\n"); return 1; }
    Mat img;
    if (!read_pgm(argv[2], &img)) { fprintf(stderr, "class_detect: cannot read %s\n", argv[2]); return 1; }

    DenseSURFFeatureExtractor extractor;
    extractor.size = Size(40, 40);
    std::vector<Rect> dense;
    extractor.ExtractPatches(dense);
    std::vector<std::vector<int>> indexes;
    cascade.GetFittedPatchIndexes(indexes);
    std::vector<std::vector<Rect>> fitted(indexes.size()), projected;
    std::vector<std::vector<std::vector<float>>> features(indexes.size());
    for (size_t p = 0; p < indexes.size(); p++) {
        features[p].resize(indexes[p].size());
        for (int idx : indexes[p]) fitted[p].push_back(dense[idx]);
    }

    extractor.IntegralImage(img);
    const int step = base > 20 ? base / 20 : 1;
    const int n_sizes = (int)std::min(std::log(img.cols / (float)base) / std::log(1.1), std::log(img.rows / (float)base) / std::log(1.1));
    long long visited = 0, passed_prefilter = 0, weak_evals = 0, raw = 0;
    const size_t n_stages = cascade.stage_classifiers.size();
    for (int i = 0; i <= n_sizes; i++) {
        const int l = (int)(base * std::pow(1.1, i));
        Rect win(0, 0, l, l);
        for (int y = 0; y <= img.rows - l; y += step) {
            win.y = y;
            int stride = 1;
            for (win.x = 0; win.x <= img.cols - l; win.x += stride * step) {
                visited++;
                if (!(extractor.sum(win) > win.area() * 6)) { stride = 2; continue; }
                passed_prefilter++;
                extractor.ProjectPatches(win, fitted, projected);
                double score = 0.0;
                size_t p = 0;
                for (; p < n_stages; p++) {
                    for (size_t q = 0; q < projected[p].size(); q++) extractor.CalcFeature(projected[p][q], features[p][q]);
                    weak_evals += (long long)projected[p].size();
                    score = cascade.stage_classifiers[p]->Predict2(features[p]);
                    if (score < cascade.stage_classifiers[p]->theta) break;
                }
                score = (score + (double)p + 1) / (double)n_stages;
                if (p == n_stages) { printf("%d %d %d %.17g\n", win.x, win.y, l, score); raw++; }
                stride = score < 0.5 ? 2 : 1;
            }
        }
    }
    printf("counters %lld %lld %lld %lld\n", visited, passed_prefilter, weak_evals, raw);

    // CascadeClassifier::Predict on the training-side layout (x indexed by pool patch, CascadeClassifier.cpp:58-67 as
    // FillNegSamples calls it, DenseSURFFeatureExtractor.cpp:160-176): all 608 projected pool descriptors of a few windows
    // through ExtractFeatures, then the cascade's verdict and every stage's StageClassifier::Predict.
    for (int k = 4; k < argc; k += 3) {
        if (k + 2 >= argc) break;
        const Rect w(atoi(argv[k]), atoi(argv[k + 1]), atoi(argv[k + 2]), atoi(argv[k + 2]));
        std::vector<Rect> pool_projected;
        std::vector<std::vector<float>> x;
        extractor.ProjectPatches(w, dense, pool_projected);
        extractor.ExtractFeatures(pool_projected, x);
        printf("predict %d %d %d %d", w.x, w.y, w.width, cascade.Predict(x) ? 1 : 0);
        for (auto& st : cascade.stage_classifiers) printf(" %.9g", st->Predict(x));
        printf("\n");
    }
    return 0;
}
