// LogisticRegression / GentleAdaboost / CascadeClassifier: inference-side mirror of the reference classes.
// All arithmetic runs on the GPU through the C-ABI; these methods only marshal.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <stdexcept>
#include <string>

#include "CascadeClassifier/CascadeClassifier.h"
#include "CascadeClassifier/GentleAdaboost.h"
#include "CascadeClassifier/LogisticRegression.h"
#include "sc_access.h"

namespace sc_host {
sc_handle* default_handle() {
    static sc_handle* h = nullptr;
    if (!h) {
        const int rc = sc_create(0, &h);
        if (rc != SC_OK || !h) {
            h = nullptr;
            throw std::runtime_error("surfcascade_b200: no usable CUDA device (sc_create -> " + std::to_string(rc) + "); this library has no CPU path");
        }
    }
    return h;
}
}  // namespace sc_host

// ---- LogisticRegression (reference: CascadeClassifier/LogisticRegression.cpp:15-33,46-68) --------------------
LogisticRegression::LogisticRegression(int patch_index) : patch_index(patch_index) { memset(w, 0, sizeof(w)); }

void LogisticRegression::set_weights(const float* w33, double bias) {
    memcpy(w, w33, sizeof(w));
    bias_ = bias;
}

float LogisticRegression::Predict(std::vector<float>& x) {
    float p = 0.f;
    // The reference reads x[0..31] unchecked and cannot fail; a GPU error or a short descriptor is reported as an exception
    // (never abort(): the caller's process is not ours to end).
    if (x.size() < 32) throw std::invalid_argument("LogisticRegression::Predict: descriptor must hold 32 floats");
    if (sc_weak_predict(sc_host::default_handle(), w, &bias_, x.data(), 1, &p) != SC_OK)
        throw std::runtime_error(std::string("LogisticRegression::Predict: ") + sc_last_error(sc_host::default_handle()));
    return p;
}

// ---- GentleAdaboost (reference: CascadeClassifier/GentleAdaboost.cpp:233-267) ------------------------------
float GentleAdaboost::mean_probability(const std::vector<const float*>& descriptors) {
    const int n = (int)weak_classifiers.size();
    std::vector<float> w((size_t)n * 33), x((size_t)n * 32), p(n);
    std::vector<double> b(n);
    for (int i = 0; i < n; i++) {
        memcpy(&w[(size_t)i * 33], weak_classifiers[i]->w, 33 * sizeof(float));
        memcpy(&x[(size_t)i * 32], descriptors[i], 32 * sizeof(float));
        b[i] = weak_classifiers[i]->bias_;
    }
    float mean = 0.f;
    if (sc_stage_predict(sc_host::default_handle(), w.data(), b.data(), x.data(), n, &mean) != SC_OK)
        throw std::runtime_error(std::string("GentleAdaboost: ") + sc_last_error(sc_host::default_handle()));
    return mean;
}

float GentleAdaboost::Predict(std::vector<std::vector<float>>& x) {
    std::vector<const float*> d;
    for (auto& wk : weak_classifiers) d.push_back(x[wk->patch_index].data());
    return mean_probability(d);
}

float GentleAdaboost::Predict2(std::vector<std::vector<float>>& x) {
    if (x.size() < weak_classifiers.size()) throw std::invalid_argument("GentleAdaboost::Predict2: one descriptor per weak classifier expected");
    std::vector<const float*> d;
    for (size_t i = 0; i < weak_classifiers.size(); i++) d.push_back(x[i].data());
    return mean_probability(d);
}

void GentleAdaboost::GetFittedPatchIndexes(std::vector<int>& patch_indexes) {
    for (auto& wk : weak_classifiers) patch_indexes.push_back(wk->patch_index);
}

// ---- CascadeClassifier (reference: CascadeClassifier/CascadeClassifier.cpp:58-96) ---------------------------
bool CascadeClassifier::Predict(std::vector<std::vector<float>>& x) {
    for (auto& st : stage_classifiers)
        if (st->Predict(x) < st->theta) return false;
    return true;
}

bool CascadeClassifier::Predict2(std::vector<std::vector<std::vector<float>>>& x, double& score) {
    size_t passed = 0;
    while (passed < stage_classifiers.size()) {
        score = stage_classifiers[passed]->Predict2(x[passed]);
        if (score < stage_classifiers[passed]->theta) break;
        passed++;
    }
    score = (score + (double)passed + 1) / (double)stage_classifiers.size();
    return passed == stage_classifiers.size();
}

void CascadeClassifier::GetFittedPatchIndexes(std::vector<std::vector<int>>& patch_indexes) {
    for (auto& st : stage_classifiers) {
        patch_indexes.emplace_back();
        st->GetFittedPatchIndexes(patch_indexes.back());
    }
}

void CascadeClassifier::Print() { std::cout << "FPR: " << FPR << ", TPR:" << TPR << std::endl; }
