// Abstract cascade stage (inference side).  Mirrors the reference's StageClassifier
// (CascadeClassifier/StageClassifier.h:10-34): public FPR / TPR / theta and the Predict / Predict2 /
// GetFittedPatchIndexes virtuals.  Train / Evaluate / SearchTheta (StageClassifier.cpp:10-70) are the
// trainer's and are out of scope here; the fields they fill are kept so that model files round-trip.
#ifndef STAGECLASSIFIER_H
#define STAGECLASSIFIER_H

#include <memory>
#include <vector>

class StageClassifier
{
    float search_step = 0.01f;
    float auc_step = 0.05f;
    float TPR_min;

protected:
    int n_total = 0;
    int n_pos = 0;
    int n_neg = 0;

public:
    float FPR = 0.f;
    float TPR = 0.f;
    float theta = 0.f;

    explicit StageClassifier(float TPR_min_perstage) : TPR_min(TPR_min_perstage) {}
    virtual ~StageClassifier() {}
    virtual float Predict(std::vector<std::vector<float>>& x) = 0;   // x indexed by pool patch_index (training layout)
    virtual float Predict2(std::vector<std::vector<float>>& x) = 0;  // x positional, one descriptor per weak classifier
    virtual void GetFittedPatchIndexes(std::vector<int>& patch_indexes) = 0;
    friend class Model;
};

#endif
