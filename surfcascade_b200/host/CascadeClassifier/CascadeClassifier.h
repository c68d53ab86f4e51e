// The cascade (inference side).  Mirrors the reference's CascadeClassifier
// (CascadeClassifier/CascadeClassifier.h:18-35): public FPR / TPR / stage_classifiers, Predict, Predict2,
// GetFittedPatchIndexes, Print.  CascadeClassifier::Train (CascadeClassifier.cpp:11-56) is out of scope.
#ifndef CASCADECLASSIFIER_H
#define CASCADECLASSIFIER_H

#include <memory>
#include <vector>

#include "CascadeClassifier/GentleAdaboost.h"
#include "CascadeClassifier/StageClassifier.h"

class CascadeClassifier
{
    int max_stages_num = 10;
    float FPR_target = 1e-6f;
    float TPR_min_perstage = 0.995f;

public:
    float FPR = 1.f;
    float TPR = 1.f;
    std::vector<std::shared_ptr<StageClassifier>> stage_classifiers;

    bool Predict(std::vector<std::vector<float>>& x);
    bool Predict2(std::vector<std::vector<std::vector<float>>>& x, double& score);
    void GetFittedPatchIndexes(std::vector<std::vector<int>>& patch_indexes);
    void Print();
    friend class Model;
};

#endif
