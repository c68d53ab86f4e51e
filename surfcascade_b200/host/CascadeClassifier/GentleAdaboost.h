// Boosted stage: mean of its weak classifiers' probabilities (inference side).  Mirrors the reference's
// GentleAdaboost (CascadeClassifier/GentleAdaboost.h:11-25); GentleAdaboost::Train (GentleAdaboost.cpp:11-231)
// is out of scope.
#ifndef GENTLEADABOOST_H
#define GENTLEADABOOST_H

#include <memory>
#include <vector>

#include "CascadeClassifier/StageClassifier.h"

class LogisticRegression;
namespace sc_host { struct Access; }

class GentleAdaboost : public StageClassifier
{
    float total_AUC_score = 0;
    int sample_num = 960;
    int max_iters = 100;
    std::vector<std::shared_ptr<LogisticRegression>> weak_classifiers;

    float mean_probability(const std::vector<const float*>& descriptors);

public:
    explicit GentleAdaboost(float TPR_min_perstage) : StageClassifier(TPR_min_perstage) {}
    float Predict(std::vector<std::vector<float>>& x);
    float Predict2(std::vector<std::vector<float>>& x);
    void GetFittedPatchIndexes(std::vector<int>& patch_indexes);
    void add_weak_classifier(const std::shared_ptr<LogisticRegression>& weak) { weak_classifiers.push_back(weak); }
    friend class Model;
    friend struct sc_host::Access;
};

#endif
