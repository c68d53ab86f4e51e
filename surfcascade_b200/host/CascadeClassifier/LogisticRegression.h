// Weak classifier of the SURF cascade: a 32-dim logistic regression (inference side).
// Mirrors the reference's LogisticRegression (CascadeClassifier/LogisticRegression.h:13-27): same public
// members and Predict signature.  Training (LogisticRegression::Train -> liblinear, LogisticRegression.cpp:35-44)
// is out of scope of this library; weights arrive through Model::Load or set_weights().
// Predict runs on the GPU through the C-ABI hook sc_weak_predict (no CPU arithmetic path).
#ifndef LOGISTICREGRESSION_H
#define LOGISTICREGRESSION_H

#include <vector>

class Model;
namespace sc_host { struct Access; }

class LogisticRegression
{
    // what liblinear's `parameter` / `model` carried in the reference, as far as model.cfg stores it (Model.cpp:61-76)
    double eps_ = 0.01;
    double C_ = 0.1;
    int nr_class_ = 2;
    int nr_feature_ = 32;
    double bias_ = 1.0;
    int label_[2] = {1, -1};
    float w[33];

public:
    int patch_index;

    explicit LogisticRegression(int patch_index);
    void set_weights(const float* w33, double bias);
    float Predict(std::vector<float>& x);

    friend class Model;
    friend class GentleAdaboost;
    friend struct sc_host::Access;
};

#endif
