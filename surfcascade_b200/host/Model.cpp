// model.cfg <-> CascadeClassifier graph.  Schema and field order follow the reference's Model::Save
// (Model.cpp:26-79); error behaviour follows Model::Load (Model.cpp:103-116,188-191): I/O and parse errors are
// logged and return EXIT_FAILURE, a missing setting silently ends the load keeping the stages completed so far.
#include "Model.h"

#include <cstdlib>
#include <memory>

#include "CascadeClassifier/CascadeClassifier.h"
#include "CascadeClassifier/GentleAdaboost.h"
#include "CascadeClassifier/LogisticRegression.h"
#include "LOG.h"
#include "cfgfile.h"

using sccfg::Node;

Model::~Model() {}

int Model::Save(CascadeClassifier& cc) {
    sccfg::File file;
    Node& top = file.root().add("cascade_classifier", Node::Group);
    top.add("max_stages_num", Node::Int).set(cc.max_stages_num);
    top.add("FPR_target", Node::Float).set((double)cc.FPR_target);
    top.add("TPR_min_perstage", Node::Float).set((double)cc.TPR_min_perstage);
    top.add("FPR", Node::Float).set((double)cc.FPR);
    top.add("TPR", Node::Float).set((double)cc.TPR);
    Node& stages = top.add("stage_classifiers", Node::List);
    for (auto& sp : cc.stage_classifiers) {
        GentleAdaboost* st = static_cast<GentleAdaboost*>(sp.get());
        Node& g = stages.add(Node::Group);
        g.add("search_step", Node::Float).set((double)st->search_step);
        g.add("auc_step", Node::Float).set((double)st->auc_step);
        g.add("TPR_min", Node::Float).set((double)st->TPR_min);
        g.add("n_total", Node::Int).set(st->n_total);
        g.add("n_pos", Node::Int).set(st->n_pos);
        g.add("n_neg", Node::Int).set(st->n_neg);
        g.add("FPR", Node::Float).set((double)st->FPR);
        g.add("TPR", Node::Float).set((double)st->TPR);
        g.add("theta", Node::Float).set((double)st->theta);
        g.add("total_AUC_score", Node::Float).set((double)st->total_AUC_score);
        g.add("sample_num", Node::Int).set(st->sample_num);
        g.add("max_iters", Node::Int).set(st->max_iters);
        Node& weaks = g.add("weak_classifiers", Node::List);
        for (auto& wk : st->weak_classifiers) {
            Node& q = weaks.add(Node::Group);
            q.add("patch_index", Node::Int).set(wk->patch_index);
            q.add("eps", Node::Float).set(wk->eps_);
            q.add("C", Node::Float).set(wk->C_);
            q.add("nr_class", Node::Int).set(wk->nr_class_);
            q.add("nr_feature", Node::Int).set(wk->nr_feature_);
            q.add("bias", Node::Float).set(wk->bias_);
            Node& w = q.add("w", Node::Array);
            for (int k = 0; k < wk->nr_feature_ + 1 && k < 33; k++) w.add(Node::Float).set((double)wk->w[k]);
            Node& lab = q.add("label", Node::Array);
            lab.add(Node::Int).set(wk->label_[0]);
            lab.add(Node::Int).set(wk->label_[1]);
        }
    }
    try {
        file.write(model_cfg);
        LOG_INFO("New model successfully written to: " << model_cfg);
    } catch (const sccfg::IoError&) {
        LOG_ERROR("I/O error while writing file: " << model_cfg);
        return EXIT_FAILURE;
    }
    return EXIT_SUCCESS;
}

int Model::Load(CascadeClassifier& cc) {
    sccfg::File file;
    try {
        file.read(model_cfg);
    } catch (const sccfg::IoError&) {
        LOG_ERROR("I/O error while reading file.");
        return EXIT_FAILURE;
    } catch (const sccfg::ParseError& pe) {
        LOG_ERROR("Parse error at " << model_cfg << ":" << pe.line << " - " << pe.what());
        return EXIT_FAILURE;
    }
    try {
        const Node& top = file.root()["cascade_classifier"];
        cc.max_stages_num = top["max_stages_num"].asInt();
        cc.FPR_target = top["FPR_target"].asFloat();
        cc.TPR_min_perstage = top["TPR_min_perstage"].asFloat();
        cc.FPR = top["FPR"].asFloat();
        cc.TPR = top["TPR"].asFloat();
        const Node& stages = top["stage_classifiers"];
        for (int i = 0; i < stages.length(); i++) {
            const Node& g = stages[i];
            std::shared_ptr<GentleAdaboost> st(new GentleAdaboost(cc.TPR_min_perstage));
            st->search_step = g["search_step"].asFloat();
            st->auc_step = g["auc_step"].asFloat();
            st->TPR_min = g["TPR_min"].asFloat();
            st->n_total = g["n_total"].asInt();
            st->n_pos = g["n_pos"].asInt();
            st->n_neg = g["n_neg"].asInt();
            st->FPR = g["FPR"].asFloat();
            st->TPR = g["TPR"].asFloat();
            st->theta = g["theta"].asFloat();
            st->total_AUC_score = g["total_AUC_score"].asFloat();
            st->sample_num = g["sample_num"].asInt();
            st->max_iters = g["max_iters"].asInt();
            const Node& weaks = g["weak_classifiers"];
            for (int j = 0; j < weaks.length(); j++) {
                const Node& q = weaks[j];
                std::shared_ptr<LogisticRegression> wk(new LogisticRegression(0));
                wk->patch_index = q["patch_index"].asInt();
                wk->eps_ = q["eps"].asDouble();
                wk->C_ = q["C"].asDouble();
                wk->nr_class_ = q["nr_class"].asInt();
                wk->nr_feature_ = q["nr_feature"].asInt();
                wk->bias_ = q["bias"].asDouble();
                const Node& w = q["w"];
                // the detector indexes w[0..32]; a shorter array in the file leaves the tail zero instead of reading
                // past a heap block as the reference would
                for (int k = 0; k < w.length() && k < 33; k++) wk->w[k] = w[k].asFloat();
                const Node& lab = q["label"];
                wk->label_[0] = lab[0].asInt();
                wk->label_[1] = lab[1].asInt();
                st->weak_classifiers.push_back(wk);
            }
            cc.stage_classifiers.push_back(st);
        }
    } catch (const sccfg::NotFound&) {
        // ignored, like the reference (Model.cpp:188-191)
    }
    return EXIT_SUCCESS;
}
