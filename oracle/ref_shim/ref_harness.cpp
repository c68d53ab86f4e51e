// TEST INFRASTRUCTURE ONLY (oracle build) -- never linked into the product library.
//
// C-ABI harness around the UNMODIFIED reference classes, compiled from the sources where they
// lie under /root/reference (see oracle/Makefile).  The reference's detect loop and train branch
// live inline in a Win32-only main(); the Makefile lifts those line ranges, verbatim, into
// generated include files under oracle/_ref/gen/ (never committed):
//   detect_setup.inc  = ObjDetector.cpp:107-144   (pool, model load, fitted patches, step)
//   detect_loop.inc   = ObjDetector.cpp:174-220   (scale loop, prefilter, stages, multi rule) with the
//                       literal base window 70 replaced by `base` and REF_* counter hooks appended
//   train_body.inc    = ObjDetector.cpp:66-91     (positives, CascadeClassifier::Train, Model::Save)
// Everything else below is new code: marshalling, timing and counters.
#include <opencv2/opencv.hpp>

#include <omp.h>
#include <chrono>
#include <cstdint>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "CascadeClassifier/CascadeClassifier.h"
#include "CascadeClassifier/GentleAdaboost.h"
#include "CascadeClassifier/LogisticRegression.h"
#include "FeatureExtractors/DenseSURFFeatureExtractor.h"
#include "LOG.h"
#include "Model.h"
#include "linear.h"

using std::ifstream;
using std::ios;
using std::ofstream;
using std::string;
using std::vector;

#ifndef min
#define min(a, b) (((a) < (b)) ? (a) : (b))  // the reference gets this macro from <windows.h>
#endif

namespace {

struct CoutSilencer {
    std::streambuf* old;
    std::ostringstream sink;
    explicit CoutSilencer(bool on) : old(0) { if (on) old = std::cout.rdbuf(sink.rdbuf()); }
    ~CoutSilencer() { if (old) std::cout.rdbuf(old); }
};

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

Mat wrap_u8(const uint8_t* img, int W, int H) { return Mat(H, W, CV_8UC1, (void*)img); }

void release_integral(DenseSURFFeatureExtractor& ex) {
    delete[] ex.sumtab;  // the reference leaks this per frame in detect (DenseSURFFeatureExtractor.cpp:82)
    ex.sumtab = 0;
}

}  // namespace

extern "C" {

// Counter slots written by ref_detect_frames (per frame, SC_REF_NCOUNTERS int64 each).
enum { REF_C_VISITED = 0, REF_C_PREFILTER = 1, REF_C_WEAK = 2, REF_C_RAW = 3, REF_C_REACH0 = 4, REF_NCOUNTERS = 4 + 16 };

int ref_num_counters() { return REF_NCOUNTERS; }

// ExtractPatches on a tw x th template -> [n][4] = x,y,w,h.  Returns the pool size.
int ref_pool_patches(int tw, int th, int* out, int cap) {
    DenseSURFFeatureExtractor ex;
    ex.size = Size(tw, th);
    vector<Rect> p;
    ex.ExtractPatches(p);
    for (int i = 0; i < (int)p.size() && i < cap; i++) {
        out[4 * i + 0] = p[i].x; out[4 * i + 1] = p[i].y; out[4 * i + 2] = p[i].width; out[4 * i + 3] = p[i].height;
    }
    return (int)p.size();
}

// ProjectPatches(win, patches) for a tw x tw template.
int ref_project(int tw, const int* win, const int* patches, int n, int* out) {
    DenseSURFFeatureExtractor ex;
    ex.size = Size(tw, tw);
    vector<Rect> a(n), b(n);
    for (int i = 0; i < n; i++) a[i] = Rect(patches[4 * i], patches[4 * i + 1], patches[4 * i + 2], patches[4 * i + 3]);
    ex.ProjectPatches(Rect(win[0], win[1], win[2], win[3]), a, b);
    for (int i = 0; i < n; i++) { out[4 * i] = b[i].x; out[4 * i + 1] = b[i].y; out[4 * i + 2] = b[i].width; out[4 * i + 3] = b[i].height; }
    return 0;
}

// IntegralImage -> out[(H+1)*(W+1)*8] float32, channel-interleaved.
int ref_integral(const uint8_t* img, int W, int H, float* out) {
    DenseSURFFeatureExtractor ex;
    ex.IntegralImage(wrap_u8(img, W, H));
    memcpy(out, ex.summat.data, (size_t)(H + 1) * (W + 1) * 8 * sizeof(float));
    release_integral(ex);
    return 0;
}

// T2bFilter alone -> planar u8 [8][H][W].
int ref_channels(const uint8_t* img, int W, int H, uint8_t* out) {
    DenseSURFFeatureExtractor ex;
    ex.T2bFilter(wrap_u8(img, W, H), out);
    return 0;
}

// CalcFeature for n rects -> out[n][32]; sum() for n rects -> sums[n].  Either output may be null.
int ref_features(const uint8_t* img, int W, int H, const int* rects, int n, float* out, float* sums) {
    DenseSURFFeatureExtractor ex;
    ex.IntegralImage(wrap_u8(img, W, H));
    vector<float> f(32);
    for (int i = 0; i < n; i++) {
        Rect r(rects[4 * i], rects[4 * i + 1], rects[4 * i + 2], rects[4 * i + 3]);
        if (out) { ex.CalcFeature(r, f); memcpy(out + 32 * (size_t)i, f.data(), 32 * sizeof(float)); }
        if (sums) sums[i] = ex.sum(r);
    }
    release_integral(ex);
    return 0;
}

// Stage scores of a loaded model on explicit windows {x,y,l}, every stage evaluated (no early
// exit): out[n][n_stages].  Uses ProjectPatches + CalcFeature + Predict2 exactly like the scan.
int ref_stage_scores(const uint8_t* img, int W, int H, const char* model_cfg, int tmpl, const int* wins, int n, float* out, int max_stages) {
    CoutSilencer quiet(true);
    DenseSURFFeatureExtractor ex;
    ex.size = Size(tmpl, tmpl);
    vector<Rect> pool;
    ex.ExtractPatches(pool);
    CascadeClassifier cc;
    Model model(model_cfg);
    if (model.Load(cc) != EXIT_SUCCESS) return -1;
    vector<vector<int> > idx;
    cc.GetFittedPatchIndexes(idx);
    vector<vector<Rect> > fitted(idx.size());
    for (size_t s = 0; s < idx.size(); s++)
        for (size_t q = 0; q < idx[s].size(); q++) fitted[s].push_back(pool[idx[s][q]]);
    vector<vector<Rect> > proj(fitted);
    ex.IntegralImage(wrap_u8(img, W, H));
    const int S = (int)fitted.size();
    for (int i = 0; i < n; i++) {
        Rect win(wins[3 * i], wins[3 * i + 1], wins[3 * i + 2], wins[3 * i + 2]);
        ex.ProjectPatches(win, fitted, proj);
        for (int s = 0; s < S && s < max_stages; s++) {
            vector<vector<float> > feats(proj[s].size(), vector<float>(32));
            for (size_t q = 0; q < proj[s].size(); q++) ex.CalcFeature(proj[s][q], feats[q]);
            out[(size_t)i * max_stages + s] = cc.stage_classifiers[s]->Predict2(feats);
        }
    }
    release_integral(ex);
    return S;
}

// Model::Load (Model.cpp:97-193) + GetFittedPatchIndexes (CascadeClassifier.cpp:83-91), flattened:
// theta[s], n_weak[s], then per weak classifier patch_index, w[33] (float, as the detector holds them)
// and bias.  Returns the number of stages, or <0.
int ref_model_load(const char* model_cfg, float* theta, int* n_weak, int max_stages, int* patch_index, float* w, double* bias, int max_weak) {
    CoutSilencer quiet(true);
    {
        std::ifstream probe(model_cfg);
        if (!probe.good()) return -1;
    }
    CascadeClassifier cc;
    Model model(model_cfg);
    if (model.Load(cc) != EXIT_SUCCESS) return -2;
    int S = (int)cc.stage_classifiers.size(), k = 0;
    if (S > max_stages) return -3;
    for (int s = 0; s < S; s++) {
        GentleAdaboost* g = static_cast<GentleAdaboost*>(cc.stage_classifiers[s].get());
        theta[s] = g->theta;
        n_weak[s] = (int)g->weak_classifiers.size();
        for (size_t q = 0; q < g->weak_classifiers.size(); q++, k++) {
            if (k >= max_weak) return -4;
            LogisticRegression* lr = g->weak_classifiers[q].get();
            patch_index[k] = lr->patch_index;
            memcpy(w + 33 * (size_t)k, lr->w, 33 * sizeof(float));
            bias[k] = lr->model_->bias;
        }
    }
    return S;
}

// Model::Load followed by Model::Save to another path (round-trip check of a product-written file).
int ref_model_resave(const char* in_cfg, const char* out_cfg) {
    CoutSilencer quiet(true);
    CascadeClassifier cc;
    Model in(in_cfg);
    if (in.Load(cc) != EXIT_SUCCESS) return -1;
    // Model::Save reads model_->w (double) and model_->param; Load only fills the float copy, so mirror it back
    for (size_t s = 0; s < cc.stage_classifiers.size(); s++) {
        GentleAdaboost* g = static_cast<GentleAdaboost*>(cc.stage_classifiers[s].get());
        for (size_t q = 0; q < g->weak_classifiers.size(); q++) {
            LogisticRegression* lr = g->weak_classifiers[q].get();
            lr->model_->w = new double[33];
            for (int i = 0; i < 33; i++) lr->model_->w[i] = lr->w[i];
        }
    }
    Model out(out_cfg);
    return out.Save(cc) == EXIT_SUCCESS ? 0 : -2;
}

// The reference detect path on a batch of equally sized frames.
//   frames        nframes pointers to H*W u8
//   base          base window side (the literal 70 at ObjDetector.cpp:104,174,180)
//   nthreads      OpenMP threads for the scale loop (ObjDetector.cpp:177)
//   do_group      also run groupRectangles (ObjDetector.cpp:224-225) into the g_* outputs
// Raw detections are appended frame after frame: det_frame/x/y/l (int32) and det_score (f64), sorted
// inside each frame by (l, y, x) so the order does not depend on the OpenMP schedule.
int ref_detect_frames(const uint8_t* const* frames, int nframes, int W, int H, const char* model_cfg, int base, int nthreads,
                      int32_t* det_frame, int32_t* det_x, int32_t* det_y, int32_t* det_l, double* det_score, int64_t cap,
                      int64_t* n_det, int64_t* counters /* [nframes][REF_NCOUNTERS] */, double* ms_integral, double* ms_scan,
                      int do_group, int32_t* g_frame, int32_t* g_rect /* [cap][4] */, double* g_score, int64_t* n_group) {
    CoutSilencer quiet(true);
    if (nthreads > 0) omp_set_num_threads(nthreads);
    Model model(model_cfg);
    int length = base;
    Rect win(0, 0, length, length);
    {
        std::ifstream probe(model_cfg);
        if (!probe.good()) return -1;
    }
#include "detect_setup.inc"
    if (cascade_classifier.stage_classifiers.empty()) return -2;

    int64_t total = 0, gtotal = 0;
    for (int j = 0; j < nframes; j++) {
        Mat img = wrap_u8(frames[j], W, H);
        long long c_visited = 0, c_prefilter = 0, c_weak = 0;
        long long c_reach[16] = {0};
#define REF_VISIT() _Pragma("omp atomic") c_visited++
#define REF_PREFILTER() _Pragma("omp atomic") c_prefilter++
#define REF_STAGE(p, nweak)                                   \
    do {                                                      \
        _Pragma("omp atomic") c_weak += (long long)(nweak);   \
        if ((p) < 16) { _Pragma("omp atomic") c_reach[(p)]++; } \
    } while (0)
        double t0 = now_ms();
        dense_surf_feature_extractor.IntegralImage(img);
        double t1 = now_ms();
#include "detect_loop.inc"
        double t2 = now_ms();
#undef REF_VISIT
#undef REF_PREFILTER
#undef REF_STAGE
        release_integral(dense_surf_feature_extractor);
        if (ms_integral) ms_integral[j] = t1 - t0;
        if (ms_scan) ms_scan[j] = t2 - t1;
        // deterministic order
        vector<size_t> order(wins.size());
        for (size_t k = 0; k < order.size(); k++) order[k] = k;
        std::sort(order.begin(), order.end(), [&](size_t a, size_t b) {
            if (wins[a].width != wins[b].width) return wins[a].width < wins[b].width;
            if (wins[a].y != wins[b].y) return wins[a].y < wins[b].y;
            return wins[a].x < wins[b].x;
        });
        for (size_t k = 0; k < order.size(); k++) {
            if (total < cap) {
                det_frame[total] = j; det_x[total] = wins[order[k]].x; det_y[total] = wins[order[k]].y;
                det_l[total] = wins[order[k]].width; det_score[total] = scores[order[k]];
            }
            total++;
        }
        if (counters) {
            int64_t* c = counters + (size_t)j * REF_NCOUNTERS;
            c[REF_C_VISITED] = c_visited; c[REF_C_PREFILTER] = c_prefilter; c[REF_C_WEAK] = c_weak; c[REF_C_RAW] = (int64_t)wins.size();
            for (int s = 0; s < 16; s++) c[REF_C_REACH0 + s] = c_reach[s];
        }
        if (do_group) {
            // grouping is order-sensitive only in its output order; feed the deterministic order
            vector<Rect> gw; vector<double> gs;
            for (size_t k = 0; k < order.size(); k++) { gw.push_back(wins[order[k]]); gs.push_back(scores[order[k]]); }
            vector<int> weights(gw.size(), 0);
            groupRectangles(gw, weights, gs, 2, 0.2);
            for (size_t k = 0; k < gw.size(); k++) {
                if (gtotal < cap) {
                    g_frame[gtotal] = j;
                    g_rect[4 * gtotal] = gw[k].x; g_rect[4 * gtotal + 1] = gw[k].y; g_rect[4 * gtotal + 2] = gw[k].width; g_rect[4 * gtotal + 3] = gw[k].height;
                    g_score[gtotal] = gs[k];
                }
                gtotal++;
            }
        }
        wins.clear();
        scores.clear();
    }
    if (n_det) *n_det = total;
    if (n_group) *n_group = gtotal;
    return 0;
}

// Timing entry (bench.py's CPU legs): the same lifted setup and loop as ref_detect_frames, but with the counter
// hooks compiled to nothing -- ref_detect_frames increments SHARED counters with `omp atomic` inside the innermost
// window loop (tens of millions of contended atomics per 1080p frame), which slows the reference down and would
// inflate every GPU / CPU ratio.  The model is loaded and the pool built ONCE per call, outside the timed per-frame
// regions; IntegralImage and the scan are timed separately per frame; `warmup` leading
// frames are run untimed (frames[0..warmup) are processed first, then all nframes are timed).
// Returns 0; n_raw[j] = raw windows of frame j (sanity check against the counting build).
int ref_detect_timed(const uint8_t* const* frames, int nframes, int W, int H, const char* model_cfg, int base, int nthreads, int warmup,
                     double* ms_integral, double* ms_scan, int64_t* n_raw) {
    CoutSilencer quiet(true);
    if (nthreads > 0) omp_set_num_threads(nthreads);
    Model model(model_cfg);
    int length = base;
    Rect win(0, 0, length, length);
    {
        std::ifstream probe(model_cfg);
        if (!probe.good()) return -1;
    }
#include "detect_setup.inc"
    if (cascade_classifier.stage_classifiers.empty()) return -2;
#define REF_VISIT() do {} while (0)
#define REF_PREFILTER() do {} while (0)
#define REF_STAGE(p, nweak) do {} while (0)
    for (int jj = -warmup; jj < nframes; jj++) {
        const int j = jj < 0 ? (jj + warmup) % nframes : jj;
        Mat img = wrap_u8(frames[j], W, H);
        double t0 = now_ms();
        dense_surf_feature_extractor.IntegralImage(img);
        double t1 = now_ms();
#include "detect_loop.inc"
        double t2 = now_ms();
        release_integral(dense_surf_feature_extractor);
        if (jj >= 0) {
            if (ms_integral) ms_integral[j] = t1 - t0;
            if (ms_scan) ms_scan[j] = t2 - t1;
            if (n_raw) n_raw[j] = (int64_t)wins.size();
        }
        wins.clear();
        scores.clear();
    }
#undef REF_VISIT
#undef REF_PREFILTER
#undef REF_STAGE
    return 0;
}

// Candidate scoring of one boosting round exactly as GentleAdaboost::Train does it (GentleAdaboost.cpp:145-148): the
// stage holds the T already chosen weak classifiers (prev_* arrays), candidate k is pushed, StageClassifier::Evaluate
// (StageClassifier.cpp:35-70) is called on the whole set, the candidate is popped.  X [N][P][32], positives first.
int ref_pool_eval(const float* X, int N, int P, int n_pos, const float* Wcand, const double* bias, const int* prev_patch, const float* prev_w,
                  const double* prev_bias, int T, float* auc) {
    vector<vector<vector<float> > > Xv(N, vector<vector<float> >(P, vector<float>(32)));
    for (int n = 0; n < N; n++)
        for (int k = 0; k < P; k++) memcpy(Xv[n][k].data(), X + ((size_t)n * P + k) * 32, 32 * sizeof(float));
    vector<bool> y(N, false);
    for (int n = 0; n < n_pos; n++) y[n] = true;
    GentleAdaboost stage(0.995f);
    stage.n_total = N; stage.n_pos = n_pos; stage.n_neg = N - n_pos;
    auto make = [](int patch, const float* w33, double b) {
        std::shared_ptr<LogisticRegression> lr(new LogisticRegression(patch));
        memcpy(lr->w, w33, 33 * sizeof(float));
        lr->model_ = new model;
        lr->model_->bias = b;
        return lr;
    };
    for (int t = 0; t < T; t++) stage.weak_classifiers.push_back(make(prev_patch[t], prev_w + 33 * (size_t)t, prev_bias[t]));
    for (int k = 0; k < P; k++) {
        stage.weak_classifiers.push_back(make(k, Wcand + 33 * (size_t)k, bias[k]));
        auc[k] = stage.Evaluate(Xv, y);
        stage.weak_classifiers.pop_back();
    }
    return 0;
}

// The reference --train branch on PGM files: `prefix` is a directory ending in '/', the two
// list files hold one file name per line relative to it.  Writes the model with Model::Save.
// The extractor keeps function-local static cursors, so this can run ONCE per process.
int ref_train(const char* prefix, const char* pos_list, const char* neg_list, const char* out_cfg, int verbose) {
    static bool used = false;
    if (used) return -100;
    used = true;
    CoutSilencer quiet(!verbose);
    Model model(out_cfg);
    string prefix_path(prefix);
    string pos_file(pos_list);
    string neg_file(neg_list);
#include "train_body.inc"
    return (int)cascade_classifier.stage_classifiers.size();
}

// DenseSURFFeatureExtractor::FillNegSamples (DenseSURFFeatureExtractor.cpp:124-195) itself, on PGM files, called
// `n_calls` times in a row on one extractor the way CascadeClassifier::Train does (CascadeClassifier.cpp:37): call c starts
// with an empty sample set and fills it to n_totals[c]; the function's static image cursor carries over from call to call.
// first != 0: every window is taken (stage 0 of the trainer); else the cascade of model_cfg decides (Predict).
// One thread, so the scale loop runs in index order.  out: [sum of counts][P][32]; counts[c] / dones[c] per call.
// The cursor is function-static in the reference, so this can run ONCE per process.
int ref_fill_neg(const char* prefix, const char* neg_list, const char* model_cfg, int tmpl, const int* n_totals, int n_calls, int first,
                 float* out, long long cap_floats, int* counts, int* dones, int* pool_size) {
    static bool used = false;
    if (used) return -100;
    used = true;
    CoutSilencer quiet(true);
    omp_set_num_threads(1);
    DenseSURFFeatureExtractor ex;
    ex.size = Size(tmpl, tmpl);
    ex.LoadFileList(neg_list, prefix, false);
    vector<Rect> patches;
    ex.ExtractPatches(patches);
    *pool_size = (int)patches.size();
    CascadeClassifier cascade;
    if (!first) {
        Model model(model_cfg);
        if (model.Load(cascade) != EXIT_SUCCESS) return -1;
    }
    long long off = 0;
    for (int c = 0; c < n_calls; c++) {
        vector<vector<vector<float>>> X;
        const bool done = ex.FillNegSamples(patches, X, n_totals[c], cascade, first != 0);
        counts[c] = (int)X.size();
        dones[c] = done ? 1 : 0;
        for (size_t i = 0; i < X.size(); i++)
            for (size_t k = 0; k < X[i].size(); k++) {
                if (off + 32 > cap_floats) return -2;
                for (int d = 0; d < 32; d++) out[off + d] = X[i][k][d];
                off += 32;
            }
    }
    return 0;
}

}  // extern "C"
