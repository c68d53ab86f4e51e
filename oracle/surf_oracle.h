/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's detection hot path.
 *
 * Plain-C restatement of mrgloom/SurfCascade's detect path,
 * used as the parity checker for the CUDA path and as the `kind: "port"` CPU baseline.  It is NOT part of
 * the product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  Parity status: PINNED -- tests/test_oracle_vs_ref.py checks every function below bit-for-bit
 * against the reference itself compiled in oracle/_ref (where /root/reference exists), and
 * tests/test_oracle_golden.py against the committed fixtures in tests/golden/ generated from that build.
 *
 * Each function cites the reference lines it follows (paths relative to /root/reference/ObjDetector).
 */
#ifndef SURF_ORACLE_H
#define SURF_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SO_DIM 32          /* n_cells * n_bins, DenseSURFFeatureExtractor.h:32-33,48 */
#define SO_MAX_STAGES 16   /* reference caps at 10, CascadeClassifier.h:20 */

/* Flattened cascade: what the detect loop reads from a loaded model. */
typedef struct {
    int n_stages;
    const float* theta;      /* [n_stages]  StageClassifier::theta */
    const int* n_weak;       /* [n_stages] */
    const int* rects;        /* [total_weak][4] template rect x,y,w,h = dense_patches[patch_index] */
    const float* w;          /* [total_weak][33] LogisticRegression::w (already cast to f32) */
    const double* bias;      /* [total_weak]   model_->bias */
} so_cascade;

typedef struct {
    int tmpl;            /* model template side, 40 (ObjDetector.cpp:112) */
    int base;            /* base window side (literal 70 at ObjDetector.cpp:104; BASELINE uses 40) */
    int step;            /* 0 -> base > 20 ? base / 20 : 1 (ObjDetector.cpp:139) */
    double scale;        /* 1.1 (ObjDetector.cpp:174,180) */
    int prefilter;       /* 6 (ObjDetector.cpp:188); < 0 disables the prefilter */
    int skip_rule;       /* 1 -> adaptive stride `multi` (ObjDetector.cpp:186,214-217); 0 -> every grid window */
    int force_all;       /* 1 -> evaluate every stage of every window (stress mode, not in the reference) */
    int nthreads;        /* OpenMP threads over scales, like ObjDetector.cpp:177 */
} so_params;

enum { SO_C_GRID = 0, SO_C_VISITED, SO_C_PREFILTER, SO_C_WEAK, SO_C_RAW, SO_C_WEAK_SQUARE, SO_C_WEAK_LONG, SO_C_REACH0, SO_NCOUNTERS = SO_C_REACH0 + SO_MAX_STAGES };

int so_pool_patches(int tw, int th, int* out, int cap);
void so_project(int tmpl, int l, const int* patch, int* out);
int so_cells(const int* rect, int* cells);
void so_channels(const uint8_t* img, int W, int H, uint8_t* out);
void so_integral(const uint8_t* img, int W, int H, float* S);
float so_window_sum(const float* S, int W, int x, int y, int w, int h);
void so_feature(const float* S, int W, const int* rect, float* out);
void so_features(const float* S, int W, const int* rects, int n, float* out, float* sums);
float so_weak(const float* w33, double bias, const float* x);
int so_num_scales(int W, int H, const so_params* p, int* sides, int cap);
float so_stage_score(const float* S, int W, const so_cascade* c, int tmpl, int stage, int x, int y, int l);

/* Per-grid-window outcome of scale index `si` (debug/parity helper): reached[k] = stage index where the
 * window was rejected (n_stages = passed all, -1 = prefilter failed), score[k] = last evaluated stage score. */
int so_grid_outcomes(const float* S, int W, int H, const so_cascade* c, const so_params* p, int si,
                     int8_t* reached, float* score, int cap);

/* The detect scan (ObjDetector.cpp:174-219) on a precomputed integral.  Detections sorted by (l, y, x). */
int64_t so_detect(const float* S, int W, int H, const so_cascade* c, const so_params* p,
                  int32_t* det_x, int32_t* det_y, int32_t* det_l, double* det_score, int64_t cap, int64_t* counters);

/* cv::groupRectangles(rects, weights = 0.., scores, thr, eps) as called at ObjDetector.cpp:224-225. */
int so_group_rectangles(const int32_t* rects, const double* scores, int n, int thr, double eps,
                        int32_t* out_rects, double* out_scores, int cap);

/* Candidate scoring for boosting (GentleAdaboost.cpp:145-148 -> StageClassifier::Evaluate, StageClassifier.cpp:35-70). */
void so_pool_eval(const float* X, int N, int P, int n_pos, const float* Wcand, const double* bias, const float* prior_sum, int T, float* auc);

#ifdef __cplusplus
}
#endif
#endif
