/* TEST INFRASTRUCTURE ONLY -- see surf_oracle.h.  Plain-C restatement of the reference's detect path.
 * Build: -O2 -msse2 -mno-fma -ffp-contract=off (every float operation below is ONE IEEE binary32
 * rounding, as the reference's SSE code produces; no contraction into FMA). */
#include "surf_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- geometry -------------------------------------------------------------------------------------- */

/* ExtractPatches, FeatureExtractors/DenseSURFFeatureExtractor.cpp:49-63; shapes :21, step/min edge .h:34-35 */
int so_pool_patches(int tw, int th, int* out, int cap) {
    static const int shape_w[3] = {2, 1, 4}, shape_h[3] = {2, 4, 1};
    int n = 0;
    for (int j = 0; j < 3; j++)
        for (int ce = 6; ce <= tw / 2; ce++) {
            int pw = shape_w[j] * ce, ph = shape_h[j] * ce;
            for (int y = 0; y + ph <= th; y += 4)
                for (int x = 0; x + pw <= tw; x += 4) {
                    if (n < cap) { out[4 * n] = x; out[4 * n + 1] = y; out[4 * n + 2] = pw; out[4 * n + 3] = ph; }
                    n++;
                }
        }
    return n;
}

/* ProjectPatches relative to the window origin, DenseSURFFeatureExtractor.cpp:486-508 (float scale,
 * float multiply, truncation; the long side is the short side times the integer aspect ratio). */
void so_project(int tmpl, int l, const int* patch, int* out) {
    float scale = (float)l / tmpl;
    out[0] = (int)(patch[0] * scale);
    out[1] = (int)(patch[1] * scale);
    if (patch[2] >= patch[3]) {
        int ratio = patch[2] / patch[3];
        out[3] = (int)(patch[3] * scale);
        out[2] = out[3] * ratio;
    } else {
        int ratio = patch[3] / patch[2];
        out[2] = (int)(patch[2] * scale);
        out[3] = out[2] * ratio;
    }
}

/* GetRectsFromPatch, DenseSURFFeatureExtractor.cpp:360-377.  cells[k] = x,y,ce,ce, row-major. */
int so_cells(const int* rect, int* cells) {
    int ce = (rect[2] == rect[3]) ? rect[2] / 2 : (rect[2] < rect[3] ? rect[2] : rect[3]);
    int nx = rect[2] / ce, ny = rect[3] / ce;
    for (int r = 0; r < ny; r++)
        for (int c = 0; c < nx; c++) {
            int k = r * nx + c;
            cells[4 * k] = rect[0] + c * ce; cells[4 * k + 1] = rect[1] + r * ce; cells[4 * k + 2] = ce; cells[4 * k + 3] = ce;
        }
    return nx * ny;
}

/* ---- channels + integral --------------------------------------------------------------------------- */

/* T2bFilter, DenseSURFFeatureExtractor.cpp:199-349: four central differences with replicated borders,
 * each split into its negative part (even channel) and positive part (odd channel); planar u8 output. */
void so_channels(const uint8_t* img, int W, int H, uint8_t* out) {
    size_t sz = (size_t)W * H;
    for (int y = 0; y < H; y++) {
        int yp = y > 0 ? y - 1 : 0, yn = y < H - 1 ? y + 1 : H - 1;
        const uint8_t *r0 = img + (size_t)yp * W, *r1 = img + (size_t)y * W, *r2 = img + (size_t)yn * W;
        for (int x = 0; x < W; x++) {
            int xp = x > 0 ? x - 1 : 0, xn = x < W - 1 ? x + 1 : W - 1;
            int d[4];
            d[0] = (int)r1[xn] - (int)r1[xp];   /* dx :224-254 */
            d[1] = (int)r2[x] - (int)r0[x];     /* dy :256-281 */
            d[2] = (int)r2[xn] - (int)r0[xp];   /* du :283-314 */
            d[3] = (int)r0[xn] - (int)r2[xp];   /* dv :316-347 */
            size_t o = (size_t)y * W + x;
            for (int k = 0; k < 4; k++) {
                out[(2 * k) * sz + o] = (uint8_t)(d[k] < 0 ? -d[k] : 0);
                out[(2 * k + 1) * sz + o] = (uint8_t)(d[k] > 0 ? d[k] : 0);
            }
        }
    }
}

/* IntegralImage, DenseSURFFeatureExtractor.cpp:65-87: per channel cv::integral(u8 -> f32) = exact row
 * prefix added, in float, to the value directly above (sequential in y; inexact past 2^24), then
 * cv::merge into 8 interleaved floats per pixel.  S is (H+1) x (W+1) x 8. */
void so_integral(const uint8_t* img, int W, int H, float* S) {
    size_t sz = (size_t)W * H, pitch = (size_t)(W + 1) * 8;
    uint8_t* ch = (uint8_t*)malloc(sz * 8);
    so_channels(img, W, H, ch);
    memset(S, 0, pitch * sizeof(float));
    for (int y = 0; y < H; y++) {
        float* cur = S + (size_t)(y + 1) * pitch;
        const float* up = S + (size_t)y * pitch;
        for (int c = 0; c < 8; c++) cur[c] = 0.f;
        for (int c = 0; c < 8; c++) {
            const uint8_t* row = ch + c * sz + (size_t)y * W;
            float run = 0.f;
            for (int x = 0; x < W; x++) {
                run += (float)row[x];
                cur[(size_t)(x + 1) * 8 + c] = up[(size_t)(x + 1) * 8 + c] + run;
            }
        }
    }
    free(ch);
}

/* ---- prefilter + descriptor ------------------------------------------------------------------------- */

#define PIX(S, W, x, y) ((S) + ((size_t)(y) * ((W) + 1) + (x)) * 8)

/* sum(), DenseSURFFeatureExtractor.cpp:351-358: channels 0-3, (A+D)-(B+C), ((s0+s1)+s2)+s3, /2 */
float so_window_sum(const float* S, int W, int x, int y, int w, int h) {
    const float *a = PIX(S, W, x, y), *d = PIX(S, W, x + w, y + h), *b = PIX(S, W, x + w, y), *c = PIX(S, W, x, y + h);
    float s[4];
    for (int k = 0; k < 4; k++) s[k] = (a[k] + d[k]) - (b[k] + c[k]);
    return (s[0] + s[1] + s[2] + s[3]) / 2;
}

/* The hadd-ordered sum of squares of Normalize, DenseSURFFeatureExtractor.cpp:427-433 / 441-451. */
static float sumsq_hadd(const float* v) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = FLT_EPSILON; /* _mm_set_ps(FLT_EPSILON,0,0,0): lane 3 */
    for (int g = 0; g < 8; g++) {
        float q0 = v[4 * g] * v[4 * g], q1 = v[4 * g + 1] * v[4 * g + 1];
        float q2 = v[4 * g + 2] * v[4 * g + 2], q3 = v[4 * g + 3] * v[4 * g + 3];
        float n0 = s0 + s1, n1 = s2 + s3, n2 = q0 + q1, n3 = q2 + q3;
        s0 = n0; s1 = n1; s2 = n2; s3 = n3;
    }
    float t0 = s0 + s1, t1 = s2 + s3; /* hadd(s,s) */
    return t0 + t1;                   /* hadd again */
}

/* CalcFeature + Normalize, DenseSURFFeatureExtractor.cpp:379-457 */
void so_feature(const float* S, int W, const int* rect, float* out) {
    int cells[16];
    so_cells(rect, cells);
    for (int k = 0; k < 4; k++) {
        int x = cells[4 * k], y = cells[4 * k + 1], ce = cells[4 * k + 2];
        const float *a = PIX(S, W, x, y), *d = PIX(S, W, x + ce, y + ce), *b = PIX(S, W, x + ce, y), *c = PIX(S, W, x, y + ce);
        for (int j = 0; j < 8; j++) out[8 * k + j] = (a[j] + d[j]) - (b[j] + c[j]);
    }
    const float theta = 2 / sqrtf((float)SO_DIM); /* DenseSURFFeatureExtractor.h:36 */
    float t = sqrtf(sumsq_hadd(out)) * theta;
    float t2 = -t;
    for (int i = 0; i < SO_DIM; i++) {
        float v = out[i];
        v = v < t ? v : t;    /* _mm_min_ps(v, t) */
        v = v > t2 ? v : t2;  /* _mm_max_ps(v, -t) */
        out[i] = v;
    }
    float inv = 1.f / sqrtf(sumsq_hadd(out));
    for (int i = 0; i < SO_DIM; i++) out[i] = out[i] * inv;
}

/* so_feature + so_window_sum over n rects (x, y, w, h): a loop around the two functions above for the Python binding */
void so_features(const float* S, int W, const int* rects, int n, float* out, float* sums) {
    for (int i = 0; i < n; i++) {
        const int* r = rects + 4 * (size_t)i;
        if (out) so_feature(S, W, r, out + 32 * (size_t)i);
        if (sums) sums[i] = so_window_sum(S, W, r[0], r[1], r[2], r[3]);
    }
}

/* LogisticRegression::Predict, CascadeClassifier/LogisticRegression.cpp:46-68 */
float so_weak(const float* w, double bias, const float* x) {
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = 0; i < SO_DIM; i += 4)
        for (int j = 0; j < 4; j++) {
            float t = w[i + j] * x[i + j];
            s[j] = t + s[j];
        }
    float z = (s[0] + s[1]) + (s[2] + s[3]);
    double prob = z;
    prob += w[SO_DIM] * bias;
    prob = 1 / (1 + exp(-prob));
    return (float)prob;
}

static int weak_base(const so_cascade* c, int stage) {
    int b = 0;
    for (int s = 0; s < stage; s++) b += c->n_weak[s];
    return b;
}

/* One stage on one window: ProjectPatches + CalcFeature per weak classifier (ObjDetector.cpp:189-195) and
 * GentleAdaboost::Predict2 (GentleAdaboost.cpp:247-261): f32 running sum in weak order, f32 divide. */
float so_stage_score(const float* S, int W, const so_cascade* c, int tmpl, int stage, int x, int y, int l) {
    int b = weak_base(c, stage), n = c->n_weak[stage];
    float sum = 0.f;
    for (int q = 0; q < n; q++) {
        int r[4];
        float f[SO_DIM];
        so_project(tmpl, l, c->rects + 4 * (size_t)(b + q), r);
        r[0] += x; r[1] += y;
        so_feature(S, W, r, f);
        sum += so_weak(c->w + 33 * (size_t)(b + q), c->bias[b + q], f);
    }
    return sum / (float)n;
}

/* ---- scan ------------------------------------------------------------------------------------------ */

static int eff_step(const so_params* p) { return p->step > 0 ? p->step : (p->base > 20 ? p->base / 20 : 1); }

/* ObjDetector.cpp:174,180: scale count from float ratios and double logs; side = (int)(base * 1.1^i) */
int so_num_scales(int W, int H, const so_params* p, int* sides, int cap) {
    double a = logf(W / (float)p->base) / log(p->scale), b = logf(H / (float)p->base) / log(p->scale);
    int n = (int)(a < b ? a : b);
    int k = 0;
    for (int i = 0; i <= n; i++, k++)
        if (k < cap) sides[k] = (int)(p->base * pow(p->scale, (double)i));
    return k;
}

typedef struct { int reached; float score; int pass_prefilter; } outcome;

/* One window: prefilter (ObjDetector.cpp:188), stages with early reject (:193-199). */
static outcome eval_window(const float* S, int W, const so_cascade* c, const so_params* p, int x, int y, int l, int64_t* cnt) {
    outcome o = {-1, 0.f, 0};
    if (p->prefilter >= 0 && !(so_window_sum(S, W, x, y, l, l) > (float)(l * l * p->prefilter))) return o;
    o.pass_prefilter = 1;
    if (cnt) cnt[SO_C_PREFILTER]++;
    int s, rejected = -1;
    for (s = 0; s < c->n_stages; s++) {
        if (cnt) {
            cnt[SO_C_WEAK] += c->n_weak[s];
            if (s < SO_MAX_STAGES) cnt[SO_C_REACH0 + s]++;
        }
        float sc = so_stage_score(S, W, c, p->tmpl, s, x, y, l);
        if (rejected < 0) o.score = sc;
        if (sc < c->theta[s] && rejected < 0) {
            rejected = s;
            if (!p->force_all) break;
        }
    }
    o.reached = rejected < 0 ? c->n_stages : rejected;
    return o;
}

int so_grid_outcomes(const float* S, int W, int H, const so_cascade* c, const so_params* p, int si, int8_t* reached, float* score, int cap) {
    int sides[128];
    int ns = so_num_scales(W, H, p, sides, 128);
    if (si >= ns) return -1;
    int l = sides[si], step = eff_step(p), k = 0;
    for (int y = 0; y <= H - l; y += step)
        for (int x = 0; x <= W - l; x += step, k++) {
            if (k >= cap) continue;
            outcome o = eval_window(S, W, c, p, x, y, l, 0);
            reached[k] = (int8_t)(o.pass_prefilter ? o.reached : -1);
            score[k] = o.score;
        }
    return k;
}

typedef struct { int32_t x, y, l; double score; } det;

static int det_cmp(const void* a, const void* b) {
    const det *p = (const det*)a, *q = (const det*)b;
    if (p->l != q->l) return p->l < q->l ? -1 : 1;
    if (p->y != q->y) return p->y < q->y ? -1 : 1;
    return p->x < q->x ? -1 : (p->x > q->x ? 1 : 0);
}

/* The scan loop, ObjDetector.cpp:174-219, including the adaptive x stride (multi) and the final score. */
int64_t so_detect(const float* S, int W, int H, const so_cascade* c, const so_params* p, int32_t* det_x, int32_t* det_y,
                  int32_t* det_l, double* det_score, int64_t cap, int64_t* counters) {
    int sides[128];
    int ns = so_num_scales(W, H, p, sides, 128);
    int step = eff_step(p);
    int64_t total[SO_NCOUNTERS];
    memset(total, 0, sizeof(total));
    det* all = 0;
    int64_t n_all = 0, cap_all = 0;
#ifdef _OPENMP
    if (p->nthreads > 0) omp_set_num_threads(p->nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < ns; i++) {
        int l = sides[i];
        int64_t cnt[SO_NCOUNTERS];
        memset(cnt, 0, sizeof(cnt));
        det* mine = 0;
        int64_t n_mine = 0, cap_mine = 0;
        for (int y = 0; y <= H - l; y += step) {
            int multi = 1;
            cnt[SO_C_GRID] += (W - l) / step + 1;
            for (int x = 0; x <= W - l; x += multi * step) {
                cnt[SO_C_VISITED]++;
                outcome o = eval_window(S, W, c, p, x, y, l, cnt);
                if (o.pass_prefilter) {
                    /* :201 score = (score + p + 1) / stages.size() in double */
                    double score = ((double)o.score + o.reached + 1) / (double)c->n_stages;
                    if (o.reached == c->n_stages) {
                        if (n_mine == cap_mine) { cap_mine = cap_mine ? cap_mine * 2 : 256; mine = (det*)realloc(mine, cap_mine * sizeof(det)); }
                        det d = {x, y, l, score};
                        mine[n_mine++] = d;
                    }
                    multi = (score < 0.5) ? 2 : 1; /* :214 */
                } else
                    multi = 2;                     /* :216-217 */
                if (!p->skip_rule) multi = 1;
            }
        }
#pragma omp critical
        {
            for (int k = 0; k < SO_NCOUNTERS; k++) total[k] += cnt[k];
            if (n_all + n_mine > cap_all) { cap_all = (n_all + n_mine) * 2 + 256; all = (det*)realloc(all, cap_all * sizeof(det)); }
            if (n_mine) memcpy(all + n_all, mine, n_mine * sizeof(det));
            n_all += n_mine;
        }
        free(mine);
    }
    total[SO_C_RAW] = n_all;
    /* split weak evaluations by patch shape (9 unique corners for 2x2, 10 for 1x4 / 4x1) */
    for (int s = 0, b = 0; s < c->n_stages && s < SO_MAX_STAGES; b += c->n_weak[s], s++)
        for (int q = 0; q < c->n_weak[s]; q++) {
            const int* r = c->rects + 4 * (size_t)(b + q);
            total[r[2] == r[3] ? SO_C_WEAK_SQUARE : SO_C_WEAK_LONG] += total[SO_C_REACH0 + s];
        }
    if (n_all) qsort(all, n_all, sizeof(det), det_cmp);
    for (int64_t k = 0; k < n_all && k < cap; k++) { det_x[k] = all[k].x; det_y[k] = all[k].y; det_l[k] = all[k].l; det_score[k] = all[k].score; }
    free(all);
    if (counters) memcpy(counters, total, sizeof(total));
    return n_all;
}

/* ---- grouping (next row N1) --------------------------------------------------------------------------- */

static int uf_find(int* p, int i) {
    while (p[i] != i) { p[i] = p[p[i]]; i = p[i]; }
    return i;
}

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* cv::groupRectangles as called at ObjDetector.cpp:224-225 (weights all 0 -> best = max score); OpenCV is
 * not vendored by the reference: restated from SURVEY.md Appendix A.6 and cross-checked against cv2. */
int so_group_rectangles(const int32_t* r, const double* scores, int n, int thr, double eps, int32_t* out_rects, double* out_scores, int cap) {
    if (n <= 0) return 0;
    int* parent = (int*)malloc(sizeof(int) * n);
    int* label = (int*)malloc(sizeof(int) * n);
    int* id = (int*)malloc(sizeof(int) * n);
    for (int i = 0; i < n; i++) { parent[i] = i; id[i] = -1; }
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++) {
            const int32_t *a = r + 4 * i, *b = r + 4 * j;
            double delta = eps * (imin(a[2], b[2]) + imin(a[3], b[3])) * 0.5;
            if (abs(a[0] - b[0]) <= delta && abs(a[1] - b[1]) <= delta && abs(a[0] + a[2] - b[0] - b[2]) <= delta &&
                abs(a[1] + a[3] - b[1] - b[3]) <= delta) {
                int ra = uf_find(parent, i), rb = uf_find(parent, j);
                if (ra != rb) parent[imax(ra, rb)] = imin(ra, rb);
            }
        }
    int k = 0;
    for (int i = 0; i < n; i++) {
        int root = uf_find(parent, i);
        if (id[root] < 0) id[root] = k++;
        label[i] = id[root];
    }
    int* acc = (int*)calloc((size_t)k * 4, sizeof(int));
    int* members = (int*)calloc(k, sizeof(int));
    double* best = (double*)malloc(sizeof(double) * k);
    for (int c = 0; c < k; c++) best[c] = DBL_MIN;
    for (int i = 0; i < n; i++) {
        int c = label[i];
        for (int t = 0; t < 4; t++) acc[4 * c + t] += r[4 * i + t];
        members[c]++;
        if (scores[i] > best[c]) best[c] = scores[i];
    }
    for (int c = 0; c < k; c++) {
        float inv = 1.f / members[c];
        for (int t = 0; t < 4; t++) acc[4 * c + t] = (int)lrint((double)(acc[4 * c + t] * inv));
    }
    int m = 0;
    for (int i = 0; i < k; i++) {
        if (members[i] <= thr) continue;
        const int* a = acc + 4 * i;
        int j = 0;
        for (; j < k; j++) {
            if (j == i || members[j] <= thr) continue;
            const int* b = acc + 4 * j;
            int dx = (int)lrint(b[2] * eps), dy = (int)lrint(b[3] * eps);
            if (a[0] >= b[0] - dx && a[1] >= b[1] - dy && a[0] + a[2] <= b[0] + b[2] + dx && a[1] + a[3] <= b[1] + b[3] + dy &&
                (members[j] > imax(3, members[i]) || members[i] < 3))
                break;
        }
        if (j == k) {
            if (m < cap) { memcpy(out_rects + 4 * m, a, 4 * sizeof(int32_t)); out_scores[m] = best[i]; }
            m++;
        }
    }
    free(parent); free(label); free(id); free(acc); free(members); free(best);
    return m;
}

/* ---- training-side pool evaluation (row A9, config C5) --------------------------------------------------- */

/* Candidate scoring of one boosting round, GentleAdaboost.cpp:145-148: the candidate weak classifier k (on pool patch
 * k) is appended to the T already chosen ones and StageClassifier::Evaluate (StageClassifier.cpp:35-70) computes the
 * AUC of GentleAdaboost::Predict (GentleAdaboost.cpp:233-245) over all samples: prob_n = (prior_n + p_k(x_n[k])) / (T+1)
 * in float, positives first; 20 float thresholds 1, 1-0.05f, ...; TPR/FPR by >=; trapezoid area in float.
 * X [N][P][32], Wcand [P][33], bias [P], prior_sum [N] (float running sum of the T chosen outputs; NULL -> 0). */
void so_pool_eval(const float* X, int N, int P, int n_pos, const float* Wcand, const double* bias, const float* prior_sum, int T, float* auc) {
    int n_neg = N - n_pos;
#pragma omp parallel for schedule(static)
    for (int k = 0; k < P; k++) {
        float area = 0.f, tpr_prev = 0.f, fpr_prev = 0.f;
        int it = 0;
        float* probs = (float*)malloc(sizeof(float) * (size_t)N);
        for (int n = 0; n < N; n++) {
            float sum = prior_sum ? prior_sum[n] : 0.f;
            sum += so_weak(Wcand + 33 * (size_t)k, bias[k], X + ((size_t)n * P + k) * SO_DIM);
            probs[n] = sum / (float)(T + 1);
        }
        for (float t = 1; t >= 0; t -= 0.05f, it++) {
            long cp = 0, cn = 0;
            for (int n = 0; n < n_pos; n++) cp += probs[n] >= t;
            for (int n = n_pos; n < N; n++) cn += probs[n] >= t;
            float tpr = cp / (float)n_pos, fpr = cn / (float)n_neg;
            if (it > 0) area += (tpr + tpr_prev) * (fpr - fpr_prev) / 2;
            tpr_prev = tpr; fpr_prev = fpr;
        }
        auc[k] = area;
        free(probs);
    }
}
