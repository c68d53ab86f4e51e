// See sc_host.h.
#include "sc_host.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <memory>

#include "CascadeClassifier/CascadeClassifier.h"
#include "CascadeClassifier/GentleAdaboost.h"
#include "CascadeClassifier/LogisticRegression.h"
#include "Model.h"
#include "sc_access.h"

namespace sc_host {

void pool_patches(int tw, int th, std::vector<sc_rect>* out) {
    // three cell layouts (columns x rows): 2x2, 1x4, 4x1; cell edge 6 .. tw/2; origin stride 4
    static const int cols[3] = {2, 1, 4}, rows[3] = {2, 4, 1};
    out->clear();
    for (int shape = 0; shape < 3; shape++)
        for (int edge = 6; edge <= tw / 2; edge++) {
            const int pw = cols[shape] * edge, ph = rows[shape] * edge;
            for (int y = 0; y + ph <= th; y += 4)
                for (int x = 0; x + pw <= tw; x += 4) out->push_back(sc_rect{x, y, pw, ph});
        }
}

sc_rect project_patch(int tmpl, int l, const sc_rect& p) {
    const float scale = (float)l / (float)tmpl;
    sc_rect r;
    r.x = (int)((float)p.x * scale);
    r.y = (int)((float)p.y * scale);
    if (p.w >= p.h) {
        r.h = (int)((float)p.h * scale);
        r.w = r.h * (p.w / p.h);
    } else {
        r.w = (int)((float)p.w * scale);
        r.h = r.w * (p.h / p.w);
    }
    return r;
}

bool project_geom(int tmpl, int l, const sc_rect& patch, const ScLayout& L, int x0, ScGeom* g) {
    sc_rect r = project_patch(tmpl, l, patch);
    r.x += x0;  // column residue of the window origin inside the layout's deinterleave period
    for (int k = 0; k < 10; k++) g->c[k] = 0;
    g->pad = 0;
    if (r.w == r.h) {
        const int ce = r.w / 2;
        if (ce < 1) return false;
        g->shape = 0;
        for (int b = 0; b < 3; b++)
            for (int a = 0; a < 3; a++) g->c[3 * b + a] = (uint32_t)(16 * sc_layout_index(L, r.x + a * ce, r.y + b * ce));
    } else {
        const int ce = std::min(r.w, r.h);
        if (ce < 1 || std::max(r.w, r.h) != 4 * ce) return false;
        g->shape = 1;
        const bool wide = r.w > r.h;
        for (int k = 0; k < 5; k++) {
            // first line: the chain of corners along the long side; second line: one cell edge across
            const int x0 = r.x + (wide ? k * ce : 0), y0 = r.y + (wide ? 0 : k * ce);
            g->c[k] = (uint32_t)(16 * sc_layout_index(L, x0, y0));
            g->c[5 + k] = (uint32_t)(16 * sc_layout_index(L, x0 + (wide ? 0 : ce), y0 + (wide ? ce : 0)));
        }
    }
    return true;
}

ScLayout make_layout(int W, int H, int sx, int sy, int min_hp) {
    ScLayout L;
    L.sx = sx < 1 ? 1 : sx;
    L.sy = sy < 1 ? 1 : sy;
    const int cols = (W + 1 + L.sx - 1) / L.sx;
    L.hp = min_hp < 8 ? 8 : min_hp;  // the scan kernels are instantiated for 256..4096 (callers reject more); hooks take any power of two
    while (L.hp < cols) L.hp *= 2;
    L.ppitch = SC_ROW_ELEMS(L.hp);
    L.prows = (H + 1 + L.sy - 1) / L.sy;
    L.pad = 0;
    L.plane4 = (long long)L.ppitch * L.prows;
    L.frame4 = (long long)L.sx * L.sy * L.plane4;
    return L;
}

void scale_ladder(int W, int H, int base, double scale, std::vector<int>* sides) {
    sides->clear();
    // the ratios are float, their logs are taken in float (MSVC resolves log(float) to the float overload),
    // the quotient and the comparison are double
    const double a = (double)logf((float)W / (float)base) / log(scale);
    const double b = (double)logf((float)H / (float)base) / log(scale);
    const int n = (int)std::min(a, b);
    for (int i = 0; i <= n; i++) sides->push_back((int)((double)base * pow(scale, (double)i)));
}

namespace {
int find_root(std::vector<int>& parent, int i) {
    while (parent[i] != i) { parent[i] = parent[parent[i]]; i = parent[i]; }
    return i;
}
bool similar(const sc_rect& a, const sc_rect& b, double eps) {
    const double delta = eps * (std::min(a.w, b.w) + std::min(a.h, b.h)) * 0.5;
    return std::abs(a.x - b.x) <= delta && std::abs(a.y - b.y) <= delta && std::abs(a.x + a.w - b.x - b.w) <= delta &&
           std::abs(a.y + a.h - b.y - b.h) <= delta;
}
int round_even(double v) { return (int)lrint(v); }
}  // namespace

void group_rectangles(std::vector<sc_rect>* rects, std::vector<double>* scores, int thr, double eps) {
    const int n = (int)rects->size();
    if (thr <= 0 || n == 0) return;
    const std::vector<sc_rect>& r = *rects;
    std::vector<int> parent(n), cls(n, -1), first(n, -1);
    for (int i = 0; i < n; i++) parent[i] = i;
    // cv::partition tests every pair; the classes are the connected components of `similar`, whatever the order the
    // pairs are met in.  similar(a, b) needs |a.x - b.x| <= delta <= eps (a.w + a.h) / 2, so with the rectangles ordered
    // by x only the partners inside that reach are tested (raw detection lists: 1080p frames with ~500 windows, ~6 x fewer
    // tests).  Class numbers below are still given in index order, as cv::partition's first-seen labels are.
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return r[a].x != r[b].x ? r[a].x < r[b].x : a < b; });
    for (int oi = 0; oi < n; oi++) {
        const int i = order[oi];
        const double reach = eps * (r[i].w + r[i].h) * 0.5;
        for (int oj = oi + 1; oj < n; oj++) {
            const int j = order[oj];
            if ((double)(r[j].x - r[i].x) > reach) break;
            const int a = find_root(parent, i), b = find_root(parent, j);
            if (a != b && similar(r[i], r[j], eps)) parent[std::max(a, b)] = std::min(a, b);  // same class already: nothing to learn
        }
    }
    int k = 0;
    for (int i = 0; i < n; i++) {
        const int root = find_root(parent, i);
        if (first[root] < 0) first[root] = k++;
        cls[i] = first[root];
    }
    struct Cluster { long long x = 0, y = 0, w = 0, h = 0; int members = 0; double best = DBL_MIN; sc_rect mean{0, 0, 0, 0}; };
    std::vector<Cluster> c(k);
    for (int i = 0; i < n; i++) {
        Cluster& q = c[cls[i]];
        q.x += r[i].x; q.y += r[i].y; q.w += r[i].w; q.h += r[i].h; q.members++;
        if ((*scores)[i] > q.best) q.best = (*scores)[i];
    }
    for (Cluster& q : c) {
        const float inv = 1.f / (float)q.members;
        q.mean = sc_rect{round_even((double)((float)(int)q.x * inv)), round_even((double)((float)(int)q.y * inv)),
                         round_even((double)((float)(int)q.w * inv)), round_even((double)((float)(int)q.h * inv))};
    }
    std::vector<sc_rect> out_r;
    std::vector<double> out_s;
    for (int i = 0; i < k; i++) {
        if (c[i].members <= thr) continue;
        const sc_rect& a = c[i].mean;
        bool swallowed = false;
        for (int j = 0; j < k && !swallowed; j++) {
            if (j == i || c[j].members <= thr) continue;
            const sc_rect& b = c[j].mean;
            const int dx = round_even(b.w * eps), dy = round_even(b.h * eps);
            swallowed = a.x >= b.x - dx && a.y >= b.y - dy && a.x + a.w <= b.x + b.w + dx && a.y + a.h <= b.y + b.h + dy &&
                        (c[j].members > std::max(3, c[i].members) || c[i].members < 3);
        }
        if (!swallowed) { out_r.push_back(a); out_s.push_back(c[i].best); }
    }
    rects->swap(out_r);
    scores->swap(out_s);
}

bool Access::flatten(CascadeClassifier& cc, int tmpl, FlatCascade* out, std::string* why) {
    std::vector<sc_rect> pool;
    pool_patches(tmpl, tmpl, &pool);
    out->theta.clear(); out->n_weak.clear(); out->rects.clear(); out->patch_index.clear(); out->w.clear(); out->bias.clear();
    for (auto& st : cc.stage_classifiers) {
        GentleAdaboost* g = dynamic_cast<GentleAdaboost*>(st.get());
        if (!g) { if (why) *why = "stage is not a GentleAdaboost"; return false; }
        out->theta.push_back(g->theta);
        out->n_weak.push_back((int)g->weak_classifiers.size());
        for (auto& wk : g->weak_classifiers) {
            if (wk->patch_index < 0 || wk->patch_index >= (int)pool.size()) { if (why) *why = "patch_index outside the template pool"; return false; }
            out->rects.push_back(pool[wk->patch_index]);
            out->patch_index.push_back(wk->patch_index);
            out->w.insert(out->w.end(), wk->w, wk->w + 33);
            out->bias.push_back(wk->bias_);
        }
    }
    if (out->theta.empty()) { if (why) *why = "model holds no stages"; return false; }
    return true;
}

bool resave_model(const std::string& in_cfg, const std::string& out_cfg) {
    CascadeClassifier cc;
    Model in(in_cfg);
    if (in.Load(cc) != EXIT_SUCCESS) return false;
    Model out(out_cfg);
    return out.Save(cc) == EXIT_SUCCESS;
}

bool load_flat_cascade(const std::string& model_cfg, int tmpl, FlatCascade* out, std::string* why) {
    CascadeClassifier cc;
    Model model(model_cfg);
    if (model.Load(cc) != EXIT_SUCCESS) { if (why) *why = "cannot read or parse " + model_cfg; return false; }
    return Access::flatten(cc, tmpl, out, why);
}

}  // namespace sc_host
