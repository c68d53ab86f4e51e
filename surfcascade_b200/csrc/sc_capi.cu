// C-ABI of the B200-native SURF-cascade detection path (include/surfcascade.h).
// Handle = one device, one stream, device buffers sized for a group of frames; no CPU fallback anywhere.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/surfcascade.h"
#include "../host/sc_host.h"
#include "sc_kernels.cuh"
#include "sc_plan.h"

static_assert(sizeof(sc_detection) == sizeof(sck::ScDetOut), "detection layout");
static_assert(sizeof(sc_detection) == 24, "detection layout");

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct HostBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMallocHost(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

bool same_params(const sc_detect_params& a, const sc_detect_params& b) {
    return a.base == b.base && a.step == b.step && a.scale == b.scale && a.prefilter == b.prefilter && a.skip_rule == b.skip_rule &&
           a.force_all_stages == b.force_all_stages && a.band_index == b.band_index && a.band_count == b.band_count;  // grouping is not part of the plan
}

}  // namespace

#ifdef SC_CHECKED
// checked build: address ranges the range-tested gathers accept: [0,1) the scan's integral buffer, [2,3) the parity hooks' one
static void checked_set_range(int which, const void* p, size_t bytes) {
    unsigned long long r[2] = {(unsigned long long)p, (unsigned long long)p + bytes};
    cudaMemcpyToSymbol(sck::sc_chk_range, r, sizeof(r), (size_t)which * 16);
}
#endif

struct sc_comm_state;   // NCCL exchange (sc_comm.inc)

struct sc_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    sc_comm_state* comm = nullptr;   // multi-GPU exchange (sc_comm_init)
    unsigned long long h2d_bytes = 0, d2h_bytes = 0;   // bytes the detect entry points copied between host and device (sc_transfer_bytes)
    int64_t launches = 0;
    int n_sms = 148;

    // cascade (host copy + device weights)
    bool have_cascade = false;
    int tmpl = 40, n_stages = 0, total_weak = 0;
    std::vector<float> theta;
    std::vector<int> n_weak, weak_base;
    std::vector<sc_rect> rects;
    std::vector<float> w;       // [total][33]
    std::vector<double> bias;   // [total]
    DevBuf d_w, d_wb;

    // plan (one cached entry)
    bool have_plan = false;
    sc_detect_params pparams{};
    ScPlan plan{};
    DevBuf d_plan, d_geom;
    // stage-0 certified fast filter (sc_plan.h ScFastParams): one parameter block per column parity
    bool allow_fast = true;     // SC_DISABLE_FAST=1 in the environment forces the exact-only kernel (A/B tests)
    bool use_fast = false;
    ScFastParams fast[2];
    // compact integer plane (sc_plan.h): distinct projected cell edges of the stage-0 patches that need a per-frame bound
    bool allow_compact = true;  // SC_DISABLE_COMPACT=1 in the environment keeps the filter on the float planes (A/B tests)
    std::vector<int> cert_ce;
    DevBuf d_cert_items, d_cert;   // [n_items] cell edges; [int_frames][n_items] bounds (k_cell_bounds)
    bool group_attr_set = false;  // k_group_frames' dynamic shared-memory limit raised on this handle's device
    int group_max = 8;  // frames per scan group (SC_GROUP_FRAMES overrides, 1..32)

    // group buffers
    int group_frames = 0;       // frames one scan group holds (records, bitmasks)
    int int_frames = 0;         // frames one integral super-group holds (images, carries, integral images)
    uint32_t rec_cap = 0;
    DevBuf d_img, d_carry, d_S, d_counters, d_det, d_detcount;
    // Scan groups alternate between two lanes (stream + private bitmasks / records) so that one group's small tail
    // kernels (later stages, replay, finalize) overlap the next group's stage 0.
    struct Lane {
        cudaStream_t st = nullptr;
        cudaEvent_t done = nullptr;
        DevBuf d_multi, d_pass, d_visited, d_start, d_rec, d_idx[2], d_small, d_chunks;  // d_chunks: work list of k_scan_odd (32-window runs of reachable odd columns)
    };
    Lane lanes[4];
    int lanes_max = 4;  // scan groups in flight, one stream each (SC_LANES overrides, 1..4; C2: 2 -> 3068, 4 -> 3095 frames/s)
    int n_lanes = 0;
    cudaEvent_t ev_integral = nullptr;
    std::vector<cudaEvent_t> ev_chunk;   // per pipeline chunk: [2c] upload done, [2c+1] integral done
    cudaStream_t copy_st = nullptr;      // host-frame uploads (sc_detect)
    HostBuf h_stage;
    std::vector<sc_counters> last_counters;
    int last_nframes = 0;
    // sc_detect_submit / sc_detect_collect: two batches in flight, so the upload of batch k+1 and the download / host work
    // of batch k-1 hide under the compute of batch k (the compute itself is serialised by the shared integral buffers)
    struct Ticket {
        bool busy = false;
        int nframes = 0;
        uint32_t det_cap = 0, eager = 0;
        DevBuf d_img, d_det, d_cnt, d_counters, d_grp;  // d_grp: segments, per-frame tables and output of the device grouping
        HostBuf h_out;  // [counters | count | first `eager` detections]  (grouped: [counters | count, overflow | first `eager` objects])
        int group_thr = 0;
        double group_eps = 0.2;
        cudaEvent_t done = nullptr;
    };
    Ticket tickets[2];

    // single-frame state for the parity hooks
    bool have_integral = false;
    int cur_W = 0, cur_H = 0;
    ScLayout hook_lay{};
    DevBuf d_hook_img, d_hook_carry, d_hook_S;
    DevBuf d_pool_w, d_pool_wb, d_pool_auc, d_pool_x, d_pool_aux;  // training-side pool evaluation
    DevBuf d_ext_img, d_ext_geom, d_ext_X;   // training-side descriptor extraction

    // optional per-kernel timing with CUDA events on the handle's stream (bench.py's roofline leg)
    bool profiling = false;
    struct Span { int kid; cudaEvent_t a, b; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> event_pool;
    double kernel_ms[16] = {0};
    int64_t kernel_launches[16] = {0};
};

namespace {

// d_small layout (uint32): [0] rec_count, [1..16] per-stage index-list counts, [17] det_count
enum { SM_REC = 0, SM_STAGE0 = 1, SM_DET = 17, SM_CHUNKS = 18, SM_CURSOR = 19, SM_WORDS = 32 };

bool any_ticket_busy(const sc_handle* h) { return h->tickets[0].busy || h->tickets[1].busy; }

int fail(sc_handle* h, int code, const std::string& msg) {
    if (h) h->err = msg;
    return code;
}

int cuda_fail(sc_handle* h, cudaError_t e, const char* what) {
    return fail(h, SC_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define SC_CUDA(h, call)                                         \
    do {                                                         \
        cudaError_t e_ = (call);                                 \
        if (e_ != cudaSuccess) return cuda_fail((h), e_, #call); \
    } while (0)

enum { K_CARRY = 0, K_WALK, K_STAGE0, K_STAGE, K_REPLAY, K_FINALIZE, K_EVENTS, K_POOL, K_STAGE0_ODD, K_GROUP, K_POOLFEAT, K_CERT, K_COUNT };
const char* const kKernelNames[K_COUNT] = {"k_strip_carry", "k_integral_walk", "k_scan_stage0", "k_scan_stage", "k_replay_rows", "k_finalize",
                                            "k_row_events", "k_pool_hist", "k_scan_stage0_odd", "k_group_frames", "k_pool_features", "k_cell_bounds"};

cudaEvent_t take_event(sc_handle* h) {
    cudaEvent_t e = nullptr;
    if (!h->event_pool.empty()) { e = h->event_pool.back(); h->event_pool.pop_back(); }
    else cudaEventCreate(&e);
    return e;
}

struct KernelSpan {
    sc_handle* h;
    int kid;
    cudaEvent_t a = nullptr;
    cudaStream_t st;
    KernelSpan(sc_handle* h_, int kid_, cudaStream_t st_ = nullptr) : h(h_), kid(kid_), st(st_ ? st_ : h_->stream) {
        h->launches++;
        if (h->profiling) { a = take_event(h); cudaEventRecord(a, st); }
    }
    ~KernelSpan() {
        if (a) {
            cudaEvent_t b = take_event(h);
            cudaEventRecord(b, st);
            h->spans.push_back(sc_handle::Span{kid, a, b});
        }
    }
};

// call after the stream has been synchronised
void drain_spans(sc_handle* h) {
    for (auto& sp : h->spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) { h->kernel_ms[sp.kid] += ms; h->kernel_launches[sp.kid]++; }
        h->event_pool.push_back(sp.a);
        h->event_pool.push_back(sp.b);
    }
    h->spans.clear();
}

sc_detect_params default_params() {
    sc_detect_params p;
    p.base = 40; p.step = 0; p.scale = 1.1; p.prefilter = 6; p.skip_rule = 1; p.force_all_stages = 0; p.band_index = 0; p.band_count = 0; p.group_threshold = 0; p.reserved = 0; p.group_eps = 0.2;
    return p;
}

// Distance budget between the fast filter's sum of weak outputs and the reference arithmetic's, stage 0
// (sc_kernels.cuh "Error budget", doubled), plus the float summations and the division by n on both sides.
double fast_margin(const sc_handle* h, int stage = 0) {
    const double u = 5.9604644775390625e-8;  // 2^-24
    const int n0 = h->n_weak[stage], wb = h->weak_base[stage];
    double margin = 0.0;
    for (int q = 0; q < n0; q++) {
        double n2 = 0.0;
        for (int i = 0; i < 32; i++) n2 += (double)h->w[(size_t)(wb + q) * 33 + i] * h->w[(size_t)(wb + q) * 33 + i];
        margin += 2.0 * (0.25 * 121.0 * u * std::sqrt(n2) + 1.5e-6);
    }
    return margin + (double)n0 * n0 * 4.0 * u + 8.0 * u * n0;
}

// Compact-plane certification item of one projected stage-0 patch (sc_kernels.cuh, k_cell_bounds): its cells are ce x ce
// (GetRectsFromPatch, DenseSURFFeatureExtractor.cpp:360-377).  255 ce^2 < 65536 needs no frame-dependent bound.
uint32_t cert_item(sc_handle* h, int l, const sc_rect& patch) {
    if (!h->allow_compact) return SC_CERT_NEVER;
    const sc_rect r = sc_host::project_patch(h->tmpl, l, patch);
    const int ce = r.w == r.h ? r.w / 2 : std::min(r.w, r.h);
    if (ce < 1) return SC_CERT_NEVER;
    if (255LL * ce * ce < (long long)SC_CELL_LIMIT) return SC_CERT_ALWAYS;
    for (size_t i = 0; i < h->cert_ce.size(); i++)
        if (h->cert_ce[i] == ce) return (uint32_t)i;
    if ((int)h->cert_ce.size() >= SC_CERT_MAX_ITEMS) return SC_CERT_NEVER;
    h->cert_ce.push_back(ce);
    return (uint32_t)(h->cert_ce.size() - 1);
}

// Scale ladder, lattice and per-(scale, weak) projected geometry.  Host arithmetic mirrors ObjDetector.cpp:139,174,180
// and ProjectPatches / GetRectsFromPatch (DenseSURFFeatureExtractor.cpp:486-508, 360-377) operation for operation.
int build_plan(sc_handle* h, int W, int H, const sc_detect_params& prm, std::vector<ScGeom>* geom_out) {
    if (!h->have_cascade) return fail(h, SC_ERR_STATE, "no cascade loaded (sc_set_cascade / sc_load_model)");
    if (W < 2 || H < 2 || prm.base < 1 || !(prm.scale > 1.0)) return fail(h, SC_ERR_INVALID, "bad frame size or scan parameters");
    const int bands = prm.band_count > 1 ? prm.band_count : 1, band = bands > 1 ? prm.band_index : 0;
    if (band < 0 || band >= bands) return fail(h, SC_ERR_INVALID, "band_index outside 0..band_count-1");
    ScPlan& p = h->plan;
    memset(&p, 0, sizeof(p));
    p.W = W; p.H = H;
    p.step = prm.step > 0 ? prm.step : (prm.base > 20 ? prm.base / 20 : 1);
    p.lay = sc_host::make_layout(W, H, 2 * p.step, p.step);
    if (p.lay.hp > 4096 || p.lay.frame4 * 16 > 0xffffffffLL) return fail(h, SC_ERR_INVALID, "frame too large for the scan layout (plane rows of at most 4096 pixels, 4 GiB per frame)");
    p.n_stages = h->n_stages; p.total_weak = h->total_weak;
    p.use_prefilter = prm.prefilter >= 0; p.skip_rule = prm.skip_rule != 0; p.force_all = prm.force_all_stages != 0;
    p.n_strips = (W + SC_STRIP - 1) / SC_STRIP;
    for (int s = 0; s < h->n_stages; s++) { p.theta[s] = h->theta[s]; p.n_weak[s] = h->n_weak[s]; p.weak_base[s] = h->weak_base[s]; }

    std::vector<int> sides;
    sc_host::scale_ladder(W, H, prm.base, prm.scale, &sides);
    if ((int)sides.size() > SC_PLAN_MAX_SCALES) return fail(h, SC_ERR_INVALID, "too many scales (max 64)");
    int nsc = 0, blocks = 0, words = 0, rows = 0;
    long long windows = 0;
    // the fast-filter variant of the stage-0 kernel is used whenever it applies (below); its tiles are taller (sc_plan.h)
    const bool fast_plan = h->allow_fast && !p.force_all && h->n_stages > 0 && h->n_weak[0] <= SC_F_MAXW;
    const int tile_y = fast_plan ? SC_TILE_Y_FAST : SC_TILE_Y_EXACT;
    for (size_t i = 0; i < sides.size(); i++) {
        const int l = sides[i];
        if (l > W || l > H) continue;
        ScScale& s = p.sc[nsc];
        s.l = l;
        s.nx = (W - l) / p.step + 1; s.ny = (H - l) / p.step + 1;
        if (s.nx > 65535 || s.ny > 65535) return fail(h, SC_ERR_INVALID, "lattice exceeds 65535 positions per axis");
        // row band of a multi-GPU split: rows [ny * band / bands, ny * (band + 1) / bands) of every scale
        s.gy0 = (int)((long long)s.ny * band / bands);
        s.ny = (int)((long long)s.ny * (band + 1) / bands) - s.gy0;
        s.wpr = (s.nx + 31) / 32;
        s.thr = (float)(l * l * (prm.prefilter >= 0 ? prm.prefilter : 0));
        s.tiles_x = ((s.nx + 1) / 2 + SC_TILE_X - 1) / SC_TILE_X;  // tiles of 64 same-parity columns
        const int tiles_y = (s.ny + tile_y - 1) / tile_y;
        s.block_base = blocks; s.word_base = words; s.row_base = rows;
        for (int ph = 0; ph < 2; ph++) {
            const int x0 = ph * p.step;
            s.pf[ph][0] = (uint32_t)(16 * sc_layout_index(p.lay, x0, 0));
            s.pf[ph][1] = (uint32_t)(16 * sc_layout_index(p.lay, x0 + l, 0));
            s.pf[ph][2] = (uint32_t)(16 * sc_layout_index(p.lay, x0, l));
            s.pf[ph][3] = (uint32_t)(16 * sc_layout_index(p.lay, x0 + l, l));
        }
        blocks += s.tiles_x * tiles_y; words += s.wpr * s.ny; rows += s.ny;
        windows += (long long)s.nx * s.ny;
        nsc++;
    }
    p.n_scales = nsc; p.blocks_per_frame = blocks; p.words_per_frame = words; p.rows_per_frame = rows; p.windows_per_frame = windows;

    ScGeom zero_geom;
    memset(&zero_geom, 0, sizeof(zero_geom));
    geom_out->assign((size_t)2 * std::max(nsc, 1) * h->total_weak, zero_geom);  // [column parity][scale][weak]
    for (int ph = 0; ph < 2; ph++)
        for (int i = 0; i < nsc; i++)
            for (int k = 0; k < h->total_weak; k++)
                if (!sc_host::project_geom(h->tmpl, p.sc[i].l, h->rects[k], p.lay, ph * p.step,
                                           &(*geom_out)[((size_t)ph * nsc + i) * h->total_weak + k]))
                    return fail(h, SC_ERR_INVALID, "weak classifier patch is not 2x2 / 4x1 / 1x4 cells after projection");

    // Certified fast arithmetic of k_scan_stage (sc_kernels.cuh): the same limits for every stage p, on the float sum of its
    // fast weak outputs.  Rejected at stage p: multi == 2  <=>  (score + p + 1) / N < 0.5  <=>  score < N / 2 - p - 1.
    p.stage_fast = (h->allow_fast && !p.force_all) ? 1 : 0;
    for (int sgi = 0; sgi < h->n_stages; sgi++) {
        const double m = fast_margin(h, sgi), ns = (double)h->n_weak[sgi], tau_s = 0.5 * h->n_stages - sgi - 1.0;
        p.fl_reject[sgi] = std::nextafterf((float)(ns * h->theta[sgi] - m), -INFINITY);
        p.fl_pass[sgi] = std::nextafterf((float)(ns * h->theta[sgi] + m), INFINITY);
        p.fl_skip[sgi] = std::nextafterf((float)(ns * tau_s - m), -INFINITY);
        p.fl_noskip[sgi] = std::nextafterf((float)(ns * tau_s + m), INFINITY);
    }
    // Certified fast filter for stage 0 (sc_kernels.cuh, "Error budget").  Limits are on the float sum of the fast weak outputs.
    const int n0 = h->n_weak[0];
    h->use_fast = fast_plan && nsc >= 1;
    h->cert_ce.clear();
    if (h->use_fast) {
        const double margin = fast_margin(h);
        const double tau = 0.5 * h->n_stages - 1.0;           // rejected at stage 0: multi == 2  <=>  score < tau (ObjDetector.cpp:201,214)
        for (int ph = 0; ph < 2; ph++) {
            ScFastParams& fp = h->fast[ph];
            memset(&fp, 0, sizeof(fp));
            fp.n_weak = n0; fp.n_scales = nsc; fp.blocks_per_frame = std::max(blocks, 1);
            fp.lim_reject = std::nextafterf((float)((double)n0 * h->theta[0] - margin), -INFINITY);
            fp.lim_skip = std::nextafterf((float)((double)n0 * tau - margin), -INFINITY);
            fp.lim_noskip = std::nextafterf((float)((double)n0 * tau + margin), INFINITY);
            fp.lim_pass = h->n_stages > 1 ? std::nextafterf((float)((double)n0 * h->theta[0] + margin), INFINITY) : INFINITY;
            for (int q = 0; q < n0; q++) {
                fp.wb[q] = (float)((double)h->w[(size_t)q * 33 + 32] * h->bias[q]);
                memcpy(fp.w[q], &h->w[(size_t)q * 33], 32 * sizeof(float));
            }
            for (int i = 0; i < nsc; i++) {
                fp.block_base[i] = p.sc[i].block_base;
                fp.row_base[i] = p.sc[i].row_base;
                for (int q = 0; q < n0; q++) {
                    const ScGeom& g = (*geom_out)[((size_t)ph * nsc + i) * h->total_weak + q];
                    for (int k = 0; k < 10; k++) fp.geom[i][q][k] = g.c[k];
                    fp.geom[i][q][10] = (uint32_t)g.shape;
                    fp.geom[i][q][11] = cert_item(h, p.sc[i].l, h->rects[q]);
                }
            }
        }
    }
    return SC_OK;
}

int ensure_plan(sc_handle* h, int W, int H, const sc_detect_params& prm) {
    if (h->have_plan && h->plan.W == W && h->plan.H == H && same_params(prm, h->pparams)) return SC_OK;
    h->have_plan = false;
    std::vector<ScGeom> geom;
    int rc = build_plan(h, W, H, prm, &geom);
    if (rc != SC_OK) return rc;
    SC_CUDA(h, h->d_plan.ensure(sizeof(ScPlan)));
    SC_CUDA(h, h->d_geom.ensure(std::max<size_t>(geom.size(), 1) * sizeof(ScGeom)));
    // synchronous copies: the plan is rebuilt only when the frame size or parameters change
    SC_CUDA(h, cudaStreamSynchronize(h->stream));
    SC_CUDA(h, cudaMemcpy(h->d_plan.p, &h->plan, sizeof(ScPlan), cudaMemcpyHostToDevice));
    if (!geom.empty()) SC_CUDA(h, cudaMemcpy(h->d_geom.p, geom.data(), geom.size() * sizeof(ScGeom), cudaMemcpyHostToDevice));
    SC_CUDA(h, h->d_cert_items.ensure(std::max<size_t>(h->cert_ce.size(), 1) * sizeof(int)));
    if (!h->cert_ce.empty()) SC_CUDA(h, cudaMemcpy(h->d_cert_items.p, h->cert_ce.data(), h->cert_ce.size() * sizeof(int), cudaMemcpyHostToDevice));
    h->pparams = prm;
    h->have_plan = true;
    h->group_frames = 0;  // buffers are re-sized for the new plan
    h->int_frames = 0;
    h->n_lanes = 0;       // and the lane count is chosen from the new plan alone (a C4-sized plan must not inherit 4 lanes)
    return SC_OK;
}

size_t align256(size_t v) { return (v + 255) / 256 * 256; }

int ensure_group_buffers(sc_handle* h, int want_frames, bool own_images) {
    const ScPlan& p = h->plan;
    const size_t per_rec = sizeof(ScRecord) + 2 * sizeof(uint32_t);
    const size_t per_frame = (size_t)p.lay.frame4 * 16 + (size_t)p.windows_per_frame * per_rec + (size_t)p.words_per_frame * 12 +
                             (size_t)p.H * p.n_strips * 32 + (size_t)p.W * p.H;
    int g = (int)std::max<size_t>(1, std::min<size_t>((size_t)h->group_max, ((size_t)6 << 30) / std::max<size_t>(per_frame, 1)));
    g = std::min(g, std::max(want_frames, 1));
    // k_scan_odd's work list packs (row index of the group) << 10 | unit: at most 2^22 lattice rows per scan group
    // (units per row <= ceil(32768 / SC_ODD_UNIT) = 256 < 1024 because a scale has at most 65535 columns)
    if ((long long)p.rows_per_frame >= (1LL << 22)) return fail(h, SC_ERR_INVALID, "too many lattice rows per frame for the scan's work list");
    while (g > 1 && (long long)g * p.rows_per_frame >= (1LL << 22)) g--;
    // the integral stage runs on a larger super-group (its warps walk rows sequentially and need many frames in flight)
    const size_t int_frame = (size_t)p.lay.frame4 * 16 + (size_t)p.H * p.n_strips * 32 + (size_t)p.W * p.H;
    int gi = (int)std::max<size_t>(1, std::min<size_t>(32, ((size_t)4 << 30) / std::max<size_t>(int_frame, 1)));
    gi = std::max(g, std::min(gi, std::max(want_frames, 1)));
    g = std::max(g, h->group_frames);  // (group_frames obeyed the same limits when it was chosen for this plan)
    gi = std::max(gi, h->int_frames);
    const unsigned long long recs = (unsigned long long)p.windows_per_frame * g;
    if (recs > 0xfffffff0ull) return fail(h, SC_ERR_INVALID, "too many windows per group");
    const int want_lanes = std::max(h->n_lanes, (want_frames > g && recs * per_rec < ((size_t)3 << 30)) ? std::min(h->lanes_max, (want_frames + g - 1) / g) : 1);
    if (g <= h->group_frames && gi <= h->int_frames && want_lanes <= h->n_lanes) return SC_OK;
    if (own_images) SC_CUDA(h, h->d_img.ensure(align256((size_t)gi * p.W * p.H)));
    SC_CUDA(h, h->d_carry.ensure(align256((size_t)gi * p.H * p.n_strips * 32)));
    SC_CUDA(h, h->d_S.ensure((size_t)gi * p.lay.frame4 * 16));
#ifdef SC_CHECKED
    checked_set_range(0, h->d_S.p, h->d_S.cap);
#endif
    SC_CUDA(h, h->d_cert.ensure(std::max<size_t>((size_t)gi * h->cert_ce.size(), 1) * 4));
    // a second lane only when more than one scan group can be in flight and the worst-case record arrays stay modest
    h->n_lanes = want_lanes;
    for (int li = 0; li < h->n_lanes; li++) {
        sc_handle::Lane& L = h->lanes[li];
        if (!L.st) SC_CUDA(h, cudaStreamCreateWithFlags(&L.st, cudaStreamNonBlocking));
        if (!L.done) SC_CUDA(h, cudaEventCreateWithFlags(&L.done, cudaEventDisableTiming));
        SC_CUDA(h, L.d_multi.ensure(align256((size_t)g * p.words_per_frame * 4 + 4)));
        SC_CUDA(h, L.d_pass.ensure(align256((size_t)g * p.words_per_frame * 4 + 4)));
        SC_CUDA(h, L.d_visited.ensure(align256((size_t)g * p.words_per_frame * 4 + 4)));
        SC_CUDA(h, L.d_start.ensure(align256((size_t)g * p.rows_per_frame * 4 + 4)));
        SC_CUDA(h, L.d_rec.ensure(align256(std::max<size_t>(recs, 1) * sizeof(ScRecord))));
        SC_CUDA(h, L.d_idx[0].ensure(align256(std::max<size_t>(recs, 1) * 4)));
        SC_CUDA(h, L.d_idx[1].ensure(align256(std::max<size_t>(recs, 1) * 4)));
        SC_CUDA(h, L.d_small.ensure(SM_WORDS * 4));
        SC_CUDA(h, L.d_chunks.ensure(align256(((size_t)g * (p.windows_per_frame / 64 + 2 * (size_t)p.rows_per_frame) + 1024) * 4)));
    }
    if (!h->ev_integral) SC_CUDA(h, cudaEventCreateWithFlags(&h->ev_integral, cudaEventDisableTiming));
    h->rec_cap = (uint32_t)recs;
    h->group_frames = g;
    h->int_frames = gi;
    return SC_OK;
}

// Channels + integral images of `n` frames (device images) into frame slots slot0.. of d_S, on the main stream
// `under_scan`: this launch will share the SMs with a scan group of the previous chunk (synchronous host path).  The tiled
// walk holds 40 KB of shared memory per CTA; next to four stage-0 CTAs that exceeds the scan's shared-memory carve-out, and
// the synchronous path dropped from 2 720 to 2 250 frames/s when its overlapped chunks used it -- those chunks keep the
// shuffle-scan form, which needs no shared memory.  Both forms produce identical bits.
int run_integral(sc_handle* h, const uint8_t* d_img, int n, int slot0, bool under_scan = false) {
    const ScPlan& p = h->plan;
    cudaStream_t st = h->stream;
    const int rows = n * p.H;
    int* carry = h->d_carry.as<int>() + (size_t)slot0 * p.H * p.n_strips * 8;
    float4* S = h->d_S.as<float4>() + (size_t)slot0 * p.lay.frame4;
    { KernelSpan ks(h, K_CARRY); sck::k_strip_carry<<<(rows + 3) / 4, 128, 0, st>>>(d_img, p.W, p.H, p.n_strips, n, carry); }
    const int warps = n * p.n_strips;
    {
        KernelSpan ks(h, K_WALK);
        if (SC_WALK_TILED && !under_scan)
            sck::k_integral_walk_tiled<<<(warps + 3) / 4, 128, 0, st>>>(d_img, p.W, p.H, p.n_strips, n, carry, S, p.lay);
        else
            sck::k_integral_walk<<<(warps + 3) / 4, 128, 0, st>>>(d_img, p.W, p.H, p.n_strips, n, carry, S, p.lay);
    }
    const int n_items = (int)h->cert_ce.size();
    if (h->use_fast && n_items > 0) {
        // per-frame cell-sum bounds that certify the compact plane for this batch's stage-0 filter (k_cell_bounds)
        uint32_t* cert = h->d_cert.as<uint32_t>() + (size_t)slot0 * n_items;
        SC_CUDA(h, cudaMemsetAsync(cert, 0, (size_t)n * n_items * 4, st));
#ifndef SC_CERT_CHUNKS
#define SC_CERT_CHUNKS 4   // CTAs per (frame, cell edge): 4 -> 0.0098, 16 -> see DESIGN.md 6b
#endif
        const int chunks = SC_CERT_CHUNKS;
        KernelSpan ks(h, K_CERT);
        sck::k_cell_bounds<<<dim3(n_items * chunks, n), 256, 0, st>>>(S, p.lay, p.W, p.H, h->d_cert_items.as<int>(), n_items, chunks, cert);
    }
    SC_CUDA(h, cudaGetLastError());
    return SC_OK;
}

// Launches the whole path for `g` frames already in d_img (device).  Detections are appended to det / det_count.
// The scan kernels are instantiated per half-row distance of the layout (sc_plan.h): 256 .. 4096 float4.
template <bool FAST, bool ALL, int NW, typename... A>
void launch_stage0_hp(int hp, int grid, size_t smem, cudaStream_t st, A... a) {
    switch (hp) {
        case 256: sck::k_scan_stage0<256, FAST, ALL, NW><<<grid, SC_TILE_THREADS, smem, st>>>(a...); break;
        case 512: sck::k_scan_stage0<512, FAST, ALL, NW><<<grid, SC_TILE_THREADS, smem, st>>>(a...); break;
        case 1024: sck::k_scan_stage0<1024, FAST, ALL, NW><<<grid, SC_TILE_THREADS, smem, st>>>(a...); break;
        case 2048: sck::k_scan_stage0<2048, FAST, ALL, NW><<<grid, SC_TILE_THREADS, smem, st>>>(a...); break;
        default: sck::k_scan_stage0<4096, FAST, ALL, NW><<<grid, SC_TILE_THREADS, smem, st>>>(a...); break;
    }
}
// fast: certified fast filter (nw0 = weak classifiers of stage 0: 2 and 3 have unrolled instantiations);
// all: exact variant that also runs stages 1..N-1 in place (force_all)
template <typename... A>
void launch_stage0(bool fast, bool all, int nw0, int hp, int grid, size_t smem, cudaStream_t st, A... a) {
    if (fast && nw0 == 3) launch_stage0_hp<true, false, 3>(hp, grid, smem, st, a...);
    else if (fast && nw0 == 2) launch_stage0_hp<true, false, 2>(hp, grid, smem, st, a...);
    else if (fast) launch_stage0_hp<true, false, 0>(hp, grid, smem, st, a...);
    else if (all) launch_stage0_hp<false, true, 0>(hp, grid, smem, st, a...);
    else launch_stage0_hp<false, false, 0>(hp, grid, smem, st, a...);
}
template <typename... A>
void launch_stage(int hp, int grid, size_t smem, cudaStream_t st, A... a) {
    switch (hp) {
        case 256: sck::k_scan_stage<256><<<grid, 128, smem, st>>>(a...); break;
        case 512: sck::k_scan_stage<512><<<grid, 128, smem, st>>>(a...); break;
        case 1024: sck::k_scan_stage<1024><<<grid, 128, smem, st>>>(a...); break;
        case 2048: sck::k_scan_stage<2048><<<grid, 128, smem, st>>>(a...); break;
        default: sck::k_scan_stage<4096><<<grid, 128, smem, st>>>(a...); break;
    }
}

// Scan of `g` frames whose integrals start at frame slot `s0` of d_S.
int run_group(sc_handle* h, sc_handle::Lane& L, int s0, int g, int frame0, sc_detection* d_det, uint32_t det_cap, uint32_t* d_det_count,
              unsigned long long* d_counters) {
    const ScPlan& p = h->plan;
    const ScPlan* dp = h->d_plan.as<ScPlan>();
    cudaStream_t st = L.st;
    uint32_t* small = L.d_small.as<uint32_t>();
    SC_CUDA(h, cudaMemsetAsync(small, 0, (SM_DET) * 4, st));  // rec + stage counts; det count is the caller's
    float4* S = h->d_S.as<float4>() + (size_t)s0 * p.lay.frame4;
    // (a row band of a tiny frame can be empty: band_count larger than a scale's lattice rows -- nothing to launch then)
    if (p.n_scales > 0 && p.n_stages > 0 && p.blocks_per_frame > 0 && p.rows_per_frame > 0) {
        const ScGeom* geom = h->d_geom.as<ScGeom>();
        const float* w = h->d_w.as<float>();
        const double* wb = h->d_wb.as<double>();
        uint32_t* multi = L.d_multi.as<uint32_t>();
        uint32_t* pass = L.d_pass.as<uint32_t>();
        uint32_t* visited = L.d_visited.as<uint32_t>();
        ScRecord* rec = L.d_rec.as<ScRecord>();
        const int n_items = (int)h->cert_ce.size();
        const uint32_t* cert = (h->use_fast && h->allow_compact) ? h->d_cert.as<uint32_t>() + (size_t)s0 * n_items : nullptr;
        // force_all: every stage of every window is evaluated inside the stage-0 tile kernel (its ALL variant), as long as
        // the whole cascade's weights and geometry fit its shared memory (200 B per weak classifier)
        const bool all_in_tile = p.force_all && !h->use_fast && p.n_stages > 1 && p.total_weak <= 200;
        {
            // even lattice columns everywhere, then the odd columns the reference's stride can reach
            const size_t smem = (size_t)(all_in_tile ? p.total_weak : p.n_weak[0]) * (SC_W_PITCH * 4 + 8 + sizeof(ScGeom));
            int* start_odd = L.d_start.as<int>();
            const int rows0 = g * p.rows_per_frame;
            {
                KernelSpan ks(h, K_STAGE0, st);
                launch_stage0(h->use_fast, all_in_tile, p.n_weak[0], p.lay.hp, g * p.blocks_per_frame, smem, st, h->fast[0], dp, S, geom, w, wb, multi, pass, rec, small + SM_REC, h->rec_cap, 0, start_odd, cert, n_items);
            }
            // odd columns: with the adaptive stride they are reachable only as ragged row suffixes -> list of 32-window runs
            // and persistent warps (k_scan_odd); without it (or without the fast filter) the tile kernel does them all
            const bool odd_list = p.skip_rule && h->use_fast;
            if (p.skip_rule) {
                if (odd_list) SC_CUDA(h, cudaMemsetAsync(small + SM_CHUNKS, 0, 8, st));
                KernelSpan ks(h, K_EVENTS, st);
                // d_chunks: [rows0] per-row run counts -> offsets | run list
                uint32_t* row_chunks = L.d_chunks.as<uint32_t>();
                sck::k_row_events<<<(rows0 + 127) / 128, 128, 0, st>>>(dp, g, multi, start_odd, d_counters, odd_list ? row_chunks : nullptr);
                if (odd_list) {
                    const int nb = (rows0 + 1023) / 1024;
                    uint32_t* blk = row_chunks + rows0;           // [nb] block sums
                    sck::k_chunk_blocksum<<<nb, 1024, 0, st>>>(row_chunks, rows0, blk);
                    sck::k_chunk_fill<<<nb, 1024, 0, st>>>(row_chunks, rows0, blk, blk + nb, small + SM_CHUNKS);
                    h->launches += 2;
                }
            } else {
                SC_CUDA(h, cudaMemsetAsync(start_odd, 0, (size_t)rows0 * 4, st));  // every odd column is visited
            }
            KernelSpan ks(h, K_STAGE0_ODD, st);
            if (odd_list) {
                const int grid = h->n_sms * SC_STAGE0_MIN_CTAS;
                switch (p.lay.hp * 8 + (p.n_weak[0] == 2 || p.n_weak[0] == 3 ? p.n_weak[0] : 0)) {
#define SC_ODD(HPV, NWV) case HPV * 8 + NWV: sck::k_scan_odd<HPV, NWV><<<grid, 256, 0, st>>>(h->fast[1], dp, S, geom, w, wb, multi, pass, rec, small + SM_REC, h->rec_cap, start_odd, \
                                                            L.d_chunks.as<uint32_t>() + rows0 + (rows0 + 1023) / 1024, small + SM_CHUNKS, small + SM_CURSOR, cert, n_items); break
#define SC_ODD3(HPV) SC_ODD(HPV, 0); SC_ODD(HPV, 2); SC_ODD(HPV, 3)
                    SC_ODD3(256); SC_ODD3(512); SC_ODD3(1024); SC_ODD3(2048); SC_ODD3(4096);
#undef SC_ODD3
#undef SC_ODD
                }
            } else {
                launch_stage0(h->use_fast, all_in_tile, p.n_weak[0], p.lay.hp, g * p.blocks_per_frame, smem, st, h->fast[1], dp, S, geom, w, wb, multi, pass, rec, small + SM_REC, h->rec_cap, 1, start_odd, cert, n_items);
            }
        }
        const int tail_grid = h->n_sms * 8;
        // with the fast filter, stage 0's undecided windows are live records: k_scan_stage(0) gives them the exact arithmetic
        const int s_begin = h->use_fast ? 0 : 1;
        for (int s = s_begin; s < p.n_stages && !all_in_tile; s++) {
            const bool first = s == s_begin || p.force_all;
            const uint32_t* in_idx = first ? nullptr : L.d_idx[(s - 1) & 1].as<uint32_t>();
            const uint32_t* in_cnt = first ? small + SM_REC : small + SM_STAGE0 + (s - 1);
            const size_t smem = (size_t)p.n_weak[s] * (SC_W_PITCH * 4 + 8);
            KernelSpan ks(h, K_STAGE, st);
            launch_stage(p.lay.hp, tail_grid, smem, st, dp, s, S, geom, w, wb, multi, rec, in_idx, in_cnt, L.d_idx[s & 1].as<uint32_t>(),
                         small + SM_STAGE0 + s, h->rec_cap);
        }
        const int rows = g * p.rows_per_frame;
        { KernelSpan ks(h, K_REPLAY, st); sck::k_replay_rows<<<(rows + 127) / 128, 128, 0, st>>>(dp, g, multi, pass, visited, d_counters); }
        { KernelSpan ks(h, K_FINALIZE, st);
          sck::k_finalize<<<h->n_sms * 4, 128, 0, st>>>(dp, rec, small + SM_REC, h->rec_cap, visited, d_counters,
                                                         reinterpret_cast<sck::ScDetOut*>(d_det), d_det_count, det_cap, frame0); }
    }
    SC_CUDA(h, cudaGetLastError());
    return SC_OK;
}

// All scan groups of one integral super-group (ni frames at slots 0..ni-1 of d_S), alternating over the lanes.  The
// integral was enqueued on the main stream; the lanes wait for it and the main stream waits for the lanes.
//
// One super-group of ni frames as a three-deep pipeline over chunks of SC_ICHUNK frames:
//   copy stream : host frames -> d_img            (host path only; `frames` non-null)
//   main stream : channels + integral of chunk c   (waits for its copy)
//   lanes       : scan groups of chunk c           (wait for its integral)
// so the upload and the integral of chunk c+1 run under the scan of chunk c.
#define SC_ICHUNK 16   // measured: 8-frame chunks lose more in the integral kernel than the upload overlap gains
int run_supergroup(sc_handle* h, const uint8_t* const* frames, int stride, const uint8_t* d_frames, int ni, int frame0, sc_detection* d_det,
                   uint32_t det_cap, uint32_t* d_det_count, unsigned long long* d_counters, uint8_t* img_buf = nullptr, bool img_buf_busy = true,
                   bool uploads_hidden = false) {
    const ScPlan& p = h->plan;
    const int lanes = h->profiling ? 1 : h->n_lanes;  // per-kernel event timing wants the kernels back to back
    // device-resident frames: one integral launch over the whole super-group is fastest (nothing to overlap it with
    // but the scan, which it only slows down); host frames: chunks, so that uploads hide under the scan
    // (uploads_hidden: another batch is computing, so this batch's uploads already run under it -- sc_detect_submit)
    const int ich = (h->profiling || !frames || uploads_hidden) ? ni : SC_ICHUNK;
    const int nch = (ni + ich - 1) / ich;
    while ((int)h->ev_chunk.size() < 2 * nch) {
        cudaEvent_t e = nullptr;
        SC_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->ev_chunk.push_back(e);
    }
    if (frames && !h->copy_st) SC_CUDA(h, cudaStreamCreateWithFlags(&h->copy_st, cudaStreamNonBlocking));
    if (frames && img_buf_busy) {
        // the image buffer may still be read by the previous super-group's integral: order the copies after it
        SC_CUDA(h, cudaEventRecord(h->ev_integral, h->stream));
        SC_CUDA(h, cudaStreamWaitEvent(h->copy_st, h->ev_integral, 0));
    }
    uint8_t* const up = img_buf ? img_buf : h->d_img.as<uint8_t>();
    for (int c = 0; c < nch; c++) {
        const int c0 = c * ich, n = std::min(ich, ni - c0);
        const uint8_t* img = d_frames ? d_frames + (size_t)c0 * p.W * p.H : up + (size_t)c0 * p.W * p.H;
        if (frames) {
            for (int k = 0; k < n; k++)
                SC_CUDA(h, cudaMemcpy2DAsync(up + (size_t)(c0 + k) * p.W * p.H, p.W, frames[c0 + k], stride, p.W, p.H,
                                             cudaMemcpyHostToDevice, h->copy_st));
            h->h2d_bytes += (unsigned long long)n * p.W * p.H;
            SC_CUDA(h, cudaEventRecord(h->ev_chunk[2 * c], h->copy_st));
            SC_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_chunk[2 * c], 0));
        }
        int rc = run_integral(h, img, n, c0, c > 0);
        if (rc != SC_OK) return rc;
        SC_CUDA(h, cudaEventRecord(h->ev_chunk[2 * c + 1], h->stream));
    }
    int k = 0;
    for (int c = 0; c < nch; c++) {
        const int c0 = c * ich, n = std::min(ich, ni - c0);
        for (int li = 0; li < lanes; li++) SC_CUDA(h, cudaStreamWaitEvent(h->lanes[li].st, h->ev_chunk[2 * c + 1], 0));
        for (int f0 = c0; f0 < c0 + n; f0 += h->group_frames, k++) {
            const int g = std::min(h->group_frames, c0 + n - f0);
            int rc = run_group(h, h->lanes[k % lanes], f0, g, frame0 + f0, d_det, det_cap, d_det_count, d_counters + (size_t)f0 * SC_CNT_STRIDE);
            if (rc != SC_OK) return rc;
        }
    }
    for (int li = 0; li < lanes; li++) {
        SC_CUDA(h, cudaEventRecord(h->lanes[li].done, h->lanes[li].st));
        SC_CUDA(h, cudaStreamWaitEvent(h->stream, h->lanes[li].done, 0));
    }
    return SC_OK;
}

void fill_counters(const sc_handle* h, const unsigned long long* raw, int nframes, sc_counters* out) {
    const ScPlan& p = h->plan;
    for (int f = 0; f < nframes; f++) {
        const unsigned long long* c = raw + (size_t)f * SC_CNT_STRIDE;
        sc_counters& o = out[f];
        memset(&o, 0, sizeof(o));
        o.grid = p.windows_per_frame;
        if (p.skip_rule) {
            for (int i = 0; i < p.n_scales; i++) o.evaluated += (int64_t)((p.sc[i].nx + 1) / 2) * p.sc[i].ny;  // even columns
            o.evaluated += (int64_t)c[SC_CNT_EVALODD];                                                          // reachable odd columns
        } else {
            o.evaluated = p.windows_per_frame;
        }
        o.visited = (int64_t)c[SC_CNT_VISITED];
        o.prefilter_pass = (int64_t)c[SC_CNT_PREFILTER];
        o.raw = (int64_t)c[SC_CNT_RAW];
        o.reach[0] = o.prefilter_pass;
        for (int s = 1; s < p.n_stages && s < SC_MAX_STAGES; s++) o.reach[s] = (int64_t)c[SC_CNT_REACH0 + s];
        for (int s = 0; s < p.n_stages && s < SC_MAX_STAGES; s++) o.weak_evals += o.reach[s] * p.n_weak[s];
    }
}

bool det_less(const sc_detection& a, const sc_detection& b) {
    if (a.frame != b.frame) return a.frame < b.frame;
    if (a.l != b.l) return a.l < b.l;
    if (a.y != b.y) return a.y < b.y;
    return a.x < b.x;
}

// Order detections by (frame, l, y, x): LSD radix sort on a packed 64-bit key (16 bits per field; frame is taken
// modulo 65536 per pass, larger batches fall back to a comparison sort).
void sort_detections(sc_detection* d, size_t n) {
    if (n < 2) return;
    bool packable = true;
    for (size_t i = 0; i < n && packable; i++)
        packable = (uint32_t)d[i].frame < 65536u && (uint32_t)d[i].l < 65536u && (uint32_t)d[i].y < 65536u && (uint32_t)d[i].x < 65536u;
    if (!packable || n < 64) { std::sort(d, d + n, det_less); return; }
    std::vector<sc_detection> tmp(n);
    std::vector<uint32_t> count(65536);
    sc_detection *src = d, *dst = tmp.data();
    for (int pass = 0; pass < 4; pass++) {  // x, y, l, frame
        std::fill(count.begin(), count.end(), 0u);
        auto digit = [pass](const sc_detection& e) -> uint32_t { return (uint32_t)(pass == 0 ? e.x : pass == 1 ? e.y : pass == 2 ? e.l : e.frame); };
        for (size_t i = 0; i < n; i++) count[digit(src[i])]++;
        uint32_t run = 0;
        for (uint32_t& c : count) { const uint32_t k = c; c = run; run += k; }
        for (size_t i = 0; i < n; i++) dst[count[digit(src[i])]++] = src[i];
        std::swap(src, dst);
    }
    // four passes: the result is back in d
}

}  // namespace

extern "C" {

#ifdef SC_CHECKED
const char* sc_version(void) { return "surfcascade-b200 0.3 (sm_100a, checked build)"; }
int sc_checked_violations(int reset) {
    unsigned int v = 0;
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(&v, sck::sc_chk_bad, 4);
    if (reset) { const unsigned int z = 0; cudaMemcpyToSymbol(sck::sc_chk_bad, &z, 4); }
    return (int)v;
}
#else
const char* sc_version(void) { return "surfcascade-b200 0.3 (sm_100a)"; }
int sc_checked_violations(int) { return -1; }   // not a checked build
#endif

int sc_create(int device, sc_handle** out) {
    if (!out) return SC_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0 || device < 0 || device >= n) return SC_ERR_CUDA;  // no CPU fallback by design
    if (cudaSetDevice(device) != cudaSuccess) return SC_ERR_CUDA;
    sc_handle* h = new sc_handle();
    h->device = device;
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return SC_ERR_CUDA; }
    cudaDeviceGetAttribute(&h->n_sms, cudaDevAttrMultiProcessorCount, device);
    const char* nofast = getenv("SC_DISABLE_FAST");
    h->allow_fast = !(nofast && nofast[0] == '1');
    const char* nocompact = getenv("SC_DISABLE_COMPACT");
    h->allow_compact = !(nocompact && nocompact[0] == '1');
    const char* sgf = getenv("SC_GROUP_FRAMES");
    if (sgf) h->group_max = std::min(32, std::max(1, atoi(sgf)));
    const char* sln = getenv("SC_LANES");
    if (sln) h->lanes_max = std::min(4, std::max(1, atoi(sln)));
    *out = h;
    return SC_OK;
}

void sc_destroy(sc_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    sc_comm_destroy(h);
    if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    for (auto& L : h->lanes) {
        if (L.st) { cudaStreamSynchronize(L.st); cudaStreamDestroy(L.st); }
        if (L.done) cudaEventDestroy(L.done);
        DevBuf* lb[] = {&L.d_multi, &L.d_pass, &L.d_visited, &L.d_start, &L.d_rec, &L.d_idx[0], &L.d_idx[1], &L.d_small, &L.d_chunks};
        for (DevBuf* b : lb) b->release();
    }
    if (h->ev_integral) cudaEventDestroy(h->ev_integral);
    for (auto& t : h->tickets) {
        if (t.done) cudaEventDestroy(t.done);
        t.d_img.release(); t.d_det.release(); t.d_cnt.release(); t.d_counters.release(); t.d_grp.release(); t.h_out.release();
    }
    for (cudaEvent_t e : h->ev_chunk) cudaEventDestroy(e);
    if (h->copy_st) { cudaStreamSynchronize(h->copy_st); cudaStreamDestroy(h->copy_st); }
    DevBuf* bufs[] = {&h->d_w, &h->d_wb, &h->d_plan, &h->d_geom, &h->d_img, &h->d_carry, &h->d_S, &h->d_counters, &h->d_det, &h->d_detcount, &h->d_pool_w, &h->d_pool_wb, &h->d_pool_auc, &h->d_pool_x, &h->d_pool_aux, &h->d_hook_img, &h->d_hook_carry, &h->d_hook_S, &h->d_ext_img, &h->d_ext_geom, &h->d_ext_X, &h->d_cert_items, &h->d_cert};
    for (DevBuf* b : bufs) b->release();
    h->h_stage.release();
    for (auto& sp : h->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (cudaEvent_t e : h->event_pool) cudaEventDestroy(e);
    delete h;
}

const char* sc_last_error(const sc_handle* h) { return h ? h->err.c_str() : "null handle"; }
void* sc_stream(sc_handle* h) { return h ? (void*)h->stream : nullptr; }
int64_t sc_launch_count(const sc_handle* h) { return h ? h->launches : 0; }

int sc_set_cascade(sc_handle* h, const sc_cascade_desc* d) {
    if (!h || !d) return SC_ERR_INVALID;
    if (any_ticket_busy(h)) return fail(h, SC_ERR_STATE, "a batch is in flight: collect it before changing the cascade");
    if (d->n_stages < 1 || d->n_stages > SC_MAX_STAGES || d->tmpl < 12) return fail(h, SC_ERR_INVALID, "n_stages must be 1..16, tmpl >= 12");
    if (!d->theta || !d->n_weak || !d->rects || !d->w || !d->bias) return fail(h, SC_ERR_INVALID, "null cascade arrays");
    SC_CUDA(h, cudaSetDevice(h->device));
    // everything is validated and staged in locals first; the handle changes only after the uploads succeeded, and from
    // here until then it holds no cascade at all (a failed call must not leave new host data beside old device weights)
    int total = 0;
    for (int s = 0; s < d->n_stages; s++) {
        if (d->n_weak[s] < 1) return fail(h, SC_ERR_INVALID, "stage without weak classifiers");
        total += d->n_weak[s];
    }
    for (int k = 0; k < total; k++) {
        const sc_rect& r = d->rects[k];
        const bool ok = r.w > 0 && r.h > 0 && r.x >= 0 && r.y >= 0 && r.x + r.w <= d->tmpl && r.y + r.h <= d->tmpl &&
                        (r.w == r.h || r.w == 4 * r.h || r.h == 4 * r.w);
        if (!ok) return fail(h, SC_ERR_INVALID, "weak classifier rect outside the template or not 2x2 / 4x1 / 1x4 cells");
    }
    std::vector<float> w36((size_t)total * SC_W_PITCH, 0.f);
    std::vector<double> wb(total);
    for (int k = 0; k < total; k++) {
        memcpy(&w36[(size_t)k * SC_W_PITCH], &d->w[(size_t)k * 33], 33 * sizeof(float));
        wb[k] = (double)d->w[(size_t)k * 33 + 32] * d->bias[k];  // prob += w[32] * bias, LogisticRegression.cpp:64
    }
    h->have_cascade = false;
    h->have_plan = false;
    SC_CUDA(h, cudaStreamSynchronize(h->stream));
    SC_CUDA(h, h->d_w.ensure(w36.size() * sizeof(float)));
    SC_CUDA(h, h->d_wb.ensure(wb.size() * sizeof(double)));
    SC_CUDA(h, cudaMemcpy(h->d_w.p, w36.data(), w36.size() * sizeof(float), cudaMemcpyHostToDevice));
    SC_CUDA(h, cudaMemcpy(h->d_wb.p, wb.data(), wb.size() * sizeof(double), cudaMemcpyHostToDevice));
    h->tmpl = d->tmpl; h->n_stages = d->n_stages; h->total_weak = total;
    h->theta.assign(d->theta, d->theta + d->n_stages);
    h->n_weak.assign(d->n_weak, d->n_weak + d->n_stages);
    h->weak_base.resize(d->n_stages);
    for (int s = 0, b = 0; s < d->n_stages; b += d->n_weak[s], s++) h->weak_base[s] = b;
    h->rects.assign(d->rects, d->rects + total);
    h->w.assign(d->w, d->w + (size_t)total * 33);
    h->bias.assign(d->bias, d->bias + total);
    h->have_cascade = true;
    return SC_OK;
}

int sc_load_model(sc_handle* h, const char* path, int tmpl) {
    if (!h || !path) return SC_ERR_INVALID;
    sc_host::FlatCascade fc;
    std::string why;
    if (!sc_host::load_flat_cascade(path, tmpl, &fc, &why)) return fail(h, SC_ERR_IO, why);
    sc_cascade_desc d;
    d.tmpl = tmpl; d.n_stages = (int)fc.theta.size();
    d.theta = fc.theta.data(); d.n_weak = fc.n_weak.data(); d.rects = fc.rects.data(); d.w = fc.w.data(); d.bias = fc.bias.data();
    return sc_set_cascade(h, &d);
}

int sc_model_flatten(const char* path, int tmpl, float* theta, int32_t* n_weak, int max_stages, sc_rect* rects, int32_t* patch_index, float* w,
                     double* bias, int max_weak, int* total_weak) {
    if (!path) return SC_ERR_INVALID;
    sc_host::FlatCascade fc;
    std::string why;
    if (!sc_host::load_flat_cascade(path, tmpl, &fc, &why)) return SC_ERR_IO;
    const int S = (int)fc.theta.size(), T = (int)fc.bias.size();
    if (total_weak) *total_weak = T;
    if (S > max_stages || T > max_weak) return SC_ERR_CAPACITY;  // the caller's arrays would hold a truncated cascade
    for (int s = 0; s < S && s < max_stages; s++) {
        if (theta) theta[s] = fc.theta[s];
        if (n_weak) n_weak[s] = fc.n_weak[s];
    }
    for (int k = 0; k < T && k < max_weak; k++) {
        if (rects) rects[k] = fc.rects[k];
        if (patch_index) patch_index[k] = fc.patch_index[k];
        if (w) memcpy(w + (size_t)k * 33, &fc.w[(size_t)k * 33], 33 * sizeof(float));
        if (bias) bias[k] = fc.bias[k];
    }
    return S;
}

int sc_model_resave(const char* in_path, const char* out_path) {
    if (!in_path || !out_path) return SC_ERR_INVALID;
    return sc_host::resave_model(in_path, out_path) ? SC_OK : SC_ERR_IO;
}

int sc_pool_patches(int tmpl, sc_rect* out, int cap) {
    std::vector<sc_rect> pool;
    sc_host::pool_patches(tmpl, tmpl, &pool);
    for (int i = 0; i < (int)pool.size() && i < cap; i++) out[i] = pool[i];
    return (int)pool.size();
}

int sc_project_patches(int tmpl, int l, const sc_rect* patches, int n, sc_rect* out) {
    if (!patches || !out || tmpl < 1) return SC_ERR_INVALID;
    for (int i = 0; i < n; i++) out[i] = sc_host::project_patch(tmpl, l, patches[i]);
    return SC_OK;
}

int sc_integral(sc_handle* h, const uint8_t* gray, int W, int H, int stride, float* out) {
    if (!h || !gray || W < 2 || H < 2 || stride < W) return fail(h, SC_ERR_INVALID, "bad image arguments");
    SC_CUDA(h, cudaSetDevice(h->device));
    h->have_integral = false;
    const int n_strips = (W + SC_STRIP - 1) / SC_STRIP;
    // the hooks keep their own single-frame integral in the step-1 layout (explicit rects sit on any pixel)
    const ScLayout L = sc_host::make_layout(W, H, 1, 1);
    if (L.frame4 * 16 > 0xffffffffLL) return fail(h, SC_ERR_INVALID, "image too large for the hooks (corner offsets are 32-bit byte offsets: 4 GiB of integral image)");
    SC_CUDA(h, h->d_hook_img.ensure(align256((size_t)W * H)));
    SC_CUDA(h, h->d_hook_carry.ensure(align256((size_t)H * n_strips * 32)));
    SC_CUDA(h, h->d_hook_S.ensure((size_t)L.frame4 * 16));
#ifdef SC_CHECKED
    checked_set_range(1, h->d_hook_S.p, h->d_hook_S.cap);
#endif
    SC_CUDA(h, cudaMemcpy2DAsync(h->d_hook_img.p, W, gray, stride, W, H, cudaMemcpyHostToDevice, h->stream));
    sck::k_strip_carry<<<(H + 3) / 4, 128, 0, h->stream>>>(h->d_hook_img.as<uint8_t>(), W, H, n_strips, 1, h->d_hook_carry.as<int>());
#if SC_WALK_TILED
    sck::k_integral_walk_tiled<<<(n_strips + 3) / 4, 128, 0, h->stream>>>(h->d_hook_img.as<uint8_t>(), W, H, n_strips, 1, h->d_hook_carry.as<int>(),
                                                                            h->d_hook_S.as<float4>(), L);
#else
    sck::k_integral_walk<<<(n_strips + 3) / 4, 128, 0, h->stream>>>(h->d_hook_img.as<uint8_t>(), W, H, n_strips, 1, h->d_hook_carry.as<int>(),
                                                                      h->d_hook_S.as<float4>(), L);
#endif
    h->launches += 2;
    SC_CUDA(h, cudaGetLastError());
    if (out) {
        // hand the caller the reference's interleaved (H+1) x (W+1) x 8 image
        DevBuf d_out;
        const size_t bytes = (size_t)(H + 1) * (W + 1) * 32;
        SC_CUDA(h, d_out.ensure(bytes));
        sck::k_export_integral<<<h->n_sms * 8, 256, 0, h->stream>>>(h->d_hook_S.as<float4>(), L, W, H, d_out.as<float4>());
        h->launches++;
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out.p, bytes, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        d_out.release();
        if (e != cudaSuccess) return cuda_fail(h, e, "sc_integral export");
    }
    SC_CUDA(h, cudaStreamSynchronize(h->stream));
    h->have_integral = true; h->cur_W = W; h->cur_H = H; h->hook_lay = L;
    return SC_OK;
}

// Parity hook for the layout the scan reads: the same two kernels as the detect path (run_integral) with the detection
// plan's deinterleave factors for lattice step `step` (columns by 2 * step, rows by step), exported to the reference layout.
int sc_integral_scan_layout(sc_handle* h, const uint8_t* gray, int W, int H, int stride, int step, float* out) {
    if (!h || !gray || !out || W < 2 || H < 2 || stride < W || step < 1 || step > 64) return fail(h, SC_ERR_INVALID, "bad image arguments");
    SC_CUDA(h, cudaSetDevice(h->device));
    const int n_strips = (W + SC_STRIP - 1) / SC_STRIP;
    const ScLayout L = sc_host::make_layout(W, H, 2 * step, step);
    if (L.hp > 4096 || L.frame4 * 16 > 0xffffffffLL) return fail(h, SC_ERR_INVALID, "frame too large for the scan layout");
    DevBuf d_img, d_carry, d_S, d_out;
    const size_t bytes = (size_t)(H + 1) * (W + 1) * 32;
    cudaError_t e = d_img.ensure(align256((size_t)W * H));
    if (e == cudaSuccess) e = d_carry.ensure(align256((size_t)H * n_strips * 32));
    if (e == cudaSuccess) e = d_S.ensure((size_t)L.frame4 * 16);
    if (e == cudaSuccess) e = d_out.ensure(bytes);
    if (e == cudaSuccess) e = cudaMemcpy2DAsync(d_img.p, W, gray, stride, W, H, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) {
        sck::k_strip_carry<<<(H + 3) / 4, 128, 0, h->stream>>>(d_img.as<uint8_t>(), W, H, n_strips, 1, d_carry.as<int>());
#if SC_WALK_TILED
        sck::k_integral_walk_tiled<<<(n_strips + 3) / 4, 128, 0, h->stream>>>(d_img.as<uint8_t>(), W, H, n_strips, 1, d_carry.as<int>(), d_S.as<float4>(), L);
#else
        sck::k_integral_walk<<<(n_strips + 3) / 4, 128, 0, h->stream>>>(d_img.as<uint8_t>(), W, H, n_strips, 1, d_carry.as<int>(), d_S.as<float4>(), L);
#endif
        sck::k_export_integral<<<h->n_sms * 8, 256, 0, h->stream>>>(d_S.as<float4>(), L, W, H, d_out.as<float4>());
        h->launches += 3;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out.p, bytes, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    d_img.release(); d_carry.release(); d_S.release(); d_out.release();
    if (e != cudaSuccess) return cuda_fail(h, e, "sc_integral_scan_layout");
    return SC_OK;
}

// Parity hook: the compact plane through the scan layout for lattice step `step`, exported in pixel order.
int sc_integral_compact(sc_handle* h, const uint8_t* gray, int W, int H, int stride, int step, uint32_t* out) {
    if (!h || !gray || !out || W < 2 || H < 2 || stride < W || step < 1 || step > 64) return fail(h, SC_ERR_INVALID, "bad image arguments");
    SC_CUDA(h, cudaSetDevice(h->device));
    const int n_strips = (W + SC_STRIP - 1) / SC_STRIP;
    const ScLayout L = sc_host::make_layout(W, H, 2 * step, step);
    if (L.hp > 4096 || L.frame4 * 16 > 0xffffffffLL) return fail(h, SC_ERR_INVALID, "frame too large for the scan layout");
    DevBuf d_img, d_carry, d_S, d_out;
    const size_t bytes = (size_t)(H + 1) * (W + 1) * 16;
    cudaError_t e = d_img.ensure(align256((size_t)W * H));
    if (e == cudaSuccess) e = d_carry.ensure(align256((size_t)H * n_strips * 32));
    if (e == cudaSuccess) e = d_S.ensure((size_t)L.frame4 * 16);
    if (e == cudaSuccess) e = d_out.ensure(bytes);
    if (e == cudaSuccess) e = cudaMemcpy2DAsync(d_img.p, W, gray, stride, W, H, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) {
        sck::k_strip_carry<<<(H + 3) / 4, 128, 0, h->stream>>>(d_img.as<uint8_t>(), W, H, n_strips, 1, d_carry.as<int>());
        // both forms of the walk write the plane: odd steps take the shuffle-scan form here so that both stay covered
        if (SC_WALK_TILED && (step & 1) == 0)
            sck::k_integral_walk_tiled<<<(n_strips + 3) / 4, 128, 0, h->stream>>>(d_img.as<uint8_t>(), W, H, n_strips, 1, d_carry.as<int>(), d_S.as<float4>(), L);
        else
            sck::k_integral_walk<<<(n_strips + 3) / 4, 128, 0, h->stream>>>(d_img.as<uint8_t>(), W, H, n_strips, 1, d_carry.as<int>(), d_S.as<float4>(), L);
        sck::k_export_compact<<<h->n_sms * 8, 256, 0, h->stream>>>(d_S.as<float4>(), L, W, H, d_out.as<uint4>());
        h->launches += 3;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out.p, bytes, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    d_img.release(); d_carry.release(); d_S.release(); d_out.release();
    if (e != cudaSuccess) return cuda_fail(h, e, "sc_integral_compact");
    return SC_OK;
}

int sc_box_sums_compact(sc_handle* h, const sc_rect* rects, int n, float* out) {
    if (!h || !rects || !out || n < 0) return SC_ERR_INVALID;
    if (!h->have_integral) return fail(h, SC_ERR_STATE, "sc_integral has not been called");
    if (n == 0) return SC_OK;
    SC_CUDA(h, cudaSetDevice(h->device));
    for (int i = 0; i < n; i++) {
        const sc_rect& r = rects[i];
        const bool inside = r.x >= 0 && r.y >= 0 && r.w > 0 && r.h > 0 && r.x + r.w <= h->cur_W && r.y + r.h <= h->cur_H;
        const bool cells = (r.w == r.h && r.w >= 2) || r.w == 4 * r.h || r.h == 4 * r.w;
        if (!inside || !cells) return fail(h, SC_ERR_INVALID, "rect outside the image or not 2x2 / 4x1 / 1x4 cells");
    }
    DevBuf d_r, d_o;
    SC_CUDA(h, d_r.ensure((size_t)n * sizeof(sc_rect)));
    SC_CUDA(h, d_o.ensure((size_t)n * 32 * sizeof(float)));
    SC_CUDA(h, cudaMemcpyAsync(d_r.p, rects, (size_t)n * sizeof(sc_rect), cudaMemcpyHostToDevice, h->stream));
    sck::k_box_sums_compact<<<(n + 127) / 128, 128, 0, h->stream>>>(h->d_hook_S.as<float4>(), h->hook_lay, d_r.as<int4>(), n, d_o.as<float>());
    h->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_o.p, (size_t)n * 32 * sizeof(float), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    d_r.release(); d_o.release();
    if (e != cudaSuccess) return cuda_fail(h, e, "sc_box_sums_compact");
    return SC_OK;
}

int sc_cell_bounds(sc_handle* h, const int32_t* ce, int n, uint32_t* out) {
    if (!h || !ce || !out || n < 1 || n > SC_CERT_MAX_ITEMS) return SC_ERR_INVALID;
    if (!h->have_integral) return fail(h, SC_ERR_STATE, "sc_integral has not been called");
    for (int i = 0; i < n; i++)
        if (ce[i] < 1) return fail(h, SC_ERR_INVALID, "cell edge must be positive");
    SC_CUDA(h, cudaSetDevice(h->device));
    DevBuf d_c, d_o;
    SC_CUDA(h, d_c.ensure((size_t)n * 4));
    SC_CUDA(h, d_o.ensure((size_t)n * 4));
    SC_CUDA(h, cudaMemcpyAsync(d_c.p, ce, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
    SC_CUDA(h, cudaMemsetAsync(d_o.p, 0, (size_t)n * 4, h->stream));
    sck::k_cell_bounds<<<dim3(n * 4, 1), 256, 0, h->stream>>>(h->d_hook_S.as<float4>(), h->hook_lay, h->cur_W, h->cur_H, d_c.as<int>(), n, 4, d_o.as<uint32_t>());
    h->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_o.p, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    d_c.release(); d_o.release();
    if (e != cudaSuccess) return cuda_fail(h, e, "sc_cell_bounds");
    return SC_OK;
}

static int features_impl(sc_handle* h, const sc_rect* rects, int n, float* out, float* sums) {
    if (!h || !rects || n < 0) return SC_ERR_INVALID;
    if (!h->have_integral) return fail(h, SC_ERR_STATE, "sc_integral has not been called");
    if (n == 0) return SC_OK;
    SC_CUDA(h, cudaSetDevice(h->device));
    for (int i = 0; i < n; i++) {
        const sc_rect& r = rects[i];
        const bool inside = r.x >= 0 && r.y >= 0 && r.w > 0 && r.h > 0 && r.x + r.w <= h->cur_W && r.y + r.h <= h->cur_H;
        const bool cells = !out || ((r.w == r.h && r.w >= 2) || r.w == 4 * r.h || r.h == 4 * r.w);
        if (!inside || !cells) return fail(h, SC_ERR_INVALID, "rect outside the image or not 2x2 / 4x1 / 1x4 cells");
    }
    DevBuf d_r, d_o, d_s;
    SC_CUDA(h, d_r.ensure((size_t)n * sizeof(sc_rect)));
    SC_CUDA(h, cudaMemcpyAsync(d_r.p, rects, (size_t)n * sizeof(sc_rect), cudaMemcpyHostToDevice, h->stream));
    if (out) SC_CUDA(h, d_o.ensure((size_t)n * 32 * sizeof(float)));
    if (sums) SC_CUDA(h, d_s.ensure((size_t)n * sizeof(float)));
    sck::k_features<<<(n + 127) / 128, 128, 0, h->stream>>>(h->d_hook_S.as<float4>(), h->hook_lay, d_r.as<int4>(), n, d_o.as<float>(), d_s.as<float>());
    h->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && out) e = cudaMemcpyAsync(out, d_o.p, (size_t)n * 32 * sizeof(float), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess && sums) e = cudaMemcpyAsync(sums, d_s.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    d_r.release(); d_o.release(); d_s.release();
    if (e != cudaSuccess) return cuda_fail(h, e, "sc_features");
    return SC_OK;
}

int sc_features(sc_handle* h, const sc_rect* rects, int n, float* out) { return out ? features_impl(h, rects, n, out, nullptr) : SC_ERR_INVALID; }
int sc_window_sum(sc_handle* h, const sc_rect* rects, int n, float* out) { return out ? features_impl(h, rects, n, nullptr, out) : SC_ERR_INVALID; }

static int stage_scores_impl(sc_handle* h, const int32_t* wins, int n, float* out, float* out2, bool fast_check);

int sc_stage_scores(sc_handle* h, const int32_t* wins, int n, float* out) { return stage_scores_impl(h, wins, n, out, nullptr, false); }

int sc_stage0_fast_check(sc_handle* h, const int32_t* wins, int n, float* fast_sum, float* exact_sum, double* margin) {
    if (!fast_sum || !exact_sum) return SC_ERR_INVALID;
    if (margin && h && h->have_cascade) *margin = fast_margin(h);
    return stage_scores_impl(h, wins, n, fast_sum, exact_sum, true);
}

static int stage_scores_impl(sc_handle* h, const int32_t* wins, int n, float* out, float* out2, bool fast_check) {
    if (!h || !wins || !out || n < 0) return SC_ERR_INVALID;
    if (!h->have_integral) return fail(h, SC_ERR_STATE, "sc_integral has not been called");
    if (!h->have_cascade) return fail(h, SC_ERR_STATE, "no cascade loaded");
    if (n == 0) return SC_OK;
    SC_CUDA(h, cudaSetDevice(h->device));
    const int W = h->cur_W, H = h->cur_H, tw = h->total_weak;
    std::vector<ScGeom> geom((size_t)n * tw);
    for (int i = 0; i < n; i++) {
        const int x = wins[3 * i], y = wins[3 * i + 1], l = wins[3 * i + 2];
        if (x < 0 || y < 0 || l < 1 || x + l > W || y + l > H) return fail(h, SC_ERR_INVALID, "window outside the image");
        for (int k = 0; k < tw; k++)
            if (!sc_host::project_geom(h->tmpl, l, h->rects[k], h->hook_lay, 0, &geom[(size_t)i * tw + k]))
                return fail(h, SC_ERR_INVALID, "degenerate projected patch");
    }
    ScPlan mini;
    memset(&mini, 0, sizeof(mini));
    mini.W = W; mini.H = H; mini.lay = h->hook_lay; mini.n_stages = h->n_stages; mini.total_weak = tw;
    for (int s = 0; s < h->n_stages; s++) { mini.theta[s] = h->theta[s]; mini.n_weak[s] = h->n_weak[s]; mini.weak_base[s] = h->weak_base[s]; }
    DevBuf d_p, d_g, d_w, d_o;
    cudaError_t e = d_p.ensure(sizeof(ScPlan));
    if (e == cudaSuccess) e = d_g.ensure(geom.size() * sizeof(ScGeom));
    if (e == cudaSuccess) e = d_w.ensure((size_t)n * 3 * sizeof(int));
    const size_t out_floats = fast_check ? (size_t)n * 2 : (size_t)n * h->n_stages;
    if (e == cudaSuccess) e = d_o.ensure(out_floats * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_p.p, &mini, sizeof(ScPlan), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_g.p, geom.data(), geom.size() * sizeof(ScGeom), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_w.p, wins, (size_t)n * 3 * sizeof(int), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) {
        if (fast_check)
            sck::k_stage0_fast_check<<<(n + 63) / 64, 64, 0, h->stream>>>(d_p.as<ScPlan>(), h->d_hook_S.as<float4>(), d_g.as<ScGeom>(), h->d_w.as<float>(),
                                                                           h->d_wb.as<double>(), d_w.as<int>(), n, d_o.as<float>(), d_o.as<float>() + n);
        else
            sck::k_stage_scores<<<(n + 63) / 64, 64, 0, h->stream>>>(d_p.as<ScPlan>(), h->d_hook_S.as<float4>(), d_g.as<ScGeom>(), h->d_w.as<float>(),
                                                                      h->d_wb.as<double>(), d_w.as<int>(), n, d_o.as<float>());
        h->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && fast_check) {
        e = cudaMemcpyAsync(out, d_o.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(out2, d_o.as<float>() + n, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, h->stream);
    } else if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_o.p, (size_t)n * h->n_stages * sizeof(float), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    d_p.release(); d_g.release(); d_w.release(); d_o.release();
    if (e != cudaSuccess) return cuda_fail(h, e, "sc_stage_scores");
    return SC_OK;
}

static int predict_impl(sc_handle* h, const float* w, const double* bias, const float* x, int n, float* out, bool stage_mean) {
    if (!h || !w || !bias || !x || !out || n < 1) return SC_ERR_INVALID;
    if (stage_mean && n > 1024) return fail(h, SC_ERR_INVALID, "a stage holds at most 1024 weak classifiers");
    SC_CUDA(h, cudaSetDevice(h->device));
    std::vector<float> w36((size_t)n * SC_W_PITCH, 0.f);
    std::vector<double> wb(n);
    for (int i = 0; i < n; i++) {
        memcpy(&w36[(size_t)i * SC_W_PITCH], w + (size_t)i * 33, 33 * sizeof(float));
        wb[i] = (double)w[(size_t)i * 33 + 32] * bias[i];
    }
    DevBuf d_w, d_b, d_x, d_o;
    cudaError_t e = d_w.ensure(w36.size() * 4);
    if (e == cudaSuccess) e = d_b.ensure((size_t)n * 8);
    if (e == cudaSuccess) e = d_x.ensure((size_t)n * 32 * 4);
    if (e == cudaSuccess) e = d_o.ensure((size_t)(n + 1) * 4);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_w.p, w36.data(), w36.size() * 4, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_b.p, wb.data(), (size_t)n * 8, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_x.p, x, (size_t)n * 32 * 4, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) {
        if (stage_mean)
            sck::k_weak_predict<<<1, 128, 0, h->stream>>>(d_w.as<float>(), d_b.as<double>(), d_x.as<float>(), n, d_o.as<float>(), d_o.as<float>() + n);
        else
            sck::k_weak_predict<<<(n + 127) / 128, 128, 0, h->stream>>>(d_w.as<float>(), d_b.as<double>(), d_x.as<float>(), n, d_o.as<float>(), nullptr);
        h->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess)
        e = stage_mean ? cudaMemcpyAsync(out, d_o.as<float>() + n, 4, cudaMemcpyDeviceToHost, h->stream)
                       : cudaMemcpyAsync(out, d_o.p, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    d_w.release(); d_b.release(); d_x.release(); d_o.release();
    if (e != cudaSuccess) return cuda_fail(h, e, "sc_weak_predict");
    return SC_OK;
}

int sc_weak_predict(sc_handle* h, const float* w, const double* bias, const float* x, int n, float* out) {
    return predict_impl(h, w, bias, x, n, out, false);
}
int sc_stage_predict(sc_handle* h, const float* w, const double* bias, const float* x, int n, float* out) {
    return predict_impl(h, w, bias, x, n, out, true);
}

static void pool_thresholds(float* thr) {
    int i = 0;
    for (float t = 1; t >= 0 && i < SC_POOL_LEVELS - 1; t -= 0.05f) thr[i++] = t;  // StageClassifier.cpp:59, auc_step = 0.05f
    while (i < SC_POOL_LEVELS - 1) thr[i++] = -1.f;
}

int sc_pool_hist_device(sc_handle* h, const float* d_X, int N, int P, const uint8_t* d_labels, const float* Wcand, const double* bias,
                        const float* d_prior_sum, int T, uint32_t* d_hist) {
    if (!h || !d_X || !d_labels || !Wcand || !bias || !d_hist || N < 1 || P < 1 || T < 0) return fail(h, SC_ERR_INVALID, "bad arguments");
    SC_CUDA(h, cudaSetDevice(h->device));
    std::vector<float> w36((size_t)P * SC_W_PITCH, 0.f);
    std::vector<double> wb(P);
    for (int k = 0; k < P; k++) {
        memcpy(&w36[(size_t)k * SC_W_PITCH], Wcand + (size_t)k * 33, 33 * sizeof(float));
        wb[k] = (double)Wcand[(size_t)k * 33 + 32] * bias[k];
    }
    float thr[SC_POOL_LEVELS - 1];
    pool_thresholds(thr);
    SC_CUDA(h, h->d_pool_w.ensure(w36.size() * 4));
    SC_CUDA(h, h->d_pool_wb.ensure((size_t)P * 8 + sizeof(thr)));
    // synchronous small uploads: the stream may still be reading the previous call's copies
    SC_CUDA(h, cudaStreamSynchronize(h->stream));
    SC_CUDA(h, cudaMemcpy(h->d_pool_w.p, w36.data(), w36.size() * 4, cudaMemcpyHostToDevice));
    SC_CUDA(h, cudaMemcpy(h->d_pool_wb.p, wb.data(), (size_t)P * 8, cudaMemcpyHostToDevice));
    float* d_thr = reinterpret_cast<float*>(h->d_pool_wb.as<unsigned char>() + (size_t)P * 8);
    SC_CUDA(h, cudaMemcpy(d_thr, thr, sizeof(thr), cudaMemcpyHostToDevice));
    const int kchunks = (P + SC_POOL_KC - 1) / SC_POOL_KC;
    const int nchunks = (N + SC_POOL_NC - 1) / SC_POOL_NC;
    const int slices = std::max(1, std::min(nchunks, (h->n_sms * 4 + kchunks - 1) / kchunks));
    {
        KernelSpan ks(h, K_POOL);
        sck::k_pool_hist<<<kchunks * slices, 256, 0, h->stream>>>(d_X, N, P, d_labels, h->d_pool_w.as<float>(), h->d_pool_wb.as<double>(), d_prior_sum,
                                                                  (float)(T + 1), d_thr, slices, d_hist);
    }
    SC_CUDA(h, cudaGetLastError());
    return SC_OK;
}

int sc_pool_auc_device(sc_handle* h, const uint32_t* d_hist, int P, int64_t n_pos, int64_t n_neg, float* auc) {
    if (!h || !d_hist || !auc || P < 1 || n_pos < 1 || n_neg < 1) return fail(h, SC_ERR_INVALID, "bad arguments");
    SC_CUDA(h, cudaSetDevice(h->device));
    SC_CUDA(h, h->d_pool_auc.ensure((size_t)P * 4));
    sck::k_pool_auc<<<(P + 127) / 128, 128, 0, h->stream>>>(d_hist, P, (float)n_pos, (float)n_neg, h->d_pool_auc.as<float>());
    h->launches++;
    SC_CUDA(h, cudaGetLastError());
    SC_CUDA(h, cudaMemcpyAsync(auc, h->d_pool_auc.p, (size_t)P * 4, cudaMemcpyDeviceToHost, h->stream));
    SC_CUDA(h, cudaStreamSynchronize(h->stream));
    drain_spans(h);
    return SC_OK;
}

int sc_pool_eval(sc_handle* h, const float* X, int N, int P, const uint8_t* labels, const float* Wcand, const double* bias, const float* prior_sum,
                 int T, float* auc) {
    if (!h || !X || !labels || !Wcand || !bias || !auc || N < 1 || P < 1) return fail(h, SC_ERR_INVALID, "bad arguments");
    SC_CUDA(h, cudaSetDevice(h->device));
    int64_t n_pos = 0;
    for (int n = 0; n < N; n++) n_pos += labels[n] != 0;
    if (n_pos == 0 || n_pos == N) return fail(h, SC_ERR_INVALID, "need both positive and negative samples");
    const size_t hist_bytes = (size_t)P * 2 * SC_POOL_LEVELS * 4;
    // stream X through a bounded device buffer in sample chunks
    const size_t row = (size_t)P * 32 * 4;
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)N, ((size_t)1 << 30) / row));
    SC_CUDA(h, h->d_pool_x.ensure((size_t)chunk * row));
    SC_CUDA(h, h->d_pool_aux.ensure(hist_bytes + (size_t)chunk * 5 + 64));
    uint32_t* d_hist = h->d_pool_aux.as<uint32_t>();
    float* d_prior = reinterpret_cast<float*>(h->d_pool_aux.as<unsigned char>() + hist_bytes);
    uint8_t* d_lab = h->d_pool_aux.as<unsigned char>() + hist_bytes + (size_t)chunk * 4;
    SC_CUDA(h, cudaMemsetAsync(d_hist, 0, hist_bytes, h->stream));
    for (int n0 = 0; n0 < N; n0 += chunk) {
        const int m = std::min(chunk, N - n0);
        SC_CUDA(h, cudaMemcpyAsync(h->d_pool_x.p, X + (size_t)n0 * P * 32, (size_t)m * row, cudaMemcpyHostToDevice, h->stream));
        SC_CUDA(h, cudaMemcpyAsync(d_lab, labels + n0, (size_t)m, cudaMemcpyHostToDevice, h->stream));
        if (prior_sum) SC_CUDA(h, cudaMemcpyAsync(d_prior, prior_sum + n0, (size_t)m * 4, cudaMemcpyHostToDevice, h->stream));
        int rc = sc_pool_hist_device(h, h->d_pool_x.as<float>(), m, P, d_lab, Wcand, bias, prior_sum ? d_prior : nullptr, T, d_hist);
        if (rc != SC_OK) return rc;
    }
    return sc_pool_auc_device(h, d_hist, P, n_pos, N - n_pos, auc);
}

// Training-side descriptor extraction (next row N3).  d_imgs: N x tmpl x tmpl u8 samples in device memory;
// d_X: [N][P][32] floats in device memory, P = size of the template pool.  Asynchronous on the handle's stream.
int sc_extract_pool_features_device(sc_handle* h, const uint8_t* d_imgs, int N, int tmpl, float* d_X) {
    if (!h || !d_imgs || !d_X || N < 1 || tmpl < 12) return fail(h, SC_ERR_INVALID, "bad arguments");
    SC_CUDA(h, cudaSetDevice(h->device));
    std::vector<sc_rect> pool;
    sc_host::pool_patches(tmpl, tmpl, &pool);
    const int P = (int)pool.size();
    const size_t smem = (size_t)(tmpl + 1) * (tmpl + 1) * 32 + (size_t)tmpl * tmpl;
    if (smem > 220 * 1024) return fail(h, SC_ERR_INVALID, "template too large for the shared-memory sample integral (side <= 82)");
    SC_CUDA(h, h->d_ext_geom.ensure((size_t)P * sizeof(sc_rect)));
    SC_CUDA(h, cudaStreamSynchronize(h->stream));  // the previous call may still read the pool
    SC_CUDA(h, cudaMemcpy(h->d_ext_geom.p, pool.data(), (size_t)P * sizeof(sc_rect), cudaMemcpyHostToDevice));
    SC_CUDA(h, cudaFuncSetAttribute(sck::k_pool_features, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        KernelSpan ks(h, K_POOLFEAT);
        sck::k_pool_features<<<N, 320, smem, h->stream>>>(d_imgs, tmpl, h->d_ext_geom.as<int4>(), P, d_X);
    }
    SC_CUDA(h, cudaGetLastError());
    return SC_OK;
}

// Host-memory variant: imgs N x tmpl x tmpl u8, X [N][P][32]; streamed through bounded device buffers.
int sc_extract_pool_features(sc_handle* h, const uint8_t* imgs, int N, int tmpl, float* X) {
    if (!h || !imgs || !X || N < 1 || tmpl < 12) return fail(h, SC_ERR_INVALID, "bad arguments");
    SC_CUDA(h, cudaSetDevice(h->device));
    std::vector<sc_rect> pool;
    sc_host::pool_patches(tmpl, tmpl, &pool);
    const size_t row = pool.size() * 32 * sizeof(float);
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)N, ((size_t)1 << 29) / row));
    SC_CUDA(h, h->d_ext_img.ensure(align256((size_t)chunk * tmpl * tmpl)));
    SC_CUDA(h, h->d_ext_X.ensure((size_t)chunk * row));
    for (int n0 = 0; n0 < N; n0 += chunk) {
        const int m = std::min(chunk, N - n0);
        SC_CUDA(h, cudaMemcpyAsync(h->d_ext_img.p, imgs + (size_t)n0 * tmpl * tmpl, (size_t)m * tmpl * tmpl, cudaMemcpyHostToDevice, h->stream));
        int rc = sc_extract_pool_features_device(h, h->d_ext_img.as<uint8_t>(), m, tmpl, h->d_ext_X.as<float>());
        if (rc != SC_OK) return rc;
        SC_CUDA(h, cudaMemcpyAsync(X + (size_t)n0 * pool.size() * 32, h->d_ext_X.p, (size_t)m * row, cudaMemcpyDeviceToHost, h->stream));
        SC_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    drain_spans(h);
    return SC_OK;
}

int sc_detect_device(sc_handle* h, const uint8_t* d_frames, int nframes, int W, int H, const sc_detect_params* params, sc_detection* d_out,
                     size_t cap, uint32_t* d_n) {
    if (!h || !d_frames || nframes < 1 || !d_out || !d_n) return fail(h, SC_ERR_INVALID, "bad arguments");
    if (any_ticket_busy(h)) return fail(h, SC_ERR_STATE, "a submitted batch is in flight (its counters are read against the current plan): collect it first");
    SC_CUDA(h, cudaSetDevice(h->device));
    const sc_detect_params prm = params ? *params : default_params();
    int rc = ensure_plan(h, W, H, prm);
    if (rc != SC_OK) return rc;
    rc = ensure_group_buffers(h, nframes, false);
    if (rc != SC_OK) return rc;
    SC_CUDA(h, h->d_counters.ensure((size_t)nframes * SC_CNT_STRIDE * 8));
    SC_CUDA(h, cudaMemsetAsync(h->d_counters.p, 0, (size_t)nframes * SC_CNT_STRIDE * 8, h->stream));
    SC_CUDA(h, cudaMemsetAsync(d_n, 0, 4, h->stream));
    const uint32_t det_cap = (uint32_t)std::min<size_t>(cap, 0xffffffffu);
    for (int i0 = 0; i0 < nframes; i0 += h->int_frames) {
        const int ni = std::min(h->int_frames, nframes - i0);
        rc = run_supergroup(h, nullptr, 0, d_frames + (size_t)i0 * W * H, ni, i0, d_out, det_cap, d_n,
                            h->d_counters.as<unsigned long long>() + (size_t)i0 * SC_CNT_STRIDE);
        if (rc != SC_OK) return rc;
    }
    SC_CUDA(h, h->h_stage.ensure((size_t)nframes * SC_CNT_STRIDE * 8));
    SC_CUDA(h, cudaMemcpyAsync(h->h_stage.p, h->d_counters.p, (size_t)nframes * SC_CNT_STRIDE * 8, cudaMemcpyDeviceToHost, h->stream));
    h->last_nframes = nframes;
    return SC_OK;
}

int sc_sync(sc_handle* h) {
    if (!h) return SC_ERR_INVALID;
    SC_CUDA(h, cudaSetDevice(h->device));
    SC_CUDA(h, cudaStreamSynchronize(h->stream));
    drain_spans(h);
    return SC_OK;
}

int sc_probe_gather(sc_handle* h, size_t table_bytes, int iters, double* gbps) {
    if (!h || !gbps || table_bytes < 4096 || iters < 1) return fail(h, SC_ERR_INVALID, "bad arguments");
    SC_CUDA(h, cudaSetDevice(h->device));
    DevBuf tab, sink;
    SC_CUDA(h, tab.ensure(table_bytes));
    SC_CUDA(h, sink.ensure(256));
    SC_CUDA(h, cudaMemsetAsync(tab.p, 0, table_bytes, h->stream));
    const uint32_t n_sectors = (uint32_t)std::min<size_t>(table_bytes / 32, 0x0fffffffu);
    const int per_thread = 64, grid = h->n_sms * 32;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int w = 0; w < 2; w++) sck::k_probe_gather<<<grid, 256, 0, h->stream>>>(tab.as<float4>(), n_sectors, per_thread, sink.as<float>());
    cudaEventRecord(a, h->stream);
    for (int i = 0; i < iters; i++) sck::k_probe_gather<<<grid, 256, 0, h->stream>>>(tab.as<float4>(), n_sectors, per_thread, sink.as<float>());
    cudaEventRecord(b, h->stream);
    cudaError_t e = cudaStreamSynchronize(h->stream);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a); cudaEventDestroy(b);
    tab.release(); sink.release();
    if (e != cudaSuccess) return cuda_fail(h, e, "sc_probe_gather");
    *gbps = (double)grid * 256 * per_thread * 32.0 * iters / (ms * 1e6);
    return SC_OK;
}

int sc_probe_stream(sc_handle* h, size_t table_bytes, int iters, int mode, double* gbps) {
    if (!h || !gbps || table_bytes < (1u << 14) || iters < 1) return fail(h, SC_ERR_INVALID, "bad arguments");
    SC_CUDA(h, cudaSetDevice(h->device));
    DevBuf tab, sink;
    SC_CUDA(h, tab.ensure(table_bytes));
    SC_CUDA(h, sink.ensure(256));
    SC_CUDA(h, cudaMemsetAsync(tab.p, 0, table_bytes, h->stream));
    uint32_t n4 = 1u << 10;
    while ((size_t)n4 * 2 * 16 <= table_bytes && n4 < (1u << 29)) n4 *= 2;  // largest power of two that fits (64 KB and less: L1-resident)
    // mode bit 0: 0 = ld.global.cg, 1 = ld.global.nc; bits 1..: independent loads in flight per thread (0 -> 4, 1 -> 8, 2 -> 16)
    const int per_thread = 256, grid = h->n_sms * 16, unroll = mode >> 1;
    mode &= 1;
    auto launch = [&]() {
        if (unroll == 2) sck::k_probe_stream<16><<<grid, 256, 0, h->stream>>>(tab.as<float4>(), n4 - 1, per_thread, mode, sink.as<float>());
        else if (unroll == 1) sck::k_probe_stream<8><<<grid, 256, 0, h->stream>>>(tab.as<float4>(), n4 - 1, per_thread, mode, sink.as<float>());
        else sck::k_probe_stream<4><<<grid, 256, 0, h->stream>>>(tab.as<float4>(), n4 - 1, per_thread, mode, sink.as<float>());
    };
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int w = 0; w < 2; w++) launch();
    cudaEventRecord(a, h->stream);
    for (int i = 0; i < iters; i++) launch();
    cudaEventRecord(b, h->stream);
    cudaError_t e = cudaStreamSynchronize(h->stream);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a); cudaEventDestroy(b);
    tab.release(); sink.release();
    if (e != cudaSuccess) return cuda_fail(h, e, "sc_probe_stream");
    *gbps = (double)grid * 256 * per_thread * 16.0 * iters / (ms * 1e6);
    return SC_OK;
}

int sc_set_profiling(sc_handle* h, int on) {
    if (!h) return SC_ERR_INVALID;
    h->profiling = on != 0;
    return SC_OK;
}

int sc_kernel_stats(sc_handle* h, int kernel_id, const char** name, double* ms, int64_t* launches, int reset) {
    if (!h || kernel_id < 0) return SC_ERR_INVALID;
    if (kernel_id >= K_COUNT) return 1;  // past the end
    if (name) *name = kKernelNames[kernel_id];
    if (ms) *ms = h->kernel_ms[kernel_id];
    if (launches) *launches = h->kernel_launches[kernel_id];
    if (reset) { h->kernel_ms[kernel_id] = 0; h->kernel_launches[kernel_id] = 0; }
    return SC_OK;
}

int sc_last_counters(sc_handle* h, sc_counters* counters, int nframes) {
    if (!h || !counters || nframes > h->last_nframes) return SC_ERR_INVALID;
    fill_counters(h, h->h_stage.as<unsigned long long>(), nframes, counters);
    return SC_OK;
}

#define SC_EAGER_DETS 16384u   // detections downloaded with the counters; more than that costs collect one extra copy

int sc_detect_submit(sc_handle* h, const uint8_t* const* frames, int nframes, int W, int H, int stride, const sc_detect_params* params, size_t cap,
                     int* ticket) {
    if (!h || !frames || nframes < 1 || !ticket || stride < W) return fail(h, SC_ERR_INVALID, "bad arguments");
    SC_CUDA(h, cudaSetDevice(h->device));
    int ti = -1;
    for (int i = 0; i < 2 && ti < 0; i++)
        if (!h->tickets[i].busy) ti = i;
    if (ti < 0) return fail(h, SC_ERR_STATE, "two batches are already in flight: collect one first");
    const bool other_busy = h->tickets[ti ^ 1].busy;
    const sc_detect_params prm = params ? *params : default_params();
    const bool same_plan = h->have_plan && h->plan.W == W && h->plan.H == H && same_params(prm, h->pparams);
    if (other_busy && (!same_plan || nframes > h->int_frames))
        return fail(h, SC_ERR_STATE, "a batch in flight uses another plan or smaller buffers: collect it first");
    int rc = ensure_plan(h, W, H, prm);
    if (rc != SC_OK) return rc;
    rc = ensure_group_buffers(h, nframes, false);
    if (rc != SC_OK) return rc;
    sc_handle::Ticket& t = h->tickets[ti];
    if (!t.done) SC_CUDA(h, cudaEventCreateWithFlags(&t.done, cudaEventDisableTiming));
    t.det_cap = (uint32_t)std::min<size_t>(std::max<size_t>(cap, 1), 0xffffffffu);
    if (prm.group_threshold > 0) t.det_cap = std::max(t.det_cap, 1u << 16);  // raw windows feed the grouping; `cap` counts objects
    t.eager = std::min(t.det_cap, SC_EAGER_DETS);
    t.nframes = nframes;
    const size_t cbytes = (size_t)nframes * SC_CNT_STRIDE * 8;
    SC_CUDA(h, t.d_img.ensure(align256((size_t)h->int_frames * W * H)));
    SC_CUDA(h, t.d_det.ensure((size_t)t.det_cap * sizeof(sc_detection)));
    SC_CUDA(h, t.d_cnt.ensure(256));
    SC_CUDA(h, t.d_counters.ensure(cbytes));
    SC_CUDA(h, t.h_out.ensure(cbytes + 16 + (size_t)t.eager * sizeof(sck::ScGroupOut)));
    SC_CUDA(h, cudaMemsetAsync(t.d_counters.p, 0, cbytes, h->stream));
    SC_CUDA(h, cudaMemsetAsync(t.d_cnt.p, 0, 4, h->stream));
    for (int i0 = 0; i0 < nframes; i0 += h->int_frames) {
        const int ni = std::min(h->int_frames, nframes - i0);
        // the ticket's image buffer is free when the call starts (its previous batch was collected); later super-groups of
        // the same call reuse it and must wait for the integral that still reads it
        rc = run_supergroup(h, frames + i0, stride, nullptr, ni, i0, t.d_det.as<sc_detection>(), t.det_cap, t.d_cnt.as<uint32_t>(),
                            t.d_counters.as<unsigned long long>() + (size_t)i0 * SC_CNT_STRIDE, t.d_img.as<uint8_t>(), i0 > 0, other_busy);
        if (rc != SC_OK) return rc;
    }
    unsigned char* ho = t.h_out.as<unsigned char>();
    SC_CUDA(h, cudaMemcpyAsync(ho, t.d_counters.p, cbytes, cudaMemcpyDeviceToHost, h->stream));
    h->d2h_bytes += cbytes;
    t.group_thr = prm.group_threshold > 0 ? prm.group_threshold : 0;
    t.group_eps = prm.group_eps;
    if (t.group_thr > 0) {
        // groupRectangles on the device: per-frame segments of the raw windows, one CTA per frame (sc_kernels.cuh)
        const size_t seg_bytes = align256((size_t)t.det_cap * sizeof(sc_detection));
        const size_t tab_bytes = align256(((size_t)3 * nframes + 8) * 4);
        const size_t out_bytes = align256((size_t)t.det_cap * sizeof(sck::ScGroupOut));
        SC_CUDA(h, t.d_grp.ensure(seg_bytes + tab_bytes + out_bytes));
        unsigned char* gb = t.d_grp.as<unsigned char>();
        sck::ScDetOut* seg = reinterpret_cast<sck::ScDetOut*>(gb);
        uint32_t* per_frame = reinterpret_cast<uint32_t*>(gb + seg_bytes);   // [nframes] counts | [nframes + 1] offsets | [nframes] fill | out_count, overflow
        uint32_t* offsets = per_frame + nframes;
        uint32_t* fill = offsets + nframes + 1;
        uint32_t* flags = fill + nframes;                                    // [0] overflow, [1] out_count
        sck::ScGroupOut* gout = reinterpret_cast<sck::ScGroupOut*>(gb + seg_bytes + tab_bytes);
        SC_CUDA(h, cudaMemsetAsync(per_frame, 0, tab_bytes, h->stream));
        const sck::ScDetOut* det = reinterpret_cast<const sck::ScDetOut*>(t.d_det.p);
        const size_t gsmem = (size_t)SC_GROUP_MAX * (8 + 8 + 6 * 4);
        if (!h->group_attr_set) {  // per handle: the attribute belongs to the handle's device
            SC_CUDA(h, cudaFuncSetAttribute(sck::k_group_frames, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
            h->group_attr_set = true;
        }
        {
            KernelSpan ks(h, K_GROUP);  // the three small table kernels ride in the same span
            sck::k_group_count<<<h->n_sms * 2, 256, 0, h->stream>>>(det, t.d_cnt.as<uint32_t>(), t.det_cap, 0, nframes, per_frame);
            sck::k_group_offsets<<<1, 32, 0, h->stream>>>(per_frame, nframes, offsets, fill, flags);
            sck::k_group_scatter<<<h->n_sms * 2, 256, 0, h->stream>>>(det, t.d_cnt.as<uint32_t>(), t.det_cap, 0, nframes, offsets, fill, seg);
            sck::k_group_frames<<<nframes, 256, gsmem, h->stream>>>(seg, offsets, 0, t.group_thr, t.group_eps, gout, flags + 1, t.det_cap);
        }
        h->launches += 3;
        SC_CUDA(h, cudaGetLastError());
        SC_CUDA(h, cudaMemcpyAsync(ho + cbytes, flags, 8, cudaMemcpyDeviceToHost, h->stream));  // overflow, object count
        SC_CUDA(h, cudaMemcpyAsync(ho + cbytes + 8, t.d_cnt.p, 4, cudaMemcpyDeviceToHost, h->stream));
        SC_CUDA(h, cudaMemcpyAsync(ho + cbytes + 16, gout, (size_t)t.eager * sizeof(sck::ScGroupOut), cudaMemcpyDeviceToHost, h->stream));
        h->d2h_bytes += 12 + (unsigned long long)t.eager * sizeof(sck::ScGroupOut);
    } else {
        SC_CUDA(h, cudaMemcpyAsync(ho + cbytes, t.d_cnt.p, 4, cudaMemcpyDeviceToHost, h->stream));
        SC_CUDA(h, cudaMemcpyAsync(ho + cbytes + 16, t.d_det.p, (size_t)t.eager * sizeof(sc_detection), cudaMemcpyDeviceToHost, h->stream));
        h->d2h_bytes += 4 + (unsigned long long)t.eager * sizeof(sc_detection);
    }
    SC_CUDA(h, cudaEventRecord(t.done, h->stream));
    t.busy = true;
    *ticket = ti;
    return SC_OK;
}

int sc_detect_collect(sc_handle* h, int ticket, sc_detection* out, size_t cap, size_t* n, sc_counters* counters) {
    if (!h || ticket < 0 || ticket > 1 || !n || (cap && !out)) return fail(h, SC_ERR_INVALID, "bad arguments");
    sc_handle::Ticket& t = h->tickets[ticket];
    if (!t.busy) return fail(h, SC_ERR_STATE, "ticket is not in flight");
    SC_CUDA(h, cudaSetDevice(h->device));
    cudaError_t e = cudaEventSynchronize(t.done);
    t.busy = false;
    if (e != cudaSuccess) return cuda_fail(h, e, "sc_detect_collect");
    if (!h->tickets[ticket ^ 1].busy) drain_spans(h);
    const size_t cbytes = (size_t)t.nframes * SC_CNT_STRIDE * 8;
    const unsigned char* ho = t.h_out.as<unsigned char>();
    if (counters) fill_counters(h, reinterpret_cast<const unsigned long long*>(ho), t.nframes, counters);
    // keep sc_last_counters working for the host path as well
    SC_CUDA(h, h->h_stage.ensure(cbytes + 16));
    memcpy(h->h_stage.p, ho, cbytes);
    h->last_nframes = t.nframes;
    if (t.group_thr > 0) {
        uint32_t fl[3];
        memcpy(fl, ho + cbytes, 12);  // overflow, objects, raw windows
        const uint32_t raw = fl[2];
        if (raw > t.det_cap) { *n = raw; return fail(h, SC_ERR_CAPACITY, "raw detection buffer too small for grouping"); }
        if (fl[0]) {
            // a frame holds more raw windows than one CTA groups (SC_GROUP_MAX): this batch is grouped on the host
            std::vector<sc_detection> rawd(raw);
            if (raw) SC_CUDA(h, cudaMemcpy(rawd.data(), t.d_det.p, (size_t)raw * sizeof(sc_detection), cudaMemcpyDeviceToHost));
            h->d2h_bytes += (unsigned long long)raw * sizeof(sc_detection);
            sort_detections(rawd.data(), raw);
            size_t k = 0, m = 0;
            while (k < raw) {
                size_t e2 = k;
                std::vector<sc_rect> r;
                std::vector<double> sc;
                while (e2 < raw && rawd[e2].frame == rawd[k].frame) { r.push_back(sc_rect{rawd[e2].x, rawd[e2].y, rawd[e2].l, rawd[e2].l}); sc.push_back(rawd[e2].score); e2++; }
                sc_host::group_rectangles(&r, &sc, t.group_thr, t.group_eps);
                for (size_t i = 0; i < r.size(); i++, m++)
                    if (m < cap) out[m] = sc_detection{rawd[k].frame, r[i].x, r[i].y, r[i].w, sc[i]};
                k = e2;
            }
            *n = m;
            return m > cap ? fail(h, SC_ERR_CAPACITY, "detection buffer too small") : SC_OK;
        }
        const uint32_t objs = fl[1];
        *n = objs;
        if (objs > cap || objs > t.det_cap) return fail(h, SC_ERR_CAPACITY, "detection buffer too small");
        std::vector<sck::ScGroupOut> g(objs);
        const uint32_t k = std::min(objs, t.eager);
        if (k) memcpy(g.data(), ho + cbytes + 16, (size_t)k * sizeof(sck::ScGroupOut));
        if (objs > k) {
            const size_t seg_bytes = align256((size_t)t.det_cap * sizeof(sc_detection)), tab_bytes = align256(((size_t)3 * t.nframes + 8) * 4);
            SC_CUDA(h, cudaMemcpy(g.data() + k, reinterpret_cast<const sck::ScGroupOut*>(t.d_grp.as<unsigned char>() + seg_bytes + tab_bytes) + k,
                                  (size_t)(objs - k) * sizeof(sck::ScGroupOut), cudaMemcpyDeviceToHost));
            h->d2h_bytes += (unsigned long long)(objs - k) * sizeof(sck::ScGroupOut);
        }
        std::sort(g.begin(), g.end(), [](const sck::ScGroupOut& a, const sck::ScGroupOut& b) { return a.frame != b.frame ? a.frame < b.frame : a.idx < b.idx; });
        for (uint32_t i = 0; i < objs; i++) out[i] = sc_detection{g[i].frame, g[i].x, g[i].y, g[i].w, g[i].score};
        return SC_OK;
    }
    uint32_t found = 0;
    memcpy(&found, ho + cbytes, 4);
    *n = found;
    if (found > cap || found > t.det_cap) return fail(h, SC_ERR_CAPACITY, "detection buffer too small");
    if (found) {
        const uint32_t k = std::min(found, t.eager);
        memcpy(out, ho + cbytes + 16, (size_t)k * sizeof(sc_detection));
        if (found > k) {
            SC_CUDA(h, cudaMemcpy(out + k, t.d_det.as<sc_detection>() + k, (size_t)(found - k) * sizeof(sc_detection), cudaMemcpyDeviceToHost));
            h->d2h_bytes += (unsigned long long)(found - k) * sizeof(sc_detection);
        }
        sort_detections(out, found);
    }
    return SC_OK;
}

int sc_transfer_bytes(sc_handle* h, uint64_t* h2d, uint64_t* d2h, int reset) {
    if (!h) return SC_ERR_INVALID;
    if (h2d) *h2d = h->h2d_bytes;
    if (d2h) *d2h = h->d2h_bytes;
    if (reset) { h->h2d_bytes = 0; h->d2h_bytes = 0; }
    return SC_OK;
}

int sc_detect(sc_handle* h, const uint8_t* const* frames, int nframes, int W, int H, int stride, const sc_detect_params* params,
              sc_detection* out, size_t cap, size_t* n, sc_counters* counters) {
    if (!h || !frames || nframes < 1 || !n || (cap && !out) || stride < W) return fail(h, SC_ERR_INVALID, "bad arguments");
    int ticket = -1;
    int rc = sc_detect_submit(h, frames, nframes, W, H, stride, params, cap, &ticket);
    if (rc != SC_OK) return rc;
    return sc_detect_collect(h, ticket, out, cap, n, counters);
}

// Hard-negative mining (next row N2): DenseSURFFeatureExtractor::FillNegSamples (DenseSURFFeatureExtractor.cpp:124-195) in its
// single-thread order.  Per image: every window of the scale ladder on a 10-pixel lattice; a window is taken when `first`
// or when the loaded cascade accepts it (CascadeClassifier::Predict == every stage score >= theta: the scan with the
// prefilter and the stride rule off); for a taken window the descriptors of ALL pool patches projected into it are
// appended to X.  Stops when `need` samples are filled; *frames_used = index after the image that completed the fill
// (the reference's static idx, :165), or nframes.
int sc_mine_negatives(sc_handle* h, const uint8_t* const* frames, const int32_t* Ws, const int32_t* Hs, const int32_t* strides, int nframes, int first,
                      int need, float* X, int* filled, int* frames_used) {
    if (!h || !frames || !Ws || !Hs || !strides || nframes < 0 || need < 0 || !X || !filled || !frames_used)
        return fail(h, SC_ERR_INVALID, "bad arguments");
    if (!first && !h->have_cascade) return fail(h, SC_ERR_STATE, "no cascade loaded");
    const int tmpl = h->have_cascade ? h->tmpl : 40;
    std::vector<sc_rect> pool;
    sc_host::pool_patches(tmpl, tmpl, &pool);
    const int P = (int)pool.size();
    sc_detect_params prm = default_params();
    prm.base = tmpl; prm.step = 10; prm.scale = 1.1; prm.prefilter = -1; prm.skip_rule = 0;
    int have = 0, used = nframes;
    std::vector<sc_detection> wins;
    std::vector<sc_rect> rects;
    DevBuf d_pool, d_wins, d_X;
    constexpr int MINE_BATCH = 16;   // equally sized consecutive images scanned per sc_detect call
    int i = 0;
    while (i < nframes && have < need) {
        const int W = Ws[i], H = Hs[i];
        if (!frames[i] || W < tmpl || H < tmpl) { i++; continue; }  // :137-138
        if (first) {
            // the first fill takes every window of the ladder: no scan; per image IntegralImage + the descriptors of host-built rects
            wins.clear();
            std::vector<int> sides;
            sc_host::scale_ladder(W, H, tmpl, 1.1, &sides);
            for (int l : sides)
                for (int y = 0; y <= H - l; y += 10)
                    for (int x = 0; x + l <= W; x += 10) wins.push_back(sc_detection{0, x, y, l, 0.0});
            const int take = (int)std::min<size_t>(wins.size(), (size_t)(need - have));
            if (take > 0) {
                int rc = sc_integral(h, frames[i], W, H, strides[i], nullptr);
                if (rc != SC_OK) return rc;
                rects.resize((size_t)take * P);
                for (int k = 0; k < take; k++)
                    for (int p = 0; p < P; p++) {
                        sc_rect r = sc_host::project_patch(tmpl, wins[k].l, pool[p]);  // ProjectPatches(win, patches, new_patches), :158
                        r.x += wins[k].x; r.y += wins[k].y;
                        rects[(size_t)k * P + p] = r;
                    }
                rc = sc_features(h, rects.data(), take * P, X + (size_t)have * P * 32);
                if (rc != SC_OK) return rc;
                have += take;
            }
            if (have == need) used = i + 1;
            i++;
            continue;
        }
        // a batch: the run of consecutive usable images of this size (the reference's order is kept: images are consumed in
        // order and the fill may stop inside the batch -- the images behind that point were scanned for nothing, never used)
        int nb = 1;
        while (nb < MINE_BATCH && i + nb < nframes && frames[i + nb] && Ws[i + nb] == W && Hs[i + nb] == H && strides[i + nb] == strides[i]) nb++;
        size_t cap = 4096, n = 0;
        for (;;) {
            wins.resize(cap);
            const int rc = sc_detect(h, frames + i, nb, W, H, strides[i], &prm, wins.data(), cap, &n, nullptr);
            if (rc == SC_ERR_CAPACITY) { cap = n + 16; continue; }
            if (rc != SC_OK) return rc;
            break;
        }
        wins.resize(n);  // sorted by (frame, l, y, x): per image the reference's loop order (:146-156)
        // the scan's integral images of the batch are still in d_S (one super-group) unless the frames are so large that the batch was split
        const bool resident = nb <= h->int_frames && h->have_plan && h->plan.W == W && h->plan.H == H;
        std::vector<sck::ScMineWin> mw;
        size_t k0 = 0;
        for (int f = 0; f < nb && have + (int)mw.size() < need; f++) {
            size_t k1 = k0;
            while (k1 < n && wins[k1].frame == f) k1++;
            const int take = (int)std::min<size_t>(k1 - k0, (size_t)(need - have - (int)mw.size()));
            for (int k = 0; k < take; k++) mw.push_back(sck::ScMineWin{f, wins[k0 + k].x, wins[k0 + k].y, wins[k0 + k].l});
            if (have + (int)mw.size() == need) used = i + f + 1;
            k0 = k1;
        }
        if (!mw.empty()) {
            SC_CUDA(h, cudaSetDevice(h->device));
            if (resident) {
                const size_t nw = mw.size();
                SC_CUDA(h, d_pool.ensure((size_t)P * sizeof(sc_rect)));
                SC_CUDA(h, d_wins.ensure(nw * sizeof(sck::ScMineWin)));
                SC_CUDA(h, d_X.ensure(nw * P * 32 * sizeof(float)));
                SC_CUDA(h, cudaMemcpyAsync(d_pool.p, pool.data(), (size_t)P * sizeof(sc_rect), cudaMemcpyHostToDevice, h->stream));
                SC_CUDA(h, cudaMemcpyAsync(d_wins.p, mw.data(), nw * sizeof(sck::ScMineWin), cudaMemcpyHostToDevice, h->stream));
                const long long threads = (long long)nw * P;
                sck::k_mine_descriptors<<<(unsigned)((threads + 127) / 128), 128, 0, h->stream>>>(h->d_S.as<float4>(), h->plan.lay, d_wins.as<sck::ScMineWin>(), (int)nw,
                                                                                           d_pool.as<int4>(), P, tmpl, d_X.as<float>());
                h->launches++;
                SC_CUDA(h, cudaGetLastError());
                SC_CUDA(h, cudaMemcpyAsync(X + (size_t)have * P * 32, d_X.p, nw * P * 32 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
                SC_CUDA(h, cudaStreamSynchronize(h->stream));
            } else {
                // very large frames: per image, through the hook layout
                size_t done = 0;
                while (done < mw.size()) {
                    const int f = mw[done].slot;
                    size_t e2 = done;
                    while (e2 < mw.size() && mw[e2].slot == f) e2++;
                    int rc = sc_integral(h, frames[i + f], W, H, strides[i], nullptr);
                    if (rc != SC_OK) return rc;
                    rects.resize((e2 - done) * P);
                    for (size_t k = done; k < e2; k++)
                        for (int p = 0; p < P; p++) {
                            sc_rect r = sc_host::project_patch(tmpl, mw[k].l, pool[p]);
                            r.x += mw[k].x; r.y += mw[k].y;
                            rects[(k - done) * P + p] = r;
                        }
                    rc = sc_features(h, rects.data(), (int)((e2 - done) * P), X + ((size_t)have + done) * P * 32);
                    if (rc != SC_OK) return rc;
                    done = e2;
                }
            }
            have += (int)mw.size();
        }
        i += nb;
    }
    d_pool.release(); d_wins.release(); d_X.release();
    *filled = have;
    *frames_used = used;
    return SC_OK;
}

int sc_group_rectangles(const sc_rect* rects, const double* scores, int n, int group_threshold, double eps, sc_rect* out_rects,
                        double* out_scores, int cap) {
    if (n < 0 || (n && (!rects || !scores))) return SC_ERR_INVALID;
    std::vector<sc_rect> r(rects, rects + n);
    std::vector<double> s(scores, scores + n);
    sc_host::group_rectangles(&r, &s, group_threshold, eps);
    for (int i = 0; i < (int)r.size() && i < cap; i++) { out_rects[i] = r[i]; out_scores[i] = s[i]; }
    return (int)r.size();
}

}  // extern "C"

#include "sc_comm.inc"
