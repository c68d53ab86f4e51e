/* surfcascade.h -- C-ABI of the B200-native SURF-cascade detection path.
 *
 * The reference (mrgloom/SurfCascade) has no plugin/FFI layer: its boundary is the C++ class surface
 * DenseSURFFeatureExtractor / CascadeClassifier / Model plus the detect loop inlined in main().  This header
 * is the thin C layer those (kept) C++ classes call; every entry point names the reference code it replaces
 * (paths relative to /root/reference/ObjDetector).  Plain pointers and sizes only; no exceptions cross it.
 *
 * Conventions: every function returns 0 on success or a negative sc_status; sc_last_error() returns a
 * human-readable reason for the last failure on that handle.  A handle owns one CUDA device, one stream and
 * all device buffers; it is thread-compatible (one thread at a time), not thread-safe.  Host buffers belong to
 * the caller.  There is NO CPU fallback: sc_create fails when no CUDA device is usable.
 */
#ifndef SURFCASCADE_H
#define SURFCASCADE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SC_DIM 32         /* descriptor length: 4 cells x 8 bins (DenseSURFFeatureExtractor.h:32-33,48) */
#define SC_MAX_STAGES 16  /* reference trains at most 10 (CascadeClassifier/CascadeClassifier.h:20) */
#define SC_MAX_SCALES 64

typedef enum {
    SC_OK = 0,
    SC_ERR_INVALID = -1,   /* bad argument */
    SC_ERR_CUDA = -2,      /* CUDA runtime failure (message in sc_last_error) */
    SC_ERR_STATE = -3,     /* call order: no cascade / no integral yet */
    SC_ERR_CAPACITY = -4,  /* output buffer too small; *n holds the required count */
    SC_ERR_IO = -5,        /* model file unreadable or malformed */
    SC_ERR_NOMEM = -6
} sc_status;

typedef struct sc_handle sc_handle;

typedef struct { int32_t x, y, w, h; } sc_rect;   /* cv::Rect */

/* Flattened cascade: what the detect loop reads out of a loaded model.  Replaces the object graph built by
 * Model::Load (Model.cpp:97-193) + CascadeClassifier::GetFittedPatchIndexes (CascadeClassifier/CascadeClassifier.cpp:83-91)
 * + the dense_patches[patch_index] lookup at ObjDetector.cpp:108-130. */
typedef struct {
    int32_t tmpl;            /* model template side; 40 at ObjDetector.cpp:112 */
    int32_t n_stages;
    const float* theta;      /* [n_stages]   StageClassifier::theta (StageClassifier.h:24) */
    const int32_t* n_weak;   /* [n_stages]   weak classifiers per stage */
    const sc_rect* rects;    /* [sum n_weak] template rect of each weak classifier's patch */
    const float* w;          /* [sum n_weak][33] LogisticRegression::w, already float32 (Model.cpp:172-175) */
    const double* bias;      /* [sum n_weak] liblinear model bias (Model.cpp:169) */
} sc_cascade_desc;

/* Scan parameters; the reference hard-codes all of them (SURVEY.md section 5 "Config / flags"). */
typedef struct {
    int32_t base;              /* base window side: literal 70 at ObjDetector.cpp:104,174,180; BASELINE uses 40 */
    int32_t step;              /* 0 -> base > 20 ? base / 20 : 1 (ObjDetector.cpp:139) */
    double scale;              /* 1.1 (ObjDetector.cpp:174,180) */
    int32_t prefilter;         /* 6 (ObjDetector.cpp:188); negative disables the prefilter */
    int32_t skip_rule;         /* 1 -> reproduce the adaptive x stride `multi` (ObjDetector.cpp:186,214-217) */
    int32_t force_all_stages;  /* 1 -> no early reject: every stage of every window (stress mode, not in the reference) */
    /* Single large frames over several GPUs: band_count > 1 restricts the scan to band band_index of
     * every scale's lattice rows [ny * i / count, ny * (i + 1) / count).  The reference's stride chain never crosses a
     * row, so the union of the bands' detections and the sum of their counters are the full scan's.  0 / 0 = whole frame. */
    int32_t band_index, band_count;
    /* group_threshold > 0: sc_detect / sc_detect_collect return GROUPED objects instead of raw windows --
     * cv::groupRectangles(wins, weights, scores, group_threshold, group_eps) of every frame's raw windows
     * (ObjDetector.cpp:224-225 uses 2, 0.2), computed on the device: {frame, x, y, l = mean side, score = best raw score},
     * per frame in the order groupRectangles emits them for the (l, y, x)-sorted window list.  0 = raw windows. */
    int32_t group_threshold;
    int32_t reserved;
    double group_eps;
} sc_detect_params;

/* One raw (ungrouped) detection: wins.push_back(win) / scores.push_back(score), ObjDetector.cpp:207-208. */
typedef struct {
    int32_t frame, x, y, l;
    double score;              /* (last stage score + n_stages + 1) / n_stages, ObjDetector.cpp:201 */
} sc_detection;

/* Per-frame work counters (same definitions as the oracle's). */
typedef struct {
    int64_t grid;              /* window origins on the step lattice, all scales */
    int64_t visited;           /* windows the reference's adaptive stride actually evaluates */
    int64_t prefilter_pass;    /* visited windows passing sum(win) > area*6 */
    int64_t weak_evals;        /* weak-classifier evaluations the reference performs on visited windows */
    int64_t raw;               /* raw detections */
    int64_t evaluated;         /* grid windows this implementation evaluated (>= visited) */
    int64_t reach[SC_MAX_STAGES]; /* visited windows reaching stage s */
} sc_counters;

/* ---- lifetime ------------------------------------------------------------------------------------- */
int sc_create(int device, sc_handle** out);
void sc_destroy(sc_handle* h);
const char* sc_last_error(const sc_handle* h);
const char* sc_version(void);
/* Checked build only (surfcascade_b200/build.py --checked -> libsurfcascade_b200_checked.so, -DSC_CHECKED): every gather from an
 * integral-image plane is range-tested on the device before it is issued; returns the number of violations counted so far
 * (process-wide: use ONE handle per process with this build), -1 in the normal build.  Stands in for compute-sanitizer's
 * memcheck on the gather addresses, which is closed on the B200 pool. */
int sc_checked_violations(int reset);

/* ---- model ---------------------------------------------------------------------------------------- */
/* Replaces ObjDetector.cpp:108-130 (pool + Model::Load + fitted patches), given an already flattened cascade. */
int sc_set_cascade(sc_handle* h, const sc_cascade_desc* desc);
/* Model::Load (Model.cpp:97-193) on a libconfig model.cfg, ExtractPatches (DenseSURFFeatureExtractor.cpp:49-63) on a
 * tmpl x tmpl template, flatten and upload. */
int sc_load_model(sc_handle* h, const char* model_cfg_path, int tmpl);
/* Host-only: Model::Load + flatten without touching a device.  Returns the number of stages (or a negative sc_status);
 * *total_weak receives the weak classifier count.  A cascade with more than max_stages stages or max_weak weak
 * classifiers returns SC_ERR_CAPACITY (nothing truncated is handed out; *total_weak still tells the size needed). */
int sc_model_flatten(const char* model_cfg_path, int tmpl, float* theta, int32_t* n_weak, int max_stages,
                     sc_rect* rects, int32_t* patch_index, float* w /* [max_weak][33] */, double* bias, int max_weak, int* total_weak);
/* Host-only: Model::Load followed by Model::Save (Model.cpp:21-95) to another path. */
int sc_model_resave(const char* in_cfg_path, const char* out_cfg_path);
/* Template pool: ExtractPatches. Returns the pool size (608 for tmpl 40); fills at most cap rects. */
int sc_pool_patches(int tmpl, sc_rect* out, int cap);
/* ProjectPatches (DenseSURFFeatureExtractor.cpp:486-508) for window side l at origin (0,0). */
int sc_project_patches(int tmpl, int l, const sc_rect* patches, int n, sc_rect* out);

/* ---- feature extraction (parity hooks and training-side extraction) ---------------------------------- */
/* DenseSURFFeatureExtractor::IntegralImage (DenseSURFFeatureExtractor.cpp:65-87, T2bFilter :199-349): gray u8
 * H x W (row stride in bytes) -> the (H+1) x (W+1) x 8 float32 interleaved integral, kept on the device for the
 * calls below and, when out is non-null, copied to the host. */
int sc_integral(sc_handle* h, const uint8_t* gray, int W, int H, int stride, float* out);
/* The same image through the memory layout the cascade scan reads for lattice step `step` (the kernels and layout of
 * sc_detect's integral stage; sc_integral keeps a step-1 layout for explicit rects), copied to the host in the reference's
 * interleaved form.  Parity hook for IntegralImage (DenseSURFFeatureExtractor.cpp:65-87) on the production path. */
int sc_integral_scan_layout(sc_handle* h, const uint8_t* gray, int W, int H, int stride, int step, float* out);
/* Parity hooks of the compact integer plane the stage-0 filter reads beside the float32 integral (16 bytes per corner;
 * csrc/sc_plan.h).  It restates the same IntegralImage (DenseSURFFeatureExtractor.cpp:65-87) as exact integers:
 * word k of pixel (X,Y) = (I[2k+1] << 16) + I[2k] mod 2^32, I[c] = exact integral of channel c.
 *   sc_integral_compact : that plane through the scan's kernels and layout for lattice step `step`, in pixel order,
 *                         out[(H+1)][(W+1)][4] uint32;
 *   sc_box_sums_compact : CalcFeature's 32 box sums (:379-415, before Normalize) of explicit rects read from the compact
 *                         plane of the current sc_integral image -- valid where every cell's sums are < 65536;
 *   sc_cell_bounds      : the per-frame certificate: for each cell edge ce[i] an upper bound of every ce x ce box sum of
 *                         every channel of the current sc_integral image (the filter uses the plane iff bound < 65536). */
int sc_integral_compact(sc_handle* h, const uint8_t* gray, int W, int H, int stride, int step, uint32_t* out);
int sc_box_sums_compact(sc_handle* h, const sc_rect* rects, int n, float* out /* [n][32] */);
int sc_cell_bounds(sc_handle* h, const int32_t* ce, int n, uint32_t* out /* [n] */);
/* DenseSURFFeatureExtractor::CalcFeature (+GetRectsFromPatch, Normalize; :360-457) on the current integral. */
int sc_features(sc_handle* h, const sc_rect* rects, int n, float* out /* [n][32] */);
/* DenseSURFFeatureExtractor::sum (:351-358). */
int sc_window_sum(sc_handle* h, const sc_rect* rects, int n, float* out /* [n] */);
/* GentleAdaboost::Predict2 (GentleAdaboost.cpp:247-261) of every stage on explicit windows {x,y,l} of the
 * current integral, no early exit: out[n][n_stages]. */
int sc_stage_scores(sc_handle* h, const int32_t* wins /* [n][3] */, int n, float* out);

/* LogisticRegression::Predict (CascadeClassifier/LogisticRegression.cpp:46-68) for n independent
 * (weights, descriptor) pairs held in host memory: w [n][33], bias [n], x [n][32] -> out [n]. */
int sc_weak_predict(sc_handle* h, const float* w, const double* bias, const float* x, int n, float* out);
/* GentleAdaboost::Predict2 (GentleAdaboost.cpp:247-261) on explicit descriptors: the n pairs are one stage's
 * weak classifiers in order; *out = float32 running sum of their probabilities / n. */
int sc_stage_predict(sc_handle* h, const float* w, const double* bias, const float* x, int n, float* out);

/* ---- training-side pool evaluation (BASELINE config 5) ------------------------------------ */
#define SC_POOL_HIST_BINS 21  /* 20 AUC thresholds (StageClassifier.cpp:59) + "below all" */
/* Candidate scoring of one boosting round: GentleAdaboost.cpp:145-148 -> StageClassifier::Evaluate
 * (StageClassifier.cpp:35-70) over GentleAdaboost::Predict (GentleAdaboost.cpp:233-245).  For every pool patch k the
 * candidate weak classifier (Wcand[k], bias[k]) is appended to the T classifiers already chosen:
 *   prob_n = float(prior_sum[n] + p_k(X[n][k])) / (T + 1),   auc[k] = trapezoid AUC over the 20 float thresholds.
 * X [N][P][32] descriptors, labels [N] (non-zero = positive), Wcand [P][33], bias [P], prior_sum [N] or null, auc [P].
 * All buffers in host memory. */
int sc_pool_eval(sc_handle* h, const float* X, int N, int P, const uint8_t* labels, const float* Wcand, const double* bias,
                 const float* prior_sum, int T, float* auc);
/* The streaming half on device-resident shards (X, labels, prior_sum in device memory): adds this shard's level
 * histograms into d_hist [P][2][SC_POOL_HIST_BINS] (uint32, caller-zeroed; class 1 = positive).  Asynchronous. */
int sc_pool_hist_device(sc_handle* h, const float* d_X, int N, int P, const uint8_t* d_labels, const float* Wcand, const double* bias,
                        const float* d_prior_sum, int T, uint32_t* d_hist);
/* The epilogue: AUC of every candidate from (all-reduced) histograms in device memory; auc [P] in host memory. */
int sc_pool_auc_device(sc_handle* h, const uint32_t* d_hist, int P, int64_t n_pos, int64_t n_neg, float* auc);

/* ---- training-side descriptor extraction (next row N3) --------------------------------------------------------- */
/* ExtractNextImageFeatures (DenseSURFFeatureExtractor.cpp:104-120): IntegralImage + CalcFeature of every template-pool
 * patch (sc_pool_patches order) for N template-sized gray samples imgs [N][tmpl][tmpl]: X [N][P][32], the training
 * matrix sc_pool_eval consumes.  The _device variant keeps samples and X in device memory and is asynchronous. */
int sc_extract_pool_features(sc_handle* h, const uint8_t* imgs, int N, int tmpl, float* X);
int sc_extract_pool_features_device(sc_handle* h, const uint8_t* d_imgs, int N, int tmpl, float* d_X);

/* ---- hard-negative mining (next row N2) ------------------------------------------------------------------------ */
/* DenseSURFFeatureExtractor::FillNegSamples (DenseSURFFeatureExtractor.cpp:124-195) in its single-thread order: per image
 * every window of the scale ladder on a 10-pixel lattice; a window is a sample when `first` is set or the loaded cascade
 * accepts it (CascadeClassifier::Predict, CascadeClassifier.cpp:58-67); a sample is the descriptors of all P pool patches
 * projected into the window, X [need][P][32].  Stops when `need` samples are filled: *filled <= need, *frames_used =
 * index after the image that completed the fill (the reference's `idx`), nframes when the images ran out. */
int sc_mine_negatives(sc_handle* h, const uint8_t* const* frames, const int32_t* W, const int32_t* H, const int32_t* stride, int nframes,
                      int first, int need, float* X, int* filled, int* frames_used);

/* ---- detection -------------------------------------------------------------------------------------- */
/* The detect path of ObjDetector.cpp:165,174-219 on a batch of equally sized gray frames held in HOST memory:
 * upload, integral, scan, adaptive-stride replay, download.  Detections are sorted by (frame, l, y, x).
 * counters may be null or point at nframes entries. */
int sc_detect(sc_handle* h, const uint8_t* const* frames, int nframes, int W, int H, int stride,
              const sc_detect_params* params, sc_detection* out, size_t cap, size_t* n, sc_counters* counters);
/* The same call split in two so that batches pipeline: submit enqueues upload + path + download of one batch and
 * returns a ticket (0 or 1) without waiting; collect waits for that batch and hands out what sc_detect would have.
 * Two batches may be in flight: the upload of batch k+1 and the download / sorting of batch k-1 then run under the
 * compute of batch k.  Host frame buffers must stay valid (and should be pinned) until the batch is collected. */
int sc_detect_submit(sc_handle* h, const uint8_t* const* frames, int nframes, int W, int H, int stride,
                     const sc_detect_params* params, size_t cap, int* ticket);
int sc_detect_collect(sc_handle* h, int ticket, sc_detection* out, size_t cap, size_t* n, sc_counters* counters);
/* Same work on frames already resident in DEVICE memory (d_frames: nframes x H x W contiguous u8), detections
 * left on the device, unsorted, in d_out (cap entries) with the count in *d_n (uint32).  Asynchronous on the
 * handle's stream; sc_sync waits.  counters (host, nframes entries or null) are valid after sc_sync. */
int sc_detect_device(sc_handle* h, const uint8_t* d_frames, int nframes, int W, int H,
                     const sc_detect_params* params, sc_detection* d_out, size_t cap, uint32_t* d_n);
int sc_sync(sc_handle* h);

/* ---- multi-GPU exchange: one process per GPU, frames sharded by the caller, NCCL over NVLink -------------
 * The reference has one host and no exchange; its counterpart here is the step before ObjDetector.cpp:224-231 (grouping and
 * output on one host): every rank's detection records travel to `root`.  NCCL is loaded at run time (libnccl.so.2, the copy
 * already in the process if there is one), so single-GPU users need no NCCL.
 *   sc_comm_unique_id : rank 0 creates the 128-byte NCCL id; the caller hands it to the other ranks (file, socket, MPI, ...).
 *   sc_comm_init      : collective over all ranks; the handle's device must be this rank's GPU.  world == 1 is allowed.
 *   sc_gather_detections : collective.  This rank contributes n_local records at `local` (local_on_device: 0 = host memory;
 *       1 = device memory, possibly still being written by work on the handle's stream, which the exchange then waits for;
 *       2 = device memory known to be complete, nothing is waited for); their frame index becomes frame * frame_mul + frame_add
 *       on the way (round-robin sharding: frame_mul = world, frame_add = rank).  Exactly n_local records per rank move to root,
 *       with their count -- no fixed-size slices.  Default transport: copy engines over NVLink into a region of root's HBM that
 *       every rank mapped with CUDA IPC at sc_comm_init (records, then an 8-byte {sequence, count} header in stream order; root
 *       polls the headers, reads exactly `count` records per rank and acknowledges) -- no kernel runs, so the exchange is not
 *       queued behind the scan kernels of the next batch (NCCL's own kernels were: csrc/sc_comm.inc).  Fallback (no peer mapping,
 *       or SC_COMM_NCCL_ONLY=1): ncclAllGather of the counts, then grouped ncclSend / ncclRecv.  On root: out (host, cap
 *       records) receives the ranks' records in rank order, *n_out their number, per_rank[world] (optional) every rank's count;
 *       SC_ERR_CAPACITY (with *n_out = needed) if cap is too small.  Other ranks: out may be null, per_rank holds only their own
 *       count.  The exchange runs on the handle's own high-priority communication stream: scan work already enqueued for the
 *       next batch (sc_detect_submit / sc_detect_device) keeps the GPU busy meanwhile.  Records may be raw windows or, with
 *       sc_detect_params.group_threshold > 0, the grouped objects of whole frames -- then only final objects cross NVLink and
 *       root has no grouping left to do.  Use one root per communicator; a rank may contribute up to SC_COMM_SLOT_RECORDS
 *       (environment, default 131072) records per call on the peer path.
 *   sc_comm_destroy   : collective teardown (also done by sc_destroy). */
#define SC_COMM_ID_BYTES 128
int sc_comm_unique_id(void* id /* SC_COMM_ID_BYTES */);
int sc_comm_init(sc_handle* h, int rank, int world, const void* id /* SC_COMM_ID_BYTES */);
int sc_gather_detections(sc_handle* h, const sc_detection* local, size_t n_local, int local_on_device, int32_t frame_mul, int32_t frame_add,
                         int root, sc_detection* out, size_t cap, size_t* n_out, size_t* per_rank);
int sc_comm_destroy(sc_handle* h);

/* Counters of the last sc_detect_device batch (after sc_sync). */
int sc_last_counters(sc_handle* h, sc_counters* counters, int nframes);
/* Bytes the host-buffer detect entry points (sc_detect, sc_detect_submit / sc_detect_collect) have copied host -> device
 * (frames) and device -> host (counters, counts, detection records) since the handle was created or last reset: counted at
 * the copies themselves. */
int sc_transfer_bytes(sc_handle* h, uint64_t* h2d, uint64_t* d2h, int reset);
/* cudaStream_t of the handle, for callers that time or order work with CUDA events. */
void* sc_stream(sc_handle* h);
/* Number of kernel launches issued by this handle so far. */
int64_t sc_launch_count(const sc_handle* h);

/* Parity hook of the stage-0 certified fast filter (csrc/sc_kernels.cuh): for explicit windows {x, y, l} on the image of the
 * last sc_integral, the float sum of the fast weak-classifier outputs of stage 0, the float sum of the reference-arithmetic
 * outputs (GentleAdaboost::Predict2's accumulator, GentleAdaboost.cpp:247-261), and the distance budget the scan's
 * decision limits are built from.  Windows further than the budget from a threshold are decided by the fast sum; all
 * others are re-evaluated exactly, so the filter never changes a result. */
int sc_stage0_fast_check(sc_handle* h, const int32_t* wins, int n, float* fast_sum, float* exact_sum, double* margin);

/* Optional per-kernel timing (CUDA events on the handle's stream around every launch of the detect path).
 * sc_kernel_stats enumerates kernels by id 0,1,..; returns 1 past the last id.  Times are accumulated at sc_sync /
 * at the end of sc_detect. */
int sc_set_profiling(sc_handle* h, int on);
/* Measurement probe, no product role: GB/s of random 32-byte sector gathers (two 16-byte loads each, like one corner
 * fetch) from a zeroed device table of table_bytes.  Tables below ~100 MB stay L2-resident on B200. */
int sc_probe_gather(sc_handle* h, size_t table_bytes, int iters, double* gbps);
/* Measurement probe, no product role: GB/s of coalesced 16-byte loads (512 contiguous bytes per warp) streaming over a
 * device table of table_bytes; mode 0 = L2-only loads (ld.global.cg), 1 = L1-allocating loads.  L2 -> SM ceiling. */
int sc_probe_stream(sc_handle* h, size_t table_bytes, int iters, int mode, double* gbps);
int sc_kernel_stats(sc_handle* h, int kernel_id, const char** name, double* ms, int64_t* launches, int reset);

/* ---- host-side grouping (next row N1) ------------------------------------------------------------------ */
/* cv::groupRectangles(wins, weights = 0.., scores, groupThreshold, eps) as called at ObjDetector.cpp:224-225. */
int sc_group_rectangles(const sc_rect* rects, const double* scores, int n, int group_threshold, double eps,
                        sc_rect* out_rects, double* out_scores, int cap);

#ifdef __cplusplus
}
#endif
#endif
