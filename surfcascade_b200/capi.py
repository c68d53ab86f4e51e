"""ctypes binding to the C-ABI of libsurfcascade_b200.so (include/surfcascade.h).

Python is plumbing here (tests, bench.py, torch device memory / streams / torch.distributed); the product is the
shared library.  There is no fallback: a missing library or a missing GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SC_LIB") or os.path.join(HERE, "libsurfcascade_b200.so")  # SC_LIB: another build of the library (tuning variants)
SC_MAX_STAGES = 16

SC_OK, SC_ERR_INVALID, SC_ERR_CUDA, SC_ERR_STATE, SC_ERR_CAPACITY, SC_ERR_IO, SC_ERR_NOMEM = 0, -1, -2, -3, -4, -5, -6


class Rect(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("w", C.c_int32), ("h", C.c_int32)]


class CascadeDesc(C.Structure):
    _fields_ = [("tmpl", C.c_int32), ("n_stages", C.c_int32), ("theta", C.POINTER(C.c_float)), ("n_weak", C.POINTER(C.c_int32)),
                ("rects", C.POINTER(Rect)), ("w", C.POINTER(C.c_float)), ("bias", C.POINTER(C.c_double))]


class DetectParams(C.Structure):
    _fields_ = [("base", C.c_int32), ("step", C.c_int32), ("scale", C.c_double), ("prefilter", C.c_int32), ("skip_rule", C.c_int32),
                ("force_all_stages", C.c_int32), ("band_index", C.c_int32), ("band_count", C.c_int32),
                ("group_threshold", C.c_int32), ("reserved", C.c_int32), ("group_eps", C.c_double)]


class Counters(C.Structure):
    _fields_ = [("grid", C.c_int64), ("visited", C.c_int64), ("prefilter_pass", C.c_int64), ("weak_evals", C.c_int64), ("raw", C.c_int64),
                ("evaluated", C.c_int64), ("reach", C.c_int64 * SC_MAX_STAGES)]


DETECTION_DTYPE = np.dtype([("frame", "<i4"), ("x", "<i4"), ("y", "<i4"), ("l", "<i4"), ("score", "<f8")])
assert DETECTION_DTYPE.itemsize == 24

EXPORTS = ["sc_create", "sc_destroy", "sc_last_error", "sc_version", "sc_set_cascade", "sc_load_model", "sc_pool_patches", "sc_project_patches",
           "sc_integral", "sc_features", "sc_window_sum", "sc_stage_scores", "sc_weak_predict", "sc_stage_predict", "sc_detect",
           "sc_detect_device", "sc_sync", "sc_last_counters", "sc_stream", "sc_launch_count", "sc_group_rectangles",
           "sc_set_profiling", "sc_kernel_stats", "sc_model_flatten", "sc_model_resave", "sc_pool_eval", "sc_pool_hist_device",
           "sc_pool_auc_device", "sc_probe_gather", "sc_probe_stream", "sc_stage0_fast_check", "sc_detect_submit", "sc_detect_collect", "sc_extract_pool_features", "sc_extract_pool_features_device", "sc_mine_negatives", "sc_integral_scan_layout",
           "sc_integral_compact", "sc_box_sums_compact", "sc_cell_bounds",
           "sc_comm_unique_id", "sc_comm_init", "sc_gather_detections", "sc_comm_destroy", "sc_transfer_bytes", "sc_checked_violations"]

_lib = None


def comm_unique_id() -> bytes:
    """128-byte NCCL id created by rank 0 (sc_comm_unique_id); hand it to the other ranks."""
    L = lib()
    buf = C.create_string_buffer(128)
    L.sc_comm_unique_id.argtypes = [C.c_char_p]
    rc = L.sc_comm_unique_id(buf)
    if rc != 0:
        raise SurfCascadeError(rc, "sc_comm_unique_id failed (NCCL not loadable?)")
    return buf.raw


class SurfCascadeError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"surfcascade error {code}: {msg}")
        self.code = code


def lib():
    """Load the shared library; raises if it has not been built (python -m surfcascade_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} not built: run `python surfcascade_b200/build.py` (needs nvcc). No CPU fallback exists.")
        L = C.CDLL(LIB_PATH)
        L.sc_last_error.restype = C.c_char_p
        L.sc_version.restype = C.c_char_p
        L.sc_stream.restype = C.c_void_p
        L.sc_launch_count.restype = C.c_int64
        L.sc_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.sc_destroy.argtypes = [C.c_void_p]
        L.sc_last_error.argtypes = [C.c_void_p]
        L.sc_stream.argtypes = [C.c_void_p]
        L.sc_launch_count.argtypes = [C.c_void_p]
        L.sc_set_cascade.argtypes = [C.c_void_p, C.POINTER(CascadeDesc)]
        L.sc_load_model.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.sc_pool_patches.argtypes = [C.c_int, C.c_void_p, C.c_int]
        L.sc_project_patches.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.sc_integral.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.sc_integral_scan_layout.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.sc_integral_compact.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.sc_box_sums_compact.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.sc_cell_bounds.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.sc_features.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.sc_window_sum.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.sc_stage_scores.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.sc_weak_predict.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.sc_stage_predict.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.sc_detect.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(DetectParams), C.c_void_p,
                                C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p]
        L.sc_detect_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(DetectParams), C.c_void_p, C.c_size_t, C.c_void_p]
        L.sc_sync.argtypes = [C.c_void_p]
        L.sc_last_counters.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.sc_set_profiling.argtypes = [C.c_void_p, C.c_int]
        L.sc_kernel_stats.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int]
        L.sc_group_rectangles.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_int]
        _lib = L
    return _lib


def params(base=40, step=0, scale=1.1, prefilter=6, skip_rule=True, force_all_stages=False, band_index=0, band_count=0, group_threshold=0,
           group_eps=0.2) -> DetectParams:
    """Defaults are BASELINE config 1/2: base 40 -> step 2, scale 1.1, prefilter 6, adaptive stride on."""
    return DetectParams(base, step, scale, prefilter, int(skip_rule), int(force_all_stages), band_index, band_count, group_threshold, 0, group_eps)


def counters_to_dict(c: Counters, n_stages: int) -> dict:
    return {"grid": c.grid, "visited": c.visited, "prefilter_pass": c.prefilter_pass, "weak_evals": c.weak_evals, "raw": c.raw,
            "evaluated": c.evaluated, "reach": [c.reach[i] for i in range(n_stages)]}


def model_flatten(path: str, tmpl: int = 40) -> dict:
    """Host-only Model::Load + flatten: {theta, n_weak, rects, patch_index, w, bias} as numpy arrays."""
    L = lib()
    L.sc_model_flatten.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                   C.POINTER(C.c_int)]
    ms, mw = SC_MAX_STAGES, 2048
    theta = np.zeros(ms, np.float32); n_weak = np.zeros(ms, np.int32); rects = np.zeros((mw, 4), np.int32); pidx = np.zeros(mw, np.int32)
    w = np.zeros((mw, 33), np.float32); bias = np.zeros(mw, np.float64)
    total = C.c_int(0)
    s = L.sc_model_flatten(path.encode(), tmpl, theta.ctypes.data, n_weak.ctypes.data, ms, rects.ctypes.data, pidx.ctypes.data, w.ctypes.data,
                           bias.ctypes.data, mw, C.byref(total))
    if s < 0:
        raise SurfCascadeError(s, f"cannot load {path}")
    k = total.value
    return {"theta": theta[:s].copy(), "n_weak": n_weak[:s].copy(), "rects": rects[:k].copy(), "patch_index": pidx[:k].copy(), "w": w[:k].copy(),
            "bias": bias[:k].copy()}


def model_resave(src: str, dst: str) -> None:
    L = lib()
    L.sc_model_resave.argtypes = [C.c_char_p, C.c_char_p]
    rc = L.sc_model_resave(src.encode(), dst.encode())
    if rc != SC_OK:
        raise SurfCascadeError(rc, f"cannot re-save {src} -> {dst}")


def pool_patches(tmpl: int = 40) -> np.ndarray:
    out = np.zeros((4096, 4), np.int32)
    n = lib().sc_pool_patches(tmpl, out.ctypes.data, 4096)
    return out[:n].copy()


def project_patches(tmpl: int, l: int, patches) -> np.ndarray:
    p = np.ascontiguousarray(patches, np.int32).reshape(-1, 4)
    out = np.zeros_like(p)
    rc = lib().sc_project_patches(tmpl, l, p.ctypes.data, len(p), out.ctypes.data)
    if rc != SC_OK:
        raise SurfCascadeError(rc, "sc_project_patches")
    return out


def group_rectangles(rects, scores, thr: int = 2, eps: float = 0.2):
    r = np.ascontiguousarray(rects, np.int32).reshape(-1, 4)
    s = np.ascontiguousarray(scores, np.float64)
    n = len(r)
    out_r = np.zeros((max(n, 1), 4), np.int32)
    out_s = np.zeros(max(n, 1), np.float64)
    m = lib().sc_group_rectangles(r.ctypes.data, s.ctypes.data, n, thr, eps, out_r.ctypes.data, out_s.ctypes.data, max(n, 1))
    if m < 0:
        raise SurfCascadeError(m, "sc_group_rectangles")
    return out_r[:m].copy(), out_s[:m].copy()


class Handle:
    """One sc_handle: a CUDA device, a stream and its device buffers."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        rc = lib().sc_create(device, C.byref(self._h))
        if rc != SC_OK:
            raise SurfCascadeError(rc, "sc_create failed: no usable CUDA device (this library has no CPU path)")
        self.device = device
        self.n_stages = 0

    def close(self):
        if self._h:
            lib().sc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != SC_OK:
            raise SurfCascadeError(rc, lib().sc_last_error(self._h).decode())

    @property
    def stream(self) -> int:
        return int(lib().sc_stream(self._h) or 0)

    @property
    def launch_count(self) -> int:
        return int(lib().sc_launch_count(self._h))

    # ---- model ----
    def load_model(self, path: str, tmpl: int = 40):
        self._check(lib().sc_load_model(self._h, path.encode(), tmpl))
        self.n_stages = -1

    def set_cascade(self, tmpl: int, theta, n_weak, rects, w, bias):
        theta = np.ascontiguousarray(theta, np.float32); n_weak = np.ascontiguousarray(n_weak, np.int32)
        rects = np.ascontiguousarray(rects, np.int32).reshape(-1, 4); w = np.ascontiguousarray(w, np.float32).reshape(-1, 33)
        bias = np.ascontiguousarray(bias, np.float64)
        d = CascadeDesc(tmpl, len(theta), theta.ctypes.data_as(C.POINTER(C.c_float)), n_weak.ctypes.data_as(C.POINTER(C.c_int32)),
                        C.cast(rects.ctypes.data, C.POINTER(Rect)), w.ctypes.data_as(C.POINTER(C.c_float)), bias.ctypes.data_as(C.POINTER(C.c_double)))
        self._check(lib().sc_set_cascade(self._h, C.byref(d)))
        self.n_stages = len(theta)

    # ---- parity hooks ----
    def integral(self, img: np.ndarray, want_output: bool = True):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        out = np.empty((h + 1, w + 1, 8), np.float32) if want_output else None
        self._check(lib().sc_integral(self._h, img.ctypes.data, w, h, w, out.ctypes.data if want_output else None))
        return out

    def integral_scan_layout(self, img: np.ndarray, step: int = 2) -> np.ndarray:
        """The integral image computed through the scan's own layout for lattice step `step` (parity hook of the detect path)."""
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        out = np.empty((h + 1, w + 1, 8), np.float32)
        self._check(lib().sc_integral_scan_layout(self._h, img.ctypes.data, w, h, w, step, out.ctypes.data))
        return out

    def integral_compact(self, img: np.ndarray, step: int = 2) -> np.ndarray:
        """The compact integer plane (csrc/sc_plan.h) through the scan's kernels and layout: uint32 [H+1][W+1][4]."""
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        out = np.empty((h + 1, w + 1, 4), np.uint32)
        self._check(lib().sc_integral_compact(self._h, img.ctypes.data, w, h, w, step, out.ctypes.data))
        return out

    def box_sums_compact(self, rects) -> np.ndarray:
        """CalcFeature's 32 box sums (before Normalize) of explicit rects, read from the compact plane of the last integral()."""
        r = np.ascontiguousarray(rects, np.int32).reshape(-1, 4)
        out = np.zeros((len(r), 32), np.float32)
        self._check(lib().sc_box_sums_compact(self._h, r.ctypes.data, len(r), out.ctypes.data))
        return out

    def cell_bounds(self, ces) -> np.ndarray:
        """Per cell edge: certified upper bound of every ce x ce box sum of every channel of the last integral()."""
        c = np.ascontiguousarray(ces, np.int32).reshape(-1)
        out = np.zeros(len(c), np.uint32)
        self._check(lib().sc_cell_bounds(self._h, c.ctypes.data, len(c), out.ctypes.data))
        return out

    def features(self, rects) -> np.ndarray:
        r = np.ascontiguousarray(rects, np.int32).reshape(-1, 4)
        out = np.zeros((len(r), 32), np.float32)
        self._check(lib().sc_features(self._h, r.ctypes.data, len(r), out.ctypes.data))
        return out

    def window_sum(self, rects) -> np.ndarray:
        r = np.ascontiguousarray(rects, np.int32).reshape(-1, 4)
        out = np.zeros(len(r), np.float32)
        self._check(lib().sc_window_sum(self._h, r.ctypes.data, len(r), out.ctypes.data))
        return out

    def stage_scores(self, wins, n_stages: int) -> np.ndarray:
        w = np.ascontiguousarray(wins, np.int32).reshape(-1, 3)
        out = np.zeros((len(w), n_stages), np.float32)
        self._check(lib().sc_stage_scores(self._h, w.ctypes.data, len(w), out.ctypes.data))
        return out

    def transfer_bytes(self, reset: bool = False) -> tuple[int, int]:
        """(host -> device, device -> host) bytes copied by the host-buffer detect entry points since creation / last reset."""
        L = lib()
        L.sc_transfer_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_int]
        a, b = C.c_uint64(0), C.c_uint64(0)
        self._check(L.sc_transfer_bytes(self._h, C.byref(a), C.byref(b), int(reset)))
        return int(a.value), int(b.value)

    # ---- multi-GPU exchange over NCCL (sc_comm_init / sc_gather_detections, include/surfcascade.h) ----
    def comm_init(self, rank: int, world: int, comm_id: bytes) -> None:
        L = lib()
        L.sc_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p]
        self._check(L.sc_comm_init(self._h, rank, world, comm_id))
        self._comm = (rank, world)

    def comm_destroy(self) -> None:
        L = lib()
        L.sc_comm_destroy.argtypes = [C.c_void_p]
        self._check(L.sc_comm_destroy(self._h))

    def gather_detections(self, local, frame_mul: int = 1, frame_add: int = 0, root: int = 0, cap: int = 1 << 20, device_ptr: int = 0, n_device: int = 0,
                          complete: bool = False, out=None):
        """Every rank's records to `root` (collective).  `local`: DETECTION_DTYPE array on the host, or device_ptr / n_device for
        records in device memory.  Returns (records, per_rank_counts) on root, (None, per_rank_counts) elsewhere."""
        rank, world = self._comm
        L = lib()
        L.sc_gather_detections.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int32, C.c_int32, C.c_int, C.c_void_p, C.c_size_t,
                                           C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        per = (C.c_size_t * world)()
        n_out = C.c_size_t(0)
        if device_ptr:
            src, n, on_dev = device_ptr, n_device, (2 if complete else 1)
        else:
            local = np.ascontiguousarray(local, DETECTION_DTYPE)
            src, n, on_dev = (local.ctypes.data if len(local) else None), len(local), 0
        if out is None:
            out = np.zeros(cap if rank == root else 0, DETECTION_DTYPE)
        elif rank == root:
            cap = len(out)
        rc = L.sc_gather_detections(self._h, src, n, on_dev, frame_mul, frame_add, root, out.ctypes.data if rank == root else None, cap, C.byref(n_out), per)
        if rc == SC_ERR_CAPACITY and rank == root:
            raise SurfCascadeError(rc, f"gather capacity {cap} < {n_out.value}")
        self._check(rc)
        return (out[:n_out.value] if rank == root else None), [int(v) for v in per]

    def stage0_fast_check(self, wins):
        """(fast_sum [n], exact_sum [n], margin) of stage 0 for windows {x, y, l} on the last integral()."""
        wins = np.ascontiguousarray(wins, np.int32).reshape(-1, 3)
        fs = np.zeros(len(wins), np.float32); es = np.zeros(len(wins), np.float32); m = C.c_double(0)
        L = lib()
        L.sc_stage0_fast_check.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        self._check(L.sc_stage0_fast_check(self._h, wins.ctypes.data, len(wins), fs.ctypes.data, es.ctypes.data, C.byref(m)))
        return fs, es, m.value

    def weak_predict(self, w, bias, x) -> np.ndarray:
        w = np.ascontiguousarray(w, np.float32).reshape(-1, 33); x = np.ascontiguousarray(x, np.float32).reshape(-1, 32)
        bias = np.ascontiguousarray(bias, np.float64).reshape(-1)
        out = np.zeros(len(w), np.float32)
        self._check(lib().sc_weak_predict(self._h, w.ctypes.data, bias.ctypes.data, x.ctypes.data, len(w), out.ctypes.data))
        return out

    def stage_predict(self, w, bias, x) -> float:
        w = np.ascontiguousarray(w, np.float32).reshape(-1, 33); x = np.ascontiguousarray(x, np.float32).reshape(-1, 32)
        bias = np.ascontiguousarray(bias, np.float64).reshape(-1)
        out = np.zeros(1, np.float32)
        self._check(lib().sc_stage_predict(self._h, w.ctypes.data, bias.ctypes.data, x.ctypes.data, len(w), out.ctypes.data))
        return float(out[0])

    # ---- training-side pool evaluation ----
    def pool_eval(self, X, labels, Wcand, bias, prior_sum=None, T: int = 0) -> np.ndarray:
        """AUC of every candidate weak classifier; X [N][P][32] host array, labels [N] (non-zero = positive)."""
        X = np.ascontiguousarray(X, np.float32)
        N, P, _ = X.shape
        labels = np.ascontiguousarray(labels, np.uint8).reshape(N)
        Wcand = np.ascontiguousarray(Wcand, np.float32).reshape(P, 33); bias = np.ascontiguousarray(bias, np.float64).reshape(P)
        ps = np.ascontiguousarray(prior_sum, np.float32).reshape(N) if prior_sum is not None else None
        auc = np.zeros(P, np.float32)
        L = lib()
        L.sc_pool_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        self._check(L.sc_pool_eval(self._h, X.ctypes.data, N, P, labels.ctypes.data, Wcand.ctypes.data, bias.ctypes.data,
                                   ps.ctypes.data if ps is not None else None, T, auc.ctypes.data))
        return auc

    def pool_hist_device(self, d_X: int, N: int, P: int, d_labels: int, Wcand, bias, d_prior: int | None, T: int, d_hist: int):
        Wcand = np.ascontiguousarray(Wcand, np.float32).reshape(P, 33); bias = np.ascontiguousarray(bias, np.float64).reshape(P)
        L = lib()
        L.sc_pool_hist_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        self._check(L.sc_pool_hist_device(self._h, d_X, N, P, d_labels, Wcand.ctypes.data, bias.ctypes.data, d_prior, T, d_hist))

    def pool_auc_device(self, d_hist: int, P: int, n_pos: int, n_neg: int) -> np.ndarray:
        auc = np.zeros(P, np.float32)
        L = lib()
        L.sc_pool_auc_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p]
        self._check(L.sc_pool_auc_device(self._h, d_hist, P, n_pos, n_neg, auc.ctypes.data))
        return auc

    def extract_pool_features(self, imgs: np.ndarray, tmpl: int = 40) -> np.ndarray:
        """imgs [N][tmpl][tmpl] u8 -> X [N][P][32] f32 (descriptors of every pool patch of every sample)."""
        imgs = np.ascontiguousarray(imgs, np.uint8)
        n = imgs.shape[0]
        P = len(pool_patches(tmpl))
        X = np.empty((n, P, 32), np.float32)
        L = lib()
        L.sc_extract_pool_features.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        self._check(L.sc_extract_pool_features(self._h, imgs.ctypes.data, n, tmpl, X.ctypes.data))
        return X

    def extract_pool_features_device(self, d_imgs: int, n: int, tmpl: int, d_X: int):
        L = lib()
        L.sc_extract_pool_features_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        self._check(L.sc_extract_pool_features_device(self._h, d_imgs, n, tmpl, d_X))

    def mine_negatives(self, frames, need: int, first: bool = False, tmpl: int = 40):
        """FillNegSamples over a list of u8 images: (X [filled][P][32], frames_used)."""
        frames = [np.ascontiguousarray(f, np.uint8) for f in frames]
        n = len(frames)
        P = len(pool_patches(tmpl))
        X = np.zeros((need, P, 32), np.float32)
        ptrs = (C.c_void_p * n)(*[f.ctypes.data for f in frames])
        Ws = (C.c_int32 * n)(*[f.shape[1] for f in frames]); Hs = (C.c_int32 * n)(*[f.shape[0] for f in frames])
        filled = C.c_int(0); used = C.c_int(0)
        L = lib()
        L.sc_mine_negatives.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int,
                                        C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        self._check(L.sc_mine_negatives(self._h, ptrs, Ws, Hs, Ws, n, int(first), need, X.ctypes.data, C.byref(filled), C.byref(used)))
        return X[:filled.value], used.value

    # ---- detection ----
    def detect(self, frames, prm: DetectParams | None = None, cap: int = 1 << 20):
        """Host frames (list of HxW u8 arrays, or an N x H x W array; pinned or pageable) -> (detections, [Counters])."""
        if isinstance(frames, np.ndarray) and frames.ndim == 3:
            frames = [frames[i] for i in range(frames.shape[0])]
        frames = [f if (f.dtype == np.uint8 and f.flags.c_contiguous) else np.ascontiguousarray(f, np.uint8) for f in frames]
        h, w = frames[0].shape
        n = len(frames)
        ptrs = (C.c_void_p * n)(*[f.ctypes.data for f in frames])
        return self.detect_ptrs(ptrs, n, w, h, w, prm, cap)

    def detect_ptrs(self, ptrs, n: int, w: int, h: int, stride: int, prm: DetectParams | None = None, cap: int = 1 << 20):
        prm = prm or params()
        out = np.zeros(cap, DETECTION_DTYPE)
        cnt = (Counters * n)()
        found = C.c_size_t(0)
        rc = lib().sc_detect(self._h, ptrs, n, w, h, stride, C.byref(prm), out.ctypes.data, cap, C.byref(found), C.byref(cnt))
        if rc == SC_ERR_CAPACITY:
            return self.detect_ptrs(ptrs, n, w, h, stride, prm, int(found.value) + 16)
        self._check(rc)
        return out[:found.value].copy(), list(cnt)

    def detect_submit(self, ptrs, n: int, w: int, h: int, stride: int, prm: DetectParams | None = None, cap: int = 1 << 20) -> int:
        """Enqueue one batch (host frame pointers, kept valid until collected); returns a ticket.  Two may be in flight."""
        prm = prm or params()
        L = lib()
        L.sc_detect_submit.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(DetectParams), C.c_size_t,
                                       C.POINTER(C.c_int)]
        t = C.c_int(-1)
        self._check(L.sc_detect_submit(self._h, ptrs, n, w, h, stride, C.byref(prm), cap, C.byref(t)))
        return t.value

    def detect_collect(self, ticket: int, n: int, cap: int = 1 << 20):
        """Wait for a submitted batch of n frames: (detections, [Counters])."""
        L = lib()
        L.sc_detect_collect.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p]
        out = np.empty(cap, DETECTION_DTYPE)
        cnt = (Counters * n)()
        found = C.c_size_t(0)
        self._check(L.sc_detect_collect(self._h, ticket, out.ctypes.data, cap, C.byref(found), C.byref(cnt)))
        return out[:found.value].copy(), list(cnt)

    def detect_device(self, d_frames_ptr: int, n: int, w: int, h: int, d_out_ptr: int, cap: int, d_count_ptr: int, prm: DetectParams | None = None):
        """Frames resident in device memory (n x h x w u8); asynchronous on the handle's stream."""
        prm = prm or params()
        self._check(lib().sc_detect_device(self._h, d_frames_ptr, n, w, h, C.byref(prm), d_out_ptr, cap, d_count_ptr))

    def set_profiling(self, on: bool):
        self._check(lib().sc_set_profiling(self._h, int(on)))

    def kernel_stats(self, reset: bool = False) -> dict:
        """{kernel name: (total ms, launches)} accumulated by the event spans since the last reset."""
        out, k = {}, 0
        while True:
            name = C.c_char_p(); ms = C.c_double(0); n = C.c_int64(0)
            rc = lib().sc_kernel_stats(self._h, k, C.byref(name), C.byref(ms), C.byref(n), int(reset))
            if rc != SC_OK:
                break
            out[name.value.decode()] = (ms.value, n.value)
            k += 1
        return out

    def probe_gather(self, table_bytes: int, iters: int = 10) -> float:
        """GB/s of random 32-byte sector gathers from a table of table_bytes (L2-resident below ~100 MB)."""
        L = lib()
        L.sc_probe_gather.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_double)]
        g = C.c_double(0)
        self._check(L.sc_probe_gather(self._h, table_bytes, iters, C.byref(g)))
        return g.value

    def probe_stream(self, table_bytes: int, iters: int = 10, mode: int = 0) -> float:
        """GB/s of coalesced 16-byte loads over a table of table_bytes; mode 0 = L2-only loads, 1 = L1-allocating."""
        L = lib()
        L.sc_probe_stream.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_double)]
        g = C.c_double(0)
        self._check(L.sc_probe_stream(self._h, table_bytes, iters, mode, C.byref(g)))
        return g.value

    def sync(self):
        self._check(lib().sc_sync(self._h))

    def last_counters(self, n: int):
        cnt = (Counters * n)()
        self._check(lib().sc_last_counters(self._h, C.byref(cnt), n))
        return list(cnt)
