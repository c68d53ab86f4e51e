"""TEST INFRASTRUCTURE ONLY -- ctypes binding to oracle/_ref/libsurfcascade_ref.so.

That library is the UNMODIFIED reference (mrgloom/SurfCascade) compiled from /root/reference by
oracle/Makefile (`make ref`).  It is the ground truth the plain-C restatement (oracle/surf_oracle.c) and
the CUDA path are pinned against, and the `kind: "reference"` CPU baseline of bench.py.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libsurfcascade_ref.so")

_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise FileNotFoundError(f"{LIB_PATH} missing: run `make -C oracle ref` where /root/reference exists")
        _lib = C.CDLL(LIB_PATH)
        _lib.ref_num_counters.restype = C.c_int
    return _lib


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint8))


def _i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(C.POINTER(C.c_int32))


def pool_patches(tmpl: int = 40) -> np.ndarray:
    """ExtractPatches (DenseSURFFeatureExtractor.cpp:49-63) -> int32 [n][4] x,y,w,h."""
    out = np.zeros((4096, 4), np.int32)
    n = lib().ref_pool_patches(tmpl, tmpl, out.ctypes.data_as(C.POINTER(C.c_int32)), 4096)
    return out[:n].copy()


def project(tmpl: int, win, patches) -> np.ndarray:
    """ProjectPatches (DenseSURFFeatureExtractor.cpp:486-508)."""
    w, wp = _i32(win)
    p, pp = _i32(patches)
    out = np.zeros_like(p)
    lib().ref_project(tmpl, wp, pp, len(p), out.ctypes.data_as(C.POINTER(C.c_int32)))
    return out


def channels(img: np.ndarray) -> np.ndarray:
    """T2bFilter (DenseSURFFeatureExtractor.cpp:199-349) -> u8 [8][H][W]."""
    img, ip = _u8(img)
    h, w = img.shape
    out = np.zeros((8, h, w), np.uint8)
    lib().ref_channels(ip, w, h, out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out


def integral(img: np.ndarray) -> np.ndarray:
    """IntegralImage (DenseSURFFeatureExtractor.cpp:65-87) -> f32 [H+1][W+1][8]."""
    img, ip = _u8(img)
    h, w = img.shape
    out = np.zeros((h + 1, w + 1, 8), np.float32)
    lib().ref_integral(ip, w, h, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def features(img: np.ndarray, rects) -> tuple[np.ndarray, np.ndarray]:
    """CalcFeature + sum() for rects [n][4] -> (f32 [n][32], f32 [n])."""
    img, ip = _u8(img)
    h, w = img.shape
    r, rp = _i32(rects)
    n = len(r)
    out = np.zeros((n, 32), np.float32)
    sums = np.zeros(n, np.float32)
    lib().ref_features(ip, w, h, rp, n, out.ctypes.data_as(C.POINTER(C.c_float)), sums.ctypes.data_as(C.POINTER(C.c_float)))
    return out, sums


def stage_scores(img: np.ndarray, model_cfg: str, wins, tmpl: int = 40, max_stages: int = 16) -> np.ndarray:
    """Every stage's Predict2 score on explicit windows [n][3]=x,y,l (no early exit) -> f32 [n][n_stages]."""
    img, ip = _u8(img)
    h, w = img.shape
    wn, wp = _i32(wins)
    n = len(wn)
    out = np.zeros((n, max_stages), np.float32)
    s = lib().ref_stage_scores(ip, w, h, model_cfg.encode(), tmpl, wp, n, out.ctypes.data_as(C.POINTER(C.c_float)), max_stages)
    if s < 0:
        raise RuntimeError(f"ref_stage_scores failed ({s})")
    return out[:, :s].copy()


@dataclass
class RefDetections:
    frame: np.ndarray
    x: np.ndarray
    y: np.ndarray
    l: np.ndarray
    score: np.ndarray
    counters: np.ndarray  # int64 [nframes][ncounters]: visited, prefilter_pass, weak_evals, raw, reach[16]
    ms_integral: np.ndarray
    ms_scan: np.ndarray
    g_frame: np.ndarray
    g_rect: np.ndarray
    g_score: np.ndarray


def detect(frames, model_cfg: str, base: int = 40, nthreads: int = 1, group: bool = True, cap: int = 1 << 22) -> RefDetections:
    """The reference detect loop (ObjDetector.cpp:107-143,174-225) on a list of equally sized u8 frames."""
    frames = [np.ascontiguousarray(f, dtype=np.uint8) for f in frames]
    h, w = frames[0].shape
    assert all(f.shape == (h, w) for f in frames)
    n = len(frames)
    ptrs = (C.POINTER(C.c_uint8) * n)(*[f.ctypes.data_as(C.POINTER(C.c_uint8)) for f in frames])
    nc = lib().ref_num_counters()
    df = np.zeros(cap, np.int32); dx = np.zeros(cap, np.int32); dy = np.zeros(cap, np.int32); dl = np.zeros(cap, np.int32)
    ds = np.zeros(cap, np.float64)
    gf = np.zeros(cap, np.int32); gr = np.zeros((cap, 4), np.int32); gs = np.zeros(cap, np.float64)
    counters = np.zeros((n, nc), np.int64)
    ms_i = np.zeros(n, np.float64); ms_s = np.zeros(n, np.float64)
    nd = C.c_int64(0); ng = C.c_int64(0)
    P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    rc = lib().ref_detect_frames(ptrs, n, w, h, model_cfg.encode(), base, nthreads,
                                 P(df, C.c_int32), P(dx, C.c_int32), P(dy, C.c_int32), P(dl, C.c_int32), P(ds, C.c_double),
                                 C.c_int64(cap), C.byref(nd), P(counters, C.c_int64), P(ms_i, C.c_double), P(ms_s, C.c_double),
                                 1 if group else 0, P(gf, C.c_int32), P(gr, C.c_int32), P(gs, C.c_double), C.byref(ng))
    if rc != 0:
        raise RuntimeError(f"ref_detect_frames failed ({rc})")
    k, g = nd.value, ng.value
    if k > cap or g > cap:
        raise RuntimeError("detection capacity exceeded")
    return RefDetections(df[:k].copy(), dx[:k].copy(), dy[:k].copy(), dl[:k].copy(), ds[:k].copy(), counters, ms_i, ms_s,
                         gf[:g].copy(), gr[:g].copy(), gs[:g].copy())


def detect_timed(frames, model_cfg: str, base: int = 40, nthreads: int = 1, warmup: int = 1):
    """Timing build of the reference detect loop: no counter hooks in the window loop, model loaded once per call,
    IntegralImage and scan timed separately per frame (ref_detect_timed).  Returns (ms_integral[n], ms_scan[n], n_raw[n])."""
    frames = [np.ascontiguousarray(f, dtype=np.uint8) for f in frames]
    h, w = frames[0].shape
    assert all(f.shape == (h, w) for f in frames)
    n = len(frames)
    ptrs = (C.POINTER(C.c_uint8) * n)(*[f.ctypes.data_as(C.POINTER(C.c_uint8)) for f in frames])
    ms_i = np.zeros(n, np.float64); ms_s = np.zeros(n, np.float64); raw = np.zeros(n, np.int64)
    P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    rc = lib().ref_detect_timed(ptrs, n, w, h, model_cfg.encode(), base, nthreads, warmup, P(ms_i, C.c_double), P(ms_s, C.c_double), P(raw, C.c_int64))
    if rc != 0:
        raise RuntimeError(f"ref_detect_timed failed ({rc})")
    return ms_i, ms_s, raw


def train(prefix: str, pos_list: str, neg_list: str, out_cfg: str, verbose: bool = False) -> int:
    """The reference --train branch (ObjDetector.cpp:66-91).  Once per process (static cursors)."""
    if not prefix.endswith("/"):
        prefix += "/"
    return lib().ref_train(prefix.encode(), pos_list.encode(), neg_list.encode(), out_cfg.encode(), 1 if verbose else 0)


def pool_eval(X, n_pos: int, Wcand, bias, prev_patch=(), prev_w=(), prev_bias=()) -> np.ndarray:
    """GentleAdaboost::Train's candidate scoring (GentleAdaboost.cpp:145-148 -> StageClassifier::Evaluate)."""
    X = np.ascontiguousarray(X, np.float32)
    N, P, _ = X.shape
    Wcand = np.ascontiguousarray(Wcand, np.float32).reshape(P, 33)
    bias = np.ascontiguousarray(bias, np.float64).reshape(P)
    pp = np.ascontiguousarray(prev_patch, np.int32).reshape(-1)
    pw = np.ascontiguousarray(prev_w, np.float32).reshape(-1, 33) if len(pp) else np.zeros((1, 33), np.float32)
    pb = np.ascontiguousarray(prev_bias, np.float64).reshape(-1) if len(pp) else np.zeros(1)
    if not len(pp):
        pp = np.zeros(1, np.int32)
        T = 0
    else:
        T = len(pp)
    auc = np.zeros(P, np.float32)
    P_ = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    lib().ref_pool_eval(P_(X, C.c_float), N, P, n_pos, P_(Wcand, C.c_float), P_(bias, C.c_double), P_(pp, C.c_int32), P_(pw, C.c_float),
                        P_(pb, C.c_double), T, P_(auc, C.c_float))
    return auc


def write_pgm(path: str, img: np.ndarray) -> None:
    img = np.ascontiguousarray(img, dtype=np.uint8)
    with open(path, "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
        f.write(img.tobytes())


def _write_pgm(path: str, img) -> None:
    img = np.ascontiguousarray(img, dtype=np.uint8)
    with open(path, "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
        f.write(img.tobytes())


_FILL_NEG_CHILD = r"""
import ctypes as C, sys, numpy as np
lib = C.CDLL(sys.argv[1])
prefix, lst, cfg, out_path = sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5]
tmpl, first = int(sys.argv[6]), int(sys.argv[7])
totals = np.array([int(v) for v in sys.argv[8].split(",")], np.int32)
cap = int(sys.argv[9])
out = np.zeros(cap, np.float32)
counts = np.zeros(len(totals), np.int32); dones = np.zeros(len(totals), np.int32); pool = C.c_int(0)
rc = lib.ref_fill_neg(prefix.encode(), lst.encode(), cfg.encode(), tmpl, totals.ctypes.data_as(C.POINTER(C.c_int)), len(totals), first,
                      out.ctypes.data_as(C.POINTER(C.c_float)), C.c_longlong(cap), counts.ctypes.data_as(C.POINTER(C.c_int)),
                      dones.ctypes.data_as(C.POINTER(C.c_int)), C.byref(pool))
if rc != 0:
    sys.exit(10 - rc)
n = int(counts.sum())
np.savez(out_path, X=out[: n * pool.value * 32].reshape(n, pool.value, 32), counts=counts, dones=dones)
"""


def fill_neg(frames, model_cfg: str, n_totals, first: bool, tmpl: int = 40):
    """DenseSURFFeatureExtractor::FillNegSamples (DenseSURFFeatureExtractor.cpp:124-195) called len(n_totals) times in a row
    on one extractor over `frames` (written as PGM files), single thread.  The reference keeps its image cursor in a
    function-local static, so every invocation runs in a fresh child process.
    Returns (list of per-call sample arrays [n][608][32], list of per-call `done` flags)."""
    import subprocess
    import sys
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        names = []
        for i, img in enumerate(frames):
            names.append(f"neg{i:04d}.pgm")
            _write_pgm(os.path.join(d, names[-1]), img)
        with open(os.path.join(d, "neg.list"), "w") as f:
            f.write("\n".join(names) + "\n")
        out_path = os.path.join(d, "out.npz")
        per_sample = 608 * 32 if tmpl == 40 else 4096 * 32
        cap = int(sum(int(v) for v in n_totals)) * per_sample + 32
        r = subprocess.run([sys.executable, "-c", _FILL_NEG_CHILD, LIB_PATH, d + "/", "neg.list", model_cfg or "", out_path, str(tmpl),
                            "1" if first else "0", ",".join(str(int(v)) for v in n_totals), str(cap)], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"ref_fill_neg child failed ({r.returncode}): {r.stderr[-400:]}")
        z = np.load(out_path)
        X, counts, dones = z["X"], z["counts"], z["dones"]
    outs, o = [], 0
    for c in counts:
        outs.append(X[o:o + int(c)].copy())
        o += int(c)
    return outs, [bool(v) for v in dones]
