"""TEST INFRASTRUCTURE ONLY -- ctypes binding to oracle/libsurf_oracle.so (the plain-C restatement).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The product (surfcascade_b200/) never does: it fails loudly when its CUDA library is missing.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

from .modelcfg import Cascade

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsurf_oracle.so")
SO_MAX_STAGES = 16
C_GRID, C_VISITED, C_PREFILTER, C_WEAK, C_RAW, C_WEAK_SQUARE, C_WEAK_LONG, C_REACH0 = range(8)
NCOUNTERS = C_REACH0 + SO_MAX_STAGES

_lib = None


def build() -> str:
    """Compile the restatement with gcc (seconds)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    return LIB_PATH


class _Cascade(C.Structure):
    _fields_ = [("n_stages", C.c_int), ("theta", C.POINTER(C.c_float)), ("n_weak", C.POINTER(C.c_int)),
                ("rects", C.POINTER(C.c_int)), ("w", C.POINTER(C.c_float)), ("bias", C.POINTER(C.c_double))]


class _Params(C.Structure):
    _fields_ = [("tmpl", C.c_int), ("base", C.c_int), ("step", C.c_int), ("scale", C.c_double), ("prefilter", C.c_int),
                ("skip_rule", C.c_int), ("force_all", C.c_int), ("nthreads", C.c_int)]


def lib():
    global _lib
    if _lib is None:
        src_m = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("surf_oracle.c", "surf_oracle.h"))
        if not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < src_m:
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.so_window_sum.restype = C.c_float
        _lib.so_weak.restype = C.c_float
        _lib.so_stage_score.restype = C.c_float
        _lib.so_detect.restype = C.c_int64
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def params(tmpl=40, base=40, step=0, scale=1.1, prefilter=6, skip_rule=True, force_all=False, nthreads=1) -> _Params:
    return _Params(tmpl, base, step, scale, prefilter, int(skip_rule), int(force_all), nthreads)


def pool_patches(tmpl: int = 40) -> np.ndarray:
    out = np.zeros((4096, 4), np.int32)
    n = lib().so_pool_patches(tmpl, tmpl, _p(out, C.c_int), 4096)
    return out[:n].copy()


def project(tmpl: int, l: int, patches) -> np.ndarray:
    p = np.ascontiguousarray(patches, np.int32)
    out = np.zeros_like(p)
    for i in range(len(p)):
        lib().so_project(tmpl, l, _p(p[i:i + 1], C.c_int), _p(out[i:i + 1], C.c_int))
    return out


def channels(img) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.zeros((8, h, w), np.uint8)
    lib().so_channels(_p(img, C.c_uint8), w, h, _p(out, C.c_uint8))
    return out


def integral(img) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.zeros((h + 1, w + 1, 8), np.float32)
    lib().so_integral(_p(img, C.c_uint8), w, h, _p(out, C.c_float))
    return out


def features(S: np.ndarray, rects) -> tuple[np.ndarray, np.ndarray]:
    S = np.ascontiguousarray(S, np.float32)
    W = S.shape[1] - 1
    r = np.ascontiguousarray(rects, np.int32)
    out = np.zeros((len(r), 32), np.float32)
    sums = np.zeros(len(r), np.float32)
    lib().so_features(_p(S, C.c_float), W, _p(r, C.c_int), len(r), _p(out, C.c_float), _p(sums, C.c_float))
    return out, sums


def weak(w33, bias: float, x) -> float:
    w33 = np.ascontiguousarray(w33, np.float32); x = np.ascontiguousarray(x, np.float32)
    return float(lib().so_weak(_p(w33, C.c_float), C.c_double(bias), _p(x, C.c_float)))


class BoundCascade:
    """A Cascade plus the pool rects, with the ctypes struct kept alive."""

    def __init__(self, c: Cascade, tmpl: int = 40):
        pool = pool_patches(tmpl)
        self.c = c
        self.theta = np.ascontiguousarray(c.theta, np.float32)
        self.n_weak = np.ascontiguousarray(c.n_weak, np.int32)
        self.rects = np.ascontiguousarray(pool[c.patch_index], np.int32)
        self.w = np.ascontiguousarray(c.w, np.float32)
        self.bias = np.ascontiguousarray(c.bias, np.float64)
        self.struct = _Cascade(len(self.theta), _p(self.theta, C.c_float), _p(self.n_weak, C.c_int), _p(self.rects, C.c_int),
                               _p(self.w, C.c_float), _p(self.bias, C.c_double))


def scales(W: int, H: int, prm: _Params) -> list[int]:
    sides = np.zeros(128, np.int32)
    n = lib().so_num_scales(W, H, C.byref(prm), _p(sides, C.c_int), 128)
    return [int(s) for s in sides[:n]]


def stage_scores(S: np.ndarray, bc: BoundCascade, wins, tmpl: int = 40) -> np.ndarray:
    S = np.ascontiguousarray(S, np.float32)
    W = S.shape[1] - 1
    wins = np.asarray(wins, np.int32)
    out = np.zeros((len(wins), bc.c.n_stages), np.float32)
    for i, (x, y, l) in enumerate(wins):
        for s in range(bc.c.n_stages):
            out[i, s] = lib().so_stage_score(_p(S, C.c_float), W, C.byref(bc.struct), tmpl, s, int(x), int(y), int(l))
    return out


def grid_outcomes(S: np.ndarray, bc: BoundCascade, prm: _Params, si: int):
    S = np.ascontiguousarray(S, np.float32)
    H, W = S.shape[0] - 1, S.shape[1] - 1
    cap = (W + 1) * (H + 1)
    reached = np.zeros(cap, np.int8); score = np.zeros(cap, np.float32)
    n = lib().so_grid_outcomes(_p(S, C.c_float), W, H, C.byref(bc.struct), C.byref(prm), si, _p(reached, C.c_int8), _p(score, C.c_float), cap)
    if n < 0:
        raise IndexError("scale index out of range")
    return reached[:n].copy(), score[:n].copy()


@dataclass
class Detections:
    x: np.ndarray
    y: np.ndarray
    l: np.ndarray
    score: np.ndarray
    counters: np.ndarray


def detect(S: np.ndarray, bc: BoundCascade, prm: _Params, cap: int = 1 << 22) -> Detections:
    S = np.ascontiguousarray(S, np.float32)
    H, W = S.shape[0] - 1, S.shape[1] - 1
    dx = np.zeros(cap, np.int32); dy = np.zeros(cap, np.int32); dl = np.zeros(cap, np.int32); ds = np.zeros(cap, np.float64)
    cnt = np.zeros(NCOUNTERS, np.int64)
    n = lib().so_detect(_p(S, C.c_float), W, H, C.byref(bc.struct), C.byref(prm), _p(dx, C.c_int32), _p(dy, C.c_int32), _p(dl, C.c_int32),
                        _p(ds, C.c_double), C.c_int64(cap), _p(cnt, C.c_int64))
    if n > cap:
        raise RuntimeError("detection capacity exceeded")
    return Detections(dx[:n].copy(), dy[:n].copy(), dl[:n].copy(), ds[:n].copy(), cnt)


def group_rectangles(rects, scores, thr: int = 2, eps: float = 0.2):
    r = np.ascontiguousarray(rects, np.int32).reshape(-1, 4)
    s = np.ascontiguousarray(scores, np.float64)
    n = len(r)
    out_r = np.zeros((max(n, 1), 4), np.int32); out_s = np.zeros(max(n, 1), np.float64)
    m = lib().so_group_rectangles(_p(r, C.c_int32), _p(s, C.c_double), n, thr, C.c_double(eps), _p(out_r, C.c_int32), _p(out_s, C.c_double), max(n, 1))
    return out_r[:m].copy(), out_s[:m].copy()


def pool_eval(X, n_pos: int, Wcand, bias, prior_sum=None, T: int = 0) -> np.ndarray:
    """AUC of every candidate weak classifier (row A9 / config C5).  X [N][P][32]."""
    X = np.ascontiguousarray(X, np.float32)
    N, P, _ = X.shape
    Wcand = np.ascontiguousarray(Wcand, np.float32).reshape(P, 33)
    bias = np.ascontiguousarray(bias, np.float64).reshape(P)
    auc = np.zeros(P, np.float32)
    ps = np.ascontiguousarray(prior_sum, np.float32) if prior_sum is not None else None
    lib().so_pool_eval(_p(X, C.c_float), N, P, n_pos, _p(Wcand, C.c_float), _p(bias, C.c_double), _p(ps, C.c_float) if ps is not None else None, T,
                       _p(auc, C.c_float))
    return auc
