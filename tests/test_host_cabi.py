"""CPU: the C-ABI library loads and exports every symbol include/surfcascade.h declares; host-only entry points
(pool, projection, grouping, model file reader / writer) against the oracle.  No GPU compute here."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import modelcfg as M
from oracle import oracle as O
from surfcascade_b200 import capi

from conftest import MODEL_C1, ROOT


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "surfcascade.h")).read()
    declared = set(re.findall(r"\b(sc_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    L = ctypes.CDLL(capi.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, f"declared in include/surfcascade.h but not exported: {missing}"
    assert declared == set(capi.EXPORTS), (declared ^ set(capi.EXPORTS))


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.SurfCascadeError):
        capi.Handle(0)


def test_pool_and_projection_match_oracle():
    pool = capi.pool_patches(40)
    assert np.array_equal(pool, O.pool_patches(40))
    for l in (40, 41, 44, 48, 53, 64, 97, 233, 1021):
        assert np.array_equal(capi.project_patches(40, l, pool), O.project(40, l, pool)), l


def test_model_reader_matches_oracle_parser():
    got = capi.model_flatten(MODEL_C1, 40)
    want = M.load(MODEL_C1)
    assert np.array_equal(got["theta"].view(np.uint32), want.theta.view(np.uint32))
    assert np.array_equal(got["n_weak"], want.n_weak) and np.array_equal(got["patch_index"], want.patch_index)
    assert np.array_equal(got["w"].view(np.uint32), want.w.view(np.uint32)) and np.array_equal(got["bias"], want.bias)
    assert np.array_equal(got["rects"], O.pool_patches(40)[want.patch_index])


def test_model_writer_round_trips(tmp_path):
    out = str(tmp_path / "resaved.cfg")
    capi.model_resave(MODEL_C1, out)
    a, b = capi.model_flatten(MODEL_C1, 40), capi.model_flatten(out, 40)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    # the oracle's parser (restating libconfig's grammar) reads the product's file to the same cascade
    c = M.load(out)
    assert np.array_equal(c.w.view(np.uint32), a["w"].view(np.uint32)) and np.array_equal(c.theta.view(np.uint32), a["theta"].view(np.uint32))
    # and so does the reference's own libconfig-based loader, where it was compiled
    from oracle import refbind as R
    if R.available():
        import ctypes as C
        th = np.zeros(16, np.float32); nw = np.zeros(16, np.int32); pi = np.zeros(1024, np.int32); w = np.zeros((1024, 33), np.float32); bb = np.zeros(1024)
        P = lambda x, t: x.ctypes.data_as(C.POINTER(t))
        s = R.lib().ref_model_load(out.encode(), P(th, C.c_float), P(nw, C.c_int), 16, P(pi, C.c_int), P(w, C.c_float), P(bb, C.c_double), 1024)
        k = int(nw[:s].sum())
        assert s == len(a["theta"]) and np.array_equal(w[:k].view(np.uint32), a["w"].view(np.uint32)) and np.array_equal(pi[:k], a["patch_index"])


def test_model_reader_error_behaviour(tmp_path):
    """Model::Load semantics (Model.cpp:103-116,188-191): unreadable / unparsable -> failure; a missing setting ends the
    load silently and keeps the stages completed so far."""
    with pytest.raises(capi.SurfCascadeError):
        capi.model_flatten(str(tmp_path / "missing.cfg"), 40)
    bad = tmp_path / "bad.cfg"
    bad.write_text("cascade_classifier : { max_stages_num = ; }")
    with pytest.raises(capi.SurfCascadeError):
        capi.model_flatten(str(bad), 40)
    text = open(MODEL_C1).read()
    # drop the LAST stage's theta: the reference keeps stages 0..n-2 and stops
    idx = text.rfind("theta = ")
    trunc = tmp_path / "trunc.cfg"
    trunc.write_text(text[:idx] + "not_theta = " + text[idx + len("theta = "):])
    full = capi.model_flatten(MODEL_C1, 40)
    part = capi.model_flatten(str(trunc), 40)
    assert len(part["theta"]) == len(full["theta"]) - 1
    assert np.array_equal(part["w"], full["w"][:len(part["w"])])
    assert len(M.load(str(trunc)).theta) == len(part["theta"])


def test_group_rectangles_matches_oracle():
    rng = np.random.default_rng(3)
    for trial in range(40):
        n = int(rng.integers(0, 80))
        base = rng.integers(0, 300, size=(max(1, n // 5), 2))
        xy = base[rng.integers(0, len(base), n)] + rng.integers(-5, 6, size=(n, 2)) if n else np.zeros((0, 2), np.int64)
        l = rng.integers(40, 70, size=(n, 1))
        rects = np.concatenate([xy, l, l], 1).astype(np.int32)
        scores = rng.random(n) + 1.0
        gr, gs = capi.group_rectangles(rects, scores)
        wr, ws = O.group_rectangles(rects, scores)
        assert np.array_equal(gr, wr) and np.array_equal(gs, ws)


def test_group_rectangles_on_dense_and_spread_sets(oracle_cascade):
    """The host grouping orders the rectangles by x and skips pairs that are out of reach or already in one class; the
    classes, their first-seen numbering, means, swallow filter and scores must stay those of the all-pairs definition
    (oracle = restatement of cv::groupRectangles): raw window lists of real frames (hundreds of windows in a few dense
    clusters), crowded random sets of mixed sizes, chains that merge late, non-square rectangles, other thresholds."""
    from surfcascade_b200 import synth
    cases = []
    for seed in (100, 101):
        img = synth.frame(480, 640, seed)
        d = O.detect(O.integral(img), oracle_cascade, O.params(base=40, nthreads=8))
        cases.append((np.stack([d.x, d.y, d.l, d.l], 1).astype(np.int32), d.score.copy(), 2, 0.2))
    rng = np.random.default_rng(7)
    for n, span, lo, hi, thr, eps in ((1500, 900, 40, 300, 2, 0.2), (800, 200, 40, 60, 3, 0.2), (600, 3000, 40, 44, 1, 0.5), (400, 100, 30, 200, 2, 0.05)):
        xy = rng.integers(0, span, size=(n, 2))
        wh = rng.integers(lo, hi, size=(n, 2))
        wh[: n // 2, 1] = wh[: n // 2, 0]  # half squares, half arbitrary
        cases.append((np.concatenate([xy, wh], 1).astype(np.int32), rng.random(n) + 1.0, thr, eps))
    # a chain along x whose links are each similar to the next only: one class, found across the ordered sweep
    k = 300
    chain = np.stack([np.arange(k) * 7, np.zeros(k, np.int64), np.full(k, 100), np.full(k, 100)], 1).astype(np.int32)
    cases.append((chain[rng.permutation(k)], rng.random(k), 2, 0.2))
    for rects, scores, thr, eps in cases:
        gr, gs = capi.group_rectangles(rects, scores, thr, eps)
        wr, ws = O.group_rectangles(rects, scores, thr, eps)
        assert np.array_equal(gr, wr) and np.array_equal(gs, ws), (len(rects), thr, eps)
        try:
            import cv2  # the real OpenCV where the wheel is present (this container): same rectangles in the same order
        except ImportError:
            continue
        cr, _ = cv2.groupRectangles(rects.tolist(), thr, eps)
        assert np.array_equal(np.array(cr, np.int32).reshape(-1, 4), gr.reshape(-1, 4)), (len(rects), thr, eps)
    assert len(cases[0][0]) > 100


def _ref_load(path):
    """The reference's own Model::Load (libconfig 1.4.9) through the harness, where it was compiled."""
    import ctypes as C
    from oracle import refbind as R
    th = np.zeros(16, np.float32); nw = np.zeros(16, np.int32); pi = np.zeros(2048, np.int32); w = np.zeros((2048, 33), np.float32); bb = np.zeros(2048)
    P = lambda x, t: x.ctypes.data_as(C.POINTER(t))
    s = R.lib().ref_model_load(path.encode(), P(th, C.c_float), P(nw, C.c_int), 16, P(pi, C.c_int), P(w, C.c_float), P(bb, C.c_double), 2048)
    k = int(nw[:max(s, 0)].sum())
    return s, th[:max(s, 0)].copy(), nw[:max(s, 0)].copy(), pi[:k].copy(), w[:k].copy(), bb[:k].copy()


def _restyle(text: str, style: int) -> str:
    """The same settings in other spellings libconfig's grammar accepts (scanner.l / grammar.y of 1.4.9)."""
    if style == 1:   # ':' for '=', comments of all three kinds, tabs, no ';' after groups and lists
        text = text.replace(" = ", " : ").replace("};", "}").replace(");", ")")
        text = "# model written by a test\n// another comment\n/* and a\n   block */\n" + text.replace("\n      ", "\n\t")
    elif style == 2:  # ',' as the setting terminator, everything on few lines
        text = text.replace(";\n", ",\n").replace("\n          ", " ")
    return text


@pytest.mark.parametrize("seed,n_weak,style", [(21, [2, 3], 0), (22, [1] * 10, 1), (23, [40, 7, 1], 2), (24, [3, 3, 3, 3], 1)])
def test_model_reader_on_synthetic_models_matches_reference_loader(tmp_path, seed, n_weak, style):
    """Row A12 beyond the one trained file: synthetic cascades (1..10 stages, up to 40 weak classifiers in a stage, weights over 60
    orders of magnitude and in every float spelling libconfig scans) in three spellings of the format, read by the product's own
    reader, by the oracle's parser and -- where it was compiled -- by the reference's Model::Load on libconfig 1.4.9."""
    from cascade_util import random_cascade, write_model_cfg
    rng = np.random.default_rng(seed)
    c = random_cascade(seed, n_weak, list(rng.uniform(0.3, 0.7, len(n_weak))))
    # stretch the weights: tiny, huge, exact integers, negative zero
    flat = c.w.reshape(-1)
    k = len(flat)
    flat[rng.integers(0, k, k // 8)] *= np.float32(1e-30)
    flat[rng.integers(0, k, k // 8)] *= np.float32(1e+30)
    flat[rng.integers(0, k, k // 16)] = np.float32(3.0)
    flat[rng.integers(0, k, k // 16)] = np.float32(-0.0)
    path = str(tmp_path / "m.cfg")
    write_model_cfg(path, c)
    text = _restyle(open(path).read(), style)
    if style == 2:  # other float spellings of the same values: "+x", "x." / ".x", exponent without a point
        text = text.replace("bias = 1.0", "bias = +1.").replace("theta = 0.", "theta = .").replace("eps = 0.01", "eps = 1e-2").replace("C = 0.1", "C = 1E-1")
    open(path, "w").write(text)
    got = capi.model_flatten(path, 40)
    want = M.load(path)
    assert len(got["theta"]) == len(n_weak) == want.n_stages
    assert np.array_equal(got["theta"].view(np.uint32), want.theta.view(np.uint32)) and np.array_equal(got["n_weak"], want.n_weak)
    assert np.array_equal(got["patch_index"], want.patch_index) and np.array_equal(got["w"].view(np.uint32), want.w.view(np.uint32))
    assert np.array_equal(got["bias"], want.bias)
    assert np.array_equal(want.w.view(np.uint32), c.w.view(np.uint32)), "%.10g text does not round-trip float32"
    from oracle import refbind as R
    if R.available():
        s, th, nw, pi, w, bb = _ref_load(path)
        assert s == len(n_weak) and np.array_equal(th.view(np.uint32), got["theta"].view(np.uint32)) and np.array_equal(nw, got["n_weak"])
        assert np.array_equal(pi, got["patch_index"]) and np.array_equal(w.view(np.uint32), got["w"].view(np.uint32)) and np.array_equal(bb, got["bias"])
        # the product's writer (Model::Save) on the same cascade, read back by the reference's loader
        out = str(tmp_path / "resaved.cfg")
        capi.model_resave(path, out)
        s2, th2, nw2, pi2, w2, bb2 = _ref_load(out)
        assert s2 == s and np.array_equal(th2.view(np.uint32), th.view(np.uint32)) and np.array_equal(nw2, nw) and np.array_equal(pi2, pi)
        assert np.array_equal(w2.view(np.uint32), w.view(np.uint32)) and np.array_equal(bb2, bb)
