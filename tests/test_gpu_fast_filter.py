"""GPU: the stage-0 certified fast filter (csrc/sc_kernels.cuh) never changes a result.

1. Its distance to the reference arithmetic stays inside the budget the decision limits are built from (measured on
   ~4*10^5 windows of textured and pure-noise frames; the bar is a quarter of the budget).
2. The scan with the filter equals the exact-only scan (SC_DISABLE_FAST=1) bit for bit: detections, scores, counters.
3. Cascades cut and re-thresholded so that the filter's three decisions (reject, skip, no-skip) all sit inside the
   score distribution still match the oracle: theta at the median stage-0 score (half of the windows go through the
   exact path), 1 / 2 / 3 stages (the `multi` rule of a stage-0 reject is never / never / score-dependent).
"""
import dataclasses
import os

import numpy as np
import pytest

from oracle import oracle as O
from surfcascade_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _windows(h, w, n, rng):
    out = []
    for _ in range(n):
        l = int(rng.integers(40, min(h, w)))
        out.append((int(rng.integers(0, w - l + 1)), int(rng.integers(0, h - l + 1)), l))
    return np.array(out, np.int32)


@pytest.mark.parametrize("kind", ["frame", "noise", "flat"])
def test_fast_sum_within_budget(gpu_handle, kind):
    h, w = 480, 640
    img = {"frame": lambda: synth.frame(h, w, 21), "noise": lambda: synth.noise_frame(h, w, 22),
           "flat": lambda: np.full((h, w), 77, np.uint8)}[kind]()
    if kind == "flat":
        img[100:140, 200:260] = 200  # mostly zero descriptors (eps-dominated norms), one edge
    gpu_handle.integral(img, want_output=False)
    wins = _windows(h, w, 150000 if kind != "flat" else 20000, np.random.default_rng(3))
    fs, es, margin = gpu_handle.stage0_fast_check(wins)
    assert np.isfinite(fs).all() and np.isfinite(es).all()
    err = np.abs(fs.astype(np.float64) - es.astype(np.float64)).max()
    assert 0 < margin < 1e-3
    assert err < margin / 4, f"fast filter error {err:.3e} vs budget {margin:.3e}"


def _exact_only_handle(model):
    os.environ["SC_DISABLE_FAST"] = "1"
    try:
        h = capi.Handle(0)
    finally:
        del os.environ["SC_DISABLE_FAST"]
    h.load_model(model, 40)
    return h


def test_filtered_scan_equals_exact_only_scan(gpu_handle, model_c1):
    hx = _exact_only_handle(model_c1)
    try:
        for frames, prm in [([synth.frame(1080, 1920, 31)], capi.params()),
                            ([synth.frame(240, 320, s) for s in range(4)] + [synth.noise_frame(240, 320, 7)], capi.params()),
                            ([synth.frame(301, 517, 4)], capi.params(base=70)),
                            ([synth.frame(200, 260, 6)], capi.params(skip_rule=False)),
                            ([synth.frame(200, 260, 6)], capi.params(prefilter=-1, step=1))]:
            a, ca = gpu_handle.detect(frames, prm)
            b, cb = hx.detect(frames, prm)
            assert a.tobytes() == b.tobytes()
            for x, y in zip(ca, cb):
                # every reference-defined counter is equal; `evaluated` is this implementation's own work count: the
                # filtered scan leaves its undecided windows to a later exact pass and, until then, treats them as "may
                # not skip" when it picks the odd columns to evaluate -- a superset of the exact-only scan's choice
                for k in ("grid", "visited", "prefilter_pass", "weak_evals", "raw"):
                    assert getattr(x, k) == getattr(y, k), k
                assert list(x.reach) == list(y.reach)
                assert x.visited <= y.evaluated <= x.evaluated <= x.grid
    finally:
        hx.close()


@pytest.mark.parametrize("n_stages,theta0", [(4, 0.29), (3, None), (3, 0.50), (2, None), (1, None), (1, 0.29)])
def test_recut_cascades_match_oracle(oracle_cascade, n_stages, theta0):
    c = oracle_cascade.c
    k = int(c.n_weak[:n_stages].sum())
    theta = c.theta[:n_stages].copy()
    if theta0 is not None:
        theta[0] = np.float32(theta0)
    cut = dataclasses.replace(c, theta=theta, n_weak=c.n_weak[:n_stages].copy(), patch_index=c.patch_index[:k].copy(), w=c.w[:k].copy(),
                              bias=c.bias[:k].copy())
    bc = O.BoundCascade(cut)
    h = capi.Handle(0)
    try:
        h.set_cascade(40, bc.theta, bc.n_weak, bc.rects, bc.w, bc.bias)
        for img in (synth.frame(240, 320, 41), synth.noise_frame(120, 160, 42)):
            dets, cnts = h.detect([img], capi.params(), cap=1 << 20)
            want = O.detect(O.integral(img), bc, O.params(base=40, nthreads=8), cap=1 << 20)
            assert cnts[0].visited == want.counters[O.C_VISITED] and cnts[0].prefilter_pass == want.counters[O.C_PREFILTER]
            assert [cnts[0].reach[s] for s in range(n_stages)] == [int(want.counters[O.C_REACH0 + s]) for s in range(n_stages)]
            assert np.array_equal(dets["x"], want.x) and np.array_equal(dets["y"], want.y) and np.array_equal(dets["l"], want.l)
            np.testing.assert_allclose(dets["score"], want.score, rtol=1e-6, atol=0)
    finally:
        h.close()


def _random_cascade(seed, n_weak, thetas, shapes="mixed"):
    """Cascade with weak classifiers on randomly drawn pool patches (squares 0-343, tall 344-475, wide 476-607)."""
    from oracle.modelcfg import Cascade
    rng = np.random.default_rng(seed)
    total = int(sum(n_weak))
    if shapes == "mixed":
        idx = np.concatenate([rng.integers(0, 344, total - 2 * (total // 3)), rng.integers(344, 476, total // 3), rng.integers(476, 608, total // 3)])
        rng.shuffle(idx)
    else:
        idx = rng.integers(0, 344, total)
    w = rng.normal(0, 1.0, (total, 33)).astype(np.float32)
    w[:, 32] = rng.normal(0, 0.3, total).astype(np.float32)
    return Cascade(np.array(thetas, np.float32), np.array(n_weak, np.int32), idx.astype(np.int32), w, np.ones(total))


@pytest.mark.parametrize("seed,n_weak,thetas,shapes", [(1, [3, 5, 4], [0.50, 0.50, 0.50], "mixed"), (2, [2, 3], [0.45, 0.52], "squares"),
                                                        (3, [6, 2], [0.48, 0.50], "mixed"), (4, [8, 3, 3], [0.49, 0.50, 0.50], "mixed"),
                                                        (5, [1], [0.55], "squares")])
def test_random_cascades_with_square_patches_match_oracle(seed, n_weak, thetas, shapes):
    """Cascades the model file does not cover: 2x2-cell patches in stage 0 (fast filter and exact path), six weak classifiers
    (the filter's limit), eight (exact-only kernel), a single weak classifier; full detect parity incl. counters."""
    bc = O.BoundCascade(_random_cascade(seed, n_weak, thetas, shapes))
    h = capi.Handle(0)
    try:
        h.set_cascade(40, bc.theta, bc.n_weak, bc.rects, bc.w, bc.bias)
        for img, prm in ((synth.frame(200, 260, 60 + seed), {}), (synth.noise_frame(130, 170, seed), {"step": 1}), (synth.frame(200, 260, 70 + seed), {"skip_rule": False})):
            dets, cnts = h.detect([img], capi.params(**prm), cap=1 << 21)
            want = O.detect(O.integral(img), bc, O.params(base=40, step=prm.get("step", 0), skip_rule=prm.get("skip_rule", True), nthreads=8), cap=1 << 21)
            assert cnts[0].visited == want.counters[O.C_VISITED] and cnts[0].prefilter_pass == want.counters[O.C_PREFILTER]
            assert [cnts[0].reach[s] for s in range(len(n_weak))] == [int(want.counters[O.C_REACH0 + s]) for s in range(len(n_weak))]
            assert np.array_equal(dets["x"], want.x) and np.array_equal(dets["y"], want.y) and np.array_equal(dets["l"], want.l)
            np.testing.assert_allclose(dets["score"], want.score, rtol=1e-6, atol=0)
        # the fast filter's distance budget holds for these weights and shapes too
        if n_weak[0] <= 6:
            img = synth.frame(240, 320, 80 + seed)
            h.integral(img, want_output=False)
            fs, es, margin = h.stage0_fast_check(_windows(240, 320, 60000, np.random.default_rng(seed)))
            assert np.abs(fs.astype(np.float64) - es.astype(np.float64)).max() < margin / 4
    finally:
        h.close()


@pytest.mark.parametrize("theta0,width", [(0.29, 1400), (0.40, 1027), (None, 1400)])
def test_wide_frames_long_odd_suffixes(oracle_cascade, theta0, width):
    """k_scan_odd's 128-window units: wide, flat frames (up to 681 lattice columns per row, several units per row suffix,
    suffixes starting at every residue of the 32-bit mask words) with stage-0 thresholds low enough that rows switch the
    stride parity early and often; also with the prefilter off (every lane of a unit survives) and the trained thresholds."""
    c = oracle_cascade.c
    theta = c.theta.copy()
    if theta0 is not None:
        theta[0] = np.float32(theta0)
    bc = O.BoundCascade(dataclasses.replace(c, theta=theta))
    h = capi.Handle(0)
    try:
        h.set_cascade(40, bc.theta, bc.n_weak, bc.rects, bc.w, bc.bias)
        for img, pf in ((synth.frame(120, width, 90), 6), (synth.noise_frame(97, width, 91), 6), (synth.frame(64, width, 92), -1)):
            dets, cnts = h.detect([img, img[:, ::-1].copy()], capi.params(prefilter=pf), cap=1 << 21)
            for f, im in enumerate((img, img[:, ::-1].copy())):
                want = O.detect(O.integral(im), bc, O.params(base=40, prefilter=pf, nthreads=8), cap=1 << 21)
                mine = dets[dets["frame"] == f]
                assert cnts[f].visited == want.counters[O.C_VISITED] and cnts[f].prefilter_pass == want.counters[O.C_PREFILTER]
                assert [cnts[f].reach[s] for s in range(4)] == [int(want.counters[O.C_REACH0 + s]) for s in range(4)]
                assert cnts[f].visited <= cnts[f].evaluated <= cnts[f].grid
                assert np.array_equal(mine["x"], want.x) and np.array_equal(mine["y"], want.y) and np.array_equal(mine["l"], want.l)
                np.testing.assert_allclose(mine["score"], want.score, rtol=1e-6, atol=0)
    finally:
        h.close()


@pytest.mark.parametrize("seed,n_weak,thetas", [(11, [2, 3, 2, 2, 2], [0.45, 0.60, 0.50, 0.50, 0.50]),
                                                 (12, [2] * 10, [0.45, 0.47, 0.48, 0.48, 0.48, 0.48, 0.48, 0.48, 0.48, 0.48]),
                                                 (13, [3, 2, 2, 2, 2, 2], [0.50, 0.50, 0.50, 0.50, 0.50, 0.50])])
def test_deep_cascades_match_oracle(seed, n_weak, thetas):
    """More than four stages: a reject at a LATER stage can still mean stride 2 (score = (s + p + 1) / N < 0.5,
    ObjDetector.cpp:201,214) -- score-dependent at stage 1 of a 5-stage cascade, always up to stage 1 / 3 of 6 / 10 stages.
    The same cascades run through the reference's own loader and loop in tests/test_oracle_vs_ref.py.  Also with every stage
    forced (the tile kernel's ALL variant sets those bits itself) and without the fast filter."""
    from cascade_util import random_cascade
    bc = O.BoundCascade(random_cascade(seed, n_weak, thetas))
    n = len(n_weak)
    handles = [capi.Handle(0)]
    os.environ["SC_DISABLE_FAST"] = "1"
    try:
        handles.append(capi.Handle(0))
    finally:
        del os.environ["SC_DISABLE_FAST"]
    try:
        for h in handles:
            h.set_cascade(40, bc.theta, bc.n_weak, bc.rects, bc.w, bc.bias)
        for img in (synth.frame(200, 260, 5), synth.frame(97, 333, 6)):
            want = O.detect(O.integral(img), bc, O.params(base=40, nthreads=8), cap=1 << 21)
            for h, prm in ((handles[0], {}), (handles[1], {}), (handles[0], {"force_all_stages": True})):
                dets, cnts = h.detect([img], capi.params(**prm), cap=1 << 21)
                assert cnts[0].visited == want.counters[O.C_VISITED] and cnts[0].prefilter_pass == want.counters[O.C_PREFILTER]
                if not prm:
                    assert [cnts[0].reach[s] for s in range(n)] == [int(want.counters[O.C_REACH0 + s]) for s in range(n)]
                assert np.array_equal(dets["x"], want.x) and np.array_equal(dets["y"], want.y) and np.array_equal(dets["l"], want.l)
                np.testing.assert_allclose(dets["score"], want.score, rtol=1e-6, atol=0)
    finally:
        for h in handles:
            h.close()
