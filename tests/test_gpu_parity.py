"""GPU parity tests: the CUDA path, called through the C-ABI, against the oracle (plain-C restatement of the
reference, pinned to the reference build in oracle/_ref) on the same seeded inputs.

Bars (BASELINE.json north_star): integral bit-exact; descriptors / weak and stage scores within 1e-5 relative (we
assert bit-exact and only fall back to the tolerance for the double-precision sigmoid); window set identical.
"""
import numpy as np
import pytest

from oracle import oracle as O
from surfcascade_b200 import capi, synth

pytestmark = pytest.mark.gpu


def stripes(h, w):
    """Columns 0,0,255,255,...: dx = +-255 on half the pixels, so the dx integrals pass 2^24 after ~263k pixels."""
    img = np.zeros((h, w), np.uint8)
    img[:, 2::4] = 255
    img[:, 3::4] = 255
    return img


@pytest.mark.parametrize("shape,kind", [((117, 203), "frame"), ((480, 640), "frame"), ((2, 2), "noise"), ((3, 33), "noise"), ((64, 31), "noise"),
                                        ((65, 32), "noise"), ((200, 97), "noise"), ((600, 700), "stripes"), ((1080, 1920), "noise"),
                                        ((1080, 1920), "frame")])
def test_integral_bit_exact(gpu_handle, shape, kind):
    h, w = shape
    img = {"frame": lambda: synth.frame(h, w, 3) if min(h, w) >= 64 else synth.noise_frame(h, w, 3), "noise": lambda: synth.noise_frame(h, w, 5),
           "stripes": lambda: stripes(h, w)}[kind]()
    got = gpu_handle.integral(img)
    want = O.integral(img)
    if kind == "stripes" or (kind == "noise" and h >= 1080):
        assert want.max() > 2 ** 24  # exercises the inexact float32 column recurrence
    assert got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("shape,kind,step", [((117, 203), "frame", 2), ((480, 640), "frame", 2), ((2, 2), "noise", 2), ((3, 33), "noise", 2), ((17, 31), "noise", 2),
                                             ((65, 32), "noise", 2), ((200, 97), "noise", 1), ((301, 517), "frame", 3), ((600, 700), "stripes", 2),
                                             ((1080, 1920), "noise", 2), ((1080, 1920), "frame", 2), ((95, 1400), "noise", 5)])
def test_integral_bit_exact_in_the_scan_layout(gpu_handle, shape, kind, step):
    """The integral stage of the detect path itself: same kernels, the scan's lattice-deinterleaved layout (columns by
    2 * step, rows by step; for step 2 the column phase of the walk runs in its plane-major lane order)."""
    h, w = shape
    img = {"frame": lambda: synth.frame(h, w, 3) if min(h, w) >= 64 else synth.noise_frame(h, w, 3), "noise": lambda: synth.noise_frame(h, w, 5),
           "stripes": lambda: stripes(h, w)}[kind]()
    got = gpu_handle.integral_scan_layout(img, step)
    want = O.integral(img)
    assert got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_features_and_window_sums_bit_exact(gpu_handle):
    img = synth.frame(240, 320, 11)
    S = O.integral(img)
    gpu_handle.integral(img, want_output=False)
    pool = O.pool_patches(40)
    rects = []
    for (x, y, l) in [(0, 0, 40), (7, 3, 44), (100, 50, 97), (0, 0, 240), (81, 1, 233), (33, 17, 121)]:
        p = O.project(40, l, pool)
        p[:, 0] += x; p[:, 1] += y
        rects.append(p)
    rects = np.concatenate(rects)
    want_f, want_s = O.features(S, rects)
    got_f = gpu_handle.features(rects)
    got_s = gpu_handle.window_sum(rects)
    assert np.array_equal(got_f.view(np.uint32), want_f.view(np.uint32))
    assert np.array_equal(got_s.view(np.uint32), want_s.view(np.uint32))


def test_weak_and_stage_predict(gpu_handle, oracle_cascade):
    rng = np.random.default_rng(0)
    c = oracle_cascade
    n = len(c.w)
    x = rng.normal(0, 0.2, size=(n, 32)).astype(np.float32)
    want = np.array([O.weak(c.w[i], c.bias[i], x[i]) for i in range(n)], np.float32)
    got = gpu_handle.weak_predict(c.w, c.bias, x)
    # float32 dot is bit-exact; the double exp may differ in the last ulp of the DOUBLE, which survives the cast to
    # float32 only on a rounding tie: tolerance 1e-5 relative per north_star, and we report exactness
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=0)
    assert (got.view(np.uint32) != want.view(np.uint32)).mean() < 0.01
    k = int(c.n_weak[0])
    acc = np.float32(0)
    for i in range(k):
        acc = np.float32(acc + want[i])
    assert abs(gpu_handle.stage_predict(c.w[:k], c.bias[:k], x[:k]) - float(np.float32(acc / np.float32(k)))) <= 1e-5 * float(acc / k)


def test_stage_scores_on_windows(gpu_handle, oracle_cascade):
    img = synth.frame(240, 320, 12)
    S = O.integral(img)
    gpu_handle.integral(img, want_output=False)
    rng = np.random.default_rng(1)
    wins = []
    for _ in range(300):
        l = int(rng.integers(40, 240))
        wins.append((int(rng.integers(0, 320 - l + 1)), int(rng.integers(0, 240 - l + 1)), l))
    wins = np.array(wins, np.int32)
    want = O.stage_scores(S, oracle_cascade, wins)
    got = gpu_handle.stage_scores(wins, oracle_cascade.c.n_stages)
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=0)
    assert (got.view(np.uint32) != want.view(np.uint32)).mean() < 0.01


def _check_detect(gpu_handle, oracle_cascade, frames, prm_kwargs, base=40):
    prm = capi.params(base=base, **prm_kwargs)
    dets, cnts = gpu_handle.detect(frames, prm)
    for f, img in enumerate(frames):
        S = O.integral(img)
        oprm = O.params(base=base, step=prm_kwargs.get("step", 0), prefilter=prm_kwargs.get("prefilter", 6),
                        skip_rule=prm_kwargs.get("skip_rule", True), force_all=prm_kwargs.get("force_all_stages", False), nthreads=8)
        want = O.detect(S, oracle_cascade, oprm)
        if prm_kwargs.get("force_all_stages", False):
            # force_all only adds work: detections and the reference-equivalent counters are those of the unforced scan
            oprm.force_all = 0
            want = O.detect(S, oracle_cascade, oprm)
        mine = dets[dets["frame"] == f]
        c = cnts[f]
        assert c.grid == want.counters[O.C_GRID]
        assert c.visited == want.counters[O.C_VISITED]
        assert c.prefilter_pass == want.counters[O.C_PREFILTER]
        n_stages = oracle_cascade.c.n_stages
        assert [c.reach[s] for s in range(n_stages)] == [int(want.counters[O.C_REACH0 + s]) for s in range(n_stages)]
        if not prm_kwargs.get("force_all_stages", False):
            assert c.weak_evals == want.counters[O.C_WEAK]
        assert c.raw == len(want.x) == len(mine)
        # identical window set, in the same (l, y, x) order; scores bit-equal (the final score is a double)
        assert np.array_equal(mine["x"], want.x) and np.array_equal(mine["y"], want.y) and np.array_equal(mine["l"], want.l)
        np.testing.assert_allclose(mine["score"], want.score, rtol=1e-6, atol=0)
    return dets, cnts


def test_detect_c1_640x480(gpu_handle, oracle_cascade):
    """BASELINE config 1: 640x480, 40x40 window, scale 1.1, step 2, the reference-trained cascade."""
    dets, cnts = _check_detect(gpu_handle, oracle_cascade, [synth.frame(480, 640, 1)], {})
    assert cnts[0].grid == 1053107 and len(dets) > 0


def test_detect_batch_mixed_frames(gpu_handle, oracle_cascade):
    frames = [synth.frame(240, 320, s) for s in range(5)] + [synth.noise_frame(240, 320, 9)]
    _check_detect(gpu_handle, oracle_cascade, frames, {})


def test_detect_odd_size_and_base70(gpu_handle, oracle_cascade):
    _check_detect(gpu_handle, oracle_cascade, [synth.frame(301, 517, 4)], {}, base=70)  # the reference's checked-in base / step 3
    _check_detect(gpu_handle, oracle_cascade, [synth.frame(97, 131, 4)], {"step": 1})


def test_detect_no_skip_rule_and_no_prefilter(gpu_handle, oracle_cascade):
    img = synth.frame(200, 260, 6)
    _check_detect(gpu_handle, oracle_cascade, [img], {"skip_rule": False})
    _check_detect(gpu_handle, oracle_cascade, [img], {"prefilter": -1})
    _check_detect(gpu_handle, oracle_cascade, [img], {"prefilter": -1, "skip_rule": False, "force_all_stages": True, "step": 1})


def test_detect_frame_smaller_than_window(gpu_handle, oracle_cascade):
    dets, cnts = gpu_handle.detect([synth.noise_frame(30, 50, 0)], capi.params(base=40))
    assert len(dets) == 0 and cnts[0].grid == 0


def test_detect_1080p_full_size_properties(gpu_handle, oracle_cascade):
    """BASELINE config 2 at full size: oracle comparison (8 host threads, ~1 s) plus size-independent properties."""
    img = synth.frame(1080, 1920, 2)
    dets, cnts = _check_detect(gpu_handle, oracle_cascade, [img], {})
    assert cnts[0].grid == 11557983
    # idempotence: the same frame twice in one batch gives the same detections
    d2, c2 = gpu_handle.detect([img, img])
    a, b = d2[d2["frame"] == 0], d2[d2["frame"] == 1]
    assert np.array_equal(a[["x", "y", "l"]], b[["x", "y", "l"]]) and np.array_equal(a["score"], b["score"])
    assert np.array_equal(a[["x", "y", "l"]], dets[["x", "y", "l"]])
    # without the adaptive stride every reference detection is still found (visited set only grows)
    d3, _ = gpu_handle.detect([img], capi.params(skip_rule=False))
    s3 = set(map(tuple, d3[["x", "y", "l"]].tolist()))
    assert set(map(tuple, dets[["x", "y", "l"]].tolist())) <= s3


def test_group_rectangles_host(gpu_handle, oracle_cascade):
    img = synth.frame(480, 640, 1)
    dets, _ = gpu_handle.detect([img])
    rects = np.stack([dets["x"], dets["y"], dets["l"], dets["l"]], 1)
    gr, gs = capi.group_rectangles(rects, dets["score"])
    wr, ws = O.group_rectangles(rects, dets["score"])
    assert np.array_equal(gr, wr) and np.array_equal(gs, ws) and len(gr) > 0


def test_submit_collect_pipeline_equals_sync(gpu_handle):
    """sc_detect_submit / sc_detect_collect with two batches in flight returns what sc_detect returns, batch by batch."""
    import ctypes
    batches = [np.ascontiguousarray(np.stack([synth.frame(240, 320, 50 + 3 * b + i) for i in range(3)])) for b in range(5)]
    want = [gpu_handle.detect(b) for b in batches]
    ptrs = [(ctypes.c_void_p * 3)(*[b.ctypes.data + i * 240 * 320 for i in range(3)]) for b in batches]
    got = []
    pending = gpu_handle.detect_submit(ptrs[0], 3, 320, 240, 320)
    for k in range(len(batches)):
        nxt = gpu_handle.detect_submit(ptrs[k + 1], 3, 320, 240, 320) if k + 1 < len(batches) else None
        got.append(gpu_handle.detect_collect(pending, 3))
        pending = nxt
    for (d1, c1), (d2, c2) in zip(want, got):
        assert d1.tobytes() == d2.tobytes() and len(d1) > 0
        assert all(bytes(a) == bytes(b) for a, b in zip(c1, c2))
    # a third submit while two are in flight is refused, not queued
    a = gpu_handle.detect_submit(ptrs[0], 3, 320, 240, 320)
    b = gpu_handle.detect_submit(ptrs[1], 3, 320, 240, 320)
    with pytest.raises(capi.SurfCascadeError):
        gpu_handle.detect_submit(ptrs[2], 3, 320, 240, 320)
    gpu_handle.detect_collect(a, 3); gpu_handle.detect_collect(b, 3)


def _host_groups(dets, n_frames, thr=2, eps=0.2):
    out = []
    for f in range(n_frames):
        d = dets[dets["frame"] == f]
        gr, gs = O.group_rectangles(np.stack([d["x"], d["y"], d["l"], d["l"]], 1), d["score"], thr, eps) if len(d) else (np.zeros((0, 4), np.int32), np.zeros(0))
        out += [(f, int(r[0]), int(r[1]), int(r[2]), float(s)) for r, s in zip(gr, gs)]
    return out


def test_device_grouping_equals_group_rectangles(gpu_handle, oracle_cascade):
    """Next row N1 on the device: sc_detect with group_threshold returns cv::groupRectangles(raw, 2, 0.2) of every frame
    (oracle restatement pinned to cv2), in the same order, with the same rounded mean rects and best scores."""
    frames = [synth.frame(480, 640, 1), synth.frame(240, 320, 3), synth.noise_frame(120, 160, 4), synth.frame(480, 640, 8)]
    frames = [np.pad(f, ((0, 480 - f.shape[0]), (0, 640 - f.shape[1]))) for f in frames]
    raw, _ = gpu_handle.detect(frames)
    for thr, eps in ((2, 0.2), (1, 0.2), (3, 0.35)):
        got, cnts = gpu_handle.detect(frames, capi.params(group_threshold=thr, group_eps=eps))
        want = _host_groups(raw, len(frames), thr, eps)
        assert [(int(g["frame"]), int(g["x"]), int(g["y"]), int(g["l"]), float(g["score"])) for g in got] == want
        assert sum(c.raw for c in cnts) == len(raw)
    assert len(want) > 0


def test_device_grouping_falls_back_to_host_when_a_frame_is_crowded(oracle_cascade):
    """More raw windows in a frame than one CTA groups (2048): the batch is grouped on the host, same result."""
    c = oracle_cascade.c
    k = int(c.n_weak[0])
    h = capi.Handle(0)
    try:
        pool = O.pool_patches(40)
        h.set_cascade(40, np.array([0.20], np.float32), c.n_weak[:1], pool[c.patch_index[:k]], c.w[:k], c.bias[:k])
        img = synth.frame(240, 320, 41)
        raw, _ = h.detect([img], capi.params(), cap=1 << 20)
        assert len(raw) > 2048
        got, _ = h.detect([img], capi.params(group_threshold=2), cap=1 << 20)
        want = _host_groups(raw, 1)
        assert [(int(g["frame"]), int(g["x"]), int(g["y"]), int(g["l"]), float(g["score"])) for g in got] == want and len(want) > 0
    finally:
        h.close()
