"""CPU: the oracle against the reference build itself (oracle/_ref), on fresh seeded inputs.  Skipped where the
reference was never compiled (it needs /root/reference at build time; the .so travels with the snapshot)."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import refbind as R
from surfcascade_b200 import synth

from conftest import MODEL_C1

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("shape,seed", [((64, 64), 0), ((117, 203), 1), ((255, 257), 2)])
def test_integral_and_channels(shape, seed):
    for img in (synth.frame(*shape, seed), synth.noise_frame(*shape, seed)):
        assert np.array_equal(O.channels(img), R.channels(img))
        assert np.array_equal(bits(O.integral(img)), bits(R.integral(img)))


def test_opencv_standin_matches_cv2():
    """The stand-in cv::integral / cv::groupRectangles used to compile the reference against real OpenCV (cv2 wheel)."""
    cv2 = pytest.importorskip("cv2")
    img = synth.noise_frame(300, 500, 3)
    ch = R.channels(img)
    S = R.integral(img)
    for c in range(8):
        assert np.array_equal(bits(S[:, :, c]), bits(cv2.integral(ch[c], sdepth=cv2.CV_32F)))
    rng = np.random.default_rng(0)
    for trial in range(60):
        n = int(rng.integers(1, 60))
        base = rng.integers(0, 200, size=(max(1, n // 6), 2))
        r = np.concatenate([base[rng.integers(0, len(base), n)] + rng.integers(-6, 7, size=(n, 2)), rng.integers(30, 60, size=(n, 1))], 1)
        rects = np.stack([r[:, 0], r[:, 1], r[:, 2], r[:, 2]], 1).astype(np.int32)
        want, _ = cv2.groupRectangles(rects.tolist(), 2, 0.2)
        got, _ = O.group_rectangles(rects, np.ones(n))
        assert sorted(map(tuple, np.asarray(want).reshape(-1, 4).tolist())) == sorted(map(tuple, got.tolist()))


def test_features_bit_exact(oracle_cascade):
    img = synth.frame(200, 300, 5)
    S = O.integral(img)
    pool = O.pool_patches(40)
    for (x, y, l) in [(0, 0, 40), (3, 5, 53), (60, 0, 200), (11, 13, 171)]:
        pr = R.project(40, [x, y, l, l], pool)
        po = O.project(40, l, pool)
        po[:, 0] += x; po[:, 1] += y
        assert np.array_equal(pr, po)
        fr, sr = R.features(img, pr)
        fo, so = O.features(S, po)
        assert np.array_equal(bits(fr), bits(fo)) and np.array_equal(bits(sr), bits(so))


@pytest.mark.parametrize("shape,seed,base,threads", [((120, 160), 0, 40, 1), ((240, 320), 1, 40, 4), ((301, 517), 2, 70, 2), ((480, 640), 3, 40, 8)])
def test_detect_identical(oracle_cascade, shape, seed, base, threads):
    img = synth.frame(*shape, seed)
    r = R.detect([img], MODEL_C1, base=base, nthreads=threads)
    d = O.detect(O.integral(img), oracle_cascade, O.params(base=base, nthreads=threads))
    assert np.array_equal(d.x, r.x) and np.array_equal(d.y, r.y) and np.array_equal(d.l, r.l) and np.array_equal(d.score, r.score)
    c = r.counters[0]
    assert (d.counters[O.C_VISITED], d.counters[O.C_PREFILTER], d.counters[O.C_WEAK], d.counters[O.C_RAW]) == (c[0], c[1], c[2], c[3])


def test_pool_eval():
    """Row A9 / config 5: the oracle's candidate scoring against the reference's own StageClassifier::Evaluate."""
    rng = np.random.default_rng(0)
    pool = O.pool_patches(40)
    P = 40
    sel = np.linspace(0, 607, P).astype(int)
    N, n_pos = 90, 40
    X = np.zeros((N, P, 32), np.float32)
    for n in range(N):
        img = synth.positive(n) if n < n_pos else np.ascontiguousarray(synth.negative_frame(n)[:40, :40])
        X[n] = O.features(O.integral(img), pool[sel])[0]
    W = rng.normal(0, 1.2, size=(P, 33)).astype(np.float32)
    b = np.ones(P)
    assert np.array_equal(bits(O.pool_eval(X, n_pos, W, b)), bits(R.pool_eval(X, n_pos, W, b)))
    prev_patch = [3, 17, 5]
    prev_w = rng.normal(0, 1, size=(3, 33)).astype(np.float32)
    prior = np.zeros(N, np.float32)
    for t in range(3):
        for n in range(N):
            prior[n] = np.float32(prior[n] + np.float32(O.weak(prev_w[t], 1.0, X[n, prev_patch[t]])))
    assert np.array_equal(bits(O.pool_eval(X, n_pos, W, b, prior, 3)), bits(R.pool_eval(X, n_pos, W, b, prev_patch, prev_w, [1.0] * 3)))


@pytest.mark.parametrize("seed,n_weak,thetas", [(11, [2, 3, 2, 2, 2], [0.45, 0.60, 0.50, 0.50, 0.50]),
                                                 (12, [2] * 10, [0.45, 0.47, 0.48, 0.48, 0.48, 0.48, 0.48, 0.48, 0.48, 0.48]),
                                                 (13, [3, 2, 2, 2, 2, 2], [0.50, 0.50, 0.50, 0.50, 0.50, 0.50])])
def test_detect_identical_on_deep_synthetic_cascades(tmp_path, seed, n_weak, thetas):
    """The stride rule multi = (score < 0.5 ? 2 : 1) with score = (s + p + 1) / N (ObjDetector.cpp:201,214) depends on the
    stage count: the trained 4-stage cascade only ever skips after a stage-0 reject.  With 5 stages a stage-1 reject skips
    iff s < 0.5, with 6 and 10 stages rejects up to stage 1 / 3 always skip and later ones never: cascades written in the
    reference's model.cfg schema, loaded by its own Model::Load, run through its own detect loop, against the oracle."""
    from cascade_util import random_cascade, write_model_cfg
    from oracle import modelcfg
    c = random_cascade(seed, n_weak, thetas)
    cfg = str(tmp_path / "deep.cfg")
    write_model_cfg(cfg, c)
    back = modelcfg.load(cfg)  # what both loaders see: %.10g text -> double -> float32
    assert back.n_stages == len(n_weak) and np.array_equal(back.patch_index, c.patch_index) and np.array_equal(back.theta, c.theta)
    assert np.array_equal(back.w, c.w)
    bc = O.BoundCascade(back)
    for shape, fs in (((200, 260), 5), ((97, 333), 6)):
        img = synth.frame(*shape, fs)
        r = R.detect([img], cfg, base=40, nthreads=2, group=False)
        d = O.detect(O.integral(img), bc, O.params(base=40, nthreads=2), cap=1 << 21)
        assert np.array_equal(d.x, r.x) and np.array_equal(d.y, r.y) and np.array_equal(d.l, r.l) and np.array_equal(d.score, r.score)
        cr = r.counters[0]
        assert (d.counters[O.C_VISITED], d.counters[O.C_PREFILTER], d.counters[O.C_WEAK], d.counters[O.C_RAW]) == (cr[0], cr[1], cr[2], cr[3])
        reach = d.counters[O.C_REACH0:O.C_REACH0 + len(n_weak)]
        assert d.counters[O.C_VISITED] < d.counters[O.C_GRID] and (reach > 0).all()  # strides of 2 happen, every stage is entered


@pytest.mark.parametrize("first,totals", [(True, [700, 300]), (False, [60, 25]), (False, [100000])])
def test_fill_neg_samples_restatement(tmp_path, oracle_cascade, first, totals):
    """Next row N2: the reference's own DenseSURFFeatureExtractor::FillNegSamples (:124-195), run on PGM files in a child
    process (its image cursor is a function-local static), against the restatement the GPU mining test checks against:
    same windows in the same order, bit-identical 608 x 32 samples, the same `done`, and -- through a second call on the same
    extractor, the way CascadeClassifier::Train calls it -- the same cursor (the image after the one that completed a call)."""
    import dataclasses
    from cascade_util import fill_neg_restated, write_model_cfg
    frames = [synth.negative_frame(20, 120, 160), np.zeros((30, 30), np.uint8), synth.frame(150, 200, 21), synth.negative_frame(22, 240, 320),
              synth.frame(97, 131, 23)]
    cfg = ""
    bc = oracle_cascade
    if not first:
        # the trained cascade rejects almost everything on these frames: its first stage with a lower threshold accepts some
        c = oracle_cascade.c
        k = int(c.n_weak[0])
        cut = dataclasses.replace(c, theta=np.array([0.40], np.float32), n_weak=c.n_weak[:1].copy(), patch_index=c.patch_index[:k].copy(),
                                  w=c.w[:k].copy(), bias=c.bias[:k].copy())
        cfg = str(tmp_path / "stage0.cfg")
        write_model_cfg(cfg, cut)
        from oracle import modelcfg
        bc = O.BoundCascade(modelcfg.load(cfg))
    got, dones = R.fill_neg(frames, cfg, totals, first)
    start = 0
    for call, need in enumerate(totals):
        want, used = fill_neg_restated(frames[start:], need, first, bc)
        assert got[call].shape == want.shape, (call, got[call].shape, want.shape)
        assert np.array_equal(bits(got[call]), bits(want))
        assert dones[call] == (len(want) == need)
        start += used
    assert len(got[0]) > 0


def test_paper_shaped_cascade_reference_vs_restatement():
    """tests/golden/model_paper8.cfg (8 stages, 2 / 3 / 5 / 8 / 12 / 16 / 24 / 32 weak classifiers) through the reference's
    own Model::Load and detect loop against the restatement: window set, scores and counters.  (The GPU path is checked against the
    restatement on this model in tests/test_gpu_big_pins.py.)"""
    import os
    from oracle import modelcfg
    model = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "model_paper8.cfg")
    c = modelcfg.load(model)
    assert c.n_stages == 8 and c.n_weak.tolist() == [2, 3, 5, 8, 12, 16, 24, 32]
    bc = O.BoundCascade(c)
    for img in (synth.frame(240, 320, 1), synth.frame(150, 400, 2)):
        d = O.detect(O.integral(img), bc, O.params(base=40))
        r = R.detect([img], model, base=40, nthreads=1, group=False)
        assert len(d.x) == len(r.x)
        order = np.lexsort((r.x, r.y, r.l))
        assert np.array_equal(d.x, r.x[order]) and np.array_equal(d.y, r.y[order]) and np.array_equal(d.l, r.l[order])
        assert np.array_equal(d.score, r.score[order])
        assert r.counters[0, :3].tolist() == [d.counters[O.C_VISITED], d.counters[O.C_PREFILTER], d.counters[O.C_WEAK]]
        assert d.counters[O.C_REACH0 + 3] > 0   # deep stages are entered
