"""GPU parity tests of the compact integer plane (csrc/sc_plan.h): the 16-byte-per-corner restatement of IntegralImage
(DenseSURFFeatureExtractor.cpp:65-87) that the stage-0 fast filter reads, its certificate (k_cell_bounds), and the equality
of the scan's results with the float-plane-only scan and with the oracle.

Bars: integer work, so everything here is bit-exact.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import modelcfg
from oracle import oracle as O
from surfcascade_b200 import capi, synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODEL = os.path.join(ROOT, "tests", "golden", "model_c1.cfg")


def exact_integral(img):
    """Exact integer integral of the eight channels, int64 [H+1][W+1][8] (channels from the oracle's T2bFilter restatement)."""
    ch = O.channels(img).astype(np.int64)  # [8][H][W]
    h, w = img.shape
    S = np.zeros((h + 1, w + 1, 8), np.int64)
    S[1:, 1:, :] = np.cumsum(np.cumsum(ch, axis=1), axis=2).transpose(1, 2, 0)
    return S


def packed(S):
    return ((S[:, :, 1::2] << 16) + S[:, :, 0::2]).astype(np.uint64).astype(np.uint32)  # mod 2^32


def stripes(h, w):
    img = np.zeros((h, w), np.uint8)
    img[:, 2::4] = 255
    img[:, 3::4] = 255
    return img


@pytest.mark.parametrize("shape,kind,step", [((117, 203), "frame", 2), ((480, 640), "frame", 2), ((2, 2), "noise", 2), ((3, 33), "noise", 2),
                                             ((65, 32), "noise", 1), ((200, 97), "noise", 1), ((301, 517), "frame", 3), ((600, 700), "stripes", 2),
                                             ((1080, 1920), "noise", 2), ((1080, 1920), "frame", 2), ((95, 1400), "noise", 5)])
def test_compact_plane_is_the_exact_integral(gpu_handle, shape, kind, step):
    """Both forms of the walk kernel (tiled for even steps, shuffle scans for odd ones), the scan's own layout; values past
    2^24 and 2^32 included (the plane is defined modulo 2^32)."""
    h, w = shape
    img = {"frame": lambda: synth.frame(h, w, 3) if min(h, w) >= 64 else synth.noise_frame(h, w, 3), "noise": lambda: synth.noise_frame(h, w, 5),
           "stripes": lambda: stripes(h, w)}[kind]()
    got = gpu_handle.integral_compact(img, step)
    want = packed(exact_integral(img))
    assert got.shape == want.shape
    assert np.array_equal(got, want)


def test_box_sums_from_the_compact_plane_equal_the_float_path(gpu_handle):
    """Where every cell sum is below 65536 and the integrals are below 2^23, the compact box sums are the reference's float box
    sums: checked against exact integers, and through Normalize against the oracle's CalcFeature."""
    img = synth.frame(240, 320, 11)
    S = exact_integral(img)
    assert S.max() < 2 ** 23
    gpu_handle.integral(img, want_output=False)
    pool = O.pool_patches(40)
    rects = []
    for (x, y, l) in [(0, 0, 40), (7, 3, 44), (100, 50, 97), (0, 0, 240), (81, 1, 233), (33, 17, 121)]:
        pr = O.project(40, l, pool)
        pr[:, 0] += x; pr[:, 1] += y
        rects.append(pr)
    rects = np.concatenate(rects)
    got = gpu_handle.box_sums_compact(rects)
    want = np.zeros_like(got)
    for i, (x, y, w, h) in enumerate(rects):
        cells = np.zeros(16, np.int32)
        n = O.lib().so_cells(np.array([x, y, w, h], np.int32).ctypes.data_as(O.C.POINTER(O.C.c_int)), cells.ctypes.data_as(O.C.POINTER(O.C.c_int)))
        assert n == 4
        for k in range(4):
            cx, cy, cw, chh = cells[4 * k:4 * k + 4]
            want[i, 8 * k:8 * k + 8] = (S[cy + chh, cx + cw] + S[cy, cx] - S[cy, cx + cw] - S[cy + chh, cx]).astype(np.float32)
    small = want.max(axis=1) < 65536
    assert small.sum() > 3000
    assert np.array_equal(got[small].view(np.uint32), want[small].view(np.uint32))


@pytest.mark.parametrize("kind", ["frame", "noise", "stripes"])
def test_cell_bounds_are_upper_bounds_and_not_loose(gpu_handle, kind):
    h, w = 360, 500
    img = {"frame": lambda: synth.frame(h, w, 21), "noise": lambda: synth.noise_frame(h, w, 22), "stripes": lambda: stripes(h, w)}[kind]()
    gpu_handle.integral(img, want_output=False)
    S = exact_integral(img)
    ces = [1, 7, 16, 17, 23, 40, 64, 65, 100, 177, 300, 360]
    got = gpu_handle.cell_bounds(ces).astype(np.int64)
    for ce, b in zip(ces, got):
        c = min(ce, h, w)  # cells larger than the image do not exist; the bound then covers the whole image
        true_max = (S[c:, c:] + S[:-c, :-c] - S[c:, :-c] - S[:-c, c:]).max() if (ce <= h and ce <= w) else 0
        assert b >= true_max, (ce, b, true_max)
        g = max(ce // 4, 16)
        area_ratio = ((ce + g) / ce) ** 2
        assert b <= 255 * (ce + g) ** 2
        if kind == "noise" and 17 <= ce <= 100:
            assert b <= 1.35 * area_ratio * true_max, (ce, b, true_max)  # uniform noise: the bound is the area ratio away


def run_child(env_extra, code):
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


_CHILD = r"""
import sys, json, numpy as np
sys.path.insert(0, '.')
from surfcascade_b200 import capi, synth
h = capi.Handle(0); h.load_model('tests/golden/model_c1.cfg', 40)
frames = [synth.frame(540, 960, 31), synth.frame(540, 960, 32), synth.noise_frame(300, 400, 33)]
out = []
for group in (frames[:2], frames[2:]):
    dets, cnts = h.detect(group, capi.params())
    out.append({"dets": [[int(d["frame"]), int(d["x"]), int(d["y"]), int(d["l"]), float(d["score"])] for d in dets],
                "cnt": [[c.visited, c.prefilter_pass, c.weak_evals, c.raw, c.evaluated] + [c.reach[i] for i in range(4)] for c in cnts]})
print(json.dumps(out))
"""


def test_scan_with_and_without_the_compact_plane_is_identical():
    """SC_DISABLE_COMPACT=1 keeps the stage-0 filter on the float planes: detections, scores and every counter must be equal."""
    a = run_child({"SC_DISABLE_COMPACT": "0"}, _CHILD)
    b = run_child({"SC_DISABLE_COMPACT": "1"}, _CHILD)
    assert a == b
    assert '"dets": [[' in a


@pytest.mark.parametrize("seed,shape", [(41, (480, 640)), (42, (1080, 1920))])
def test_detect_parity_with_compact_plane_in_use(gpu_handle, seed, shape):
    """Detect parity against the oracle on frames where the certificate holds for most scales (the compact path runs)."""
    h, w = shape
    img = synth.frame(h, w, seed)
    gpu_handle.load_model(MODEL, 40)
    dets, cnts = gpu_handle.detect([img], capi.params())
    bc = O.BoundCascade(modelcfg.load(MODEL))
    want = O.detect(O.integral(img), bc, O.params(base=40, nthreads=8))
    assert np.array_equal(dets["x"], want.x) and np.array_equal(dets["y"], want.y) and np.array_equal(dets["l"], want.l)
    assert np.allclose(dets["score"], want.score, rtol=1e-6, atol=0)
    assert cnts[0].visited == want.counters[O.C_VISITED] and cnts[0].weak_evals == want.counters[O.C_WEAK]
    assert cnts[0].prefilter_pass == want.counters[O.C_PREFILTER]
    # the certificate of this frame: most projected cell edges of the stage-0 patches qualify
    gpu_handle.integral(img, want_output=False)
    ces = sorted({int(min(r[2], r[3]) if r[2] != r[3] else r[2] // 2) for l in O.scales(w, h, O.params(base=40))
                  for r in O.project(40, l, bc.rects[:3])})
    b = gpu_handle.cell_bounds(ces)
    assert (b < 65536).mean() > 0.5
