"""GPU: BASELINE configs 3-5 at (near) full size through size-independent properties; the oracle comparison at these
sizes would take minutes of CPU."""
import numpy as np
import pytest

from surfcascade_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _key(d):
    return d[["x", "y", "l"]].tolist()


def test_c3_many_1080p_frames_are_independent(gpu_handle):
    """Config 3 (sharded 1080p batch): 40 frames cycling over 3 distinct ones cross the 32-frame integral super-group and
    the 8-frame scan groups; equal frames must give equal detections and counters wherever they sit in the batch."""
    base = [synth.frame(1080, 1920, 200 + i) for i in range(3)]
    frames = [base[i % 3] for i in range(40)]
    dets, cnts = gpu_handle.detect(frames)
    ref = {}
    for f in range(40):
        mine = dets[dets["frame"] == f]
        sig = (_key(mine), mine["score"].tolist(), cnts[f].visited, cnts[f].prefilter_pass, cnts[f].weak_evals, cnts[f].raw, cnts[f].evaluated)
        assert ref.setdefault(f % 3, sig) == sig
        assert cnts[f].grid == 11557983 and cnts[f].visited <= cnts[f].evaluated <= cnts[f].grid
    assert all(len(ref[k][0]) > 0 for k in ref)


def test_c4_4k_step1_forced_stages(gpu_handle):
    """Config 4: 3840x2160, step 1, prefilter off, skip rule off, every stage of every window forced (242,480,344 windows).
    Forcing only adds work: the detections must equal the unforced scan's."""
    img = synth.frame(2160, 3840, 300, n_objects=12)
    forced = capi.params(step=1, prefilter=-1, skip_rule=False, force_all_stages=True)
    plain = capi.params(step=1, prefilter=-1, skip_rule=False, force_all_stages=False)
    d1, c1 = gpu_handle.detect([img], forced, cap=1 << 22)
    d2, c2 = gpu_handle.detect([img], plain, cap=1 << 22)
    assert c1[0].grid == 242480344 == c1[0].evaluated == c1[0].visited == c1[0].prefilter_pass
    assert _key(d1) == _key(d2) and np.array_equal(d1["score"], d2["score"]) and len(d1) > 0
    assert [c1[0].reach[s] for s in range(4)] == [c2[0].reach[s] for s in range(4)]
    # default parameters on the same frame: the adaptive stride visits a subset, detections are a subset
    d3, c3 = gpu_handle.detect([img], capi.params(step=1), cap=1 << 22)
    assert set(map(tuple, _key(d3))) <= set(map(tuple, _key(d1)))
    assert c3[0].visited < c3[0].grid


def test_c5_pool_eval_at_scale(gpu_handle):
    """Config 5 shape at 1/25 of the sample count: 4,000 samples x 608 candidates x 32 floats (311 MB) streamed once.
    Properties: a candidate whose weights separate the classes gets AUC ~ 0.95 (the reference's 20-threshold ROC never
    reaches (1,1)); negating the weights mirrors the ROC; a constant classifier scores exactly 0.5."""
    rng = np.random.default_rng(0)
    N, P = 4000, 608
    n_pos = N // 2
    X = rng.normal(0, 0.15, size=(N, P, 32)).astype(np.float32)
    X[:n_pos, :, 0] += 0.5    # feature 0 separates the classes
    labels = np.zeros(N, np.uint8); labels[:n_pos] = 1
    W = np.zeros((P, 33), np.float32)
    W[0::3, 0] = 8.0          # good candidates
    W[1::3, 0] = -8.0         # inverted candidates
    auc = gpu_handle.pool_eval(X, labels, W, np.ones(P))
    # all-zero weights give p = 0.5 for every sample: one ROC step from (0,0) to (1,1), area exactly 0.5
    assert (auc[0::3] > 0.9).all() and (auc[1::3] < 0.1).all() and (auc[2::3] == 0.5).all()


@pytest.mark.parametrize("bands", [2, 3, 8])
def test_row_bands_partition_the_scan(gpu_handle, bands):
    """Multi-GPU split of a single frame: every rank scans one band of each scale's lattice rows.  The union
    of the bands' detections and the sum of their counters must be the whole-frame scan's, bit for bit."""
    img = synth.frame(1080, 1920, 77)
    full, cf = gpu_handle.detect([img])
    parts, tot = [], dict(grid=0, visited=0, prefilter_pass=0, weak_evals=0, raw=0)
    for k in range(bands):
        d, c = gpu_handle.detect([img], capi.params(band_index=k, band_count=bands))
        parts.append(d)
        for key in tot:
            tot[key] += getattr(c[0], key)
    got = np.sort(np.concatenate(parts), order=["frame", "l", "y", "x"])
    assert got.tobytes() == full.tobytes() and len(full) > 0
    assert tot == {key: getattr(cf[0], key) for key in tot}
    # the bands are disjoint in y per scale
    for l in np.unique(full["l"]):
        ys = [set(p[p["l"] == l]["y"].tolist()) for p in parts]
        assert all(not (ys[a] & ys[b]) for a in range(bands) for b in range(a + 1, bands))
