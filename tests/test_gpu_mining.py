"""GPU: hard-negative mining (next row N2) == FillNegSamples in its single-thread order, restated from pinned oracle pieces
(the cascade's accept decision is the oracle's detect with the prefilter and stride rule off on a 10-pixel lattice; the
samples are the oracle's descriptors of all 608 pool patches projected into each taken window)."""
import numpy as np
import pytest

from oracle import oracle as O
from surfcascade_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _fill_neg(frames, need, first, bc):
    pool = O.pool_patches(40)
    out, used = [], len(frames)
    for i, img in enumerate(frames):
        if len(out) >= need:
            break
        H, W = img.shape
        if W < 40 or H < 40:
            continue
        S = O.integral(img)
        prm = O.params(base=40, step=10, prefilter=-1, skip_rule=False)
        if first:
            wins = [(x, y, l) for l in O.scales(W, H, prm) for y in range(0, H - l + 1, 10) for x in range(0, W - l + 1, 10)]
        else:
            d = O.detect(S, bc, prm)
            wins = list(zip(d.x.tolist(), d.y.tolist(), d.l.tolist()))
        for (x, y, l) in wins[: need - len(out)]:
            r = O.project(40, l, pool)
            r[:, 0] += x; r[:, 1] += y
            out.append(O.features(S, r)[0])
        if len(out) == need:
            used = i + 1
    return (np.stack(out) if out else np.zeros((0, 608, 32), np.float32)), used


@pytest.mark.parametrize("first,need", [(True, 700), (False, 60), (False, 100000)])
def test_mine_negatives_matches_fill_neg_samples(gpu_handle, oracle_cascade, first, need):
    frames = [synth.negative_frame(20, 120, 160), np.zeros((30, 30), np.uint8), synth.frame(150, 200, 21), synth.negative_frame(22, 240, 320)]
    if not first:
        # the trained cascade rejects almost everything on these frames: cut it to its first stage so that samples exist
        import dataclasses
        c = oracle_cascade.c
        k = int(c.n_weak[0])
        bc = O.BoundCascade(dataclasses.replace(c, theta=np.array([0.40], np.float32), n_weak=c.n_weak[:1].copy(), patch_index=c.patch_index[:k].copy(),
                                                w=c.w[:k].copy(), bias=c.bias[:k].copy()))
        h = capi.Handle(0)
        h.set_cascade(40, bc.theta, bc.n_weak, bc.rects, bc.w, bc.bias)
    else:
        bc, h = oracle_cascade, gpu_handle
    try:
        want, used_w = _fill_neg(frames, need, first, bc)
        got, used_g = h.mine_negatives(frames, need, first)
        assert got.shape == want.shape and used_g == used_w
        assert len(want) > 0 and (need > 50000) == (len(want) < need)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    finally:
        if h is not gpu_handle:
            h.close()
