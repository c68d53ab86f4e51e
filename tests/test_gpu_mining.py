"""GPU: hard-negative mining (next row N2) == FillNegSamples in its single-thread order.  The checker is the restatement in
tests/cascade_util.py (the cascade's accept decision is the oracle's detect with the prefilter and stride rule off on a
10-pixel lattice; the samples are the oracle's descriptors of all 608 pool patches projected into each taken window), which
tests/test_oracle_vs_ref.py pins against the reference's own FillNegSamples."""
import numpy as np
import pytest

from cascade_util import fill_neg_restated as _fill_neg
from oracle import oracle as O
from surfcascade_b200 import capi, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("first,need,layout", [(True, 700, "mixed"), (False, 60, "mixed"), (False, 100000, "mixed"), (False, 300, "runs"), (False, 1200, "runs"), (False, 492, "runs"), (False, 100000, "runs"),
                                               (True, 900, "runs")])
def test_mine_negatives_matches_fill_neg_samples(gpu_handle, oracle_cascade, first, need, layout):
    if layout == "mixed":   # every image its own size: one-image batches
        frames = [synth.negative_frame(20, 120, 160), np.zeros((30, 30), np.uint8), synth.frame(150, 200, 21), synth.negative_frame(22, 240, 320)]
    else:                   # runs of equally sized images: scanned as batches, the fill stops inside a batch (need = 300, 1200) or exactly at an image boundary (492)
        frames = [synth.negative_frame(30, 120, 160), synth.frame(120, 160, 31), synth.negative_frame(32, 120, 160), np.zeros((30, 30), np.uint8),
                  synth.frame(150, 200, 33), synth.frame(150, 200, 34), synth.negative_frame(35, 120, 160)]
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
