"""Test helpers: synthetic cascades and a writer of the reference's model.cfg schema (Model::Save, Model.cpp:21-95), so that the
reference's own loader and detect loop can run cascades its trainer never produced."""
from __future__ import annotations

import numpy as np

from oracle.modelcfg import Cascade


def random_cascade(seed: int, n_weak, thetas, shapes: str = "mixed", w_sigma: float = 1.0) -> Cascade:
    """Weak classifiers on randomly drawn pool patches (squares 0-343, tall 344-475, wide 476-607), Gaussian weights."""
    rng = np.random.default_rng(seed)
    total = int(sum(n_weak))
    if shapes == "mixed":
        idx = np.concatenate([rng.integers(0, 344, total - 2 * (total // 3)), rng.integers(344, 476, total // 3), rng.integers(476, 608, total // 3)])
        rng.shuffle(idx)
    else:
        idx = rng.integers(0, 344, total)
    w = rng.normal(0, w_sigma, (total, 33)).astype(np.float32)
    w[:, 32] = rng.normal(0, 0.3, total).astype(np.float32)
    return Cascade(np.array(thetas, np.float32), np.array(n_weak, np.int32), idx.astype(np.int32), w, np.ones(total))


def _flt(v) -> str:
    """libconfig's float writer: %.10g, and a float must show a '.' or an exponent (libconfig.c:212-243)."""
    s = "%.10g" % float(v)
    return s if any(ch in s for ch in ".e") and "inf" not in s and "nan" not in s else s + ".0"


def write_model_cfg(path: str, c: Cascade) -> None:
    """Every setting Model::Load reads, in Model::Save's order; the training statistics carry placeholder values."""
    out = ["cascade_classifier : ", "{", "  max_stages_num = 10;", "  FPR_target = 9.999999975e-07;", "  TPR_min_perstage = 0.9950000048;",
           "  FPR = 1e-06;", "  TPR = 0.98;", "  stage_classifiers = ( "]
    k = 0
    stages = []
    for s in range(c.n_stages):
        st = ["    {", "      search_step = 0.009999999776;", "      auc_step = 0.05000000075;", "      TPR_min = 0.9950000048;", "      n_total = 1600;",
              "      n_pos = 800;", "      n_neg = 800;", "      FPR = 0.01;", "      TPR = 0.995;", f"      theta = {_flt(c.theta[s])};",
              "      total_AUC_score = 0.99;", "      sample_num = 960;", "      max_iters = 100;", "      weak_classifiers = ( "]
        weak = []
        for _ in range(int(c.n_weak[s])):
            ws = ", ".join(_flt(x) for x in c.w[k])
            weak.append("\n".join(["        {", f"          patch_index = {int(c.patch_index[k])};", "          eps = 0.01;", "          C = 0.1;",
                                   "          nr_class = 2;", "          nr_feature = 32;", f"          bias = {_flt(c.bias[k])};",
                                   f"          w = [ {ws} ];", "          label = [ 1, -1 ];", "        }"]))
            k += 1
        st.append(", \n".join(weak))
        st += ["      );", "    }"]
        stages.append("\n".join(st))
    out.append(", \n".join(stages))
    out += ["  );", "};", ""]
    with open(path, "w") as f:
        f.write("\n".join(out))


def fill_neg_restated(frames, need, first, bc):
    """FillNegSamples (DenseSURFFeatureExtractor.cpp:124-195) in its single-thread order, restated from the oracle's pieces:
    every window of the scale ladder on a 10-pixel lattice (prefilter and stride rule off), taken when `first` or when the
    cascade accepts it, its sample = the descriptors of all 608 pool patches projected into it; stops at `need` samples and
    reports how many images were consumed (the reference's static cursor idx = i + 1).  Pinned against the reference's own
    function in tests/test_oracle_vs_ref.py::test_fill_neg_samples_restatement."""
    from oracle import oracle as O
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
