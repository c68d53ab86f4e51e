"""Generates tests/golden/model_paper8.cfg: SURVEY.md 8(d)'s "paper-shaped" cascade -- 8 stages with 2, 3, 5, 8, 12, 16, 24, 32
weak classifiers (the stage sizes of Li & Zhang's SURF cascade), patches drawn uniformly from the 608-patch pool, weights
w ~ N(0, 1.5^2), numpy default_rng(7) -- so that the work of a scan does not depend on what the reference's trainer happened to
select.  The survey fixes no thresholds; they are set here, deterministically, from the cascade's own score distribution on one
seeded frame (synth.frame(480, 640, 7), every 3rd lattice window of every scale): stage 0 passes 10 % of the windows that reach
it, stage 1 35 %, every later stage 50 % -- a cascade whose stage 0 is much less selective than model_c1.cfg's (0.3 %), i.e. one
that keeps the later stages busy.  Scores come from the plain-C oracle.

    python tests/golden/make_paper8_model.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cascade_util import write_model_cfg  # noqa: E402
from oracle import oracle  # noqa: E402
from oracle.modelcfg import Cascade  # noqa: E402
from surfcascade_b200 import synth  # noqa: E402

N_WEAK = [2, 3, 5, 8, 12, 16, 24, 32]
PASS = [0.10, 0.35, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5]
OUT = os.path.join(ROOT, "tests", "golden", "model_paper8.cfg")


def main():
    rng = np.random.default_rng(7)
    total = sum(N_WEAK)
    idx = rng.integers(0, 608, total).astype(np.int32)
    w = rng.normal(0.0, 1.5, (total, 33)).astype(np.float32)
    bias = np.ones(total)
    pool = oracle.pool_patches(40)
    img = synth.frame(480, 640, 7)
    S = oracle.integral(img)
    H, W = img.shape
    wins = []
    for l in oracle.scales(W, H, oracle.params(base=40)):
        xs = np.arange(0, W - l + 1, 6); ys = np.arange(0, H - l + 1, 6)
        X, Y = np.meshgrid(xs, ys)
        wins.append(np.stack([X.ravel(), Y.ravel(), np.full(X.size, l)], 1))
    wins = np.concatenate(wins).astype(np.int32)
    proj = {int(l): oracle.project(40, int(l), pool) for l in np.unique(wins[:, 2])}
    alive = np.arange(len(wins))
    theta, k = [], 0
    for s, nw in enumerate(N_WEAK):
        acc = np.zeros(len(alive), np.float32)
        for q in range(nw):
            rects = np.stack([proj[int(l)][idx[k]] for l in wins[alive, 2]]).astype(np.int32)
            rects[:, 0] += wins[alive, 0]; rects[:, 1] += wins[alive, 1]
            f, _ = oracle.features(S, rects)
            p = np.array([oracle.weak(w[k], 1.0, f[i]) for i in range(len(f))], np.float32)
            acc = (acc + p).astype(np.float32)
            k += 1
        score = (acc / np.float32(nw)).astype(np.float32)
        t = np.float32(np.quantile(score, 1.0 - PASS[s]))
        theta.append(t)
        alive = alive[score >= t]
        print(f"stage {s}: {nw} weak, theta {t:.6f}, {len(alive)} windows pass")
    write_model_cfg(OUT, Cascade(np.array(theta, np.float32), np.array(N_WEAK, np.int32), idx, w, bias))
    print("wrote", OUT)


if __name__ == "__main__":
    main()
