"""Generates tests/golden/big_pins.npz: oracle results at BASELINE.json's full sizes, produced HERE by the reference itself
(oracle/_ref, the unmodified reference sources) where it has the mode, else by the plain-C restatement pinned to it:

  c2_*   the eight 1920x1080 frames bench.py times (synth.frame seeds 100..107), default scan: raw windows, scores, counters
         -- from oracle/_ref's lifted detect loop (ObjDetector.cpp:174-219);
  c4_*   3840x2160 (seed 300, 12 objects), step 1, prefilter off, stride rule off: raw windows, scores and reach counters.
         The forced-stages mode of config 4 is not in the reference; its detections and reach counters are by definition the
         unforced scan's, which the C restatement computes (so_detect; OpenMP over scales);
  c5_auc config 5 at full size, 100,000 samples x 608 candidates x 32 floats (7.78 GB, numpy default_rng(55) normals with a
         class shift, generated in 10 chunks of 10,000 samples so the GPU test can stream the same chunks): AUC per candidate
         from the restatement of StageClassifier::Evaluate (so_pool_eval), which tests/test_oracle_vs_ref.py pins to the reference.

    python tests/golden/make_big_pins.py      (about 10 minutes on 8 cores, 9 GB of RAM)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import modelcfg, oracle, refbind  # noqa: E402
from surfcascade_b200 import synth  # noqa: E402

MODEL = os.path.join(ROOT, "tests", "golden", "model_c1.cfg")
OUT = os.path.join(ROOT, "tests", "golden", "big_pins.npz")
C5_N, C5_P, C5_CHUNK = 100000, 608, 10000


def c5_chunk(k):
    """Samples [k * C5_CHUNK, (k + 1) * C5_CHUNK): positives first (n_pos = N / 2).  Feature 0 of every candidate separates the
    classes a little; the candidates' weights differ, so their AUCs spread over 0.3 .. 0.9."""
    rng = np.random.default_rng(5500 + k)
    X = rng.normal(0.0, 0.15, size=(C5_CHUNK, C5_P, 32)).astype(np.float32)
    if (k + 1) * C5_CHUNK <= C5_N // 2:
        X[:, :, 0] += 0.2
    return X


def c5_weights():
    rng = np.random.default_rng(55)
    W = rng.normal(0.0, 1.0, size=(C5_P, 33)).astype(np.float32)
    W[:, 0] = np.linspace(-6.0, 6.0, C5_P).astype(np.float32)
    return W


def main():
    nthreads = os.cpu_count() or 1
    out = {}
    bc = oracle.BoundCascade(modelcfg.load(MODEL))
    # ---- C2: the bench frames, by the reference itself
    t0 = time.time()
    frames = [synth.frame(1080, 1920, 100 + i) for i in range(8)]
    if refbind.available():
        r = refbind.detect(frames, MODEL, base=40, nthreads=nthreads, group=False)
        order = np.lexsort((r.x, r.y, r.l, r.frame))
        out["c2_frame"], out["c2_x"], out["c2_y"], out["c2_l"], out["c2_score"] = (a[order] for a in (r.frame, r.x, r.y, r.l, r.score))
        out["c2_counters"] = r.counters[:, :4].astype(np.int64)   # visited, prefilter_pass, weak_evals, raw
        out["c2_source"] = np.array("oracle/_ref (reference sources)")
    else:
        fr, xs, ys, ls, ss, cn = [], [], [], [], [], []
        for f, img in enumerate(frames):
            d = oracle.detect(oracle.integral(img), bc, oracle.params(base=40, nthreads=nthreads))
            fr.append(np.full(len(d.x), f, np.int32)); xs.append(d.x); ys.append(d.y); ls.append(d.l); ss.append(d.score)
            cn.append([d.counters[oracle.C_VISITED], d.counters[oracle.C_PREFILTER], d.counters[oracle.C_WEAK], len(d.x)])
        out["c2_frame"], out["c2_x"], out["c2_y"], out["c2_l"], out["c2_score"] = (np.concatenate(a) for a in (fr, xs, ys, ls, ss))
        out["c2_counters"] = np.array(cn, np.int64)
        out["c2_source"] = np.array("oracle/surf_oracle.c (restatement)")
    print(f"C2: {len(out['c2_x'])} raw windows on 8 frames, {time.time() - t0:.0f} s, {out['c2_source']}", flush=True)
    # ---- C4: 4K, step 1, no prefilter, no stride rule
    t0 = time.time()
    img = synth.frame(2160, 3840, 300, n_objects=12)
    d = oracle.detect(oracle.integral(img), bc, oracle.params(base=40, step=1, prefilter=-1, skip_rule=False, nthreads=nthreads), cap=1 << 23)
    out["c4_x"], out["c4_y"], out["c4_l"], out["c4_score"] = d.x, d.y, d.l, d.score
    out["c4_reach"] = d.counters[oracle.C_REACH0:oracle.C_REACH0 + bc.c.n_stages].astype(np.int64)
    out["c4_grid"] = np.int64(d.counters[oracle.C_GRID])
    print(f"C4: {len(d.x)} raw windows of {int(out['c4_grid'])}, reach {out['c4_reach'].tolist()}, {time.time() - t0:.0f} s", flush=True)
    # ---- C5: full-size pool evaluation
    t0 = time.time()
    X = np.empty((C5_N, C5_P, 32), np.float32)
    for k in range(C5_N // C5_CHUNK):
        X[k * C5_CHUNK:(k + 1) * C5_CHUNK] = c5_chunk(k)
    W = c5_weights()
    out["c5_auc"] = oracle.pool_eval(X, C5_N // 2, W, np.ones(C5_P))
    print(f"C5: AUC of {C5_P} candidates over {C5_N} samples, range {out['c5_auc'].min():.3f} .. {out['c5_auc'].max():.3f}, {time.time() - t0:.0f} s", flush=True)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
