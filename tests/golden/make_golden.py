"""Generates the golden vectors in tests/golden/ from the REFERENCE ITSELF (oracle/_ref: the unmodified
mrgloom/SurfCascade sources compiled by oracle/Makefile).  Run where /root/reference exists:

    make -C oracle ref && python tests/golden/make_golden.py

The reference has no tests or fixtures of its own; these files pin the oracle
(oracle/surf_oracle.c) and, through it, the CUDA path.  Inputs are regenerated from seeds at test time
(surfcascade_b200/synth.py), only outputs are stored.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import refbind as R  # noqa: E402
from surfcascade_b200 import synth  # noqa: E402

MODEL = os.path.join(HERE, "model_c1.cfg")


def stripes(h, w):
    img = np.zeros((h, w), np.uint8)
    img[:, 2::4] = 255
    img[:, 3::4] = 255
    return img


def main():
    out = {}
    # 1. pool + projection geometry
    pool = R.pool_patches(40)
    out["pool"] = pool
    for l in (40, 44, 97, 233):
        out[f"project_l{l}"] = R.project(40, [0, 0, l, l], pool)
    # 2. channels + integral on a small seeded frame (bit patterns)
    img = synth.frame(67, 101, 7)
    out["channels_67x101_s7"] = R.channels(img)
    out["integral_67x101_s7"] = R.integral(img)
    # 3. descriptors + prefilter sums: all 608 pool patches projected at 3 window sides
    img2 = synth.frame(240, 320, 11)
    rects = []
    for (x, y, l) in [(0, 0, 40), (7, 3, 44), (81, 1, 233)]:
        p = R.project(40, [x, y, l, l], pool)
        rects.append(p)
    rects = np.concatenate(rects)
    f, s = R.features(img2, rects)
    out["feat_rects_240x320_s11"] = rects
    out["feat_240x320_s11"] = f
    out["sums_240x320_s11"] = s
    # 4. stage scores on explicit windows
    rng = np.random.default_rng(1)
    wins = []
    for _ in range(200):
        l = int(rng.integers(40, 240))
        wins.append((int(rng.integers(0, 320 - l + 1)), int(rng.integers(0, 240 - l + 1)), l))
    wins = np.array(wins, np.int32)
    out["stage_wins_240x320_s11"] = wins
    out["stage_scores_240x320_s11"] = R.stage_scores(img2, MODEL, wins)
    # 5. the detect loop: raw windows, scores, counters, grouped output
    for name, frame, base in [("160x120_s3", synth.frame(120, 160, 3), 40), ("640x480_s1", synth.frame(480, 640, 1), 40),
                              ("517x301_s4_b70", synth.frame(301, 517, 4), 70), ("noise_320x240_s9", synth.noise_frame(240, 320, 9), 40)]:
        r = R.detect([frame], MODEL, base=base, nthreads=1)
        out[f"det_{name}_xyl"] = np.stack([r.x, r.y, r.l], 1).astype(np.int32)
        out[f"det_{name}_score"] = r.score
        out[f"det_{name}_counters"] = r.counters[0]
        out[f"det_{name}_grect"] = r.g_rect
        out[f"det_{name}_gscore"] = r.g_score
    # 6. float32 integral past 2^24: digest + sampled rows of a 600x700 stripes frame and a 1080p noise frame
    for name, frame in [("stripes_600x700", stripes(600, 700)), ("noise_1080p_s5", synth.noise_frame(1080, 1920, 5))]:
        S = R.integral(frame)
        assert S.max() > 2 ** 24
        out[f"big_{name}_sha256"] = np.frombuffer(hashlib.sha256(S.tobytes()).digest(), np.uint8)
        out[f"big_{name}_lastrow"] = S[-1, ::37].copy()
        out[f"big_{name}_max"] = np.array([S.max()], np.float32)
    # 7. the model as the reference loads it
    import ctypes as C
    th = np.zeros(16, np.float32); nw = np.zeros(16, np.int32); pi = np.zeros(1024, np.int32); w = np.zeros((1024, 33), np.float32); b = np.zeros(1024)
    P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    S_ = R.lib().ref_model_load(MODEL.encode(), P(th, C.c_float), P(nw, C.c_int), 16, P(pi, C.c_int), P(w, C.c_float), P(b, C.c_double), 1024)
    k = int(nw[:S_].sum())
    out["model_theta"] = th[:S_]; out["model_n_weak"] = nw[:S_]; out["model_patch_index"] = pi[:k]; out["model_w"] = w[:k]; out["model_bias"] = b[:k]
    path = os.path.join(HERE, "golden_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
