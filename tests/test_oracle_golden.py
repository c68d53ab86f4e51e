"""CPU: the oracle (oracle/surf_oracle.c, oracle/modelcfg.py) against the committed golden vectors, which were
produced by the reference itself (tests/golden/make_golden.py -> oracle/_ref).  Bit-exact throughout."""
import hashlib
import os

import numpy as np
import pytest

from oracle import modelcfg as M
from oracle import oracle as O
from surfcascade_b200 import synth

from conftest import GOLDEN, MODEL_C1


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "golden_ref.npz"))


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def test_pool_and_projection(gold):
    pool = O.pool_patches(40)
    assert len(pool) == 608  # 344 + 132 + 132, DenseSURFFeatureExtractor.cpp:49-63
    assert np.array_equal(pool, gold["pool"])
    for l in (40, 44, 97, 233):
        assert np.array_equal(O.project(40, l, pool), gold[f"project_l{l}"])


def test_channels_and_integral(gold):
    img = synth.frame(67, 101, 7)
    assert np.array_equal(O.channels(img), gold["channels_67x101_s7"])
    assert np.array_equal(bits(O.integral(img)), bits(gold["integral_67x101_s7"]))


def test_descriptors_and_prefilter_sums(gold):
    S = O.integral(synth.frame(240, 320, 11))
    f, s = O.features(S, gold["feat_rects_240x320_s11"])
    assert np.array_equal(bits(f), bits(gold["feat_240x320_s11"]))
    assert np.array_equal(bits(s), bits(gold["sums_240x320_s11"]))


def test_stage_scores(gold, oracle_cascade):
    S = O.integral(synth.frame(240, 320, 11))
    got = O.stage_scores(S, oracle_cascade, gold["stage_wins_240x320_s11"])
    assert np.array_equal(bits(got), bits(gold["stage_scores_240x320_s11"]))


@pytest.mark.parametrize("name,shape,seed,base,noise", [("160x120_s3", (120, 160), 3, 40, False), ("640x480_s1", (480, 640), 1, 40, False),
                                                        ("517x301_s4_b70", (301, 517), 4, 70, False), ("noise_320x240_s9", (240, 320), 9, 40, True)])
def test_detect_loop(gold, oracle_cascade, name, shape, seed, base, noise):
    img = synth.noise_frame(*shape, seed) if noise else synth.frame(*shape, seed)
    d = O.detect(O.integral(img), oracle_cascade, O.params(base=base, nthreads=4))
    xyl = gold[f"det_{name}_xyl"]
    assert np.array_equal(np.stack([d.x, d.y, d.l], 1), xyl)
    assert np.array_equal(d.score, gold[f"det_{name}_score"])
    c = gold[f"det_{name}_counters"]  # visited, prefilter_pass, weak_evals, raw, reach[16]
    assert d.counters[O.C_VISITED] == c[0] and d.counters[O.C_PREFILTER] == c[1] and d.counters[O.C_WEAK] == c[2] and d.counters[O.C_RAW] == c[3]
    assert np.array_equal(d.counters[O.C_REACH0:O.C_REACH0 + 16], c[4:20])
    gr, gs = O.group_rectangles(np.stack([d.x, d.y, d.l, d.l], 1), d.score)
    assert sorted(map(tuple, gr.tolist())) == sorted(map(tuple, gold[f"det_{name}_grect"].tolist()))
    assert sorted(gs.tolist()) == sorted(gold[f"det_{name}_gscore"].tolist())


def test_integral_past_2_24(gold):
    img = np.zeros((600, 700), np.uint8)
    img[:, 2::4] = 255
    img[:, 3::4] = 255
    S = O.integral(img)
    assert S.max() > 2 ** 24 and S.max() == gold["big_stripes_600x700_max"][0]
    assert hashlib.sha256(S.tobytes()).digest() == gold["big_stripes_600x700_sha256"].tobytes()
    # the exact integer integral differs once a column passes 2^24: the float recurrence is what the reference computes
    exact = np.cumsum(np.cumsum(O.channels(img)[1].astype(np.int64), 0), 1)
    assert (S[1:, 1:, 1] != exact.astype(np.float32)).any()


def test_integral_1080p_noise_digest(gold):
    S = O.integral(synth.noise_frame(1080, 1920, 5))
    assert hashlib.sha256(S.tobytes()).digest() == gold["big_noise_1080p_s5_sha256"].tobytes()
    assert np.array_equal(bits(S[-1, ::37]), bits(gold["big_noise_1080p_s5_lastrow"]))


def test_model_parse_matches_reference_loader(gold):
    c = M.load(MODEL_C1)
    assert np.array_equal(bits(c.theta), bits(gold["model_theta"])) and np.array_equal(c.n_weak, gold["model_n_weak"])
    assert np.array_equal(c.patch_index, gold["model_patch_index"]) and np.array_equal(bits(c.w), bits(gold["model_w"]))
    assert np.array_equal(c.bias, gold["model_bias"])


def test_scale_ladder_and_grid_sizes():
    """SURVEY.md Appendix D closed forms."""
    prm = O.params(base=40)
    assert len(O.scales(640, 480, prm)) == 27 and len(O.scales(1920, 1080, prm)) == 35 and len(O.scales(3840, 2160, prm)) == 42
    assert O.scales(1920, 1080, prm)[-1] == 1021
    grid = lambda W, H, step, sides: sum(((W - l) // step + 1) * ((H - l) // step + 1) for l in sides)
    assert grid(640, 480, 2, O.scales(640, 480, prm)) == 1053107
    assert grid(1920, 1080, 2, O.scales(1920, 1080, prm)) == 11557983
    assert grid(3840, 2160, 1, O.scales(3840, 2160, prm)) == 242480344
    assert grid(640, 480, 3, O.scales(640, 480, O.params(base=70))) == 302685
