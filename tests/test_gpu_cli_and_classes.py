"""GPU: the kept C++ surface -- the ObjDetector command line (detect branch) end to end against the oracle."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from oracle import refbind as R
from surfcascade_b200 import build, synth

from conftest import MODEL_C1

pytestmark = pytest.mark.gpu


def test_objdetector_cli_matches_oracle_grouped_output(tmp_path, oracle_cascade):
    assert os.path.exists(build.CLI), "ObjDetector binary not built"
    names = []
    for s in range(3):
        p = str(tmp_path / f"img{s}.pgm")
        R.write_pgm(p, synth.frame(240, 320, 20 + s))
        names.append(p)
    out = str(tmp_path / "det.txt")
    subprocess.check_call([build.CLI, "--detect", "--model", MODEL_C1, "--base", "40", "--out", out] + names, stdout=subprocess.DEVNULL)
    lines = open(out).read().split("\n")
    k = 0
    for s, name in enumerate(names):
        assert lines[k] == name
        n = int(lines[k + 1])
        got = [tuple(float(v) for v in lines[k + 2 + i].split()) for i in range(n)]
        k += 2 + n
        d = O.detect(O.integral(synth.frame(240, 320, 20 + s)), oracle_cascade, O.params(base=40))
        wr, ws = O.group_rectangles(np.stack([d.x, d.y, d.l, d.l], 1), d.score)
        want = [tuple(map(float, r)) + (float(f"{sc:.6g}"),) for r, sc in zip(wr.tolist(), ws.tolist())]  # ostream << double prints 6 significant digits
        assert sorted(got) == sorted(want)


def test_kept_classes_run_the_reference_detect_branch(tmp_path, oracle_cascade):
    """tests/cpp/class_detect.cpp is the reference's detect branch (ObjDetector.cpp:107-143,174-219) written against the kept
    classes -- Model::Load, ExtractPatches, GetFittedPatchIndexes, IntegralImage, sum, ProjectPatches, CalcFeature,
    stage_classifiers[p]->Predict2 / ->theta, and CascadeClassifier::Predict on the pool layout: its windows, scores and
    counters must be the oracle's."""
    assert os.path.exists(build.CLASS_TEST), "class_detect binary not built"
    img = synth.frame(120, 160, 77)
    pgm = str(tmp_path / "f.pgm")
    R.write_pgm(pgm, img)
    wins = [(0, 0, 40), (12, 6, 44), (40, 20, 97), (61, 3, 70)]
    S = O.integral(img)
    d = O.detect(S, oracle_cascade, O.params(base=40))
    if len(d.x):
        wins.append((int(d.x[0]), int(d.y[0]), int(d.l[0])))
    out = subprocess.run([build.CLASS_TEST, MODEL_C1, pgm, "40"] + [str(v) for w in wins for v in w], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln.split() for ln in out.stdout.strip().split("\n")]
    dets = sorted((int(a[2]), int(a[1]), int(a[0]), float(a[3])) for a in lines if a[0] not in ("counters", "predict"))
    want = sorted(zip(d.l.tolist(), d.y.tolist(), d.x.tolist(), d.score.tolist()))
    assert [t[:3] for t in dets] == [t[:3] for t in want]
    assert np.allclose([t[3] for t in dets], [t[3] for t in want], rtol=1e-6, atol=0)
    cnt = [a for a in lines if a[0] == "counters"][0]
    assert [int(v) for v in cnt[1:]] == [int(d.counters[O.C_VISITED]), int(d.counters[O.C_PREFILTER]), int(d.counters[O.C_WEAK]), len(d.x)]
    assert len(d.x) > 0
    # CascadeClassifier::Predict / StageClassifier::Predict on the 608-descriptor pool layout
    pred = [a for a in lines if a[0] == "predict"]
    assert len(pred) == len(wins)
    sc = O.stage_scores(S, oracle_cascade, np.array(wins, np.int32))
    theta = oracle_cascade.theta
    for a, w, row in zip(pred, wins, sc):
        assert tuple(int(v) for v in a[1:4]) == w
        got = np.array([float(v) for v in a[5:]], np.float32)
        assert np.allclose(got, row, rtol=1e-5, atol=0)
        # Predict stops at the first failing stage: verdict = every stage score >= theta
        assert int(a[4]) == int(all(r >= t for r, t in zip(row, theta)))
