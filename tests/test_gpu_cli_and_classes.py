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
