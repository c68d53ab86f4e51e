"""GPU: the checked build (-DSC_CHECKED, libsurfcascade_b200_checked.so) runs the detect path on awkward shapes -- ragged tails,
frames barely larger than the window, odd lattice steps, forced stages, row bands, the paper-shaped cascade, the parity hooks --
with every gather from an integral-image plane range-tested on the device: no violation may be counted, and the results must
equal the normal build's.  compute-sanitizer is closed on the B200 pool; this is the memcheck of the gather addresses."""
import json
import os
import subprocess
import sys

import pytest

from surfcascade_b200 import build

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_CHILD = r"""
import sys, json, ctypes, numpy as np
sys.path.insert(0, '.')
from surfcascade_b200 import capi, synth
L = capi.lib()
L.sc_checked_violations.argtypes = [ctypes.c_int]
h = capi.Handle(0)   # ONE handle: the accepted address ranges are process-wide
out = []
for model, n_st in (('tests/golden/model_c1.cfg', 4), ('tests/golden/model_paper8.cfg', 8)):
    h.load_model(model, 40)
    for (hh, ww, seed) in ((41, 41, 1), (97, 203, 2), (240, 321, 3), (480, 640, 4), (133, 900, 5)):
        img = synth.frame(hh, ww, seed) if min(hh, ww) >= 64 else synth.noise_frame(hh, ww, seed)
        for prm in (capi.params(), capi.params(step=1), capi.params(base=70, step=3), capi.params(skip_rule=False, prefilter=-1),
                    capi.params(band_index=1, band_count=3)):
            d, c = h.detect([img, img[::-1].copy()], prm)
            out.append([len(d), int(c[0].visited), int(c[0].weak_evals)])
    img = synth.frame(150, 220, 9)
    d, c = h.detect([img], capi.params(step=1, prefilter=-1, skip_rule=False, force_all_stages=True))
    out.append([len(d), int(c[0].visited), int(c[0].weak_evals)])
    S = h.integral(img)
    f = h.features(np.array([[0, 0, 40, 40], [180, 110, 40, 40], [3, 7, 20, 80], [100, 20, 80, 20]], np.int32))
    sc = h.stage_scores(np.array([[0, 0, 40], [70, 0, 150], [180, 110, 40]], np.int32), n_st)
    out.append([float(S.sum()), float(f.sum()), float(sc.sum())])
print(json.dumps({"version": L.sc_version().decode(), "violations": int(L.sc_checked_violations(0)), "out": out}))
"""


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, "-c", _CHILD], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().split("\n")[-1])


def test_checked_build_counts_no_out_of_range_gather():
    if not os.path.exists(build.LIB_CHECKED):
        pytest.skip("checked library not built (python surfcascade_b200/build.py --checked)")
    chk = _run({"SC_LIB": build.LIB_CHECKED})
    ref = _run({"SC_LIB": ""})
    assert "checked build" in chk["version"] and "checked" not in ref["version"]
    assert chk["violations"] == 0
    assert ref["violations"] == -1
    assert chk["out"] == ref["out"] and len(chk["out"]) > 50
