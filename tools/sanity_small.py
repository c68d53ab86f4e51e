"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from surfcascade_b200 import capi, synth
MODEL = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "model_c1.cfg")
h = capi.Handle(0); h.load_model(MODEL, 40)
frames = [synth.frame(120, 160, 0), synth.noise_frame(120, 160, 1), synth.frame(120, 160, 2)]
d, c = h.detect(frames)
g, _ = h.detect(frames, capi.params(group_threshold=2))
b, _ = h.detect(frames[:1], capi.params(band_index=1, band_count=2))
n, _ = h.detect(frames[:1], capi.params(skip_rule=False, prefilter=-1))
f, _ = h.detect(frames[:1], capi.params(step=1, force_all_stages=True))
X = h.extract_pool_features(np.stack([synth.positive(s) for s in range(3)]))
M, used = h.mine_negatives(frames[:2], 5, first=True)
rng = np.random.default_rng(0)
Xp = rng.normal(0, 0.2, (64, 16, 32)).astype(np.float32); lab = np.zeros(64, np.uint8); lab[:32] = 1
auc = h.pool_eval(Xp, lab, rng.normal(0, 1, (16, 33)).astype(np.float32), np.ones(16))
print("ok", len(d), len(g), len(b), len(n), len(f), X.shape, M.shape, used, auc[:2])
h.close()
