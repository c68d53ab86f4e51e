"""Timing of sc_mine_negatives (next row N2) on the C1 training shape: 64 negative frames of 320x240, stage-0-only cascade cut so
that samples exist, host frames in, X [n][608][32] out.  Prints windows scanned / s and samples / s."""
import dataclasses, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import modelcfg
from oracle import oracle as O
from surfcascade_b200 import capi, synth
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
c = modelcfg.load(os.path.join(ROOT, "tests", "golden", "model_c1.cfg"))
k = int(c.n_weak[0])
bc = O.BoundCascade(dataclasses.replace(c, theta=np.array([0.40], np.float32), n_weak=c.n_weak[:1].copy(), patch_index=c.patch_index[:k].copy(), w=c.w[:k].copy(), bias=c.bias[:k].copy()))
h = capi.Handle(0)
h.set_cascade(40, bc.theta, bc.n_weak, bc.rects, bc.w, bc.bias)
frames = [synth.negative_frame(100 + i, 240, 320) for i in range(64)]
grid = sum(((320 - l) // 10 + 1) * ((240 - l) // 10 + 1) for l in O.scales(320, 240, O.params(base=40)))
h.mine_negatives(frames[:16], 100000)   # warm-up: plan, buffers
t0 = time.perf_counter()
X, used = h.mine_negatives(frames, 10 ** 6)
dt = time.perf_counter() - t0
print(json.dumps({"frames": len(frames), "frame": "320x240", "lattice_windows_per_frame": grid, "samples": len(X), "frames_used": used, "seconds": round(dt, 4),
                  "lattice_windows_per_s": round(len(frames) * grid / dt), "samples_per_s": round(len(X) / dt), "descriptor_MB_out": round(X.nbytes / 1e6, 1)}))
