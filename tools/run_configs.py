"""Times BASELINE configs C1, C3 (per-GPU share), C4 and C5 on one B200 and prints a markdown table (kept in profiles/)."""
import ctypes, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from surfcascade_b200 import capi, synth

MODEL = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "model_c1.cfg")
h = capi.Handle(0); h.load_model(MODEL, 40)
stream = torch.cuda.ExternalStream(h.stream)
rows = []

def timed_device(frames_dev, n, W, H, prm, reps=5, cap=1 << 22):
    d_out = torch.zeros(cap * 24, dtype=torch.uint8, device="cuda"); d_cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    for _ in range(2):
        h.detect_device(frames_dev.data_ptr(), n, W, H, d_out.data_ptr(), cap, d_cnt.data_ptr(), prm)
    h.sync()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        h.detect_device(frames_dev.data_ptr(), n, W, H, d_out.data_ptr(), cap, d_cnt.data_ptr(), prm)
    e1.record(stream); h.sync()
    return e0.elapsed_time(e1) / reps, h.last_counters(n), int(d_cnt.item())

# C1: one 640x480 frame (latency) and a batch of 64
f = np.stack([synth.frame(480, 640, 1 + i) for i in range(64)])
fd = torch.from_numpy(f).cuda()
ms, c, nd = timed_device(fd, 1, 640, 480, capi.params(), reps=20)
rows.append(("C1 single 640x480 frame (latency)", f"{ms:.3f} ms/frame", f"{1e3/ms:.0f} frames/s", f"grid {c[0].grid}, visited {c[0].visited}, evaluated {c[0].evaluated}, raw {nd}"))
ms, c, nd = timed_device(fd, 64, 640, 480, capi.params(), reps=5)
rows.append(("C1 batch of 64 640x480 frames", f"{ms/64:.4f} ms/frame", f"{64e3/ms:.0f} frames/s", f"{64e3/ms*c[0].grid/1e9:.2f} G grid windows/s"))
# C3 share: 128 1080p frames per GPU (1024 / 8)
f = np.stack([synth.frame(1080, 1920, 100 + i % 8) for i in range(128)])
fd = torch.from_numpy(f).cuda()
ms, c, nd = timed_device(fd, 128, 1920, 1080, capi.params(), reps=3)
rows.append(("C3 per-GPU share: 128 x 1080p frames", f"{ms:.1f} ms/batch", f"{128e3/ms:.0f} frames/s", f"raw detections {nd} ({nd*24/1e3:.0f} KB to gather), {128e3/ms*c[0].grid/1e9:.2f} G grid windows/s"))
# C4: 4K, step 1, forced stages
f = synth.frame(2160, 3840, 300, n_objects=12)[None]
fd = torch.from_numpy(f).cuda()
ms, c, nd = timed_device(fd, 1, 3840, 2160, capi.params(step=1, prefilter=-1, skip_rule=False, force_all_stages=True), reps=3)
we = c[0].grid * 22
rows.append(("C4 3840x2160 step 1, all 4 stages (22 weak) forced", f"{ms:.1f} ms/frame", f"{c[0].grid/ms/1e6:.2f} G windows/s", f"{we/ms/1e6:.1f} G weak evaluations/s"))
ms, c, nd = timed_device(fd, 1, 3840, 2160, capi.params(step=1), reps=3)
rows.append(("4K step 1, reference semantics (prefilter, early reject, stride)", f"{ms:.1f} ms/frame", f"{c[0].grid/ms/1e6:.2f} G grid windows/s", f"visited {c[0].visited}, evaluated {c[0].evaluated}"))
del fd
# SURVEY.md 8(d)'s paper-shaped cascade (tests/golden/model_paper8.cfg: 8 stages, 2/3/5/8/12/16/24/32 weak classifiers) on C2 and C4
PAPER = os.path.join(os.path.dirname(MODEL), "model_paper8.cfg")
if os.path.exists(PAPER):
    h.load_model(PAPER, 40)
    f = np.stack([synth.frame(1080, 1920, 100 + i % 8) for i in range(32)])
    fd = torch.from_numpy(f).cuda()
    ms, c, nd = timed_device(fd, 32, 1920, 1080, capi.params(), reps=3)
    rows.append(("C2 with the paper-shaped 8-stage cascade (102 weak; stage 0: 2 weak, passes ~10 %)", f"{ms/32:.3f} ms/frame", f"{32e3/ms:.0f} frames/s",
                 f"visited {c[0].visited}, weak evaluations (reference count) {c[0].weak_evals} = {c[0].weak_evals*32/ms/1e6:.1f} G/s, reach {[c[0].reach[i] for i in range(8)]}, raw {nd//32}/frame"))
    del fd
    f = synth.frame(2160, 3840, 300, n_objects=12)[None]
    fd = torch.from_numpy(f).cuda()
    ms, c, nd = timed_device(fd, 1, 3840, 2160, capi.params(step=1, prefilter=-1, skip_rule=False, force_all_stages=True), reps=2)
    we = c[0].grid * 102
    rows.append(("C4 with the paper-shaped cascade: 3840x2160 step 1, all 8 stages (102 weak) forced", f"{ms:.1f} ms/frame", f"{c[0].grid/ms/1e6:.2f} G windows/s", f"{we/ms/1e6:.1f} G exact weak evaluations/s"))
    del fd
    h.load_model(MODEL, 40)
# C5: 100k samples x 608 candidates x 32 floats resident in HBM (7.78 GB)
N, P = 100000, 608
X = torch.randn(N, P, 32, device="cuda") * 0.2
lab = torch.zeros(N, dtype=torch.uint8, device="cuda"); lab[:N // 2] = 1
hist = torch.zeros(P * 2 * 21, dtype=torch.int32, device="cuda")
W = np.random.default_rng(0).normal(0, 1, size=(P, 33)).astype(np.float32); b = np.ones(P)
torch.cuda.synchronize()
h.pool_hist_device(X.data_ptr(), N, P, lab.data_ptr(), W, b, None, 0, hist.data_ptr()); h.sync()
h.set_profiling(True); h.kernel_stats(reset=True)
for _ in range(3):
    h.pool_hist_device(X.data_ptr(), N, P, lab.data_ptr(), W, b, None, 0, hist.data_ptr())
h.sync()
ms = h.kernel_stats()["k_pool_hist"][0] / 3
gb = N * P * 128 / 1e9
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
rows.append(("C5 pool evaluation 100k x 608 x 32 f32 (7.78 GB in HBM)", f"{ms:.2f} ms", f"{gb/ms*1e3:.0f} GB/s", f"{gb/ms*1e3/peak:.2f} of HBM copy peak {peak:.0f} GB/s"))
# N3: descriptors of all 608 pool patches for 100k 40x40 samples, written straight into the C5 matrix on the device
h.set_profiling(False)
imgs = torch.randint(0, 256, (N, 40, 40), dtype=torch.uint8, device="cuda")
h.extract_pool_features_device(imgs.data_ptr(), N, 40, X.data_ptr()); h.sync()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(stream)
h.extract_pool_features_device(imgs.data_ptr(), N, 40, X.data_ptr())
e1.record(stream); h.sync()
ms = e0.elapsed_time(e1)
rows.append(("N3 pool-feature extraction: 100k 40x40 samples -> X 100k x 608 x 32 f32 (7.78 GB written)", f"{ms:.2f} ms", f"{N/ms*1e3/1e6:.2f} M samples/s",
             f"{gb/ms*1e3:.0f} GB/s of X written = {gb/ms*1e3/peak:.2f} of HBM copy peak"))
print("| config | time | rate | notes |\n|---|---|---|---|")
for r in rows:
    print("| " + " | ".join(r) + " |")
