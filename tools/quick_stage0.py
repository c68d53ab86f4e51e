"""Quick per-kernel timing of the C2 scan on one GPU (tuning aid): N frames resident in HBM, event spans inside the library."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from surfcascade_b200 import capi, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
tag = sys.argv[2] if len(sys.argv) > 2 else ""
MODEL = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", sys.argv[3] if len(sys.argv) > 3 else "model_c1.cfg")
h = capi.Handle(0); h.load_model(MODEL, 40)
base = [synth.frame(1080, 1920, 100 + i) for i in range(4)]
fd = torch.from_numpy(np.stack([base[i % 4] for i in range(n)])).cuda()
cap = 1 << 18
d_out = torch.zeros(cap * 24, dtype=torch.uint8, device="cuda"); d_cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
stream = torch.cuda.ExternalStream(h.stream)
for _ in range(2):
    h.detect_device(fd.data_ptr(), n, 1920, 1080, d_out.data_ptr(), cap, d_cnt.data_ptr(), capi.params())
h.sync()
reps = 5
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(reps):
    h.detect_device(fd.data_ptr(), n, 1920, 1080, d_out.data_ptr(), cap, d_cnt.data_ptr(), capi.params())
e1.record(stream); h.sync()
total = e0.elapsed_time(e1) / reps / n
h.set_profiling(True); h.kernel_stats(reset=True)
for _ in range(reps):
    h.detect_device(fd.data_ptr(), n, 1920, 1080, d_out.data_ptr(), cap, d_cnt.data_ptr(), capi.params())
h.sync()
st = h.kernel_stats(reset=True)
c = h.last_counters(n)[0]
print(json.dumps({"tag": tag, "frames": n, "ms_per_frame": round(total, 4), "fps": round(1e3 / total, 1), "raw": int(d_cnt.item()),
                  "visited": c.visited, "weak": c.weak_evals,
                  "kernels_ms_per_frame": {k: round(v[0] / reps / n, 4) for k, v in st.items() if v[1]}}))
