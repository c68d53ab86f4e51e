"""Per-kernel timing of config C4 (one 3840x2160 frame, step 1, all stages forced) on one GPU (tuning aid)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from surfcascade_b200 import capi, synth

tag = sys.argv[1] if len(sys.argv) > 1 else ""
MODEL = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "model_c1.cfg")
h = capi.Handle(0); h.load_model(MODEL, 40)
W, H = 3840, 2160
fd = torch.from_numpy(synth.frame(H, W, 7)[None]).cuda()
cap = 1 << 22
d_out = torch.zeros(cap * 24, dtype=torch.uint8, device="cuda"); d_cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
prm = capi.params(step=1, prefilter=-1, skip_rule=False, force_all_stages=True)
stream = torch.cuda.ExternalStream(h.stream)
h.detect_device(fd.data_ptr(), 1, W, H, d_out.data_ptr(), cap, d_cnt.data_ptr(), prm); h.sync()
reps = 2
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(reps):
    h.detect_device(fd.data_ptr(), 1, W, H, d_out.data_ptr(), cap, d_cnt.data_ptr(), prm)
e1.record(stream); h.sync()
total = e0.elapsed_time(e1) / reps
h.set_profiling(True); h.kernel_stats(reset=True)
h.detect_device(fd.data_ptr(), 1, W, H, d_out.data_ptr(), cap, d_cnt.data_ptr(), prm); h.sync()
st = h.kernel_stats(reset=True)
c = h.last_counters(1)[0]
print(json.dumps({"tag": tag, "ms": round(total, 2), "raw": int(d_cnt.item()), "weak": c.weak_evals,
                  "kernels_ms": {k: round(v[0], 3) for k, v in st.items() if v[1]}}))
