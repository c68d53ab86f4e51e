import numpy as np, sys, ctypes
sys.path.insert(0, '.')
import torch
from surfcascade_b200 import capi, synth
torch.cuda.set_device(0)
h = capi.Handle(0); h.load_model('tests/golden/model_c1.cfg', 40)
W,H=1920,1080
frames = [synth.frame(H,W,100+i) for i in range(2)]
d, c = h.detect(frames)
print("numpy", len(d), [(x.visited, x.prefilter_pass, x.raw, x.reach[1]) for x in c], flush=True)
arr = np.stack(frames)
ht = torch.from_numpy(arr).pin_memory()
print("pinned equal", np.array_equal(ht.numpy(), arr))
ptrs = (ctypes.c_void_p * 2)(*[ht.data_ptr() + i*W*H for i in range(2)])
d, c = h.detect_ptrs(ptrs, 2, W, H, W)
print("pinned", len(d), [(x.visited, x.prefilter_pass, x.raw, x.reach[1]) for x in c], flush=True)
dt = ht.to('cuda:0')
torch.cuda.synchronize()
print("dev equal", np.array_equal(dt.cpu().numpy(), arr))
d_out = torch.zeros(65536*24, dtype=torch.uint8, device='cuda:0'); d_cnt = torch.zeros(1, dtype=torch.int32, device='cuda:0')
torch.cuda.synchronize()
h.detect_device(dt.data_ptr(), 2, W, H, d_out.data_ptr(), 65536, d_cnt.data_ptr())
h.sync()
c = h.last_counters(2)
print("device", int(d_cnt.item()), [(x.visited, x.prefilter_pass, x.raw, x.reach[1]) for x in c], flush=True)
