import numpy as np, sys, ctypes
sys.path.insert(0, '.')
import torch
from surfcascade_b200 import capi, synth
from oracle import oracle as O
mode = sys.argv[1]
if mode != 'noset': torch.cuda.set_device(0)
h = capi.Handle(0); h.load_model('tests/golden/model_c1.cfg', 40)
W,H=1920,1080
frames = [synth.frame(H,W,100+i) for i in range(2)]
def show(tag, c): print(mode, tag, [(x.grid, x.visited, x.prefilter_pass, x.raw, x.reach[1]) for x in c], flush=True)
d, c = h.detect(frames); show('numpy', c)
arr = np.stack(frames)
ht = torch.from_numpy(arr).pin_memory()
d, c = h.detect(frames); show('numpy after pin', c)
d, c = h.detect([arr[0], arr[1]]); show('arr rows', c)
ptrs = (ctypes.c_void_p * 2)(*[arr.ctypes.data + i*W*H for i in range(2)])
d, c = h.detect_ptrs(ptrs, 2, W, H, W); show('arr ptrs', c)
v = ht.numpy()
d, c = h.detect([v[0], v[1]]); show('pinned view', c)
ptrs = (ctypes.c_void_p * 2)(*[ht.data_ptr() + i*W*H for i in range(2)])
print(hex(ht.data_ptr()), hex(v.ctypes.data), v.strides)
d, c = h.detect_ptrs(ptrs, 2, W, H, W); show('pinned ptrs', c)
