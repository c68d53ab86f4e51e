import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, '.')
import torch
from surfcascade_b200 import capi, synth
W,H,B=1920,1080,32
h = capi.Handle(0); h.load_model('tests/golden/model_c1.cfg', 40)
fr = np.ascontiguousarray(np.stack([synth.frame(H,W,100+i%8) for i in range(B)]))
ht = torch.from_numpy(fr).pin_memory(); dt = ht.cuda()
d_out = torch.zeros((1<<16)*24, dtype=torch.uint8, device='cuda'); d_cnt = torch.zeros(1, dtype=torch.int32, device='cuda')
ptrs = (ctypes.c_void_p * B)(*[ht.data_ptr() + i*W*H for i in range(B)])
def t(fn, n=5):
    fn(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/n*1e3
print("torch H2D 66MB pinned ms", t(lambda: dt.copy_(ht, non_blocking=True)))
def dev():
    h.detect_device(dt.data_ptr(), B, W, H, d_out.data_ptr(), 1<<16, d_cnt.data_ptr()); h.sync()
print("device path ms", t(dev))
print("host path ms", t(lambda: h.detect_ptrs(ptrs, B, W, H, W, None, 1<<16)))
pg = np.array(fr)  # pageable
ptrs2 = (ctypes.c_void_p * B)(*[pg.ctypes.data + i*W*H for i in range(B)])
print("host path pageable ms", t(lambda: h.detect_ptrs(ptrs2, B, W, H, W, None, 1<<16)))
# host path with cap small
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); h.detect_ptrs(ptrs, B, W, H, W, None, 1<<16); pr.disable()
pstats.Stats(pr).sort_stats('cumtime').print_stats(8)
