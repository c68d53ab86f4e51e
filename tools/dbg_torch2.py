import numpy as np, sys, ctypes
sys.path.insert(0, '.')
from surfcascade_b200 import capi, synth
from oracle import oracle as O
mode = sys.argv[1]
W,H=1920,1080
frames = [synth.frame(H,W,100+i) for i in range(2)]
h = capi.Handle(0); h.load_model('tests/golden/model_c1.cfg', 40)
def show(tag, c): print(mode, tag, [(x.visited, x.prefilter_pass, x.raw, x.reach[1]) for x in c], flush=True)
if mode == 'twice':
    d,c = h.detect(frames); show('first', c)
    d,c = h.detect(frames); show('second', c)
    S = h.integral(frames[0]); print('integral ok', np.array_equal(S, O.integral(frames[0])))
    d,c = h.detect(frames); show('third', c)
elif mode == 'torch_first':
    import torch
    x = torch.zeros(10, device='cuda:0'); torch.cuda.synchronize()
    d,c = h.detect(frames); show('after torch init', c)
    S = h.integral(frames[0]); print('integral ok', np.array_equal(S, O.integral(frames[0])))
elif mode == 'cudahostalloc':
    import torch
    d,c = h.detect(frames); show('before', c)
    ht = torch.from_numpy(np.stack(frames)).pin_memory()
    d,c = h.detect(frames); show('numpy after pin', c)
    S = h.integral(frames[0]); print('integral ok', np.array_equal(S, O.integral(frames[0])))
    d,c = h.detect([ht.numpy()[0], ht.numpy()[1]]); show('pinned via numpy view', c)
