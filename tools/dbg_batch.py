import numpy as np, sys
sys.path.insert(0, '.')
from surfcascade_b200 import capi, synth
h = capi.Handle(0); h.load_model('./tests/golden/model_c1.cfg', 40)
frames = [synth.frame(1080,1920,100+i) for i in range(3)]
for n in (1,2,3,8,9,16):
    fs = [frames[i%3] for i in range(n)]
    d, c = h.detect(fs)
    print(n, len(d), [(x.visited, x.prefilter_pass, x.raw, x.reach[1]) for x in c][:4], flush=True)
