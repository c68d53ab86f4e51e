"""L2 / HBM sector-gather ceiling (MEASURED_PEAKS.json has no L2 figure).  Random 32-byte gathers, two
16-byte loads each like one corner fetch of the scan, from tables of growing size (sc_probe_gather).  JSON line out."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from surfcascade_b200 import capi
h = capi.Handle(0)
out = {f"gather32B_{mb}MB_GBps": round(h.probe_gather(mb << 20, 10), 1) for mb in (8, 32, 64, 96, 256, 1024)}
for mode, tag in ((0, "cg"), (1, "nc")):
    for mb in (16, 32, 64, 512):
        out[f"stream16B_{tag}_{mb}MB_GBps"] = round(h.probe_stream(mb << 20, 10, mode), 1)
print(json.dumps(out))
