"""One rank of the C-ABI exchange test (tests/test_gpu_comm.py): python tools/comm_worker.py RANK WORLD IDFILE OUT.npz [grouped]

Every rank scans its share of a batch of frames (frame f -> rank f mod WORLD) on its own GPU, then all ranks call
sc_gather_detections; rank 0 writes what it received.  No torch.distributed: the NCCL id travels through a file."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from surfcascade_b200 import capi, synth  # noqa: E402

rank, world, idfile, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
grouped = len(sys.argv) > 5 and sys.argv[5] == "grouped"
MODEL = os.path.join(ROOT, "tests", "golden", "model_c1.cfg")
h = capi.Handle(rank % max(1, int(os.environ.get("SC_TEST_GPUS", str(world)))))
h.load_model(MODEL, 40)
if rank == 0:
    with open(idfile + ".tmp", "wb") as f:
        f.write(capi.comm_unique_id())
    os.rename(idfile + ".tmp", idfile)
else:
    t0 = time.time()
    while not os.path.exists(idfile):
        if time.time() - t0 > 120:
            raise SystemExit("no NCCL id")
        time.sleep(0.05)
comm_id = open(idfile, "rb").read()
h.comm_init(rank, world, comm_id)
n_frames = 6
frames = [synth.frame(240, 320, 60 + f) for f in range(n_frames)]
mine = [frames[f] for f in range(n_frames) if f % world == rank]
prm = capi.params(group_threshold=2, group_eps=0.2) if grouped else capi.params()
results = []
for rep in range(2):  # two rounds: the second one reuses the communication buffers
    dets, _ = h.detect(mine, prm) if mine else (np.zeros(0, capi.DETECTION_DTYPE), None)
    got, per = h.gather_detections(dets, frame_mul=world, frame_add=rank, root=0, cap=1 << 16)
    results.append((got, per))
if rank == 0:
    np.savez(out, dets0=results[0][0], dets1=results[1][0], per=np.array(results[1][1]))
h.comm_destroy()
h.close()
print("rank", rank, "done", results[1][1])
