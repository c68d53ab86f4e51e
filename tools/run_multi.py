"""BASELINE configs C4 and C5 on N GPUs of one box (one process per GPU; C3 is bench.py --workload c3):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/run_multi.py

C4: ONE 3840x2160 frame, step 1, all stages forced: every rank computes the integral image itself and scans its band of
    every scale's lattice rows (sc_detect_params.band_index / band_count); the bands' detections are gathered on rank 0 with
    sc_gather_detections (C-ABI).
C5: pool evaluation sharded by sample, one all-reduce of the level histograms.
Times are CUDA-event times on the handle's stream, max over ranks.  One JSON line on rank 0."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from surfcascade_b200 import capi, synth

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
json_fd = os.dup(1); os.dup2(2, 1)
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
MODEL = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "model_c1.cfg")
h = capi.Handle(local); h.load_model(MODEL, 40)
stream = torch.cuda.ExternalStream(h.stream, device=dev)


def max_over_ranks(x):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


out = {"n_gpus": world, "C3": "see bench.py --workload c3 (1024-frame batch, host frames in, grouped objects gathered on rank 0 through sc_gather_detections)"}
cap = 1 << 18
d_cnt = torch.zeros(1, dtype=torch.int32, device=dev)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
if world > 1:
    ids = [capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    h.comm_init(rank, world, ids[0])     # the path's own exchange (C-ABI): sc_gather_detections
gbuf = np.zeros((1 << 20) if rank == 0 else 0, capi.DETECTION_DTYPE)
# ---- C4 ------------------------------------------------------------------------------------------------------
img = torch.from_numpy(synth.frame(2160, 3840, 300, n_objects=12)[None]).to(dev)
cap4 = 1 << 20
d_out4 = torch.zeros(cap4 * 24, dtype=torch.uint8, device=dev)
prm = capi.params(step=1, prefilter=-1, skip_rule=False, force_all_stages=True, band_index=rank, band_count=world)
h.detect_device(img.data_ptr(), 1, 3840, 2160, d_out4.data_ptr(), cap4, d_cnt.data_ptr(), prm); h.sync()
barrier()
e0.record(stream)
reps = 3
n_gathered = 0
for _ in range(reps):
    h.detect_device(img.data_ptr(), 1, 3840, 2160, d_out4.data_ptr(), cap4, d_cnt.data_ptr(), prm)
    if world > 1:
        h.sync()
        got, per = h.gather_detections(None, frame_mul=1, frame_add=0, root=0, device_ptr=d_out4.data_ptr(), n_device=int(d_cnt.item()), complete=True, out=gbuf)
        n_gathered = sum(per)
e1.record(stream); h.sync(); barrier()
ms4 = max_over_ranks(e0.elapsed_time(e1)) / reps
c = h.last_counters(1)[0]
tot = torch.tensor([c.grid, int(d_cnt.item())], dtype=torch.int64, device=dev)
if world > 1:
    dist.all_reduce(tot)
out["C4"] = {"ms_per_frame": round(ms4, 2), "windows": int(tot[0].item()), "windows_per_s": round(int(tot[0].item()) / ms4 * 1e3 / 1e9, 3), "unit": "G windows/s",
             "raw_detections": int(tot[1].item()), "gathered_on_rank0": n_gathered, "split": f"{world} row bands per scale, integral replicated"}
# ---- C5 ------------------------------------------------------------------------------------------------------
# candidate scoring of one boosting round: X [100k][608][32] sharded by sample; every rank streams its shard into the
# 608 x 2 x 21 level histograms, ONE all-reduce (NCCL) of 102 KB, AUC of every candidate on every rank
del d_out4, img
torch.cuda.empty_cache()
N5, P5 = 100000, 608
n_loc = len(range(rank, N5, world))
X5 = torch.randn(n_loc, P5, 32, device=dev) * 0.2
lab5 = (torch.arange(n_loc, device=dev) % 2).to(torch.uint8)
hist = torch.zeros(P5 * 2 * 21, dtype=torch.int32, device=dev)
W5 = np.random.default_rng(0).normal(0, 1, size=(P5, 33)).astype(np.float32); b5 = np.ones(P5)
n_pos = torch.tensor([int(lab5.sum().item()), n_loc], dtype=torch.int64, device=dev)
if world > 1:
    dist.all_reduce(n_pos)
for it in range(2):   # first pass warms up
    hist.zero_()
    barrier()
    e0.record(stream)
    h.pool_hist_device(X5.data_ptr(), n_loc, P5, lab5.data_ptr(), W5, b5, None, 0, hist.data_ptr())
    if world > 1:
        with torch.cuda.stream(stream):
            dist.all_reduce(hist)
    e1.record(stream); h.sync()
    auc = h.pool_auc_device(hist.data_ptr(), P5, int(n_pos[0].item()), int((n_pos[1] - n_pos[0]).item()))
    barrier()
ms5 = max_over_ranks(e0.elapsed_time(e1))
out["C5"] = {"ms": round(ms5, 3), "GB_streamed": round(N5 * P5 * 128 / 1e9, 2), "GBps_aggregate": round(N5 * P5 * 128 / ms5 / 1e6, 0), "samples_per_rank": n_loc,
             "auc_mean": round(float(auc.mean()), 4), "exchange": "one all-reduce of 608 x 2 x 21 int32 level counts"}
if rank == 0:
    os.write(json_fd, (json.dumps(out) + "\n").encode())
if world > 1:
    h.comm_destroy()
    dist.destroy_process_group()
