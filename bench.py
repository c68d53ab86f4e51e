#!/usr/bin/env python
"""bench.py -- 1080p full-scale-range SURF-cascade detection throughput (BASELINE.json metric, config C2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

A step is one pass of the whole hot path (channels + integral + cascade scan + stride replay + detection
output) over one batch of B synthetic 1920x1080 frames per GPU.  `value` is frames/s with the frames already
resident in HBM (sc_detect_device); `e2e` is the same metric through the host-buffer C-ABI call sc_detect
(pinned host frames -> H2D -> path -> D2H detections) -- the number to compare with the reference arm.
Under torchrun (N > 1) every rank runs the same work on its own GPU (frames sharded by rank, weak scaling) and
the detections of every step are gathered over NCCL; times are CUDA-event/device based, max over ranks.

`--impl reference` times the reference's own CPU code (oracle/_ref: the unmodified reference sources compiled
by oracle/Makefile; falls back to the plain-C port oracle/libsurf_oracle.so) on all host cores.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 1920, 1080
GOLDEN = os.path.join(ROOT, "tests", "golden")
MODEL = os.path.join(GOLDEN, "model_c1.cfg")
N_UNIQUE = 8  # distinct synthetic frames; batches cycle through them with different offsets
# --workload: c2 is BASELINE.json's configs[1], the one `metric` is quoted on (default; the driver's run).  c2_paper8 is the same
# scan with SURVEY.md 8(d)'s paper-shaped cascade.  c3 is configs[2]: ONE batch of 1,024 frames per step, sharded by frame over
# the ranks (strong scaling), host frames in, grouped objects on rank 0 out.
WORKLOADS = {
    "c2": {"model": MODEL, "cascade": "tests/golden/model_c1.cfg (4 stages, 3/6/7/6 weak classifiers)",
           "workload": "C2: 1920x1080 synthetic frames, base 40, step 2, scale 1.1 (35 scales, 11,557,983 grid windows/frame), reference-trained cascade model_c1.cfg"},
    "c2_paper8": {"model": os.path.join(GOLDEN, "model_paper8.cfg"), "cascade": "tests/golden/model_paper8.cfg (8 stages, 2/3/5/8/12/16/24/32 weak classifiers, seed 7)",
                  "workload": "C2 with the paper-shaped cascade: 1920x1080 synthetic frames, base 40, step 2, scale 1.1 (35 scales, 11,557,983 grid windows/frame), model_paper8.cfg"},
    "c3": {"model": MODEL, "cascade": "tests/golden/model_c1.cfg (4 stages, 3/6/7/6 weak classifiers)",
           "workload": "C3: one batch of 1024 1920x1080 synthetic frames per step sharded by frame over the GPUs, base 40, step 2, scale 1.1, host frames in, grouped objects gathered on rank 0"},
}
C3_FRAMES = 1024


def config_of(name: str) -> dict:
    w = WORKLOADS[name]
    return {"workload": w["workload"], "frame": "1920x1080", "base": 40, "step": 2, "scale": 1.1, "cascade": w["cascade"]}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark(self):
        """Start of the region whose samples count (the sampler itself is started earlier: nvidia-smi takes a while to come up)."""
        self.t_mark = time.time()

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t_mark = getattr(self, "t_mark", 0.0)
        inside = [ln for (t, ln) in self.lines if t >= t_mark]
        if not inside:  # a region shorter than the sampling period: the samples closest to it
            inside = [ln for (_, ln) in self.lines[-2:]]
        for ln in inside:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx = float(parts[1])
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.lower().startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def make_frames(n: int):
    from surfcascade_b200 import synth
    return [synth.frame(H, W, 100 + i) for i in range(n)]


# ------------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU detect path on the host cores
# ------------------------------------------------------------------------------------------------------------
def cpu_reference_timing(frames, threads: int, warmup: int = 1, model: str = MODEL):
    """Per-frame (IntegralImage ms, scan ms) of the reference CPU path over `frames`, after `warmup` untimed frames.
    oracle/_ref: the reference's own classes and lifted detect loop WITHOUT counter hooks, model loaded once, the two
    phases timed separately (ref_detect_timed); else the plain-C port, timed the same way."""
    from oracle import refbind
    if refbind.available():
        ms_i, ms_s, raw = refbind.detect_timed(frames, model, base=40, nthreads=threads, warmup=warmup)
        return ms_i, ms_s, "reference", int(raw.sum())
    from oracle import modelcfg, oracle
    bc = oracle.BoundCascade(modelcfg.load(model))
    prm = oracle.params(base=40, nthreads=threads)
    ms_i, ms_s, raw = [], [], 0
    for k in range(-warmup, len(frames)):
        f = frames[k % len(frames)]
        t0 = time.perf_counter()
        S = oracle.integral(f)
        t1 = time.perf_counter()
        d = oracle.detect(S, bc, prm)
        t2 = time.perf_counter()
        if k >= 0:
            ms_i.append(1e3 * (t1 - t0)); ms_s.append(1e3 * (t2 - t1)); raw += len(d.x)
    return np.array(ms_i), np.array(ms_s), "port", raw


def cpu_baseline_block(n_all: int, n_single: int, threads: int, model: str = MODEL) -> dict:
    """`cpu_baseline` of a bench line: n_all frames on all host threads and n_single frames on one thread."""
    fr = make_frames(max(n_all, n_single, 2))
    ms_i, ms_s, kind, raw = cpu_reference_timing([fr[i % len(fr)] for i in range(n_all)], threads, warmup=1, model=model)
    tot = ms_i + ms_s
    out = {"value": 1e3 * len(tot) / float(tot.sum()), "unit": "frames/s", "cores": threads, "kind": kind, "cpu_model": cpu_model(),
           "median_ms_per_frame": float(np.median(tot)), "median_ms_integral": float(np.median(ms_i)), "median_ms_scan": float(np.median(ms_s)),
           "raw_detections": raw}
    single = None
    if n_single > 0 and threads > 1:
        si, ss, _, _ = cpu_reference_timing([fr[i % len(fr)] for i in range(n_single)], 1, warmup=0, model=model)
        single = float(np.median(si + ss))
        out["single_thread"] = {"frames_per_s": 1e3 / single, "median_ms_integral": float(np.median(si)), "median_ms_scan": float(np.median(ss)),
                                "parallel_efficiency": (single / float(np.median(tot))) / threads}
    out["sample"] = (f"{n_all} frames of the workload after 1 warm-up frame on all {threads} host threads (OpenMP over scales as in "
                     f"ObjDetector.cpp:177, OMP_PROC_BIND=spread), value = frames / summed per-frame time; hook-free timing build of the "
                     f"reference (no counters in the window loop), model loaded once, IntegralImage and scan timed separately"
                     + (f"; {n_single} frame(s) on 1 thread" if single else ""))
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ.setdefault("OMP_PROC_BIND", "spread")
    threads = os.cpu_count() or 1
    model = WORKLOADS[args.workload]["model"]
    fr = make_frames(min(max(args.steps, 1), N_UNIQUE))
    frames = [fr[i % len(fr)] for i in range(args.steps)]
    ms_i, ms_s, kind, raw = cpu_reference_timing(frames, threads, warmup=args.warmup, model=model)
    tot = ms_i + ms_s
    t_total = float(tot.sum()) / 1e3
    fps = args.steps / t_total
    single = cpu_reference_timing(frames[:1], 1, warmup=0, model=model) if threads > 1 else None
    line = {"impl": "reference", "metric": "1080p full-scale-range detection throughput", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_of(args.workload),
            "run": {"frames_per_step": 1, "threads": threads},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind, "cpu_model": cpu_model(),
                             "median_ms_per_frame": float(np.median(tot)), "median_ms_integral": float(np.median(ms_i)),
                             "median_ms_scan": float(np.median(ms_s)), "raw_detections": raw,
                             "single_thread_frames_per_s": (1e3 / float((single[0] + single[1])[0])) if single else None,
                             "sample": f"{args.steps} step(s) of 1 frame of the workload after {args.warmup} warm-up frame(s), all {threads} host threads "
                                       "(OpenMP over scales as in ObjDetector.cpp:177); hook-free timing build of the reference, model loaded once, "
                                       "IntegralImage and scan timed separately; value = frames / summed per-frame time"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
def run_ours(args):
    os.environ.setdefault("OMP_PROC_BIND", "spread")  # for the cpu_baseline leg (read when the OpenMP runtime starts)
    import torch
    import torch.distributed as dist
    from surfcascade_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # libraries (NCCL's version banner) write to fd 1: keep stdout for the ONE JSON line
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this implementation has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    wl = WORKLOADS[args.workload]
    c3 = args.workload == "c3"
    B = args.batch  # frames per library call
    n_sub = (C3_FRAMES // world) // B if c3 else 1  # C3: this rank's share of the 1024-frame batch goes through n_sub pipelined calls
    if c3 and n_sub * B * world != C3_FRAMES:
        raise SystemExit("bench.py: --batch x ranks must divide 1024 for --workload c3")
    no_exchange = os.environ.get("SC_BENCH_NO_EXCHANGE") == "1"  # scaling diagnosis only: the line says so
    h = capi.Handle(local)
    h.load_model(wl["model"], 40)
    hx = None  # the exchange has its own handle (a handle is thread-compatible, not thread-safe) and its own host thread
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)  # plumbing: barriers, the max over ranks, the NCCL id
        ids = [capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        hx = capi.Handle(local)
        hx.comm_init(rank, world, ids[0])  # the path's own exchange: sc_gather_detections (C-ABI)
    prm = capi.params()
    prm_g = capi.params(group_threshold=2, group_eps=0.2)
    stream = torch.cuda.ExternalStream(h.stream, device=dev)
    side = torch.cuda.Stream(device=dev)

    # inputs: N_UNIQUE distinct frames; rank r / step s read a rotated batch so consecutive steps differ
    base_frames = make_frames(N_UNIQUE)
    n_sets = 5  # 5 x B x 2 MB of input (> 126 MB L2 for B >= 16); intermediates are 100 MB per frame
    host_sets, dev_sets = [], []
    for s in range(n_sets):
        arr = np.ascontiguousarray(np.stack([base_frames[(rank + s + i) % N_UNIQUE] for i in range(B)]))
        ht = torch.from_numpy(arr).pin_memory()
        host_sets.append(ht)
        if not c3:
            dev_sets.append(ht.to(dev))
    cap = 1 << 16 if not c3 else 1 << 18
    NB = 4  # output buffers in flight: the exchange thread may lag the enqueueing thread by up to NB - 1 steps
    d_out = [torch.zeros(cap * 24, dtype=torch.uint8, device=dev) for _ in range(NB)]
    d_cnt = [torch.zeros(1, dtype=torch.int32, device=dev) for _ in range(NB)]
    cnt_host = torch.zeros(NB, dtype=torch.int32).pin_memory()
    evs = [torch.cuda.Event() for _ in range(NB)]
    gbuf = np.zeros(cap * world if rank == 0 else 0, capi.DETECTION_DTYPE)
    xstat = {"calls": 0, "records": 0, "local": 0}

    # The exchange runs on its own host thread (and its own handle / stream): sc_gather_detections is host-synchronous -- root
    # waits for every rank's header -- and with it on the enqueueing thread every rank's host was tied to the slowest rank at
    # every step with one step of lookahead; on 8 GPUs and 16 shared host cores that halved the throughput (14.6 k frames/s).
    import queue
    import threading
    jobs = queue.Queue()
    free = [threading.Semaphore(1) for _ in range(NB)]
    xerr = []

    blk_ev = torch.cuda.Event(blocking=True)

    def exchange_worker():
        torch.cuda.set_device(local)
        while True:
            job = jobs.get()
            if job is None:
                return
            try:
                if job[0] == "dev":
                    k = job[1]
                    with torch.cuda.stream(side):
                        side.wait_event(evs[k])
                        cnt_host[k:k + 1].copy_(d_cnt[k], non_blocking=True)
                        blk_ev.record(side)
                    blk_ev.synchronize()   # blocking-sync event: this thread sleeps, it does not spin on a core the enqueueing threads need
                    n = min(int(cnt_host[k]), cap)
                    got, per = hx.gather_detections(None, frame_mul=world, frame_add=rank, root=0, device_ptr=d_out[k].data_ptr(), n_device=n, complete=True, out=gbuf)
                    xstat["calls"] += 1; xstat["records"] += sum(per); xstat["local"] += n
                    free[k].release()
                else:
                    got, per = hx.gather_detections(job[1], frame_mul=world, frame_add=rank, root=0, out=gbuf)
                    xstat["host_records"] = xstat.get("host_records", 0) + sum(per)
            except Exception as e:  # noqa: BLE001
                xerr.append(e)
                if job[0] == "dev":
                    free[job[1]].release()
            finally:
                jobs.task_done()

    worker = None
    if world > 1 and not no_exchange:
        worker = threading.Thread(target=exchange_worker, daemon=True)
        worker.start()

    def step_device(s):
        k = s % NB
        x = dev_sets[s % n_sets]
        if worker:
            free[k].acquire()   # buffer k's previous records have been exchanged
        h.detect_device(x.data_ptr(), B, W, H, d_out[k].data_ptr(), cap, d_cnt[k].data_ptr(), prm)
        if worker:
            # the one exchange step of the path: this step's records travel while the next steps compute
            evs[k].record(stream)
            jobs.put(("dev", k))

    def flush_exchange():
        if worker:
            jobs.join()
            if xerr:
                raise xerr[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps):
        """Device time of `steps` steps (CUDA events on the handle's stream; the last exchange has completed when the closing
        event is recorded), max over ranks."""
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for s in range(steps):
            fn(s)
        flush_exchange()
        e1.record(stream)
        h.sync()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    def ptrs_of(s):
        x = host_sets[s % n_sets]
        return (ctypes.c_void_p * B)(*[x.data_ptr() + i * W * H for i in range(B)])

    n_calls = args.steps * n_sub
    keep = [ptrs_of(s) for s in range(n_calls)]

    def pipelined_pass(prm_x):
        """`steps` steps through sc_detect_submit / sc_detect_collect, two batches in flight, every step's result gathered on rank 0
        (N > 1; C3: a step is n_sub calls and ONE gather of the whole batch's objects); returns (wall seconds bracketed by syncs,
        results on this rank, records rank 0 received)."""
        barrier()
        t0 = time.perf_counter()
        n_out = n_got = 0
        acc = []
        pend = h.detect_submit(keep[0], B, W, H, W, prm_x, cap)
        for s in range(n_calls):
            nxt = h.detect_submit(keep[s + 1], B, W, H, W, prm_x, cap) if s + 1 < n_calls else None
            out, _ = h.detect_collect(pend, B, cap)
            n_out += len(out)
            if c3:
                out = out.copy()
                out["frame"] += (s % n_sub) * B   # frame index inside this rank's share of the batch
                acc.append(out)
            if worker and (s + 1) % n_sub == 0:
                jobs.put(("host", np.concatenate(acc) if c3 else out.copy()))
                acc = []
            pend = nxt
        flush_exchange()
        torch.cuda.synchronize()
        n_got = xstat.pop("host_records", 0)
        return time.perf_counter() - t0, n_out, n_got

    clocks = ClockSampler(local)
    cnts = None
    stats = {}
    launches = 0
    ms = ms_prof = None
    if rank == 0:
        clocks.start()   # before the warm-up steps; only samples taken after mark() are reported
    if not c3:
        for s in range(args.warmup):
            step_device(s)
        flush_exchange()
        h.sync()
        cnts = h.last_counters(B)
        if rank == 0:
            clocks.mark()
        launches0 = h.launch_count
        xstat.update(calls=0, records=0, local=0)
        ms = timed(step_device, args.steps)
        launches = h.launch_count - launches0
        xrec = dict(xstat)
        clk = clocks.stop() if rank == 0 else {}
        fps = world * B * args.steps / (ms / 1e3)
        # per-kernel split of one more timed pass (same steps) with event spans inside the library
        h.set_profiling(True)
        h.kernel_stats(reset=True)
        ms_prof = timed(step_device, args.steps)
        stats = h.kernel_stats(reset=True)
        h.set_profiling(False)

    # e2e: host buffers through the C-ABI (H2D and D2H of every step inside the timed region), wall clock bracketed by syncs.
    # The steps go through sc_detect_submit / sc_detect_collect with two batches in flight, as a streaming caller would:
    # every step still uploads its own frames from pinned host memory and downloads its own detections and counters.
    if c3:
        for _ in range(max(args.warmup, 1)):
            pipelined_pass(prm_g)            # untimed: allocates the tickets' buffers, warms the clocks
        if rank == 0:
            clocks.mark()
        launches0 = h.launch_count
        h.transfer_bytes(reset=True)
        e2e_grouped_s, n_obj, n_got_g = pipelined_pass(prm_g)
        launches = h.launch_count - launches0
        h2d_b, d2h_b = h.transfer_bytes()
        e2e_grouped_s = max_over_ranks(e2e_grouped_s)
        clk = clocks.stop() if rank == 0 else {}
        e2e_s = e2e_grouped_s
        fps = world * B * n_calls / e2e_s
        ms = 1e3 * e2e_s
        e2e_sync_s = None
        nd = n_got = 0
        group_ms = 0.0
        cnts = h.last_counters(min(B, 32))
    else:
        for s in range(min(args.warmup, 2)):
            ptrs = ptrs_of(s)
            h.detect_ptrs(ptrs, B, W, H, W, prm, cap)
        pipelined_pass(prm)                  # untimed: the first use of the two tickets allocates their device / pinned buffers
        h.transfer_bytes(reset=True)
        e2e_s, nd, n_got = pipelined_pass(prm)
        h2d_b, d2h_b = h.transfer_bytes()
        # the same pipeline with groupRectangles(2, 0.2) of every frame done on the device (grouped objects out instead of raw windows)
        pipelined_pass(prm_g)                # allocates the grouping buffers of both tickets
        e2e_grouped_s, n_obj, n_got_g = pipelined_pass(prm_g)
        h.set_profiling(True); h.kernel_stats(reset=True)
        pipelined_pass(prm_g)                # per-kernel event spans (one scan lane): only the grouping kernels' time is read
        group_ms = h.kernel_stats(reset=True).get("k_group_frames", (0.0, 0))[0]
        h.set_profiling(False)
        # the plain synchronous call, one batch at a time, for comparison
        barrier()
        t0 = time.perf_counter()
        for s in range(args.steps):
            h.detect_ptrs(keep[s], B, W, H, W, prm, cap)
        torch.cuda.synchronize()
        e2e_sync_s = max_over_ranks(time.perf_counter() - t0)
        e2e_s = max_over_ranks(e2e_s)
        e2e_grouped_s = max_over_ranks(e2e_grouped_s)
    e2e_fps = world * B * n_calls / e2e_s

    if rank == 0:
        peak, peak_src = load_peaks()
        frames_timed = B * n_calls
        c = cnts[0]
        grid = c.grid
        mean = lambda f: float(np.mean([f(x) for x in cnts]))
        weak_ref = mean(lambda x: x.weak_evals)
        vis_ref = mean(lambda x: x.visited)
        line = {
            "metric": "1080p full-scale-range detection throughput", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if c3 else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": config_of(args.workload),
            "run": {"frames_per_step_per_gpu": B * n_sub, "frames_per_library_call": B, "parallelism": f"frames sharded over {world} GPU(s)",
                    "l2": f"inputs rotate over {n_sets} batches ({n_sets * B * W * H / 1e6:.0f} MB) and each step writes {B * 99.7:.0f} MB of integral images: larger than the 126 MB L2"},
            "windows_per_s": {"grid": fps * grid, "reference_visited": fps * vis_ref},
            "work_per_frame": {"grid_windows": grid, "visited": vis_ref, "prefilter_pass": mean(lambda x: x.prefilter_pass),
                               "weak_evals_reference": weak_ref, "raw_detections": mean(lambda x: x.raw)},
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": h2d_b // args.steps, "d2h_bytes_per_step": d2h_b // args.steps,
                    "bytes_counted_by": "the library at its cudaMemcpy calls (sc_transfer_bytes), this rank",
                    "api": "sc_detect_submit / sc_detect_collect, two batches in flight (pinned host frames in, sorted detections + counters out)"
                           + (", then sc_gather_detections of every step's records to rank 0 over NCCL" if world > 1 else "")},
            "gpu_launches": int(launches),
            "clocks": clk,
        }
        if no_exchange:
            line["note_no_exchange"] = "SC_BENCH_NO_EXCHANGE=1: the per-step gather was skipped (scaling diagnosis, not a valid multi-GPU line)"
        if world > 1:
            line["exchange"] = {"api": "sc_gather_detections (C-ABI; NCCL bootstrap, then per step copy-engine writes over NVLink into rank 0's IPC-mapped HBM region: exactly n records + {sequence, count} header per rank, library's own stream)",
                                "e2e_records_to_rank0_per_step": (n_got_g if c3 else n_got) / args.steps,
                                "grouped_objects_to_rank0_per_step": n_got_g / args.steps}
        if c3:
            line["e2e"]["objects_per_frame"] = n_obj / max(B * n_calls, 1)
            line["note"] = ("C3 is end to end by definition (host frames in, grouped objects on rank 0 out): `value` is the e2e figure; groupRectangles(2, 0.2) "
                            "runs on the device of the rank that owns the frame, only grouped objects cross NVLink")
        else:
            line["e2e"]["synchronous_sc_detect"] = world * B * args.steps / e2e_sync_s
            line["e2e"]["with_device_grouping"] = {"value": world * B * args.steps / e2e_grouped_s, "unit": "frames/s", "objects_per_frame": n_obj / max(B * args.steps, 1),
                                                   "k_group_frames_ms_per_frame": group_ms / max(B * args.steps, 1),
                                                   "note": "groupRectangles(raw, 2, 0.2) per frame on the device (next row N1); same pipelined calls, grouped objects out"}
            if world > 1:
                line["exchange"]["device_resident_records_to_rank0_per_step"] = xrec["records"] / max(xrec["calls"], 1)
                line["exchange"]["device_resident_bytes_per_rank_per_step"] = 24 * xrec["local"] / max(xrec["calls"], 1)
                line["exchange"]["host_thread"] = "sc_gather_detections runs on its own host thread and handle; up to 3 steps of lookahead for the enqueueing thread"
            # algorithmic gather bytes of the scan per frame: 32 B x (4 per prefilter + 9|10 corners per weak eval)
            ev_ms, ev_n = stats.get("k_scan_stage0", (0.0, 0))       # even lattice columns, one launch per 8-frame scan group
            odd_ms, _ = stats.get("k_scan_stage0_odd", (0.0, 0))      # reachable odd columns
            st0_ms = ev_ms + odd_ms
            walk_ms, walk_n = stats.get("k_integral_walk", (0.0, 0))
            carry_ms, _ = stats.get("k_strip_carry", (0.0, 0))
            total_k_ms = sum(v[0] for v in stats.values()) or 1.0
            # 32 B x (4 x prefilters + C_shape x weak evaluations) with the REFERENCE's counts (visited windows);
            # C_shape = 10 (1x4 / 4x1 patches; 9 for 2x2 -- model_c1.cfg's stage 0 is all 1x4 / 4x1, later stages are < 1 % of the evaluations)
            alg_scan_bytes = 32.0 * (4.0 * vis_ref + 10.0 * weak_ref)
            scan_gbs = alg_scan_bytes * frames_timed / (st0_ms / 1e3) / 1e9 if st0_ms else None
            alg_int_bytes = W * H + 32.0 * (W + 1) * (H + 1)
            int_gbs = alg_int_bytes * frames_timed / ((walk_ms + carry_ms) / 1e3) / 1e9 if walk_ms else None
            traffic_scan = traffic_int = l2_bytes_ev = ncu_pct = None
            traffic_file = None
            group = 8  # frames per scan-group launch, the unit the ncu capture in profiles/ was taken on
            for name in ("r2_traffic.json", "r1_traffic.json"):  # per-launch DRAM and L2->L1 bytes from the committed ncu --set full capture
                try:
                    tj = json.load(open(os.path.join(ROOT, "profiles", name)))
                    traffic_scan = tj["dram_bytes_per_launch"]["k_scan_stage0_even"]
                    traffic_int = tj["dram_bytes_per_launch"]["k_integral_walk"]
                    l2_bytes_ev = tj["l2_to_l1_bytes_per_launch"]["k_scan_stage0_even"]
                    ncu_pct = tj["ncu_pct_of_peak"]["k_scan_stage0_even"]
                    traffic_file = "profiles/" + name
                    break
                except Exception:
                    continue
            try:  # L2 -> SM reference point measured live: coalesced 16-byte loads streaming over an L2-resident 32 MB table
                l2_peak = max(h.probe_stream(32 << 20, 5, m) for m in (1, 3))
            except Exception:
                l2_peak = None
            sm_mhz = (clk.get("sm_mhz") or clk.get("sm_max_mhz") or 1965.0) if isinstance(clk, dict) else 1965.0
            l1_peak = 18944.0 * sm_mhz * 1e6 / 1e9  # GB/s: 128 B/clk/SM x 148 SMs
            ev_launch_ms = ev_ms / ev_n if ev_n else None
            l2_meas_gbs = (l2_bytes_ev / (ev_launch_ms / 1e3) / 1e9) if (l2_bytes_ev and ev_launch_ms and B % group == 0) else None
            line["roofline"] = {"kernel": "k_scan_stage0 (even columns) + k_scan_odd (reachable odd columns)", "bound": "l1",
                                "achieved": scan_gbs, "peak": l1_peak, "unit": "GB/s", "frac": (scan_gbs / l1_peak) if (scan_gbs and l1_peak) else None,
                                "peak_source": "DERIVED, not from MEASURED_PEAKS.json (which has HBM and bf16 figures only, no L1 / L2 one): ncu's peak_sustained of the L1TEX data pipe, "
                                               "derived__l1tex__lsu_writeback_bytes_mem_lgds.sum.peak_sustained = 18944 B/clk on 148 SMs (128 B/clk/SM), x the SM clock sampled in this run",
                                "traffic": traffic_scan,
                                "traffic_note": f"dram__bytes_read+write of one k_scan_stage0 launch (even columns of an 8-frame scan group), ncu --set full, {traffic_file}: the integral images are read from HBM about once, everything else is cache traffic",
                                "algorithmic_bytes_per_launch": alg_scan_bytes * group,
                                "ncu_pct_of_peak_even_launch": ncu_pct,
                                "l2_to_l1_bytes_per_launch": l2_bytes_ev, "l2_to_l1_achieved_gbs": l2_meas_gbs,
                                "stream_probe_gbs": l2_peak,
                                "hbm_peak": peak, "achieved_over_hbm_peak": (scan_gbs / peak) if scan_gbs else None,
                                "note": "gather-bound scan (SURVEY.md 8d): `achieved` = algorithmic corner bytes 32 B x (4 x reference-visited windows + 10 x reference weak evaluations) per frame / CUDA-event time of both stage-0 kernels. They are served by L1 and L2, not HBM (" + peak_src + ", given for scale only), so the roof is the L1TEX data pipe every 16-byte-per-lane load goes through. Since round 2 most corners are read from the compact integer plane (16 bytes per corner instead of the 32 the algorithmic figure counts), which is why `achieved` can pass what the pipe would deliver at 32 bytes per corner; ncu figures of the even-column launch are in ncu_pct_of_peak_even_launch. `l2_to_l1_achieved_gbs` = sectors L2 delivered in that launch (ncu l1tex__m_xbar2l1tex_read_bytes) / the launch's live duration; `stream_probe_gbs` = sc_probe_stream, coalesced 16-byte loads over a 32 MB table, measured live (LSU-bound: a reference point, not a ceiling)",
                                "share_of_step": st0_ms / total_k_ms}
            line["roofline_integral"] = {"kernel": "k_strip_carry+k_integral_walk", "bound": "hbm", "achieved": int_gbs, "peak": peak, "unit": "GB/s",
                                         "frac": (int_gbs / peak) if int_gbs else None, "traffic": traffic_int,
                                         "traffic_note": "dram bytes of one 8-frame k_integral_walk launch (ncu); bench launches cover 32 frames",
                                         "share_of_step": (walk_ms + carry_ms) / total_k_ms, "algorithmic_bytes_per_frame": alg_int_bytes,
                                         "note": "algorithmic bytes per SURVEY.md 8d (u8 in + 32 B per integral pixel out); the kernel also writes the 16-byte compact plane the scan reads (+50 % bytes)"}
            line["kernel_ms_per_frame"] = {k: v[0] / frames_timed for k, v in stats.items()}
            line["profiled_pass_ms_per_step"] = ms_prof / args.steps
        cpu = None
        if world == 1:
            try:
                cpu = cpu_baseline_block(5, 2, os.cpu_count() or 1, wl["model"])
            except Exception as e:  # the checker is optional for the product arm
                cpu = {"value": None, "unit": "frames/s", "cores": 0, "kind": "unavailable", "sample": repr(e)}
        line["cpu_baseline"] = cpu
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        if worker:
            jobs.put(None)
            worker.join()
        hx.comm_destroy()
        hx.close()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32, help="1080p frames per step per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS), help="c2 = BASELINE.json's configs[1] (default); c2_paper8; c3 = 1024-frame batch, strong scaling")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
