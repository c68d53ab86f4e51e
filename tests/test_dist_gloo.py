"""CPU, world_size 2, gloo: frame sharding and the detection gather (the N > 1 path's host logic)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from conftest import MODEL_C1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_frames, q):
    import torch
    import torch.distributed as dist
    from oracle import modelcfg, oracle
    from surfcascade_b200 import capi, synth
    from surfcascade_b200 import dist as scdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bc = oracle.BoundCascade(modelcfg.load(MODEL_C1))
    mine = scdist.shard_frames(n_frames, rank, world)
    recs = []
    for k, f in enumerate(mine):  # the oracle stands in for the GPU here: this test is about the exchange, not the kernels
        d = oracle.detect(oracle.integral(synth.frame(120, 160, f)), bc, oracle.params(base=40))
        r = np.zeros(len(d.x), capi.DETECTION_DTYPE)
        r["frame"] = k; r["x"] = d.x; r["y"] = d.y; r["l"] = d.l; r["score"] = d.score
        recs.append(r)
    local = np.concatenate(recs) if recs else np.zeros(0, capi.DETECTION_DTYPE)
    allr = scdist.gather_detections(local, rank, world)
    # fixed-capacity device-style gather (what bench.py runs over NCCL)
    cap = 4096
    buf = torch.zeros(cap * 24, dtype=torch.uint8)
    buf[:len(local) * 24] = torch.from_numpy(np.frombuffer(local.tobytes(), np.uint8).copy())
    bufs, counts = scdist.gather_records(buf, torch.tensor([len(local)], dtype=torch.int32))
    q.put((rank, allr.tobytes(), counts.tolist(), bytes(bufs[rank][:len(local) * 24].numpy().tobytes()) == local.tobytes()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    from oracle import modelcfg, oracle
    from surfcascade_b200 import capi, synth
    n_frames, world = 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    bc = oracle.BoundCascade(modelcfg.load(MODEL_C1))
    want = []
    for f in range(n_frames):
        d = oracle.detect(oracle.integral(synth.frame(120, 160, f)), bc, oracle.params(base=40))
        r = np.zeros(len(d.x), capi.DETECTION_DTYPE)
        r["frame"] = f; r["x"] = d.x; r["y"] = d.y; r["l"] = d.l; r["score"] = d.score
        want.append(r)
    want = np.sort(np.concatenate(want), order=["frame", "l", "y", "x"])
    per_rank = [sum(len(want[want["frame"] == f]) for f in range(r, n_frames, world)) for r in range(world)]
    for rank, blob, counts, own_ok in results:
        got = np.frombuffer(blob, capi.DETECTION_DTYPE)
        assert np.array_equal(got, want), f"rank {rank}"
        assert counts == per_rank and own_ok


def test_shard_frames_round_robin():
    from surfcascade_b200 import dist as scdist
    assert scdist.shard_frames(10, 1, 4) == [1, 5, 9]
    assert sorted(sum((scdist.shard_frames(1024, r, 8) for r in range(8)), [])) == list(range(1024))
    assert all(len(scdist.shard_frames(1024, r, 8)) == 128 for r in range(8))


def _band_worker(rank, world, port, q):
    import torch.distributed as dist
    from oracle import modelcfg, oracle
    from surfcascade_b200 import capi, synth
    from surfcascade_b200 import dist as scdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    H, W, base, step = 200, 260, 40, 2
    bc = oracle.BoundCascade(modelcfg.load(MODEL_C1))
    d = oracle.detect(oracle.integral(synth.frame(H, W, 6)), bc, oracle.params(base=base))
    # this rank's share of ONE frame: the lattice rows of its band, per scale (the oracle stands in for the GPU scan)
    keep = np.zeros(len(d.x), bool)
    for i in range(len(d.x)):
        ny = (H - int(d.l[i])) // step + 1
        y0, y1 = scdist.band_rows(ny, rank, world)
        keep[i] = y0 <= int(d.y[i]) // step < y1
    r = np.zeros(int(keep.sum()), capi.DETECTION_DTYPE)
    r["frame"] = 0; r["x"] = d.x[keep]; r["y"] = d.y[keep]; r["l"] = d.l[keep]; r["score"] = d.score[keep]
    allr = scdist.gather_detections(r, rank, world, renumber=False)   # same frame on every rank
    q.put((rank, allr.tobytes(), int(keep.sum())))
    dist.barrier()
    dist.destroy_process_group()


def test_single_frame_row_bands_world2():
    """One frame split by lattice rows over two ranks: the bands partition every scale's rows, the gather restores the scan."""
    from oracle import modelcfg, oracle
    from surfcascade_b200 import capi, synth
    from surfcascade_b200 import dist as scdist
    for ny in (1, 2, 7, 521):
        for world in (2, 3, 8):
            rows = [scdist.band_rows(ny, r, world) for r in range(world)]
            assert rows[0][0] == 0 and rows[-1][1] == ny and all(rows[i][1] == rows[i + 1][0] for i in range(world - 1))
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_band_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    bc = oracle.BoundCascade(modelcfg.load(MODEL_C1))
    d = oracle.detect(oracle.integral(synth.frame(200, 260, 6)), bc, oracle.params(base=40))
    want = np.zeros(len(d.x), capi.DETECTION_DTYPE)
    want["x"] = d.x; want["y"] = d.y; want["l"] = d.l; want["score"] = d.score
    want = np.sort(want, order=["frame", "l", "y", "x"])
    assert sum(k for _, _, k in results) == len(want) and all(k > 0 for _, _, k in results)
    for rank, blob, _ in results:
        assert np.array_equal(np.frombuffer(blob, capi.DETECTION_DTYPE), want)
