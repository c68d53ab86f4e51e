"""Multi-GPU plumbing: frames are independent units, sharded round-robin over ranks; the only
exchange step of the path is gathering the variable-length detection records for host-side grouping.
torch.distributed is the transport (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .capi import DETECTION_DTYPE

RECORD_BYTES = DETECTION_DTYPE.itemsize  # 24: {i32 frame, i32 x, i32 y, i32 l, f64 score}


def shard_frames(n_frames: int, rank: int, world: int) -> list[int]:
    """Frame f -> rank f mod world."""
    return list(range(rank, n_frames, world))


def to_global_frames(dets: np.ndarray, rank: int, world: int) -> np.ndarray:
    """Local frame index k on `rank` is global frame k * world + rank."""
    out = dets.copy()
    out["frame"] = dets["frame"] * world + rank
    return out


def band_rows(ny: int, rank: int, world: int) -> tuple[int, int]:
    """Lattice rows [y0, y1) of a scale with ny rows that rank `rank` scans when ONE frame is split over `world` GPUs
    (sc_detect_params.band_index / band_count; the reference's stride chain never crosses a row, ObjDetector.cpp:182-186)."""
    return ny * rank // world, ny * (rank + 1) // world


def band_params(rank: int, world: int, **kw):
    """Scan parameters of rank `rank` for a single frame split by lattice rows over `world` GPUs."""
    from . import capi
    return capi.params(band_index=rank, band_count=world, **kw)


def gather_records(buf: torch.Tensor, count: torch.Tensor, group=None) -> tuple[torch.Tensor, torch.Tensor]:
    """All-gather fixed-capacity record buffers and their fill counts.

    buf: uint8 [cap * 24] (device or host), count: int32 [1].  Returns (bufs [world, cap*24], counts [world]).
    Counts go first so a consumer can size follow-up work; both collectives are launched back to back."""
    world = dist.get_world_size(group)
    counts = [torch.empty_like(count) for _ in range(world)]
    bufs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(counts, count, group=group)
    dist.all_gather(bufs, buf, group=group)
    return torch.stack(bufs), torch.cat(counts)


def gather_detections(local: np.ndarray, rank: int, world: int, device: torch.device | None = None, group=None,
                      renumber: bool = True) -> np.ndarray:
    """Host-level helper: every rank passes its detections (local frame indices); every rank gets all of them with
    global frame indices, sorted by (frame, l, y, x).  renumber=False keeps the frame indices (ranks that scanned row
    bands of the same frames)."""
    device = device or torch.device("cpu")
    g = to_global_frames(local, rank, world) if renumber else local
    n = torch.tensor([len(g)], dtype=torch.int32, device=device)
    ns = [torch.empty_like(n) for _ in range(world)]
    dist.all_gather(ns, n, group=group)
    cap = max(int(x.item()) for x in ns)
    buf = torch.zeros(max(cap, 1) * RECORD_BYTES, dtype=torch.uint8)
    if len(g):
        buf[:len(g) * RECORD_BYTES] = torch.from_numpy(np.frombuffer(g.tobytes(), np.uint8).copy())
    buf = buf.to(device)
    bufs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf, group=group)
    parts = []
    for r in range(world):
        k = int(ns[r].item())
        if k:
            parts.append(np.frombuffer(bufs[r].cpu().numpy().tobytes()[:k * RECORD_BYTES], DETECTION_DTYPE))
    allr = np.concatenate(parts) if parts else np.zeros(0, DETECTION_DTYPE)
    return np.sort(allr, order=["frame", "l", "y", "x"])
