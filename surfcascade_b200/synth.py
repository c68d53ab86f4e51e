"""Seeded synthetic workloads for the SURF-cascade detection path.

There is no dataset in the reference (its paths are hard-coded D:/FaceData/..., ObjDetector.cpp:62-64,146)
and no network here, so every frame, positive sample and negative frame is generated from numpy
`default_rng(seed)`.  Pure numpy; shared by tests, bench.py and the golden-vector scripts.

  frame(h, w, seed)            natural-image-like u8 frame: band-limited noise under a smooth contrast
                               envelope (flat regions fail the reference's gradient prefilter,
                               ObjDetector.cpp:188) plus planted objects at several scales
  noise_frame(h, w, seed)      un-blurred uniform noise: drives the float32 integral past 2^24 (H1)
  positive(seed) / negative_frame(seed)   training set for the reference trainer (config C1)
"""
from __future__ import annotations

import numpy as np

TEMPLATE = 40  # the reference's model template side (ObjDetector.cpp:112)


def _box_blur(a: np.ndarray, k: int) -> np.ndarray:
    """k x k mean filter with edge replication (float32 in/out)."""
    r = k // 2
    p = np.pad(a, r, mode="edge")
    c = np.cumsum(np.cumsum(p, axis=0, dtype=np.float64), axis=1)
    c = np.pad(c, ((1, 0), (1, 0)))
    h, w = a.shape
    s = c[k:k + h, k:k + w] - c[:h, k:k + w] - c[k:k + h, :w] + c[:h, :w]
    return (s / (k * k)).astype(np.float32)


def _smooth_field(h: int, w: int, rng: np.random.Generator, cell: int) -> np.ndarray:
    """Low-frequency field in [0,1]: coarse uniform grid, bilinear upsample."""
    gh, gw = h // cell + 2, w // cell + 2
    g = rng.random((gh, gw), dtype=np.float32)
    ys = np.arange(h, dtype=np.float32) / cell
    xs = np.arange(w, dtype=np.float32) / cell
    y0 = ys.astype(np.int32); x0 = xs.astype(np.int32)
    fy = (ys - y0)[:, None]; fx = (xs - x0)[None, :]
    a = g[y0][:, x0]; b = g[y0][:, x0 + 1]; c = g[y0 + 1][:, x0]; d = g[y0 + 1][:, x0 + 1]
    return (a * (1 - fy) * (1 - fx) + b * (1 - fy) * fx + c * fy * (1 - fx) + d * fy * fx).astype(np.float32)


def texture(h: int, w: int, rng: np.random.Generator) -> np.ndarray:
    """Band-limited noise (uniform u8, two 3x3 box passes) under a smooth contrast envelope, float32."""
    n = rng.integers(0, 256, size=(h, w), dtype=np.uint8).astype(np.float32)
    n = _box_blur(_box_blur(n, 3), 3) - 127.5
    env = _smooth_field(h, w, rng, 48)
    env = np.clip((env - 0.35) * 2.2, 0.0, 1.0)  # ~1/3 of the area is (nearly) flat
    lum = 96.0 + 64.0 * _smooth_field(h, w, rng, 96)
    return lum + 1.0 * env * n


def object_template() -> np.ndarray:
    """Fixed 40x40 'object': a bright oval with two dark blobs and a dark bar (float32, ~[-1,1])."""
    t = TEMPLATE
    yy, xx = np.mgrid[0:t, 0:t].astype(np.float32)
    cy = cx = (t - 1) / 2.0
    oval = np.exp(-(((xx - cx) / 15.0) ** 2 + ((yy - cy) / 18.0) ** 2) ** 2)
    eye = lambda ex, ey: np.exp(-(((xx - ex) / 3.5) ** 2 + ((yy - ey) / 2.5) ** 2))
    bar = np.exp(-(((xx - cx) / 7.0) ** 4 + ((yy - 29.0) / 2.0) ** 2))
    nose = np.exp(-(((xx - cx) / 1.8) ** 2 + ((yy - 21.0) / 5.0) ** 2))
    return (oval - 1.1 * eye(13.0, 14.0) - 1.1 * eye(26.0, 14.0) - 0.9 * bar - 0.35 * nose).astype(np.float32)


def _resize_bilinear(a: np.ndarray, side: int) -> np.ndarray:
    h, w = a.shape
    ys = (np.arange(side, dtype=np.float32) + 0.5) * h / side - 0.5
    xs = (np.arange(side, dtype=np.float32) + 0.5) * w / side - 0.5
    ys = np.clip(ys, 0, h - 1); xs = np.clip(xs, 0, w - 1)
    y0 = np.minimum(ys.astype(np.int32), h - 2); x0 = np.minimum(xs.astype(np.int32), w - 2)
    fy = (ys - y0)[:, None]; fx = (xs - x0)[None, :]
    return (a[y0][:, x0] * (1 - fy) * (1 - fx) + a[y0][:, x0 + 1] * (1 - fy) * fx +
            a[y0 + 1][:, x0] * fy * (1 - fx) + a[y0 + 1][:, x0 + 1] * fy * fx).astype(np.float32)


def _render_object(side: int, rng: np.random.Generator, contrast: float) -> np.ndarray:
    """Object at `side` px with +-5% jitter, returned as an additive float32 patch."""
    pad = 4
    big = np.pad(object_template(), pad, mode="constant")
    dx, dy = rng.integers(-2, 3, size=2)
    big = np.roll(big, (int(dy), int(dx)), axis=(0, 1))[pad:-pad, pad:-pad]
    return contrast * _resize_bilinear(big, side)


def positive(seed: int) -> np.ndarray:
    """One 40x40 u8 positive: object over a weak texture, random contrast, N(0,8^2) noise, +-2 px jitter."""
    rng = np.random.default_rng(1_000_003 + seed)
    bg = texture(TEMPLATE, TEMPLATE, rng)
    bg = 0.5 * (bg - bg.mean()) + 110.0 + 30.0 * rng.random()
    contrast = 22.0 + 50.0 * rng.random()
    img = bg + _render_object(TEMPLATE, rng, contrast) + rng.normal(0.0, 8.0, size=(TEMPLATE, TEMPLATE))
    return np.ascontiguousarray(np.clip(np.rint(img), 0, 255).astype(np.uint8))


def _distractor(side: int, rng: np.random.Generator, contrast: float) -> np.ndarray:
    """Object-like clutter for the negatives: flipped, half-erased or inverted objects."""
    o = _render_object(side, rng, contrast)
    kind = int(rng.integers(0, 6))
    if kind == 0:
        o = o[::-1, :]
    elif kind == 1:
        o = o.copy(); o[:, side // 2:] = 0
    elif kind == 2:
        o = -o
    elif kind == 3:
        o = o.T
    elif kind == 4:
        o = o.copy(); o[side // 2:, :] = 0
    else:
        o = np.roll(o, side // 3, axis=int(rng.integers(0, 2)))
    return o


def negative_frame(seed: int, h: int = 240, w: int = 320) -> np.ndarray:
    """u8 negative frame: the texture plus object-like distractors (never the upright object)."""
    rng = np.random.default_rng(2_000_003 + seed)
    img = texture(h, w, rng)
    for _ in range(14):
        side = int(rng.integers(40, min(h, w) // 2))
        y = int(rng.integers(0, h - side)); x = int(rng.integers(0, w - side))
        img[y:y + side, x:x + side] += _distractor(side, rng, 35.0 + 40.0 * rng.random())
    img += rng.normal(0.0, 4.0, size=(h, w))
    return np.ascontiguousarray(np.clip(np.rint(img), 0, 255).astype(np.uint8))


def frame(h: int, w: int, seed: int, n_objects: int = 8) -> np.ndarray:
    """u8 test frame: texture, distractors and `n_objects` planted objects at three scales."""
    rng = np.random.default_rng(3_000_003 + seed)
    img = texture(h, w, rng)
    m = min(h, w)
    for _ in range(max(2, n_objects // 2)):
        side = int(rng.integers(40, max(41, m // 3)))
        y = int(rng.integers(0, h - side)); x = int(rng.integers(0, w - side))
        img[y:y + side, x:x + side] += _distractor(side, rng, 35.0 + 40.0 * rng.random())
    sides = [48, max(56, m // 6), max(64, m // 3)]
    for k in range(n_objects):
        side = min(sides[k % 3], m - 8)
        y = int(rng.integers(0, h - side)); x = int(rng.integers(0, w - side))
        img[y:y + side, x:x + side] += _render_object(side, rng, 55.0 + 30.0 * rng.random())
    img += rng.normal(0.0, 4.0, size=(h, w))
    return np.ascontiguousarray(np.clip(np.rint(img), 0, 255).astype(np.uint8))


def noise_frame(h: int, w: int, seed: int) -> np.ndarray:
    """Un-blurred uniform u8 noise: every channel integral crosses 2^24 at >= 480p."""
    return np.random.default_rng(4_000_003 + seed).integers(0, 256, size=(h, w), dtype=np.uint8)
