"""Synthetic workloads of BASELINE.json's configs (shapes and point formulas from SURVEY.md §8d, which lifts
them from the reference's test/benchmark.js and test/test.js) and the frame-sharding rule for multi-GPU runs."""
from __future__ import annotations

import math

import numpy as np


def shard_range(n_units: int, rank: int, world: int):
    """Block partition of independent frames over ranks: frames[r*N/G : (r+1)*N/G] (no data-path collective)."""
    lo = (n_units * rank) // world
    hi = (n_units * (rank + 1)) // world
    return lo, hi


def projective_1080p():
    """Config 2: projective 4-point warp, 1920x1080 RGBA8; dst = [[w/10,0],[w/10,h],[w,h/4],[w,3h/4]]
    (benchmark.js:282-283 shape) -> output window xOff=192, yOff=0, 1728x1080."""
    w, h = 1920, 1080
    src = np.array([0, 0, 0, h, w, 0, w, h], np.float64)
    dst = np.array([w / 10, 0, w / 10, h, w, h / 4, w, 3 * h / 4], np.float64)
    return dict(name="projective 4-point warp, 1920x1080 RGBA8 -> 1728x1080", kind=1, W=w, H=h, src=src, dst=dst,
                x_off=192, y_off=0, o_w=1728, o_h=1080)


def projective_1080p_generic():
    """Diagnostic twin of config 2 with "un-nice" destination points (no coordinate lands exactly on a pixel
    boundary, and the perspective has both an x and a y component)."""
    wl = projective_1080p()
    wl = dict(wl)
    wl["dst"] = wl["dst"] + np.array([0.37, 0.11, -0.23, 0.41, 0.19, -0.31, -0.13, 0.29])
    wl["name"] = "projective 4-point warp (generic points), 1920x1080 RGBA8 -> 1728x1080 window"
    return wl


def affine_1080p():
    """Affine variant of config 2 (same frame and output window sizes): dst = [[w/10,0],[w/10,h],[w,0]]."""
    w, h = 1920, 1080
    src = np.array([0, 0, 0, h, w, 0], np.float64)
    dst = np.array([w / 10, 0, w / 10, h, w, 0], np.float64)
    return dict(name="affine 3-point warp, 1920x1080 RGBA8 -> 1728x1080", kind=0, W=w, H=h, src=src, dst=dst,
                x_off=192, y_off=0, o_w=1728, o_h=1080)


def affine_1080p_rot90():
    """Diagnostic: 90-degree rotation (x' = h - y, y' = x) with a 10 % shrink so the inverse loop is dispatched; every
    warp's gathers walk DOWN a source column (worst case for coalescing, the case TMA-staged source tiles would fix)."""
    w, h = 1920, 1080
    src = np.array([0, 0, 0, h, w, 0], np.float64)
    dst = np.array([0.9 * h, 0, 0, 0, 0.9 * h, 0.9 * w], np.float64)
    return dict(name="affine 90-degree rotation x0.9, 1920x1080 RGBA8 -> 972x1728", kind=0, W=w, H=h, src=src, dst=dst,
                x_off=0, y_off=0, o_w=972, o_h=1728)


def affine_256():
    """Config 1: affine 3-point warp, 256x256 (benchmark.js:204-205 shape)."""
    w = h = 256
    src = np.array([0, 0, 0, h, w, 0], np.float64)
    dst = np.array([0, h / 2, w / 2, h * 0.8, w / 2, 0], np.float64)
    return dict(name="affine 3-point warp, 256x256 RGBA8", kind=0, W=w, H=h, src=src, dst=dst)


def grid_mesh(nx: int, ny: int, w: float, h: float):
    """Regular nx x ny point grid with the explicit triangle list of SURVEY §8d (cell (i,j): [p00,p10,p01],
    [p10,p11,p01]) — a regular grid is Delaunay-degenerate, so the triangles are given, not derived."""
    xs = np.arange(nx) * (w / (nx - 1))
    ys = np.arange(ny) * (h / (ny - 1))
    pts = np.array([[x, y] for y in ys for x in xs], np.float32)
    tris = []
    for j in range(ny - 1):
        for i in range(nx - 1):
            p00, p10, p01, p11 = j * nx + i, j * nx + i + 1, (j + 1) * nx + i, (j + 1) * nx + i + 1
            tris += [[p00, p10, p01], [p10, p11, p01]]
    return pts, np.array(tris, np.uint32)


def piecewise_sinusoid(nx: int, ny: int, w: int, h: int, phase: float = 0.0, amplitude: float | None = None):
    """Configs 3/4: dst = (x, A + y + A*sin(2*pi*2x/w + phase)), A = h/20 (test.js:133-144 shape)."""
    A = h / 20.0 if amplitude is None else amplitude
    src, tris = grid_mesh(nx, ny, w, h)
    dst = src.copy()
    dst[:, 1] = (A + src[:, 1].astype(np.float64)
                 + A * np.sin(2 * math.pi * 2 * src[:, 0].astype(np.float64) / w + phase)).astype(np.float32)
    return src, dst, tris


def piecewise_extent(dst: np.ndarray):
    """Piecewise output window (H.js:706-710): offsets = round(min), size = round(max) - round(min)."""
    def r(v):
        f = math.floor(v)
        return f + 1 if v - f >= 0.5 else f
    d = np.asarray(dst, np.float64).reshape(-1, 2)
    x0, y0, x1, y1 = r(d[:, 0].min()), r(d[:, 1].min()), r(d[:, 0].max()), r(d[:, 1].max())
    return int(x0), int(y0), int(x1 - x0), int(y1 - y0)


def video_stream(n_frames: int, w: int = 1920, h: int = 1080, seed: int = 5):
    """Config 5: 30 points (6x5 grid, interior points jittered +-2 % by rng seed 5), a fixed triangulation, and per-frame
    destiny points dst_f = src + 0.03*(w,h)*(sin(w1 f + th_i), cos(w2 f + th_i)); the output window changes every frame."""
    rng = np.random.default_rng(seed)
    src, tris = grid_mesh(6, 5, w, h)
    src = src.astype(np.float64)
    interior = (src[:, 0] > 0) & (src[:, 0] < w) & (src[:, 1] > 0) & (src[:, 1] < h)
    src[interior] += rng.uniform(-0.02, 0.02, (int(interior.sum()), 2)) * [w, h]
    src = src.astype(np.float32)
    theta = rng.uniform(0, 2 * math.pi, len(src))
    f = np.arange(n_frames)[:, None]
    dx = 0.03 * w * np.sin(0.11 * f + theta[None, :])
    dy = 0.03 * h * np.cos(0.07 * f + theta[None, :])
    dst = (src[None, :, :].astype(np.float64) + np.stack([dx, dy], axis=2)).astype(np.float32)
    return src, dst, tris


def u64_to_halves(v: int):
    """A 64-bit checksum as two 32-bit halves (low, high): what an int64 all-reduce can sum over ranks without overflow
    (at most 2^32 * world per half)."""
    v &= 0xFFFFFFFFFFFFFFFF
    return v & 0xFFFFFFFF, v >> 32


def halves_to_u64(lo_sum: int, hi_sum: int) -> int:
    """Sum mod 2^64 of the checksums whose halves were summed separately."""
    return (lo_sum + (hi_sum << 32)) & 0xFFFFFFFFFFFFFFFF
