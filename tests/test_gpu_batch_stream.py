"""Parity of the batched / streamed entry points (per-frame solves, forward batches incl. the lattice plan, streamed
piecewise frames with device-side windows, checksums, the piecewise host-to-host pipe) with the CPU oracle: bit-exact.
Run on the B200 box:  python -m pytest tests -m gpu -x -q"""
import ctypes as C
import math

import numpy as np
import pytest
import torch

import homography_js_b200 as hg
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _rand_img(seed, w, h):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 4), dtype=np.uint8)


def _diff(a, b):
    a = np.asarray(a).reshape(-1, 4)
    b = np.asarray(b).reshape(-1, 4)
    return int((a != b).any(axis=1).sum())


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _grid_mesh(nx, ny, w, h):
    return hg.workloads.grid_mesh(nx, ny, w, h)


def _oracle_piecewise(img, W, H, src, dst, tris, smm):
    xo, yo, oW, oH = hg.workloads.piecewise_extent(dst)
    fwd = O.piecewise_matrices(src, dst, tris)
    imap = O.build_index_map(dst, tris, oW, yo, oW * oH)
    out = O.warp_inverse_piecewise(img, W, H, imap, O.inverse_matrices(fwd), xo, yo, oW, oH, smm[0], smm[1], threads=4)
    return (xo, yo, oW, oH), out


# ------------------------------------------------------------------ checksum
def test_checksum_frames_matches_numpy_definition(ctx):
    rng = np.random.default_rng(3)
    frames, bufs, want = [], [], []
    for (w, h) in ((1, 1), (7, 3), (64, 64), (333, 129), (1920, 17)):
        a = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        t = _dev(a.reshape(-1))
        bufs.append(t)
        frames.append(hg.HgFrame(None, t.data_ptr(), 0, 0, 0, 0, w, h))
        want.append(hg._abi.checksum_reference(a))
    torch.cuda.synchronize()
    got = ctx.checksum_frames(frames)
    assert [int(v) for v in got] == want
    # position-sensitive: swapping two pixels changes it
    a = rng.integers(0, 256, (4, 4, 4), dtype=np.uint8)
    b = a.copy()
    b[0, 0], b[3, 3] = a[3, 3].copy(), a[0, 0].copy()
    assert hg._abi.checksum_reference(a) != hg._abi.checksum_reference(b)


# ------------------------------------------------------------------ per-frame solve + warp batches
@pytest.mark.parametrize("kind", [0, 1])
def test_inverse_points_batch_solves_every_frame_on_the_device(ctx, kind):
    rng = np.random.default_rng(40 + kind)
    W, H, F = 160, 120, 9
    n = 6 if kind == 0 else 8
    base = np.array([0, 0, 0, H, W, 0, W, H], np.float64)[:n]
    imgs = [_rand_img(100 + f, W, H) for f in range(F)]
    srcs = [_dev(i.reshape(-1)) for i in imgs]
    dsts, frames, outs, wins = [], [], [], []
    for f in range(F):
        d = base + rng.uniform(-25, 25, n)
        if f == 3:
            d = base.copy()          # identity
        dsts.append(d)
        fwd = O.calculate_transform_matrix("projective" if kind else "affine", base, d)
        lim = [int(v) for v in O.transform_limits(fwd, W, H)]
        lim[2], lim[3] = max(lim[2], 1), max(lim[3], 1)
        wins.append(lim)
        o = torch.zeros(lim[2] * lim[3] * 4, dtype=torch.uint8, device="cuda")
        outs.append(o)
        frames.append(hg.HgFrame(srcs[f].data_ptr(), o.data_ptr(), W, H, lim[0], lim[1], lim[2], lim[3]))
    torch.cuda.synchronize()
    ctx.warp_inverse_points_batch(kind, np.stack(dsts), np.tile(base, (F, 1)), frames)
    ctx.synchronize()
    for f in range(F):
        inv = O.calculate_transform_matrix("projective" if kind else "affine", dsts[f], base)
        xo, yo, oW, oH = wins[f]
        want = O.warp_inverse_geometric(imgs[f], W, H, inv, xo, yo, oW, oH)
        assert _diff(outs[f].cpu().numpy(), want) == 0, f


# ------------------------------------------------------------------ forward batches
def _lattice_matrices():
    return [np.array(m, np.float32) for m in (
        [1, 0, 0, 1, 100, 50],            # translation (test.js:167-192 shape)
        [1, 0, 0, 1, 12.5, -7.25],        # half-integer translation: round(x + 12.5) = x + 13
        [1, 0, 0, 1, 0, 0],               # identity
        [-1, 0, 0, 1, 59, 0],             # mirror in x
        [1, 0, 0, -1, 3, 39.75],          # mirror in y
        [0, 1, -1, 0, 39.5, 2],           # quarter turn
        [0, -1, 1, 0, 0, 59],             # quarter turn the other way
        [-1, 0, 0, -1, 59.3, 39.6],       # half turn
        [1, 0, 0, 1, 1e-3, 0],            # tiny translation: NOT a lattice plan (general path), same answer
        [1, 0, 0, 1, 300, 0],             # everything lands outside / wraps: general path
    )]


def test_forward_batch_lattice_and_general_frames_bit_exact(ctx):
    W, H = 60, 40
    rng = np.random.default_rng(77)
    mats = _lattice_matrices()
    ang = 0.05
    mats += [np.array([math.cos(ang), math.sin(ang), -math.sin(ang), math.cos(ang), 2.2, 1.1], np.float32),   # collisions + holes
             np.array([0.6, 0, 0, 0.6, 5, 5], np.float32), np.array([1.5, 0.2, -0.1, 1.4, -3, 4], np.float32),
             np.full(6, np.nan, np.float32)]
    F = len(mats)
    imgs = [_rand_img(200 + f, W, H) for f in range(F)]
    srcs = [_dev(i.reshape(-1)) for i in imgs]
    frames, outs, wins = [], [], []
    for f in range(F):
        if f < 8:
            lim = [int(v) for v in O.transform_limits(mats[f], W, H)]
        else:
            lim = [int(rng.integers(-10, 10)), int(rng.integers(-10, 10)), int(rng.integers(20, 90)), int(rng.integers(20, 70))]
        wins.append(lim)
        o = torch.full((lim[2] * lim[3] * 4,), 77, dtype=torch.uint8, device="cuda")   # stale bytes must be overwritten
        outs.append(o)
        frames.append(hg.HgFrame(srcs[f].data_ptr(), o.data_ptr(), W, H, *lim))
    torch.cuda.synchronize()
    for rep in range(2):   # the second pass runs on winner planes handed back by the first
        ctx.warp_forward_batch(0, np.stack(mats), frames)
        ctx.synchronize()
        for f in range(F):
            want = O.warp_forward_geometric(imgs[f], W, H, mats[f], *wins[f])
            assert _diff(outs[f].cpu().numpy(), want) == 0, (rep, f, mats[f])


def test_forward_single_frame_lattice_plan_with_odd_windows(ctx):
    """The lattice plan must hold for windows that do not come from calculateTransformLimits: wrap-prone windows fall
    back to the general path, the others stay exact."""
    W, H = 37, 23
    img = _rand_img(5, W, H)
    ctx.image_set(img, W, H)
    rng = np.random.default_rng(6)
    for m in _lattice_matrices():
        for _ in range(6):
            xo, yo = int(rng.integers(-15, 15)), int(rng.integers(-15, 15))
            oW, oH = int(rng.integers(1, 80)), int(rng.integers(1, 60))
            got = ctx.warp_forward_matrix(m, xo, yo, oW, oH)
            want = O.warp_forward_geometric(img, W, H, m, xo, yo, oW, oH)
            assert _diff(got, want) == 0, (m, xo, yo, oW, oH)


def test_forward_projective_batch_bit_exact(ctx):
    W, H, F = 50, 30, 3
    imgs = [_rand_img(300 + f, W, H) for f in range(F)]
    srcs = [_dev(i.reshape(-1)) for i in imgs]
    mats = np.stack([np.array([1.1, 0.05, 2.0, -0.03, 0.95, 1.0, 1e-3, -5e-4]) * (1 + 0.01 * f) for f in range(F)])
    outs = [torch.zeros(80 * 60 * 4, dtype=torch.uint8, device="cuda") for _ in range(F)]
    frames = [hg.HgFrame(srcs[f].data_ptr(), outs[f].data_ptr(), W, H, -5, -5, 80, 60) for f in range(F)]
    torch.cuda.synchronize()
    ctx.warp_forward_batch(1, mats, frames)
    ctx.synchronize()
    for f in range(F):
        assert _diff(outs[f].cpu().numpy(), O.warp_forward_geometric(imgs[f], W, H, mats[f], -5, -5, 80, 60)) == 0


def test_piecewise_forward_batch_bit_exact(ctx):
    rng = np.random.default_rng(560)
    W, H, F = 240, 180, 5
    img = _rand_img(71, W, H)
    src, tris = _grid_mesh(6, 5, W, H)
    src = (src + rng.uniform(-4, 4, src.shape)).astype(np.float32)
    smm = [int(v) for v in O.minmax_xy(src)]
    ctx.image_set(img, W, H)
    ctx.piecewise_set_mesh(src, tris)
    dsts, frames, outs, wins = [], [], [], []
    for f in range(F):
        dst = (src * rng.uniform(0.85, 1.0) + rng.uniform(-8, 8, src.shape) + 10).astype(np.float32)
        dsts.append(dst)
        mm = O.minmax_xy(dst)
        win = [int(mm[0]), int(mm[1]), int(mm[2] - mm[0]), int(mm[3] - mm[1])]
        wins.append(win)
        o = torch.full((win[2] * win[3] * 4,), 9, dtype=torch.uint8, device="cuda")
        outs.append(o)
        frames.append(hg.HgFrame(None, o.data_ptr(), 0, 0, *win))
    torch.cuda.synchronize()
    ctx.warp_piecewise_forward_batch(np.stack(dsts), frames, smm[0], smm[1], smm[2], smm[3])
    ctx.synchronize()
    mw = smm[2] - smm[0]
    fmap = O.build_index_map(src, tris, mw, smm[1], mw * (smm[3] - smm[1]))
    for f in range(F):
        fwd = O.piecewise_matrices(src, dsts[f], tris)
        want = O.warp_forward_piecewise(img, W, H, fmap, fwd, *wins[f], smm[0], smm[1], smm[2], smm[3])
        assert _diff(outs[f].cpu().numpy(), want) == 0, f


# ------------------------------------------------------------------ streamed piecewise frames
def _stream_case(ctx, n_frames, n_slots, w=200, h=120, n_src=0, first=0, seed=5, bad=()):
    src_pts, dst_all, tris = hg.workloads.video_stream(first + n_frames, w, h, seed=seed)
    dst_all = dst_all[first:].copy()
    for b in bad:
        dst_all[b, 3, 0] = np.nan
    smm = [int(O.js_round(float(src_pts[:, 0].min()))), int(O.js_round(float(src_pts[:, 1].min())))]
    imgs = [_rand_img(900 + k, w, h) for k in range(max(n_src, 1))]
    ring_src = _dev(np.stack([i.reshape(-1) for i in imgs]))
    max_w, max_h = int(w * 1.1) + 8, int(h * 1.1) + 8
    slot = ctx.stream_slot_bytes(max_w, max_h)
    ring = torch.zeros(n_slots * slot, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.piecewise_set_mesh(src_pts, tris)
    if n_src:
        info = ctx.warp_piecewise_stream(dst_all, first, smm[0], smm[1], ring.data_ptr(), n_slots, max_w, max_h,
                                         src_ring_dev=ring_src.data_ptr(), n_src=n_src, src_w=w, src_h=h)
    else:
        ctx.image_set(imgs[0], w, h)
        info = ctx.warp_piecewise_stream(dst_all, first, smm[0], smm[1], ring.data_ptr(), n_slots, max_w, max_h)
    host = ring.cpu().numpy()
    checked = 0
    for f in range(n_frames):
        I = info[f]
        assert I.slot == (first + f) % n_slots
        if f in bad:
            assert I.status == 2
            continue
        assert I.status == 0
        win, want = _oracle_piecewise(imgs[(first + f) % max(n_src, 1)], w, h, src_pts, dst_all[f], tris, smm)
        assert (I.x_off, I.y_off, I.o_w, I.o_h) == win, (f, win)
        if f + n_slots < n_frames:
            continue   # its slot was reused by a later frame of the same call
        got = host[I.slot * slot: I.slot * slot + win[2] * win[3] * 4]
        assert _diff(got, want) == 0, f
        checked += 1
    return checked


def test_stream_windows_and_pixels_bit_exact(ctx):
    assert _stream_case(ctx, n_frames=12, n_slots=16) == 12


@pytest.mark.parametrize("mode", [1, 2])
def test_stream_under_both_binning_passes(ctx, mode):
    """span + run passes over global bins (1) and the one-pass band binning through shared memory (2): same bytes, and
    neither hands a frame to the map-based path."""
    ctx.debug_piecewise_binning(mode)
    try:
        f0, g0 = ctx.debug_piecewise_stats()
        assert _stream_case(ctx, n_frames=23, n_slots=8, n_src=3, first=7) == 8
        f1, g1 = ctx.debug_piecewise_stats()
        assert (f1 - f0, g1 - g0) == (23, 0)
    finally:
        ctx.debug_piecewise_binning(0)


def test_stream_wraps_the_ring_and_reads_a_source_ring(ctx):
    assert _stream_case(ctx, n_frames=23, n_slots=8, n_src=3, first=1000) == 8


def test_two_lane_chunk_pipeline_gives_the_same_frames(monkeypatch):
    """HG_PW_LANES=1 (the default for meshes binned by the span + run passes, forced here for both binning passes): binning
    passes of chunk k+1 on a second stream beside the pixel kernel of chunk k, two scratch sets.  Chunks of two frames so that a short stream / batch runs through both lanes several times, ring wrap
    and skipped frames included; both binning passes."""
    monkeypatch.setenv("HG_PW_LANES", "1")
    monkeypatch.setenv("HG_PW_CHUNK", "2")
    c = hg.Context(0)
    try:
        for mode in (2, 1):
            c.debug_piecewise_binning(mode)
            assert _stream_case(c, n_frames=11, n_slots=8, n_src=2, first=3) == 8
            assert _stream_case(c, n_frames=7, n_slots=16, bad=(4,)) == 6
        c.debug_piecewise_binning(0)
        # the batch entry point: nine 4:3 frames with their own windows
        w, h, n = 160, 120, 9
        img = _rand_img(77, w, h)
        src_pts, dst_all, tris = hg.workloads.video_stream(n, w, h, seed=9)
        smm = [int(O.js_round(float(src_pts[:, 0].min()))), int(O.js_round(float(src_pts[:, 1].min())))]
        c.image_set(img, w, h)
        c.piecewise_set_mesh(src_pts, tris)
        wins = [hg.workloads.piecewise_extent(dst_all[f]) for f in range(n)]
        outs = [torch.zeros(wn[2] * wn[3] * 4, dtype=torch.uint8, device="cuda") for wn in wins]
        torch.cuda.synchronize()
        frames = [hg.HgFrame(None, outs[f].data_ptr(), 0, 0, *wins[f]) for f in range(n)]
        c.warp_piecewise_inverse_batch(dst_all, frames, smm[0], smm[1])
        for f in range(n):
            win, want = _oracle_piecewise(img, w, h, src_pts, dst_all[f], tris, smm)
            assert tuple(win) == tuple(wins[f])
            assert _diff(outs[f].cpu().numpy(), want) == 0, f
    finally:
        c.close()


def test_fine_mesh_stream_takes_two_lanes_by_default(ctx):
    """A mesh of more than 2,048 triangles is binned by the span + run passes, and a stream longer than one chunk of 64 frames
    then runs in two lanes without any switch set (pw_two_lanes): every frame against the oracle, none through the map path."""
    w, h, n = 1100, 300, 70      # cells of 33 x 9 pixels: two cells per 64-column bin, well inside its eight entries
    img = _rand_img(41, w, h)
    src_pts, _, tris = hg.workloads.piecewise_sinusoid(34, 34, w, h)
    assert len(tris) > 2048
    dst_all = np.stack([hg.workloads.piecewise_sinusoid(34, 34, w, h, phase=2 * np.pi * f / n)[1] for f in range(n)])
    smm = [0, 0]
    max_w, max_h = w + 8, int(h * 1.1) + 16
    slot = ctx.stream_slot_bytes(max_w, max_h)
    n_slots = 80
    ring = torch.zeros(n_slots * slot, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.image_set(img, w, h)
    ctx.piecewise_set_mesh(src_pts, tris)
    f0, g0 = ctx.debug_piecewise_stats()
    info = ctx.warp_piecewise_stream(dst_all, 0, smm[0], smm[1], ring.data_ptr(), n_slots, max_w, max_h)
    f1, g1 = ctx.debug_piecewise_stats()
    assert (f1 - f0, g1 - g0) == (n, 0)
    host = ring.cpu().numpy()
    for f in range(n):
        I = info[f]
        assert I.status == 0 and I.slot == f
        win, want = _oracle_piecewise(img, w, h, src_pts, dst_all[f], tris, smm)
        assert (I.x_off, I.y_off, I.o_w, I.o_h) == win, (f, win)
        assert _diff(host[I.slot * slot: I.slot * slot + win[2] * win[3] * 4], want) == 0, f


def test_stream_skips_frames_without_a_window(ctx):
    assert _stream_case(ctx, n_frames=6, n_slots=6, bad=(2,)) == 5


def test_stream_general_path_fallback_matches(ctx):
    ctx.debug_force_general(True)
    try:
        assert _stream_case(ctx, n_frames=5, n_slots=4) == 4
    finally:
        ctx.debug_force_general(False)


def test_stream_frames_match_their_checksums(ctx):
    """What bench.py does at full size: per-frame 64-bit checksums of the ring slots against the oracle's frames."""
    w, h, n = 160, 96, 10
    src_pts, dst_all, tris = hg.workloads.video_stream(n, w, h)
    smm = [int(O.js_round(float(src_pts[:, 0].min()))), int(O.js_round(float(src_pts[:, 1].min())))]
    img = _rand_img(44, w, h)
    max_w, max_h = 200, 128
    slot = ctx.stream_slot_bytes(max_w, max_h)
    ring = torch.zeros(n * slot, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.image_set(img, w, h)
    ctx.piecewise_set_mesh(src_pts, tris)
    info = ctx.warp_piecewise_stream(dst_all, 0, smm[0], smm[1], ring.data_ptr(), n, max_w, max_h)
    frames = [hg.HgFrame(None, ring.data_ptr() + I.slot * slot, 0, 0, I.x_off, I.y_off, I.o_w, I.o_h) for I in info]
    got = ctx.checksum_frames(frames)
    for f in range(n):
        _, want = _oracle_piecewise(img, w, h, src_pts, dst_all[f], tris, smm)
        assert int(got[f]) == hg._abi.checksum_reference(want), f


# ------------------------------------------------------------------ piecewise host-to-host pipe
def test_piecewise_pipe_frames_bit_exact(ctx):
    w, h, n = 200, 120, 7
    src_pts, dst_all, tris = hg.workloads.video_stream(n, w, h)
    smm = [int(O.js_round(float(src_pts[:, 0].min()))), int(O.js_round(float(src_pts[:, 1].min())))]
    imgs = [_rand_img(700 + f, w, h) for f in range(n)]
    ctx.image_set(imgs[0], w, h)
    ctx.piecewise_set_mesh(src_pts, tris)
    max_w, max_h = 240, 150
    pipe = hg.Pipe.piecewise(ctx, w, h, max_w, max_h, depth=3)
    h_in = [ctx.pinned_array(w * h * 4) for _ in range(n)]
    h_out = [ctx.pinned_array(max_w * max_h * 4) for _ in range(n)]
    wins = []
    for f in range(n):
        h_in[f][:] = imgs[f].reshape(-1)
        # even frames bring their own image, odd frames warp the context image (the reference's video protocol)
        t, win = pipe.submit_piecewise(h_in[f].ctypes.data if f % 2 == 0 else None, dst_all[f], smm[0], smm[1], h_out[f].ctypes.data)
        wins.append(win)
    pipe.flush()
    for f in range(n):
        win, want = _oracle_piecewise(imgs[f] if f % 2 == 0 else imgs[0], w, h, src_pts, dst_all[f], tris, smm)
        assert wins[f] == win
        assert _diff(h_out[f][: win[2] * win[3] * 4], want) == 0, f
    bad = dst_all[0].copy()
    bad[:, 0] += 5000 * np.arange(len(bad))   # a window far larger than the pipe's maximum
    with pytest.raises(hg.HgError):
        pipe.submit_piecewise(None, bad, smm[0], smm[1], h_out[0].ctypes.data)
    pipe.close()


def test_pcie_probe_reports_three_positive_rates(ctx):
    a, b, c2 = ctx.pcie_probe(8 << 20, 4)
    assert a > 0.5 and b > 0.5 and c2 > 0.5


def test_forward_translation_shift_copy_every_alignment(ctx):
    """Translations between images with 16-byte aligned rows take the 128-bit shifted-copy form of the lattice kernel: every
    source alignment (x0 mod 4), windows hanging over every image edge, half-pixel translations."""
    W, H = 64, 20
    img = _rand_img(12, W, H)
    ctx.image_set(img, W, H)
    for e in list(range(-9, 10)) + [2.5, -3.5, 70, -70]:
        for f_, yo, oW, oH in ((0, 0, 64, 20), (3, -2, 72, 28), (-1.5, 5, 32, 12)):
            m = np.array([1, 0, 0, 1, e, f_], np.float32)
            got = ctx.warp_forward_matrix(m, 0, yo, oW, oH)
            want = O.warp_forward_geometric(img, W, H, m, 0, yo, oW, oH)
            assert _diff(got, want) == 0, (e, f_, yo, oW, oH)
