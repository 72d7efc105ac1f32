"""Parity of the CUDA path (through the C ABI) with the CPU oracle: bit-exact for every byte.
Run on the B200 box:  python -m pytest tests -m gpu -x -q"""
import math

import numpy as np
import pytest

import flows
import homography_js_b200 as hg
from oracle import oracle as O
from oracle.homography_ref import RefHomography, RefImageData

pytestmark = pytest.mark.gpu


def _rand_img(seed, w, h):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 4), dtype=np.uint8)


def _diff(a, b):
    a = a.reshape(-1, 4)
    b = b.reshape(-1, 4)
    return int((a != b).any(axis=1).sum())


# ------------------------------------------------------------------ solves (K5)
def test_affine_solve_bit_exact(ctx):
    rng = np.random.default_rng(21)
    for k in range(400):
        s = rng.uniform(-500, 4000, 6)
        d = rng.uniform(-500, 4000, 6)
        if k % 4 == 0:
            s, d = s.astype(np.float32).astype(np.float64), d.astype(np.float32).astype(np.float64)
        if k % 50 == 7:
            s[2:4] = s[0:2]  # degenerate: zero determinant -> Inf / NaN entries, not an error
        got, want = ctx.solve_affine(s, d), O.affine_from_triangles(s, d)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (k, got, want)


def test_inverse_affine_bit_exact(ctx):
    rng = np.random.default_rng(22)
    for k in range(300):
        m = rng.uniform(-3, 3, 6).astype(np.float32)
        m[4:] *= 500
        if k % 60 == 3:
            m[:4] = [1, 2, 2, 4]  # singular
        assert np.array_equal(ctx.inverse_affine(m).view(np.uint32), O.inverse_affine(m).view(np.uint32)), k


def test_projective_solve_bit_exact(ctx):
    rng = np.random.default_rng(23)
    for k in range(400):
        s = rng.uniform(0, 4000, 8)
        d = rng.uniform(0, 4000, 8)
        if k % 3 == 0:
            s, d = s.astype(np.float32).astype(np.float64), d.astype(np.float32).astype(np.float64)
        if k % 5 == 0:  # axis-aligned rectangle as source: many exact zeros / pivot ties
            w, h = rng.integers(100, 4000, 2)
            s = np.array([0, 0, 0, h, w, 0, w, h], np.float64)
        if k % 67 == 11:
            s[6:8] = s[0:2]  # degenerate quad
        got, want = ctx.solve_projective(s, d), O.projective_from_squares(s, d)
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), (k, got, want)


def test_limits_bit_exact(ctx):
    rng = np.random.default_rng(24)
    for k in range(200):
        if k % 2:
            m = rng.uniform(-2, 2, 6).astype(np.float32)
            m[4:] *= 300
        else:
            m = np.concatenate([rng.uniform(-2, 2, 6), rng.uniform(-1e-3, 1e-3, 2)])
            m[2] *= 300
            m[5] *= 300
        w, h = rng.integers(1, 4000, 2)
        got, want = ctx.transform_limits(m, w, h), O.transform_limits(m, w, h)
        assert np.array_equal(got, want, equal_nan=True), (k, got, want)
    nan_m = np.full(6, np.nan, np.float32)
    assert np.isnan(ctx.transform_limits(nan_m, 10, 10)).all()


def test_solve_with_limits_matches_separate_calls(ctx):
    s = np.array([0, 0, 0, 1080, 1920, 0, 1920, 1080], np.float64)
    d = np.array([192, 0, 192, 1080, 1920, 270, 1920, 810], np.float64)
    m, lim = ctx.solve_with_limits(hg._abi.HG_PROJECTIVE, s, d, 1920, 1080)
    assert np.array_equal(m, O.projective_from_squares(s, d))
    assert np.array_equal(lim, O.transform_limits(m, 1920, 1080))
    assert list(lim) == [192.0, 0.0, 1728.0, 1080.0]


# ------------------------------------------------------------------ inverse affine / projective warp (K1/K2)
def _geo_case(ctx, img, inv, xo, yo, oW, oH):
    H, W = img.shape[:2]
    ctx.image_set(img, W, H)
    got = ctx.warp_inverse_matrix(inv, xo, yo, oW, oH)
    want = O.warp_inverse_geometric(img, W, H, inv, xo, yo, oW, oH, threads=4)
    assert _diff(got, want) == 0, f"{_diff(got, want)} of {oW * oH} pixels differ"


@pytest.mark.parametrize("seed", range(12))
def test_affine_inverse_warp_random(ctx, seed):
    rng = np.random.default_rng(100 + seed)
    W, H = int(rng.integers(5, 300)), int(rng.integers(5, 300))
    img = _rand_img(seed, W, H)
    ang = rng.uniform(0, 2 * math.pi)
    sc = rng.uniform(0.3, 3.0)
    inv = np.array([math.cos(ang) * sc, math.sin(ang) * sc, -math.sin(ang) * sc * rng.uniform(0.5, 1.5),
                    math.cos(ang) * sc, rng.uniform(-W, W), rng.uniform(-H, H)], np.float32)
    _geo_case(ctx, img, inv, int(rng.integers(-50, 50)), int(rng.integers(-50, 50)),
              int(rng.integers(1, 400)), int(rng.integers(1, 400)))


@pytest.mark.parametrize("deg", [90.0, 97.5, 270.0, 262.0])
def test_affine_quarter_turn_maps_take_the_tall_kernel(ctx, deg):
    """Maps near a quarter turn run warp_inverse_geo_affine_tall_kernel (2 x 64 thread layout, 8 CTAs per SM): same pixels."""
    W, H = 333, 251
    img = _rand_img(int(deg), W, H)
    a = math.radians(deg)
    inv = np.array([math.cos(a), math.sin(a), -math.sin(a), math.cos(a), W / 2 + 3.25, H / 2 - 7.5], np.float32)
    for oW, oH in ((260, 341), (257, 130), (3, 400)):
        _geo_case(ctx, img, inv, -oW // 2, -oH // 2, oW, oH)


@pytest.mark.parametrize("seed", range(12))
def test_projective_inverse_warp_random(ctx, seed):
    rng = np.random.default_rng(200 + seed)
    W, H = int(rng.integers(5, 300)), int(rng.integers(5, 300))
    img = _rand_img(seed, W, H)
    s = np.array([0, 0, 0, H, W, 0, W, H], np.float64)
    d = s + rng.uniform(-0.3, 0.3, 8) * max(W, H)
    inv = O.projective_from_squares(d, s)
    lim = O.transform_limits(O.projective_from_squares(s, d), W, H)
    xo, yo, oW, oH = [int(v) for v in lim]
    oW, oH = max(1, min(oW, 1500)), max(1, min(oH, 1500))
    _geo_case(ctx, img, inv, xo, yo, oW, oH)


def test_half_integer_and_boundary_coordinates(ctx):
    """Coordinates that land EXACTLY on k+0.5 (Math.round ties up), on 0 and on W (Q1/Q2): scale 0.5 and
    integer shifts make every decision boundary exact; also exercises the projective exact-division path."""
    img = _rand_img(5, 64, 48)
    for inv in (np.array([0.5, 0, 0, 0.5, 0, 0], np.float32),
                np.array([0.5, 0, 0, 0.5, -0.5, 31.5], np.float32),
                np.array([-1, 0, 0, -1, 64, 48], np.float32),
                np.array([0.25, 0, 0, 0.75, 0.125, -0.25], np.float32)):
        _geo_case(ctx, img, inv, -8, -8, 160, 130)
        h8 = np.array([inv[0], inv[2], inv[4], inv[1], inv[3], inv[5], 0.0, 0.0], np.float64)
        _geo_case(ctx, img, h8, -8, -8, 160, 130)
    # a projective map with exact dyadic quotients: denominators are powers of two
    _geo_case(ctx, img, np.array([1, 0, 0, 0, 1, 0, 0, 0.0], np.float64), 0, 0, 64, 48)
    _geo_case(ctx, img, np.array([0.5, 0, 0.5, 0, 0.5, 0.5, 0, 0.0], np.float64), -3, -3, 140, 110)


def test_degenerate_matrices_give_transparent_output(ctx):
    img = _rand_img(6, 32, 32)
    for inv in (np.full(6, np.nan, np.float32), np.array([np.inf, 0, 0, 1, 0, 0], np.float32),
                np.full(8, np.nan), np.array([1, 0, 0, 0, 1, 0, 1e308, 1e308]), np.zeros(8)):
        _geo_case(ctx, img, inv, 0, 0, 40, 40)


def test_horizon_crossing_projective(ctx):
    """Denominator changes sign inside the output window (huge / negative / infinite quotients)."""
    img = _rand_img(7, 200, 100)
    inv = np.array([1.0, 0.1, 5.0, 0.05, 1.0, 3.0, -0.01, 0.002], np.float64)
    _geo_case(ctx, img, inv, -20, -20, 300, 200)


def test_output_tail_not_multiple_of_four(ctx):
    img = _rand_img(8, 17, 13)
    inv = np.array([1, 0, 0, 1, 0, 0], np.float32)
    for oW, oH in ((1, 1), (3, 1), (5, 7), (17, 13), (19, 3)):
        _geo_case(ctx, img, inv, 0, 0, oW, oH)


def test_inverse_points_entry_solves_on_device(ctx, golden):
    src = (golden["src_points"].reshape(-1) * 400).astype(np.float32).astype(np.float64)
    dst = (golden["dst_points"].reshape(-1) * 400).astype(np.float32).astype(np.float64)
    ctx.image_set(golden["src"], 400, 400)
    got = ctx.warp_inverse_points(hg._abi.HG_PROJECTIVE, dst, src, 0, 200, 400, 200)
    assert np.array_equal(got.reshape(200, 400, 4), golden["out"])


def test_config2_full_size_projective_1080p(ctx):
    """BASELINE config 2 at full size (1920x1080 -> 1728x1080), every pixel compared."""
    w, h = 1920, 1080
    img = _rand_img(2, w, h)
    s = np.array([0, 0, 0, h, w, 0, w, h], np.float64)
    d = np.array([w / 10, 0, w / 10, h, w, h / 4, w, 3 * h / 4], np.float64)
    fwd, lim = ctx.solve_with_limits(hg._abi.HG_PROJECTIVE, s, d, w, h)
    assert list(lim) == [192.0, 0.0, 1728.0, 1080.0]
    ctx.image_set(img, w, h)
    got = ctx.warp_inverse_points(hg._abi.HG_PROJECTIVE, d, s, 192, 0, 1728, 1080)
    want = O.warp_inverse_geometric(img, w, h, O.projective_from_squares(d, s), 192, 0, 1728, 1080, threads=8)
    assert _diff(got, want) == 0


def test_identity_at_4k_is_a_copy(ctx):
    """Size-independent property at 3840x2160: the identity warp returns the image."""
    w, h = 3840, 2160
    img = _rand_img(3, w, h)
    ctx.image_set(img, w, h)
    out = ctx.warp_inverse_matrix(np.array([1, 0, 0, 1, 0, 0], np.float32), 0, 0, w, h)
    assert np.array_equal(out.reshape(h, w, 4), img)
    out = ctx.warp_inverse_matrix(np.array([1, 0, 0, 0, 1, 0, 0, 0], np.float64), 0, 0, w, h)
    assert np.array_equal(out.reshape(h, w, 4), img)
    # integer translation: out[y, x] = img[y + 5, x + 7] inside, transparent outside
    out = ctx.warp_inverse_matrix(np.array([1, 0, 0, 1, 7, 5], np.float32), 0, 0, w, h).reshape(h, w, 4)
    assert np.array_equal(out[: h - 5, : w - 7], img[5:, 7:])
    assert not out[h - 5:].any() and not out[:, w - 7:].any()


def test_batch_entry_matches_single_frames(ctx):
    rng = np.random.default_rng(31)
    W, H = 160, 120
    imgs = [_rand_img(40 + k, W, H) for k in range(3)]
    n = 6
    dev_src = [ctx.dev_alloc(W * H * 4) for _ in imgs]
    for p, im in zip(dev_src, imgs):
        ctx.memcpy_h2d(p, im.ctypes.data, im.nbytes)
    mats, frames, outs, shapes = [], [], [], []
    for f in range(n):
        s = np.array([0, 0, 0, H, W, 0, W, H], np.float64)
        d = s + rng.uniform(-0.2, 0.2, 8) * W
        mats.append(O.projective_from_squares(d, s))
        xo, yo, oW, oH = [int(v) for v in O.transform_limits(O.projective_from_squares(s, d), W, H)]
        p = ctx.dev_alloc(oW * oH * 4)
        outs.append(p)
        shapes.append((xo, yo, oW, oH))
        frames.append(hg.HgFrame(dev_src[f % 3], p, W, H, xo, yo, oW, oH))
    ctx.warp_inverse_batch(hg._abi.HG_PROJECTIVE, np.stack(mats), frames)
    for f in range(n):
        xo, yo, oW, oH = shapes[f]
        got = np.empty(oW * oH * 4, np.uint8)
        ctx.memcpy_d2h(got.ctypes.data, outs[f], got.nbytes)
        ctx.synchronize()
        want = O.warp_inverse_geometric(imgs[f % 3], W, H, mats[f], xo, yo, oW, oH)
        assert _diff(got, want) == 0, f
    for p in dev_src + outs:
        ctx.dev_free(p)


# ------------------------------------------------------------------ piecewise (K3/K4/K5)
def _grid_mesh(nx, ny, w, h):
    xs = np.arange(nx) * (w / (nx - 1))
    ys = np.arange(ny) * (h / (ny - 1))
    pts = np.array([[x, y] for y in ys for x in xs], np.float32)
    tris = []
    for j in range(ny - 1):
        for i in range(nx - 1):
            p00, p10, p01, p11 = j * nx + i, j * nx + i + 1, (j + 1) * nx + i, (j + 1) * nx + i + 1
            tris += [[p00, p10, p01], [p10, p11, p01]]
    return pts, np.array(tris, np.uint32)


def test_piecewise_matrices_bit_exact(ctx):
    rng = np.random.default_rng(51)
    src, tris = _grid_mesh(9, 7, 640, 480)
    dst = (src + rng.uniform(-20, 20, src.shape)).astype(np.float32)
    ctx.piecewise_set_mesh(src, tris)
    fwd, inv = ctx.piecewise_matrices(dst, want_inverse=True)
    want = O.piecewise_matrices(src, dst, tris)
    assert np.array_equal(fwd.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(inv.view(np.uint32), O.inverse_matrices(want).view(np.uint32))


@pytest.mark.parametrize("seed", range(8))
def test_index_map_bit_exact(ctx, seed):
    """A5 incl. the quirks: no x offset (Q4), negative relative fill index (Q5), overlaps (Q7)."""
    rng = np.random.default_rng(300 + seed)
    if seed < 4:
        src, tris = _grid_mesh(6, 5, 200, 150)
        pts = (src + rng.uniform(-12, 12, src.shape) + rng.uniform(-30, 30, 2)).astype(np.float32)
    else:
        pts = rng.uniform(-20, 220, (12, 2)).astype(np.float32)
        if seed % 2:
            pts = np.round(pts)
        tris = rng.integers(0, 12, (15, 3)).astype(np.uint32)  # arbitrary, overlapping, possibly degenerate
    ctx.piecewise_set_mesh(pts, tris)
    mm = O.minmax_xy(pts)
    mw = float(mm[2] - mm[0]) if seed % 2 == 0 else 211.0
    yoff = float(mm[1])
    length = int(mw * (mm[3] - mm[1] + 3))
    got = ctx.build_index_map(pts, mw, yoff, length)
    want = O.build_index_map(pts, tris, mw, yoff, length)
    assert np.array_equal(got, want), f"{(got != want).sum()} map entries differ"


@pytest.mark.parametrize("seed", range(6))
def test_piecewise_inverse_warp_bit_exact(ctx, seed):
    rng = np.random.default_rng(400 + seed)
    W, H = 320, 200
    img = _rand_img(60 + seed, W, H)
    src, tris = _grid_mesh(7, 5, W, H)
    dst = src.copy()
    dst[:, 0] = dst[:, 0] * rng.uniform(0.8, 2.2) + rng.uniform(-15, 15, len(dst))
    dst[:, 1] = dst[:, 1] * rng.uniform(0.8, 2.2) + rng.uniform(-15, 15, len(dst)) + rng.uniform(0, 40)
    dst = dst.astype(np.float32)
    mm = O.minmax_xy(dst)
    xo, yo, oW, oH = int(mm[0]), int(mm[1]), int(mm[2] - mm[0]), int(mm[3] - mm[1])
    smm = O.minmax_xy(src)
    ctx.image_set(img, W, H)
    ctx.piecewise_set_mesh(src, tris)
    got = ctx.warp_piecewise_inverse(dst, xo, yo, oW, oH, int(smm[0]), int(smm[1]))
    fwd = O.piecewise_matrices(src, dst, tris)
    imap = O.build_index_map(dst, tris, oW, yo, oW * oH)
    want = O.warp_inverse_piecewise(img, W, H, imap, O.inverse_matrices(fwd), xo, yo, oW, oH, int(smm[0]), int(smm[1]), threads=4)
    assert _diff(got, want) == 0, f"{_diff(got, want)} of {oW * oH} pixels differ"


def test_config3_full_size_piecewise_4k(ctx):
    """BASELINE config 3: 10x10 grid (162 triangles), 3840x2160, sinusoidal destiny points, all pixels."""
    w, h = 3840, 2160
    img = _rand_img(3, w, h)
    src, tris = _grid_mesh(10, 10, w, h)
    A = 108.0
    dst = src.copy()
    dst[:, 1] = (A + src[:, 1] + A * np.sin(2 * np.pi * 2 * src[:, 0].astype(np.float64) / w)).astype(np.float32)
    mm = O.minmax_xy(dst)
    xo, yo, oW, oH = int(mm[0]), int(mm[1]), int(mm[2] - mm[0]), int(mm[3] - mm[1])
    assert oH > h
    ctx.image_set(img, w, h)
    ctx.piecewise_set_mesh(src, tris)
    got = ctx.warp_piecewise_inverse(dst, xo, yo, oW, oH, 0, 0)
    fwd = O.piecewise_matrices(src, dst, tris)
    imap = O.build_index_map(dst, tris, oW, yo, oW * oH)
    want = O.warp_inverse_piecewise(img, w, h, imap, O.inverse_matrices(fwd), xo, yo, oW, oH, 0, 0, threads=8)
    assert _diff(got, want) == 0


# ------------------------------------------------------------------ the class surface over CUDA
@pytest.mark.parametrize("flow", flows.ALL, ids=lambda f: f.__name__)
def test_reference_test_page_flows_over_cuda(ctx, flow, golden):
    ref_res, ref = flow(lambda *a: RefHomography(*a), RefImageData(golden["src"].reshape(-1).copy(), 400, 400))
    got_res, got = flow(lambda *a: hg.Homography(*a, context=ctx), hg.ImageData(golden["src"].reshape(-1).copy(), 400, 400))
    assert got.last_path == ref.last_path
    for r, g in zip(ref_res, got_res):
        assert (g.width, g.height) == (r.width, r.height)
        assert _diff(g.data, r.data) == 0


def test_config1_affine_256_over_cuda(ctx):
    """BASELINE.json configs[0] (affine 3-point warp, 256 x 256) through the class surface over CUDA."""
    ref_res, ref = flows.config1(lambda *a: RefHomography(*a), RefImageData)
    got_res, got = flows.config1(lambda *a: hg.Homography(*a, context=ctx), hg.ImageData)
    assert got.last_path == ref.last_path == "inverse_geometric"
    assert (got_res[0].width, got_res[0].height) == (ref_res[0].width, ref_res[0].height) == (256, 205)
    assert _diff(got_res[0].data, ref_res[0].data) == 0


@pytest.mark.parametrize("flow", flows.CSS, ids=lambda f: f.__name__)
def test_css_matrix_strings_over_cuda(ctx, flow):
    """getTransformationMatrixAsCSS (H.js:548): the device-solved matrices print the same strings as the reference's."""
    assert flow(lambda *a: hg.Homography(*a, context=ctx)) == flow(lambda *a: RefHomography(*a))


def test_node_golden_over_cuda(ctx, golden):
    res, _ = flows.node_test(lambda *a: hg.Homography(*a, context=ctx), hg.ImageData(golden["src"].reshape(-1).copy(), 400, 400))
    assert np.array_equal(res[0].as_array(), golden["out"])


def test_unsupported_and_state_errors(ctx):
    c2 = hg.Context(0)
    with pytest.raises(hg.HgError, match="no image set"):
        c2.warp_inverse_matrix(np.array([1, 0, 0, 1, 0, 0], np.float32), 0, 0, 4, 4)
    c2.image_set(_rand_img(1, 8, 8), 8, 8)
    with pytest.raises(hg.HgError) as e:
        c2.warp_inverse_matrix(np.array([1, 0, 0, 1, 0, 0], np.float32), 1 << 20, 0, 4, 4)
    assert e.value.status == hg._abi.HG_ERR_UNSUPPORTED
    with pytest.raises(hg.HgError):
        c2.warp_inverse_matrix(np.array([1, 0, 0, 1, 0, 0], np.float32), 0, 0, 0, 4)
    c2.close()


def test_pipelined_host_stream_matches_oracle(ctx):
    """hg_pipe_*: independent frames, each with its own image and its own points, more frames than slots."""
    rng = np.random.default_rng(71)
    W, H = 200, 150
    pipe = hg.Pipe(ctx, hg._abi.HG_PROJECTIVE, W, H, 400, 300, depth=3)
    n = 8
    imgs = [_rand_img(80 + k, W, H) for k in range(n)]
    outs, wants = [], []
    for k in range(n):
        s = np.array([0, 0, 0, H, W, 0, W, H], np.float64)
        d = s + rng.uniform(-0.15, 0.15, 8) * W
        xo, yo, oW, oH = [int(v) for v in O.transform_limits(O.projective_from_squares(s, d), W, H)]
        oW, oH = min(oW, 400), min(oH, 300)
        out = np.zeros(oW * oH * 4, np.uint8)
        outs.append(out)
        pipe.submit(imgs[k].ctypes.data, d, s, xo, yo, oW, oH, out.ctypes.data)
        wants.append(O.warp_inverse_geometric(imgs[k], W, H, O.projective_from_squares(d, s), xo, yo, oW, oH))
    pipe.flush()
    for k in range(n):
        assert _diff(outs[k], wants[k]) == 0, k
    pipe.close()


# ------------------------------------------------------------------ forward scatter (A3 / A4)
@pytest.mark.parametrize("seed", range(8))
def test_forward_geometric_bit_exact(ctx, seed):
    """_geometricWarp: last-writer-wins collisions (minification), wrapped / dropped writes (Q3)."""
    rng = np.random.default_rng(500 + seed)
    W, H = int(rng.integers(8, 200)), int(rng.integers(8, 200))
    img = _rand_img(seed, W, H)
    ctx.image_set(img, W, H)
    ang = rng.uniform(0, 2 * math.pi)
    sc = [1.0, 0.6, 1.7, 0.35, 1.0, 2.5, 0.9, 1.2][seed]
    if seed in (0, 4):
        fwd = np.array([1, 0, 0, 1, rng.integers(-30, 30), rng.integers(-30, 30)], np.float32)   # pure translation
    else:
        fwd = np.array([math.cos(ang) * sc, math.sin(ang) * sc, -math.sin(ang) * sc, math.cos(ang) * sc,
                        rng.uniform(-40, 40), rng.uniform(-40, 40)], np.float32)
    xo, yo = int(rng.integers(-20, 20)), int(rng.integers(-20, 20))
    oW, oH = int(rng.integers(4, 260)), int(rng.integers(4, 260))
    got = ctx.warp_forward_matrix(fwd, xo, yo, oW, oH)
    want = O.warp_forward_geometric(img, W, H, fwd, xo, yo, oW, oH)
    assert _diff(got, want) == 0


def test_forward_geometric_projective_and_degenerate(ctx):
    img = _rand_img(9, 60, 40)
    ctx.image_set(img, 60, 40)
    for fwd in (np.array([1.1, 0.05, 2.0, -0.03, 0.95, 1.0, 1e-3, -5e-4], np.float64),
                np.full(6, np.nan, np.float32), np.array([1e9, 0, 0, 1, 0, 0], np.float32),
                np.array([1, 0, 0, 1, 1e12, 0], np.float32)):
        got = ctx.warp_forward_matrix(fwd, -5, -5, 90, 70)
        want = O.warp_forward_geometric(img, 60, 40, fwd, -5, -5, 90, 70)
        assert _diff(got, want) == 0


@pytest.mark.parametrize("seed", range(6))
def test_forward_piecewise_bit_exact(ctx, seed):
    rng = np.random.default_rng(550 + seed)
    W, H = 240, 180
    img = _rand_img(70 + seed, W, H)
    src, tris = _grid_mesh(6, 5, W, H)
    if seed % 2:
        src = (src + rng.uniform(-6, 6, src.shape)).astype(np.float32)   # bbox leaves the image: OOB source reads
    dst = (src * rng.uniform(0.85, 1.0) + rng.uniform(-8, 8, src.shape) + 10).astype(np.float32)
    mm = O.minmax_xy(dst)
    xo, yo, oW, oH = int(mm[0]), int(mm[1]), int(mm[2] - mm[0]), int(mm[3] - mm[1])
    smm = [int(v) for v in O.minmax_xy(src)]
    ctx.image_set(img, W, H)
    ctx.piecewise_set_mesh(src, tris)
    got = ctx.warp_piecewise_forward(dst, xo, yo, oW, oH, smm[0], smm[1], smm[2], smm[3])
    fwd = O.piecewise_matrices(src, dst, tris)
    mw = smm[2] - smm[0]
    fmap = O.build_index_map(src, tris, mw, smm[1], mw * (smm[3] - smm[1]))
    want = O.warp_forward_piecewise(img, W, H, fmap, fwd, xo, yo, oW, oH, smm[0], smm[1], smm[2], smm[3])
    assert _diff(got, want) == 0


# ------------------------------------------------------------------ fused (map-free) piecewise path
BINNINGS = {"span": 1, "band": 2}   # hg_debug_piecewise_binning: the two ways the fused path bins triangle rows


def _pw_case(ctx, img, src, dst, tris, force_general=False, binning=None, window=None):
    H, W = img.shape[:2]
    mm = O.minmax_xy(dst)
    xo, yo, oW, oH = window or (int(mm[0]), int(mm[1]), int(mm[2] - mm[0]), int(mm[3] - mm[1]))
    if oW < 1 or oH < 1:
        return
    smm = O.minmax_xy(src)
    ctx.image_set(img, W, H)
    ctx.piecewise_set_mesh(src, tris)
    ctx.debug_force_general(force_general)
    ctx.debug_piecewise_binning(BINNINGS.get(binning, 0))
    try:
        got = ctx.warp_piecewise_inverse(dst, xo, yo, oW, oH, int(smm[0]), int(smm[1]))
    finally:
        ctx.debug_force_general(False)
        ctx.debug_piecewise_binning(0)
    fwd = O.piecewise_matrices(src, dst, tris)
    imap = O.build_index_map(dst, tris, oW, yo, oW * oH)
    want = O.warp_inverse_piecewise(img, W, H, imap, O.inverse_matrices(fwd), xo, yo, oW, oH, int(smm[0]), int(smm[1]), threads=4)
    assert _diff(got, want) == 0, f"{_diff(got, want)} of {oW * oH} pixels differ (force_general={force_general}, binning={binning})"


@pytest.mark.parametrize("seed", range(10))
@pytest.mark.parametrize("path", ["span", "band", "general"])
def test_piecewise_irregular_frames_fused_and_general(ctx, seed, path):
    """Offsets != 0 (Q4: spans spill over row ends), ~~minY < round(minY) (Q5: fills wrap to the END of the map),
    negative coordinates, non-multiple-of-4 widths, folded meshes with many overlaps (bin overflow -> fallback)."""
    rng = np.random.default_rng(700 + seed)
    W, H = 300, 220
    img = _rand_img(90 + seed, W, H)
    src, tris = _grid_mesh(8, 6, W, H)
    dst = src.astype(np.float64)
    kind = seed % 5
    if kind == 0:      # positive offset: xOff > 0 -> every span is shifted and spills
        dst = dst * 1.3 + [37.6, 21.7]
    elif kind == 1:    # negative offset (video-style jitter around the frame)
        dst = dst + rng.uniform(-0.03, 0.03, dst.shape) * [W, H] - [11.3, 7.8]
    elif kind == 2:    # strong fold: triangles overlap heavily
        dst = dst + rng.uniform(-60, 60, dst.shape) + 30
    elif kind == 3:    # fractional minima with frac >= .5 (Q5) and odd output width
        dst = dst * [1.117, 0.93] + [0.6, 0.7]
    else:              # minification far below 1/1.2
        dst = dst * 0.31 + [3.3, 2.2]
    _pw_case(ctx, img, src, dst.astype(np.float32), tris, path == "general", path)


@pytest.mark.parametrize("binning", ["span", "band"])
def test_piecewise_windows_unrelated_to_the_mesh(ctx, binning):
    """The C ABI takes any window: cropped inside the mesh, far to its right / below it (spans land many rows away from
    their own, or nowhere), a one-row and a one-column map, a window whose width is smaller than the mesh offset."""
    rng = np.random.default_rng(4242)
    W, H = 200, 150
    img = _rand_img(17, W, H)
    src, tris = _grid_mesh(7, 6, W, H)
    dst = (src * 1.1 + rng.uniform(-4, 4, src.shape) + [25.3, 11.6]).astype(np.float32)
    f0, g0 = ctx.debug_piecewise_stats()
    for window in ((60, 40, 90, 70), (0, 0, 300, 200), (-40, -30, 120, 100), (150, 100, 64, 300), (30, 50, 1, 80),
                   (30, 50, 130, 1), (200, 12, 17, 160), (26, 12, 7, 5), (-500, 0, 800, 90), (0, -400, 250, 700)):
        _pw_case(ctx, img, src, dst, tris, binning=binning, window=window)
    f1, g1 = ctx.debug_piecewise_stats()
    assert f1 - f0 >= 3, "every window fell back to the map-based path"


def test_piecewise_random_triangle_soup(ctx):
    """Arbitrary (non-mesh) triangles incl. degenerate and repeated ones."""
    rng = np.random.default_rng(808)
    W, H = 160, 120
    img = _rand_img(12, W, H)
    src = rng.uniform(0, [W, H], (14, 2)).astype(np.float32)
    dst = (src + rng.uniform(-25, 25, src.shape) + 12).astype(np.float32)
    tris = rng.integers(0, 14, (20, 3)).astype(np.uint32)
    for fg, binning in ((False, "span"), (False, "band"), (True, None)):
        _pw_case(ctx, img, src, dst, tris, fg, binning)


@pytest.mark.parametrize("binning", ["span", "band"])
def test_config4_mesh_one_frame_4k(ctx, binning):
    """Config 4 mesh (64x64 points, 7,938 triangles) on a 3840x2160 frame, one frame, all pixels."""
    import homography_js_b200 as hgm
    w, h = 3840, 2160
    img = _rand_img(4, w, h)
    src, dst, tris = hgm.workloads.piecewise_sinusoid(64, 64, w, h, phase=0.7)
    f0, g0 = ctx.debug_piecewise_stats()
    _pw_case(ctx, img, src, dst, tris, binning=binning)
    f1, g1 = ctx.debug_piecewise_stats()
    assert (f1 - f0, g1 - g0) == (1, 0), "the frame fell back to the general path"


def test_piecewise_batch_matches_oracle(ctx):
    import homography_js_b200 as hgm
    w, h = 480, 270
    n = 5
    imgs = [_rand_img(30 + k, w, h) for k in range(2)]
    dev_src = [ctx.dev_alloc(w * h * 4) for _ in imgs]
    for p, im in zip(dev_src, imgs):
        ctx.memcpy_h2d(p, im.ctypes.data, im.nbytes)
    src, _, tris = hgm.workloads.piecewise_sinusoid(12, 9, w, h)
    ctx.piecewise_set_mesh(src, tris)
    dsts, frames, outs, shapes = [], [], [], []
    for f in range(n):
        _, dst, _ = hgm.workloads.piecewise_sinusoid(12, 9, w, h, phase=2 * math.pi * f / n)
        if f == 3:
            dst = (dst + np.float32(13.4)).astype(np.float32)   # an irregular frame inside the batch
        xo, yo, oW, oH = hgm.workloads.piecewise_extent(dst)
        p = ctx.dev_alloc(oW * oH * 4)
        dsts.append(dst); outs.append(p); shapes.append((xo, yo, oW, oH))
        frames.append(hg.HgFrame(dev_src[f % 2], p, w, h, xo, yo, oW, oH))
    ctx.warp_piecewise_inverse_batch(np.stack(dsts), frames, 0, 0)
    for f in range(n):
        xo, yo, oW, oH = shapes[f]
        got = np.empty(oW * oH * 4, np.uint8)
        ctx.memcpy_d2h(got.ctypes.data, outs[f], got.nbytes)
        ctx.synchronize()
        fwd = O.piecewise_matrices(src, dsts[f], tris)
        imap = O.build_index_map(dsts[f], tris, oW, yo, oW * oH)
        want = O.warp_inverse_piecewise(imgs[f % 2], w, h, imap, O.inverse_matrices(fwd), xo, yo, oW, oH, 0, 0, threads=4)
        assert _diff(got, want) == 0, f
    for p in dev_src + outs:
        ctx.dev_free(p)


def test_config5_video_stream_variable_windows(ctx):
    """Config 5 shape (scaled down 4x): 30-point mesh, per-frame destiny points, a different output window per
    frame (negative offsets -> Q4 spill on every frame); all frames through one batch call, bit-exact, and the
    fused path must take them (no silent fallback to the map)."""
    import homography_js_b200 as hgm
    w, h, n = 480, 270, 12
    img = _rand_img(5, w, h)
    src, dst, tris = hgm.workloads.video_stream(n, w, h)
    smm = [int(v) for v in O.minmax_xy(src)]
    ctx.image_set(img, w, h)
    ctx.piecewise_set_mesh(src, tris)
    frames, outs, shapes = [], [], []
    for f in range(n):
        xo, yo, oW, oH = hgm.workloads.piecewise_extent(dst[f])
        p = ctx.dev_alloc(oW * oH * 4)
        outs.append(p); shapes.append((xo, yo, oW, oH))
        frames.append(hg.HgFrame(None, p, 0, 0, xo, yo, oW, oH))
    assert len(set(shapes)) > 1
    f0, g0 = ctx.debug_piecewise_stats()
    ctx.warp_piecewise_inverse_batch(dst, frames, smm[0], smm[1])
    f1, g1 = ctx.debug_piecewise_stats()
    assert (f1 - f0, g1 - g0) == (n, 0), "frames fell back to the general path"
    for f in range(n):
        xo, yo, oW, oH = shapes[f]
        got = np.empty(oW * oH * 4, np.uint8)
        ctx.memcpy_d2h(got.ctypes.data, outs[f], got.nbytes)
        ctx.synchronize()
        fwd = O.piecewise_matrices(src, dst[f], tris)
        imap = O.build_index_map(dst[f], tris, oW, yo, oW * oH)
        want = O.warp_inverse_piecewise(img, w, h, imap, O.inverse_matrices(fwd), xo, yo, oW, oH, smm[0], smm[1], threads=4)
        assert _diff(got, want) == 0, f
    for p in outs:
        ctx.dev_free(p)


def test_folded_mesh_overflows_bins_and_falls_back(ctx):
    """More than 8 overlapping spans in one 64-pixel block: the fused path must refuse and the general path answer."""
    rng = np.random.default_rng(99)
    W, H = 128, 96
    img = _rand_img(13, W, H)
    src = rng.uniform(0, [W, H], (40, 2)).astype(np.float32)
    dst = (rng.uniform(0, [W, H], (40, 2)) * 0.9 + 3).astype(np.float32)   # unrelated to src: everything overlaps
    tris = rng.integers(0, 40, (60, 3)).astype(np.uint32)
    f0, g0 = ctx.debug_piecewise_stats()
    _pw_case(ctx, img, src, dst, tris)
    f1, g1 = ctx.debug_piecewise_stats()
    assert g1 - g0 == 1 and f1 == f0


def test_piecewise_extents_on_device(ctx):
    """A9 for piecewise frames: difference of ROUNDED extrema (Q13), incl. .5 ties and negative values."""
    import homography_js_b200 as hgm
    rng = np.random.default_rng(61)
    d = rng.uniform(-50, 2000, (17, 33, 2)).astype(np.float32)
    d[0, 0] = [-2.5, 7.5]      # Math.round ties toward +inf
    d[1, :, 0] = 12.5
    got = ctx.piecewise_extents(d)
    for f in range(d.shape[0]):
        mm = O.minmax_xy(d[f])
        assert list(got[f]) == [mm[0], mm[1], mm[2] - mm[0], mm[3] - mm[1]], f
    _, dst, _ = hgm.workloads.video_stream(5, 480, 270)
    got = ctx.piecewise_extents(dst)
    for f in range(5):
        assert tuple(int(v) for v in got[f]) == hgm.workloads.piecewise_extent(dst[f])


# ------------------------------------------------------------------ doubled-coordinate pixel loop (geo_fast_body) edge cases
def test_fast_body_tiny_images_and_wrap_column(ctx):
    """1- and 2-pixel dimensions (the end-pixel interior test must never apply), and maps whose rounded coordinate
    reaches column W / row H (flat-index wrap and past-the-end reads, Q2) for affine and projective."""
    for W, H in ((1, 1), (1, 7), (9, 1), (2, 2), (3, 2)):
        img = _rand_img(40 + W + H, W, H)
        for inv in (np.array([1, 0, 0, 1, 0, 0], np.float32), np.array([0.25, 0, 0, 0.25, -0.3, -0.2], np.float32),
                    np.array([0.3, 0.1, -0.1, 0.3, 0.49, 0.51], np.float32)):
            _geo_case(ctx, img, inv, -5, -5, 37, 29)
            h8 = np.array([inv[0], inv[2], inv[4], inv[1], inv[3], inv[5], 1e-3, -2e-3], np.float64)
            _geo_case(ctx, img, h8, -5, -5, 37, 29)
    img = _rand_img(50, 40, 30)
    # sx in [W - 0.5, W) and sy in [H - 0.5, H) on whole rows / columns
    _geo_case(ctx, img, np.array([1, 0, 0, 1, 0.75, 0.75], np.float32), -4, -4, 60, 50)
    _geo_case(ctx, img, np.array([1, 0, 0.75, 0, 1, 0.75, 1e-4, 1e-4], np.float64), -4, -4, 60, 50)


def test_fast_body_far_offsets_and_large_coordinates(ctx):
    """Windows far from the origin (|offset| up to 2^18) and steep scales: coordinates beyond the +-2^18 range of the
    doubled fixed-point layout must read as outside, never alias into the image."""
    img = _rand_img(51, 64, 48)
    for xo, yo in ((-(1 << 18), -(1 << 18)), ((1 << 18) - 300, (1 << 18) - 200), (-(1 << 18), 100)):
        for inv in (np.array([1, 0, 0, 1, -xo, -yo], np.float32), np.array([3.5, 0.25, -0.5, 2.75, 7, 9], np.float32)):
            _geo_case(ctx, img, inv, xo, yo, 300, 200)
        h8 = np.array([1, 0, -xo, 0, 1, -yo, 1e-7, -1e-7], np.float64)
        _geo_case(ctx, img, h8, xo, yo, 300, 200)
    # a quad row whose end pixels are inside while the map is steep: 40x minification
    _geo_case(ctx, img, np.array([40, 0, 0, 40, 0, 0], np.float32), -2, -2, 20, 20)
    _geo_case(ctx, img, np.array([0.01, 0, 0, 0.01, 10, 10], np.float32), 0, 0, 500, 300)


def test_fast_body_denominator_range_switches_mode(ctx):
    """Projective frames on both sides of geo_fast_mode's conditions (denominator range / sign over the window) give
    the oracle's bytes: strong perspective, denominators near 1/64 and 64, negative denominators."""
    img = _rand_img(52, 160, 120)
    for h6, h7 in ((2e-3, 1e-3), (-3e-3, 0.0), (0.0, 4e-3), (0.2, 0.0), (-0.0035, -0.0035), (1e-2, -1e-2)):
        inv = np.array([1.1, 0.05, 3.0, -0.04, 0.9, 2.0, h6, h7], np.float64)
        _geo_case(ctx, img, inv, -10, -10, 260, 200)
    # every denominator negative over the window (sign flips both numerators' roles)
    _geo_case(ctx, img, np.array([-1.0, 0, -5.0, 0, -1.0, -4.0, -1e-3, -1e-3], np.float64) * 1.0, 2000, 2000, 200, 150)


def test_staged_tma_kernel_parity():
    """The opt-in TMA-staged kernel (HG_GEO_STAGED=1: tensor maps, producer warp, shared-memory ring) gives the same
    bytes as the oracle on maps that exercise every tile class: interior, image border, wrap column, rotation,
    minification beyond the box budget, horizon."""
    import os
    os.environ["HG_GEO_STAGED"] = "1"
    try:
        c = hg.Context(0)
    finally:
        del os.environ["HG_GEO_STAGED"]
    try:
        for W, H in ((256, 192), (120, 90), (644, 100)):
            img = _rand_img(60 + W, W, H)
            cases = [np.array([1, 0, 0, 1, 3, 2], np.float32), np.array([0.9, 0.2, -0.2, 0.9, 10, -5], np.float32),
                     np.array([0, 1, -1, 0, W, 0], np.float32), np.array([2.5, 0, 0, 2.5, 0, 0], np.float32),
                     np.array([1.05, 0.02, 4, 0.01, 1.2, 3, 4e-4, 2e-4], np.float64),
                     np.array([1.0, 0.1, 5.0, 0.05, 1.0, 3.0, -0.01, 0.002], np.float64)]
            for inv in cases:
                for (xo, yo, oW, oH) in ((-20, -16, W + 60, H + 40), (0, 0, W, H), (5, 7, 129, 65)):
                    c.image_set(img, W, H)
                    got = c.warp_inverse_matrix(inv, xo, yo, oW, oH)
                    want = O.warp_inverse_geometric(img, W, H, inv, xo, yo, oW, oH)
                    assert _diff(got, want) == 0, (W, H, inv.tolist(), xo, yo, oW, oH)
    finally:
        c.close()


def test_node_test_flow_from_png_file_to_png_file(golden):
    """test/nodeTest.js end to end without a canvas: PNG bytes in (hg_png_decode), projective warp on the GPU, PNG bytes
    out (hg_png_encode) — the decoded result is the reference's golden transformedImage.png."""
    src_png = hg._abi.png_encode(np.asarray(golden["src"], np.uint8).reshape(400, 400, 4))
    h = hg.Homography()
    h.setReferencePoints(golden["src_points"].tolist(), golden["dst_points"].tolist())
    h.setImage(hg.ImageData.from_png(src_png))
    out_png = h.warp().to_png()
    got = hg._abi.png_decode(out_png)
    assert got.shape == (200, 400, 4)
    assert np.array_equal(got.reshape(-1), np.asarray(golden["out"]).reshape(-1))
