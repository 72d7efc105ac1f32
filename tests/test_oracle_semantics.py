"""JS-number semantics and solver behaviour of the oracle, checked against independent statements
(pure-Python restatements for small cases, numpy.linalg for the solves).  CPU only."""
import math

import numpy as np
import pytest

from oracle import oracle as O


@pytest.mark.parametrize("x,want", [(0.5, 1), (1.5, 2), (2.5, 3), (-0.5, 0), (-1.5, -1), (-2.5, -2), (0.49999999999999994, 0),
                                     (-0.49999999999999994, 0), (2.4, 2), (-2.6, -3), (4503599627370497.0, 4503599627370497.0),
                                     (1e300, 1e300)])
def test_math_round(x, want):
    assert O.js_round(x) == want


def test_math_round_nan_inf():
    assert math.isnan(O.js_round(float("nan")))
    assert O.js_round(float("inf")) == float("inf") and O.js_round(float("-inf")) == float("-inf")


@pytest.mark.parametrize("x,want", [(2.9, 2), (-2.9, -2), (float("nan"), 0), (float("inf"), 0), (4294967296.0 + 5.5, 5),
                                     (2147483648.0, -2147483648), (-2147483649.0, 2147483647)])
def test_toint32(x, want):
    assert O.js_toint32(x) == want


def _py_lu_solve(A, b):
    """numeric.js LU + LUsolve (H.js:1664-1749) in plain Python floats (IEEE double, unfused)."""
    n = len(A)
    A = [list(r) for r in A]
    P = [0] * n
    for k in range(n):
        Pk, mx = k, abs(A[k][k])
        for j in range(k + 1, n):
            if mx < abs(A[j][k]):
                mx, Pk = abs(A[j][k]), j
        P[k] = Pk
        if Pk != k:
            A[k], A[Pk] = A[Pk], A[k]
        for i in range(k + 1, n):
            A[i][k] /= A[k][k]
        for i in range(k + 1, n):
            for j in range(k + 1, n):
                A[i][j] -= A[i][k] * A[k][j]
    x = list(b)
    for i in range(n):
        if P[i] != i:
            x[i], x[P[i]] = x[P[i]], x[i]
        for j in range(i):
            x[i] -= x[j] * A[i][j]
    for i in range(n - 1, -1, -1):
        for j in range(i + 1, n):
            x[i] -= x[j] * A[i][j]
        x[i] /= A[i][i]
    return x


def _py_projective(s, d):
    A = []
    for p in range(4):
        sx, sy, dx, dy = s[2 * p], s[2 * p + 1], d[2 * p], d[2 * p + 1]
        A.append([sx, sy, 1, 0, 0, 0, -dx * sx, -dx * sy])
        A.append([0, 0, 0, sx, sy, 1, -dy * sx, -dy * sy])
    return _py_lu_solve(A, list(d))


def test_projective_solve_matches_python_restatement_bitwise():
    rng = np.random.default_rng(7)
    for _ in range(300):
        s = rng.uniform(0, 2000, 8).astype(np.float32).astype(np.float64)
        d = rng.uniform(0, 2000, 8).astype(np.float32).astype(np.float64)
        got = O.projective_from_squares(s, d)
        want = np.array(_py_projective(list(s), list(d)))
        assert np.array_equal(got, want, equal_nan=True)


def test_projective_solve_maps_points():
    s = np.array([0, 0, 0, 1080, 1920, 0, 1920, 1080], np.float64)
    d = np.array([192, 0, 192, 1080, 1920, 270, 1920, 810], np.float64)
    h = O.projective_from_squares(s, d)
    Hm = np.array([[h[0], h[1], h[2]], [h[3], h[4], h[5]], [h[6], h[7], 1.0]])
    for k in range(4):
        q = Hm @ np.array([s[2 * k], s[2 * k + 1], 1.0])
        assert np.allclose(q[:2] / q[2], d[2 * k:2 * k + 2], atol=1e-8)


def test_affine_solve_matches_python_restatement_bitwise():
    rng = np.random.default_rng(8)
    for _ in range(500):
        s = rng.uniform(-500, 3000, 6).astype(np.float32).astype(np.float64)
        d = rng.uniform(-500, 3000, 6).astype(np.float32).astype(np.float64)
        sE, sF = s[4], s[5]
        sA, sB, sC, sD = s[0] - sE, s[1] - sF, s[2] - sE, s[3] - sF
        dE, dF = d[4], d[5]
        dA, dB, dC, dD = d[0] - dE, d[1] - dF, d[2] - dE, d[3] - dF
        den = sA * sD - sB * sC
        iA, iB, iC, iD = sD / den, sB / -den, sC / -den, sA / den
        iE, iF = (sD * sE - sC * sF) / -den, (sB * sE - sA * sF) / den
        want = np.array([dA * iA + dC * iB, dB * iA + dD * iB, dA * iC + dC * iD, dB * iC + dD * iD,
                         dA * iE + dC * iF + dE, dB * iE + dD * iF + dF]).astype(np.float32)
        assert np.array_equal(O.affine_from_triangles(s, d), want)


def test_inverse_warp_matches_python_loop_small():
    """_inverseGeometricWarp restated as the literal double loop in Python, 37x23 image, both kinds."""
    rng = np.random.default_rng(3)
    W, H = 37, 23
    img = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
    flat = img.reshape(-1)
    cases = [np.array([0.9, 0.1, -0.2, 1.1, 3.0, -2.0], np.float32),
             np.array([1.01, 0.02, -3.0, 0.03, 0.98, 2.0, 1e-4, -2e-4], np.float64)]
    for m in cases:
        xo, yo, oW, oH = -3, -4, 45, 31
        want = np.zeros(oW * oH * 4, np.uint8)
        for y in range(yo, yo + oH):
            for x in range(xo, xo + oW):
                if m.size == 6:
                    mm = [float(v) for v in m]
                    sx = mm[0] * x + mm[2] * y + mm[4]
                    sy = mm[1] * x + mm[3] * y + mm[5]
                else:
                    den = m[6] * x + m[7] * y + 1
                    sx = (m[0] * x + m[1] * y + m[2]) / den
                    sy = (m[3] * x + m[4] * y + m[5]) / den
                if 0 <= sx < W and 0 <= sy < H:
                    si = int(math.floor(sy + 0.5)) * W * 4 + int(math.floor(sx + 0.5)) * 4
                    di = ((y - yo) * oW + (x - xo)) * 4
                    for c in range(4):
                        want[di + c] = flat[si + c] if si + c < flat.size else 0
        got = O.warp_inverse_geometric(img, W, H, m, xo, yo, oW, oH)
        assert np.array_equal(got, want)


def _py_fill(tri, idx, mw, yoff, arr):
    """fillTriangle (H.js:1111) in plain Python with TypedArray.fill semantics."""
    x0, y0, x1, y1, x2, y2 = [float(np.float32(v)) for v in tri]
    minY = int(math.trunc(min(y0, y1, y2)))
    maxY = math.ceil(max(y0, y1, y2))

    def seg(xa, ya, xb, yb):
        if xb != xa:
            m = (yb - ya) / (xb - xa)
            return m, ya - xa * m, min(ya, yb), max(ya, yb)
        return math.inf, xa, min(ya, yb), max(ya, yb)

    segs = [seg(x0, y0, x1, y1), seg(x0, y0, x2, y2), seg(x1, y1, x2, y2)]
    n = len(arr)

    def bound(v):
        if v != v:
            v = 0
        if v == math.inf:
            return n
        if v == -math.inf:
            return 0
        v = math.trunc(v)
        return max(n + v, 0) if v < 0 else min(v, n)

    def jsround(v):
        if v in (math.inf, -math.inf) or v != v:
            return v
        return math.floor(v + 0.5)

    for y in range(minY, maxY):
        mn, mx = math.inf, -math.inf
        for m, b, lo, hi in segs:
            if lo <= y <= hi:
                if m == math.inf:
                    x = b
                elif m == 0:
                    continue
                else:
                    x = (y - b) / m
                mn, mx = min(mn, x), max(mx, x)
        k0, k1 = bound((y - yoff) * mw + jsround(mn)), bound((y - yoff) * mw + jsround(mx))
        for k in range(k0, k1):
            arr[k] = idx


def test_index_map_matches_python_restatement_including_quirks():
    rng = np.random.default_rng(11)
    for trial in range(40):
        n_pts = 7
        pts = rng.uniform(-4, 40, (n_pts, 2)).astype(np.float32)
        if trial % 3 == 0:
            pts = np.round(pts)  # integer vertices: horizontal / vertical edges, shared rows
        tris = np.array([[0, 1, 2], [2, 3, 4], [4, 5, 6], [0, 3, 6], [1, 4, 5]], np.uint32)
        mw = 30 + trial % 5
        yoff = float(O.js_round(float(pts[:, 1].min())))   # Q5: ~~minY can be yoff-1 -> negative relative index
        length = mw * 36
        want = [-1] * length
        for t, tri in enumerate(tris):
            _py_fill(pts[tri].reshape(-1), t, mw, yoff, want)
        got = O.build_index_map(pts, tris, mw, yoff, length)
        assert np.array_equal(got, np.array(want, np.int16)), trial


def test_index_map_last_triangle_wins_and_x_offset_is_ignored():
    # two identical triangles: the later id wins everywhere (Q7)
    pts = np.array([[2, 1], [12, 1], [2, 9]], np.float32)
    m = O.build_index_map(pts, np.array([0, 1, 2, 0, 1, 2], np.uint32), 16, 1, 16 * 10)
    assert set(np.unique(m)) == {-1, 1}
    # Q4: columns are absolute x, no x offset is subtracted -> span starts at column round(xmin) = 2
    row1 = m[16:32]
    assert row1[2] == 1 and row1[1] == -1


def test_division_free_decision_algorithm_against_rational_arithmetic():
    """The algorithm behind the CUDA path's exact resolution of projective pixels on a decision boundary
    (quotient_at_least, csrc/warp_geo.cuh) restated with exact-rational fmas: RN(N / D) >= b  <=>  the sign of
    fma(h, D, fma(-b, D, N)), h = half the gap below b.  (The CUDA function itself is checked on the same operands by
    tests/test_gpu_numerics.py::test_division_free_exact_decision_against_rational_arithmetic.)"""
    import struct
    from fractions import Fraction

    def fma(a, b, c):
        return float(Fraction(a) * Fraction(b) + Fraction(c))  # one correctly rounded operation

    def decide(N, D, b):
        bits = struct.unpack("<q", struct.pack("<d", b))[0]
        E = (bits >> 52) & 0x7FF
        pow2_pos = bits > 0 and (bits & 0xFFFFFFFFFFFFF) == 0
        h = struct.unpack("<d", struct.pack("<q", (E - 53 - (1 if pow2_pos else 0)) << 52))[0]
        t = fma(h, D, fma(-b, D, N))
        return t >= 0 if D > 0 else t <= 0

    rng = np.random.default_rng(9)
    n_checked = 0
    for b in [v * 0.5 for v in range(-12, 13) if v] + [512.0, 1024.0, -2048.0, 4096.5, 131072.0, -131071.5, 262143.5]:
        for _ in range(25):
            D = float(rng.uniform(1 / 64, 64)) * (1 if rng.random() < 0.5 else -1)
            if rng.random() < 0.3:
                D = float(np.float32(D))
            base = float(Fraction(b) * Fraction(D))
            for k in range(-3, 4):
                N = base
                for _ in range(abs(k)):
                    N = float(np.nextafter(N, np.inf if k > 0 else -np.inf))
                assert decide(N, D, b) == (float(Fraction(N) / Fraction(D)) >= b), (N, D, b)
                n_checked += 1
    assert n_checked > 5000


# ------------------------------------------------------------------ getTransformationMatrixAsCSS (H.js:548-586)
# Number.prototype.toFixed known answers (ECMA-262: exact binary value, ties to the larger n, sign from x < 0)
TO_FIXED_KATS = [(1.0, "1.00000"), (0.125, "0.12500"), (0.015625, "0.01563"), (-0.015625, "-0.01563"), (2.5, "2.50000"),
                 (1.005, "1.00500"), (0.000005, "0.00001"), (-0.0, "0.00000"), (-1e-7, "-0.00000"), (1e21, "1e+21"),
                 (123456.7891251, "123456.78913"), (1e-300, "0.00000"), (float("nan"), "NaN"), (float("inf"), "Infinity"),
                 (float("-inf"), "-Infinity"), (999999.999995, "999999.99999"), (0.1 + 0.2, "0.30000")]


@pytest.mark.parametrize("x,want", TO_FIXED_KATS)
def test_to_fixed_oracle_and_host(x, want):
    from oracle.homography_ref import js_to_fixed
    from homography_js_b200.homography import _js_to_fixed   # host string code of the product: no GPU involved
    assert js_to_fixed(x, 5) == want
    assert _js_to_fixed(x, 5) == want


def test_to_fixed_oracle_equals_host_on_random_doubles():
    from oracle.homography_ref import js_to_fixed
    from homography_js_b200.homography import _js_to_fixed
    rng = np.random.default_rng(77)
    xs = np.concatenate([rng.uniform(-3, 3, 2000), rng.uniform(-5000, 5000, 1000), rng.uniform(-1e-4, 1e-4, 500),
                         np.arange(-64, 64) / 64.0 + 2.0 ** -6,      # exactly representable ...5 ties
                         rng.uniform(-3, 3, 500).astype(np.float32).astype(np.float64)])
    for x in xs:
        for d in (0, 2, 5):
            assert js_to_fixed(float(x), d) == _js_to_fixed(float(x), d), (x, d)


def test_css_matrix_strings_known_answers(oracle_lib):
    from oracle.homography_ref import RefHomography
    # test/test.js:354-380 (testCSS1): affine from normalised points, element 300 x 150 -> matrix stays in the
    # normalised space (Q11); columns of [[1, .5], [.125, 1]]
    h = RefHomography("auto")
    assert h.getTransformationMatrixAsCSS([[0, 0], [0, 1], [1, 0]], [[0, 0], [1 / 2, 1], [1, 1 / 8]], 300.0, 150.0) == \
        "matrix(1.00000, 0.12500, 0.50000, 1.00000, 0.00000, 0.00000)"
    # identity projective: the 4x4 identity with toFixed(5) on the eight solved entries
    h = RefHomography("projective")
    sq = [[0, 0], [0, 1], [1, 0], [1, 1]]
    h.setSourcePoints(sq)
    h.setDestinyPoints([list(p) for p in sq])
    assert h.getTransformationMatrixAsCSS() == \
        "matrix3d(1.00000, 0.00000, 0, 0.00000, 0.00000, 1.00000, 0, 0.00000, 0, 0, 1, 0, 0.00000, 0.00000, 0, 1)"
    # test/test.js:383-425 (testCSS2): cells are h0 h3 0 h6 | h1 h4 0 h7 | 0 0 1 0 | h2 h5 0 1
    h = RefHomography("projective")
    h.setSourcePoints(sq)
    h.setDestinyPoints([[0, 0], [0, 1], [1, 0.1], [1, 1]])
    css = h.getTransformationMatrixAsCSS()
    m = h._transformMatrix
    cells = css[len("matrix3d("):-1].split(", ")
    assert len(cells) == 16 and [cells[i] for i in (2, 6, 8, 9, 10, 11, 14, 15)] == ["0", "0", "0", "0", "1", "0", "0", "1"]
    for cell, k in zip([cells[i] for i in (0, 1, 3, 4, 5, 7, 12, 13)], (0, 3, 6, 1, 4, 7, 2, 5)):
        assert abs(float(cell) - m[k]) <= 0.5e-5 + 1e-12
    # piecewise has no CSS form; missing points raise the reference's texts
    with pytest.raises(ValueError, match="srcPoints are not set"):
        RefHomography("affine").getTransformationMatrixAsCSS()
    h = RefHomography("piecewiseaffine")
    with pytest.raises(ValueError, match="Only \"affine\" or \"projective\""):
        h.setSourcePoints([[0, 0], [0, 1], [1, 0], [1, 1], [0.5, 0.5]], None, 10, 10)
        h._dstPoints = h._srcPoints.copy()
        h._transformMatrix = np.zeros(6, np.float32)
        h.getTransformationMatrixAsCSS()


# ------------------------------------------------------------------ the other three loops, restated literally in Python
def _js_round_py(v):
    if v != v or v in (math.inf, -math.inf):
        return v
    return float(math.floor(v + 0.5))


def _to_int32(v):
    if v != v or v in (math.inf, -math.inf):
        return 0
    v = int(math.trunc(v)) & 0xFFFFFFFF
    return v - (1 << 32) if v >= (1 << 31) else v


def _shl2(v):  # JS `v << 2`
    r = (_to_int32(v) << 2) & 0xFFFFFFFF
    return r - (1 << 32) if r >= (1 << 31) else r


def _typed_get(arr, i):  # typed-array read: a non-integer / out-of-range index gives undefined -> stored as 0
    return int(arr[int(i)]) if (0 <= i < len(arr) and i == math.floor(i)) else 0


def _typed_set(arr, i, v):  # typed-array write: silently dropped unless the index is an in-range integer
    if 0 <= i < len(arr) and i == math.floor(i):   # NaN and +-Inf fail the range test
        arr[int(i)] = v


def _py_forward(flat, W, H, point_fn, xo, yo, oW, oH, domain):
    """H.js:911-932 / 948-972: for every (x, y) of `domain` (raster order) scatter src pixel to round(T(x, y) - offset)."""
    dst_row = _shl2(oW)
    out = [0] * int(dst_row * oH)
    src_row = _shl2(W)
    for x, y in domain:
        p = point_fn(x, y)
        if p is None:
            continue
        idx = y * src_row + _shl2(x)
        nx, ny = _js_round_py(p[0] - xo), _js_round_py(p[1] - yo)
        nidx = ny * dst_row + _shl2(nx)
        for c in range(4):
            _typed_set(out, nidx + c, _typed_get(flat, idx + c))
    return np.array(out, np.uint8)


def test_forward_geometric_matches_python_loop_small():
    """_geometricWarp restated as the literal double loop: collisions (last writer wins), writes wrapped into neighbouring
    rows and dropped outside the buffer (Q3), NaN targets."""
    rng = np.random.default_rng(5)
    W, H = 19, 13
    img = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
    flat = img.reshape(-1)
    dom = [(x, y) for y in range(H) for x in range(W)]
    cases = [(np.array([1, 0, 0, 1, 3, 2], np.float32), (3, 2, W, H)),                # translation: a copy
             (np.array([0.6, 0.1, -0.2, 0.7, 1.5, 0.5], np.float32), (0, 0, 15, 11)),  # shrink: many collisions
             (np.array([1.3, 0, 0, 1.2, -4, -3], np.float32), (0, 0, W, H)),           # negative / overflowing columns wrap
             (np.array([0, 1, -1, 0, 12, 0], np.float32), (0, 0, H, W)),               # quarter turn
             (np.array([1.0, 0.02, -3.0, 0.03, 0.98, 2.0, 1e-3, -2e-3], np.float64), (-2, 1, 21, 14)),
             (np.array([1, 0, 0, 0, 1, 0, 0, -1.0], np.float64), (0, 0, W, H))]        # denominator 0 on row y = 1: Inf / NaN
    for m, (xo, yo, oW, oH) in cases:
        mm = [float(v) for v in m]
        if m.size == 6:
            fn = lambda x, y: (mm[0] * x + mm[2] * y + mm[4], mm[1] * x + mm[3] * y + mm[5])
        else:
            def fn(x, y):
                den = mm[6] * x + mm[7] * y + 1
                with np.errstate(all="ignore"):
                    return (float(np.float64(mm[0] * x + mm[1] * y + mm[2]) / np.float64(den)),
                            float(np.float64(mm[3] * x + mm[4] * y + mm[5]) / np.float64(den)))
        want = _py_forward(flat, W, H, fn, xo, yo, oW, oH, dom)
        got = O.warp_forward_geometric(img, W, H, m, xo, yo, oW, oH)
        assert np.array_equal(got, want), m


def test_piecewise_loops_match_python_restatement_small():
    """_inversePiecewiseAffineWarp (H.js:1029-1058) and _piecewiseAffineWarp (H.js:948-972) as literal Python loops over the
    oracle's own map and matrices (both pinned separately above): window test on [minSrc, W + minSrc), Int16 ids, the flat
    source index, reads past the image, forward collisions."""
    rng = np.random.default_rng(8)
    W, H = 24, 18
    img = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
    flat = img.reshape(-1)
    src = np.array([[0, 0], [W, 0], [0, H], [W, H], [W / 2, H / 2]], np.float32)
    tris = np.array([[0, 1, 4], [1, 3, 4], [3, 2, 4], [2, 0, 4]], np.uint32)
    for trial in range(4):
        dst = (src.astype(np.float64) * [1.4, 1.25] + rng.uniform(-2, 2, src.shape) + [3, 1]).astype(np.float32)
        fwd = O.piecewise_matrices(src, dst, tris)
        inv = O.inverse_matrices(fwd)
        mm = O.minmax_xy(dst)
        xo, yo, oW, oH = int(mm[0]), int(mm[1]), int(mm[2] - mm[0]), int(mm[3] - mm[1])
        min_sx, min_sy = (0, 0) if trial % 2 == 0 else (2, 1)       # Q14: the window moves with minSrc
        imap = O.build_index_map(dst, tris, oW, yo, oW * oH)
        want = [0] * (oW * oH * 4)
        for y in range(yo, yo + oH):
            for x in range(xo, xo + oW):
                t = int(imap[(y - yo) * oW + (x - xo)])
                if t >= 0:
                    m = [float(v) for v in inv[t]]
                    sx, sy = m[0] * x + m[2] * y + m[4], m[1] * x + m[3] * y + m[5]
                    if min_sx <= sx < W + min_sx and min_sy <= sy < H + min_sy:
                        si = _js_round_py(sy) * (W * 4) + _js_round_py(sx) * 4
                        di = ((y - yo) * oW + (x - xo)) * 4
                        for c in range(4):
                            want[di + c] = _typed_get(flat, si + c)
        got = O.warp_inverse_piecewise(img, W, H, imap, inv, xo, yo, oW, oH, min_sx, min_sy)
        assert np.array_equal(got, np.array(want, np.uint8)), trial
        # forward loop over the source-point bounding box, through the forward map
        bx0, by0, bx1, by1 = 0, 0, W, H
        fmap = O.build_index_map(src, tris, bx1 - bx0, by0, (bx1 - bx0) * (by1 - by0))

        def fn(x, y):
            t = int(fmap[(y - by0) * (bx1 - bx0) + (x - bx0)])
            if t <= -1:
                return None
            m = [float(v) for v in fwd[t]]
            return m[0] * x + m[2] * y + m[4], m[1] * x + m[3] * y + m[5]

        dom = [(x, y) for y in range(by0, by1) for x in range(bx0, bx1)]
        want_f = _py_forward(flat, W, H, fn, xo, yo, oW, oH, dom)
        got_f = O.warp_forward_piecewise(img, W, H, fmap, fwd, xo, yo, oW, oH, bx0, by0, bx1, by1)
        assert np.array_equal(got_f, want_f), trial
