"""CPU emulation of the second-generation bilinear kernel's per-pixel algorithm (csrc/bilinear.cuh:
warp_inverse_geo_bilinear2_kernel), statement by statement in numpy, against the oracle's definition.  It checks the
DESIGN on the build machine (no GPU there): flat quad -> row / column with the row wrap, the approximate reciprocal with
the exact re-computation of coordinates next to a window bound, the PRMT byte -> float conversion, the float blend and the
magic-add rounding.  The CUDA kernel itself is checked on the device by tests/test_gpu_numerics.py."""
import numpy as np

from oracle import oracle as O

MAGIC = 1572864.0
HI_ZERO = 0x41380000
NEAR = 4096


def _byte_perm(x, y, sel):
    """CUDA __byte_perm: result byte i = byte (nibble i of sel) of the 8-byte pool {x: 0-3, y: 4-7}."""
    x = np.asarray(x, np.uint32)
    y = np.broadcast_to(np.asarray(y, np.uint32), x.shape)
    pool = [(x >> (8 * i)) & 0xFF for i in range(4)] + [(y >> (8 * i)) & 0xFF for i in range(4)]
    out = np.zeros_like(x)
    for i in range(4):
        out |= pool[(sel >> (4 * i)) & 7].astype(np.uint32) << np.uint32(8 * i)
    return out


def _bilerp_px2(p00, p10, p01, p11, fx, fy):
    f32 = np.float32
    BIAS, ROUND = f32(8388608.0), f32(12582912.0)
    r = []
    for c in range(4):
        sel = 0x7440 | c
        a, b, d, e = (_byte_perm(p, 0x4B000000, sel).view(np.float32) for p in (p00, p10, p01, p11))
        fma = lambda u, v, w: (u.astype(np.float64) * v.astype(np.float64) + w.astype(np.float64)).astype(np.float32)
        top = fma(fx, (b - a).astype(f32), (a - BIAS).astype(f32))
        bot = fma(fx, (e - d).astype(f32), (d - BIAS).astype(f32))
        v = fma(fy, (bot - top).astype(f32), top)
        r.append((v + ROUND).astype(f32).view(np.uint32))
    lo, hi = _byte_perm(r[0], r[1], 0x0040), _byte_perm(r[2], r[3], 0x0040)
    return _byte_perm(lo, hi, 0x5410)


def _split(t):
    bits = np.asarray(t, np.float64).view(np.uint64)
    return (bits >> np.uint64(32)).astype(np.uint32), (bits & np.uint64(0xFFFFFFFF)).astype(np.uint32)


def _add_rd(v, c):
    """RD(v + c) for c = MAGIC: the sum lands on the 2^-32 grid of [2^20, 2^21); emulate round-down from the RN sum."""
    s = v + c
    return np.where(s - c > v, np.nextafter(s, -np.inf), s)


def emulate(img, W, H, inv, x_off, y_off, o_w, o_h, projective):
    assert o_w >= 4
    src = np.ascontiguousarray(img).reshape(-1, 4).view(np.uint32).reshape(-1)
    m = np.asarray(inv, np.float64)
    npix = o_w * o_h
    nquad = (npix + 3) // 4
    q = np.arange(nquad, dtype=np.int64)
    p0 = q * 4
    yy, xx = p0 // o_w, p0 % o_w
    out = np.zeros(nquad * 4, np.uint32)
    exact_count = 0
    for k in range(4):
        xk = xx + k
        wrap = xk >= o_w
        xk = np.where(wrap, xk - o_w, xk)
        x = (x_off + xk).astype(np.float64)
        y = (y_off + yy + wrap).astype(np.float64)
        if not projective:
            sx = (m[0] * x + m[2] * y) + m[4]
            sy = (m[1] * x + m[3] * y) + m[5]
            tx, ty = _add_rd(sx, MAGIC), _add_rd(sy, MAGIC)
        else:
            ld = np.longdouble
            r0, r1, r2 = ld(m[1]) * y + ld(m[2]), ld(m[4]) * y + ld(m[5]), ld(m[7]) * y + ld(1.0)   # per-row fma terms
            rc = (ld(1.0) / (ld(m[6]) * x + r2).astype(np.float64)).astype(np.float64)
            tx = ((ld(m[0]) * x + r0).astype(np.float64).astype(ld) * rc + ld(MAGIC)).astype(np.float64)
            ty = ((ld(m[3]) * x + r1).astype(np.float64).astype(ld) * rc + ld(MAGIC)).astype(np.float64)
            hx, lx = _split(tx)
            hy, ly = _split(ty)
            fxi, fyi = hx - np.uint32(HI_ZERO), hy - np.uint32(HI_ZERO)
            near_x = (lx + np.uint32(NEAR)) < np.uint32(2 * NEAR)
            near_y = (ly + np.uint32(NEAR)) < np.uint32(2 * NEAR)
            crit_x = ((fxi + np.uint32(1)) < 2) | ((fxi + np.uint32(1) - np.uint32(W)) < 2)
            crit_y = ((fyi + np.uint32(1)) < 2) | ((fyi + np.uint32(1) - np.uint32(H)) < 2)
            exact = (near_x & crit_x) | (near_y & crit_y)
            exact_count += int(exact.sum())
            with np.errstate(all="ignore"):
                dne = ((m[6] * x + m[7] * y) + 1.0)
                ex = _add_rd(((m[0] * x + m[1] * y) + m[2]) / dne, MAGIC)
                ey = _add_rd(((m[3] * x + m[4] * y) + m[5]) / dne, MAGIC)
            tx, ty = np.where(exact, ex, tx), np.where(exact, ey, ty)
        hx, lx = _split(tx)
        hy, ly = _split(ty)
        ux, uy = hx - np.uint32(HI_ZERO), hy - np.uint32(HI_ZERO)
        ok = (ux < W) & (uy < H) & (p0 + k < npix)
        fx = (lx.astype(np.float32) * np.float32(2.3283064365386963e-10)).astype(np.float32)
        fy = (ly.astype(np.float32) * np.float32(2.3283064365386963e-10)).astype(np.float32)
        uxc, uyc = np.where(ok, ux, 0).astype(np.int64), np.where(ok, uy, 0).astype(np.int64)
        x1, y1 = np.minimum(uxc + 1, W - 1), np.minimum(uyc + 1, H - 1)
        g = lambda r, c: np.where(ok, src[r * W + c], 0).astype(np.uint32)
        out[k::4] = _bilerp_px2(g(uyc, uxc), g(uyc, x1), g(y1, uxc), g(y1, x1), fx, fy)
    return out[:npix].view(np.uint8), exact_count


def _check(img, W, H, inv, window, projective, max_mismatch=0.02):
    got, n_exact = emulate(img, W, H, inv, *window, projective)
    want = O.warp_inverse_geometric_bilinear(img, W, H, np.asarray(inv, np.float64 if projective else np.float32), *window)
    diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
    assert diff.max() <= 1, (diff.max(), int((diff > 1).sum()))
    assert (diff > 0).mean() <= max_mismatch
    return n_exact, want


def test_emulated_kernel_generic_frames(oracle_lib):
    rng = np.random.default_rng(321)
    W, H = 180, 130
    img = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
    for trial in range(4):
        a = np.array([rng.uniform(0.4, 2.2), rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), rng.uniform(0.4, 2.2),
                      rng.uniform(-20, 20), rng.uniform(-20, 20)], np.float32)
        _check(img, W, H, a.astype(np.float64), (-10, -10, 230 + trial, 170), False)   # odd widths: quads wrap rows
        s = np.array([0, 0, 0, H, W, 0, W, H], np.float64)
        p = O.projective_from_squares(s + rng.uniform(-0.2, 0.2, 8) * W, s)
        _check(img, W, H, p, (-10, -10, 230 + trial, 170), True)
    ident = np.array([1, 0, 3, 0, 1, 2, 0.0, 1e-3])
    _, want = _check(img, W, H, ident, (0, 0, 100, 1), True)    # row y = 0: denominator 1, integer coordinates
    assert np.array_equal(want.reshape(-1, 4), img[2, 3:103].reshape(-1, 4))


def test_emulated_kernel_keeps_window_bounds_exact(oracle_lib):
    rng = np.random.default_rng(4242)
    W, H = 180, 130
    img = rng.integers(1, 256, (H, W, 4), dtype=np.uint8)
    total_exact = 0
    for h6, h7 in ((1e-3, 2e-3), (-7e-4, 1.3e-3)):
        on_x = np.array([W * h6, W * h7, float(W), 0.37 * h6, 1.0 + 0.37 * h7, 2.37, h6, h7])
        on_y = np.array([1.0 + 0.61 * h6, 0.61 * h7, 3.61, H * h6, H * h7, float(H), h6, h7])
        tiny = np.array([1e-17, -2e-17, 1e-16, 0.0, 1.0, 2.0, h6, h7])
        for inv in (on_x, on_y, tiny):
            n, want = _check(img, W, H, inv, (-5, -4, 150, 100), True, max_mismatch=0.05)
            total_exact += n
    assert total_exact > 10000   # the exact re-computation really is what decides these frames
