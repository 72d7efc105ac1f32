"""Error budget of the projective pixel loop (csrc/warp_geo.cuh, geo_fast_issue), re-derived on the CPU in exact arithmetic.

The loop does not compute the reference's quotient RN(N / D) (H.js:1401-1404); it computes an APPROXIMATE doubled quotient

    ax = fma(2h0, x0, 2h2)          once per thread (x0 = first column of the thread's quad)
    nx = fma(2h1, y, ax)            once per row, then nx += 2h0 per pixel            (same for ny and the denominator dn)
    rc = 1 / dn  within 2^-39.9     (MUFU.RCP64H + one Newton step: tests/test_gpu_numerics.py measures that bound exhaustively)
    T  = fma(nx, rc, magic)

and decides exactly (queue + quotient_at_least) whenever T lies within delta = 2^-19 of a decision boundary.  That is sound iff
|nx * rc - 2 * RN(N / D)| < delta for every pixel of every frame geo_fast_mode admits.  This test replays the arithmetic with
correctly rounded fmas built from fractions.Fraction (no GPU, no product code) on random frames that pass geo_fast_mode and
checks the bound with the reciprocal at both ends of its error interval."""
import math
from fractions import Fraction as Fr

import numpy as np
import pytest

DELTA = 2.0 ** -19
RCP_REL = 2.0 ** -39.9


def fma(a, b, c):
    return float(Fr(a) * Fr(b) + Fr(c))   # one rounding, to nearest even


def solve_h(src, dst):
    A, b = [], []
    for (x, y), (u, v) in zip(src, dst):
        A.append([x, y, 1, 0, 0, 0, -u * x, -u * y]); b.append(u)
        A.append([0, 0, 0, x, y, 1, -v * x, -v * y]); b.append(v)
    return np.linalg.solve(np.array(A, float), np.array(b, float))


def fast_mode(m, xo, yo, ow, oh):
    """geo_fast_mode of warp_geo.cuh, statement by statement."""
    if m[6] == 0.0 and m[7] == 0.0:
        return False
    X0, X1, Y0, Y1 = float(xo - 3), float(xo + ow + 3), float(yo), float(yo + oh + 15)
    dmin, dmax, pos, neg = 1e300, 0.0, True, True
    for c in range(4):
        dn = m[6] * (X1 if c & 1 else X0) + m[7] * (Y1 if c & 2 else Y0) + 1.0
        pos, neg = pos and dn > 0.0, neg and dn < 0.0
        dmin, dmax = min(dmin, abs(dn)), max(dmax, abs(dn))
    if not (pos or neg) or not dmin >= 0.015625 or not dmax <= 64.0:
        return False
    Xm, Ym, big = max(abs(X0), abs(X1)), max(abs(Y0), abs(Y1)), 16777216.0 * dmin
    return (abs(m[0]) * Xm + abs(m[1]) * Ym + abs(m[2]) < big and abs(m[3]) * Xm + abs(m[4]) * Ym + abs(m[5]) < big and
            abs(m[6]) * Xm + abs(m[7]) * Ym + 1.0 < 256.0 * dmin)


def frames(rng, n):
    out = []
    while len(out) < n:
        w, h = int(rng.integers(64, 4000)), int(rng.integers(64, 2200))
        src = [(0, 0), (0, h), (w, 0), (w, h)]
        dst = [(x + rng.uniform(-0.3, 0.3) * w, y + rng.uniform(-0.3, 0.3) * h) for x, y in src]
        try:
            m = solve_h(dst, src)   # the inverse map: output -> source
        except np.linalg.LinAlgError:
            continue
        if not np.all(np.isfinite(m)):
            continue
        xs, ys = [p[0] for p in dst], [p[1] for p in dst]
        xo, yo = int(math.floor(min(xs))), int(math.floor(min(ys)))
        ow, oh = int(math.ceil(max(xs))) - xo, int(math.ceil(max(ys))) - yo
        if rng.random() < 0.3:   # far offsets: the coordinates the bound is tightest for
            shift = int(rng.integers(-(1 << 17), 1 << 17))
            m = solve_h([(x + shift, y) for x, y in dst], src)
            xo += shift
        if ow >= 8 and oh >= 4 and fast_mode(m, xo, yo, ow, oh):
            out.append((m, xo, yo, ow, oh))
    return out


@pytest.mark.parametrize("seed", range(4))
def test_doubled_quotient_of_the_loop_stays_within_delta_of_the_reference_quotient(seed):
    rng = np.random.default_rng(1900 + seed)
    worst = 0.0
    for m, xo, yo, ow, oh in frames(rng, 40):
        h = [float(v) for v in m]
        c = [2.0 * h[0], 2.0 * h[1], 2.0 * h[2], 2.0 * h[3], 2.0 * h[4], 2.0 * h[5], h[6], h[7]]   # exact scalings
        for _ in range(12):
            x0 = float(xo + int(rng.integers(-3, ow)))       # first column of a quad (may start left of the window)
            y = float(yo + int(rng.integers(0, oh + 15)))
            ax, ay, ad = fma(c[0], x0, c[2]), fma(c[3], x0, c[5]), fma(c[6], x0, 1.0)
            nx, ny, dn = fma(c[1], y, ax), fma(c[4], y, ay), fma(c[7], y, ad)
            for k in range(4):
                x = x0 + k
                # the reference's own numerators, denominator and quotients: every product and sum rounded (H.js:1401-1404)
                Nx = (h[0] * x + h[1] * y) + h[2]
                Ny = (h[3] * x + h[4] * y) + h[5]
                D = (h[6] * x + h[7] * y) + 1.0
                qx, qy = Nx / D, Ny / D
                for num, q in ((nx, qx), (ny, qy)):
                    for e in (-RCP_REL, RCP_REL):
                        approx = Fr(num) / Fr(dn) * (1 + Fr(e))     # nx * rc with rc at either end of its interval
                        err = abs(float(approx - 2 * Fr(q)))
                        # + the rounding of T = fma(nx, rc, magic) onto the 2^-32 grid of [2^20, 2^21)
                        worst = max(worst, err + 2.0 ** -33)
                if k < 3:
                    nx, ny, dn = nx + c[0], ny + c[3], dn + c[6]
    assert worst < DELTA, worst   # (with the reciprocal at 2^-30 instead of 2^-39.9 this fails at 2^-16.8)
    assert worst < 2.0 ** -20.5      # the margin DESIGN.md states (2^-20.9 at |2v| < 2^20) with room for this sample
