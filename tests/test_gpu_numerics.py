"""Numerical guarantees the fast projective path relies on, measured on the device."""
import numpy as np
import pytest

import homography_js_b200 as hg
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def test_newton_reciprocal_error_bound_exhaustive(ctx):
    """warp_geo.cuh trusts quotients unless they are within 2^-20 of a multiple of 0.5; for coordinates below
    2^16 that needs a reciprocal with relative error < 2^-36.  Exhaustive over the 2^20 mantissa patterns the
    MUFU reads, for exponents across the guarded range [2^-500, 2^500] and both signs."""
    worst = 0.0
    for e in (1023, 1022, 1024, 1023 - 499, 1023 + 499, 900, 1100, 1, 2046 - 1):
        for neg in (False, True):
            if e in (1, 2045):
                continue  # outside the guarded range: the kernel never uses the fast path there
            err = ctx.debug_rcp_max_error(e, neg)
            assert err == err, "NaN residual"
            worst = max(worst, err)
    print(f"max |1 - d*rcp| = {worst:.3e} = 2^{np.log2(worst):.1f}")
    assert worst < 2.0 ** -38


@pytest.mark.parametrize("seed", range(6))
def test_projective_quotients_on_decision_boundaries(ctx, seed):
    """Matrices built so that MANY quotients land exactly on k and k+0.5 with a non-trivial denominator
    (denominator = 2^-j or 3/2^j on whole rows): every such pixel must take the exact-division path."""
    rng = np.random.default_rng(900 + seed)
    W, H = 96, 64
    img = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
    ctx.image_set(img, W, H)
    d0 = [0.5, 0.25, 0.75, 1.5, 2.0, 0.375][seed]
    # x' = (a x + c) / (g y + d0), y' = (b y + f) / (g y + d0) scaled so the "+1" of the reference form holds
    g = [0.0, 1 / 64, -1 / 128, 1 / 32, 0.0, 1 / 16][seed]
    h = np.array([0.5 / d0, 0, 1.0 / d0, 0, 0.25 / d0, 0.5 / d0, 0, g / d0], np.float64)
    # reference form has denominator h6 x + h7 y + 1: divide everything by d0 already done above
    got = ctx.warp_inverse_matrix(h, -4, -4, 220, 150)
    want = O.warp_inverse_geometric(img, W, H, h, -4, -4, 220, 150)
    assert np.array_equal(got, want)


def test_many_random_projective_matrices_small(ctx):
    """Differential test over 150 random homographies (incl. strong perspective) on a small image."""
    rng = np.random.default_rng(77)
    W, H = 53, 41
    img = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
    ctx.image_set(img, W, H)
    for k in range(150):
        s = np.array([0, 0, 0, H, W, 0, W, H], np.float64)
        d = s + rng.uniform(-0.45, 0.45, 8) * max(W, H)
        inv = O.projective_from_squares(d, s)
        oW, oH = int(rng.integers(1, 90)), int(rng.integers(1, 70))
        xo, yo = int(rng.integers(-20, 10)), int(rng.integers(-20, 10))
        got = ctx.warp_inverse_matrix(inv, xo, yo, oW, oH)
        want = O.warp_inverse_geometric(img, W, H, inv, xo, yo, oW, oH)
        assert np.array_equal(got, want), k


@pytest.mark.parametrize("seed", range(5))
def test_projective_x_only_denominator_mode(ctx, seed):
    """h7 == 0 (denominator a function of x only) takes the per-column reciprocal path; includes columns where
    the denominator crosses zero (horizon inside the window) and -0.0 as h7."""
    rng = np.random.default_rng(600 + seed)
    W, H = 120, 90
    img = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
    ctx.image_set(img, W, H)
    h6 = [1e-3, -2e-3, 1 / 64, -1 / 50, 3e-4][seed]
    h = np.array([rng.uniform(0.5, 1.5), rng.uniform(-0.2, 0.2), rng.uniform(-10, 10),
                  rng.uniform(-0.2, 0.2), rng.uniform(0.5, 1.5), rng.uniform(-10, 10), h6, -0.0 if seed % 2 else 0.0])
    got = ctx.warp_inverse_matrix(h, -30, -20, 260, 170)
    want = O.warp_inverse_geometric(img, W, H, h, -30, -20, 260, 170)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("kind", ["affine", "projective"])
def test_bilinear_extension_within_one_lsb(ctx, kind):
    """EXTENSION (not in the reference): bilinear sampling, <= 1 LSB per channel against the oracle's definition,
    and exactly the nearest-neighbour result when every coordinate is an integer."""
    rng = np.random.default_rng(321)
    W, H = 180, 130
    img = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
    ctx.image_set(img, W, H)
    ctx.set_sampling(hg._abi.HG_BILINEAR)
    try:
        for trial in range(6):
            if kind == "affine":
                inv = np.array([rng.uniform(0.4, 2.2), rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), rng.uniform(0.4, 2.2),
                                rng.uniform(-20, 20), rng.uniform(-20, 20)], np.float32)
            else:
                s = np.array([0, 0, 0, H, W, 0, W, H], np.float64)
                inv = O.projective_from_squares(s + rng.uniform(-0.2, 0.2, 8) * W, s)
            got = ctx.warp_inverse_matrix(inv, -10, -10, 230, 170).astype(np.int16)
            want = O.warp_inverse_geometric_bilinear(img, W, H, inv, -10, -10, 230, 170).astype(np.int16)
            diff = np.abs(got - want)
            assert diff.max() <= 1, diff.max()
            assert (diff > 0).mean() < 0.02   # differences only at rounding ties
        ident = np.array([1, 0, 0, 1, 3, 2], np.float32) if kind == "affine" else np.array([1, 0, 3, 0, 1, 2, 0, 0.0])
        got = ctx.warp_inverse_matrix(ident, 0, 0, 100, 90)
        ctx.set_sampling(hg._abi.HG_NEAREST)
        assert np.array_equal(got, ctx.warp_inverse_matrix(ident, 0, 0, 100, 90))
    finally:
        ctx.set_sampling(hg._abi.HG_NEAREST)


def test_division_free_exact_decision_against_rational_arithmetic(ctx):
    """quotient_at_least: RN(N / D) >= b decided with two fmas instead of the division.  Checked against exact
    rational arithmetic (float(Fraction) is the correctly rounded quotient) on the cases that matter: N within a few ulp of
    b * D — including exact ties at the midpoint below b —, b a power of two (the gap below it is halved), negative b,
    negative D, plus random far-away operands."""
    from fractions import Fraction
    rng = np.random.default_rng(9)
    Ns, Ds, bs = [], [], []
    halves = np.concatenate([np.arange(-40, 41) * 0.5, [512.0, 1024.0, -2048.0, 4096.5, 65535.5, 131072.0, -131071.5, 262143.5, 0.5, -0.5]])
    halves = halves[halves != 0.0]
    for b in halves:
        for _ in range(60):
            D = float(rng.uniform(1 / 64, 64)) * (1 if rng.random() < 0.5 else -1)
            if rng.random() < 0.2:
                D = float(np.float32(D))  # short mantissas make b * D exactly representable: true ties
            base = float(Fraction(b) * Fraction(D))  # RN(b * D)
            for k in (-3, -2, -1, 0, 1, 2, 3):
                N = base
                for _ in range(abs(k)):
                    N = float(np.nextafter(N, np.inf if k > 0 else -np.inf))
                Ns.append(N); Ds.append(D); bs.append(float(b))
            # the exact midpoint between pred(b) and b, times D, when representable
            pb = float(np.nextafter(b, -np.inf))
            mid = (Fraction(b) + Fraction(pb)) / 2 * Fraction(D)
            if Fraction(float(mid)) == mid:
                Ns.append(float(mid)); Ds.append(D); bs.append(float(b))
    for _ in range(4000):
        Ns.append(float(rng.uniform(-1e6, 1e6))); Ds.append(float(rng.uniform(0.02, 60)) * (1 if rng.random() < 0.5 else -1))
        bs.append(float(rng.integers(-500000, 500000)) * 0.5 or 0.5)
    got = ctx.debug_quotient_at_least(Ns, Ds, bs)
    want = np.array([float(Fraction(n) / Fraction(d)) >= b for n, d, b in zip(Ns, Ds, bs)])
    bad = np.nonzero(got != want)[0]
    assert bad.size == 0, [(Ns[i], Ds[i], bs[i], bool(got[i]), bool(want[i])) for i in bad[:5]]
    assert want.sum() > 1000 and (~want).sum() > 1000


def _bilinear_check(ctx, img, W, H, inv, window, max_mismatch=0.02):
    got = ctx.warp_inverse_matrix(inv, *window).astype(np.int16)
    want = O.warp_inverse_geometric_bilinear(img, W, H, inv, *window).astype(np.int16)
    diff = np.abs(got - want)
    assert diff.max() <= 1, (diff.max(), int((diff > 1).sum()))
    if diff.size >= 4000:   # rounding ties are rare, not impossible: only a statistic over many values means something
        assert (diff > 0).mean() <= max_mismatch
    return want


def test_bilinear_window_bounds_stay_exact_under_the_fast_reciprocal(ctx):
    """Second-generation bilinear kernel: projective frames take the reciprocal from MUFU + one Newton step; the window test
    of H.js:1001 must still be the reference's.  Matrices built so that the quotient is mathematically W (or H) for EVERY
    pixel — numerator = W * denominator up to rounding — which puts the reference's RN(N / D) on either side of the bound
    pixel by pixel: one wrong decision shows as a whole transparent / opaque pixel, far more than 1 LSB."""
    rng = np.random.default_rng(4242)
    W, H = 180, 130
    img = rng.integers(1, 256, (H, W, 4), dtype=np.uint8)   # no zero bytes: transparent vs sampled always differs
    ctx.image_set(img, W, H)
    ctx.set_sampling(hg._abi.HG_BILINEAR)
    try:
        seen_in = seen_out = 0
        for h6, h7 in ((1e-3, 2e-3), (-7e-4, 1.3e-3), (3.3e-3, -1e-3), (1 / 1024, 1 / 2048)):
            for bound_on in ("x", "y"):
                if bound_on == "x":
                    inv = np.array([W * h6, W * h7, float(W), 0.37 * h6, 1.0 + 0.37 * h7, 2.37, h6, h7])
                else:
                    inv = np.array([1.0 + 0.61 * h6, 0.61 * h7, 3.61, H * h6, H * h7, float(H), h6, h7])
                want = _bilinear_check(ctx, img, W, H, inv, (-5, -4, 150, 100), max_mismatch=0.05)
                opaque = (want.reshape(-1, 4)[:, 3] != 0).mean()
                seen_in += opaque > 0
                seen_out += opaque < 1
                # the same with the quotient sitting on 0 from both sides: numerators of a few ulp
                tiny = np.array([1e-17, -2e-17, 1e-16, 0.0, 1.0, 2.0, h6, h7]) if bound_on == "x" else \
                    np.array([1.0, 0.0, 3.0, -1e-17, 3e-17, -1e-16, h6, h7])
                _bilinear_check(ctx, img, W, H, tiny, (-5, -4, 150, 100), max_mismatch=0.05)
        assert seen_in and seen_out   # both sides of the bound actually occurred
    finally:
        ctx.set_sampling(hg._abi.HG_NEAREST)


@pytest.mark.parametrize("kind", ["affine", "projective"])
def test_bilinear_narrow_and_odd_output_widths(ctx, kind):
    """Flat quads run over row ends when oW % 4 != 0, and span several rows when oW < 4."""
    rng = np.random.default_rng(99)
    W, H = 64, 48
    img = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
    ctx.image_set(img, W, H)
    ctx.set_sampling(hg._abi.HG_BILINEAR)
    try:
        for o_w in (1, 2, 3, 4, 5, 6, 7, 9, 13, 66):
            for o_h in (1, 3, 50):
                if kind == "affine":
                    inv = np.array([0.83, 0.11, -0.07, 0.91, 1.3, 0.6], np.float32)
                else:
                    inv = np.array([0.83, -0.07, 1.3, 0.11, 0.91, 0.6, 1.5e-3, -0.8e-3])
                _bilinear_check(ctx, img, W, H, inv, (-2, -1, o_w, o_h), max_mismatch=0.05)
    finally:
        ctx.set_sampling(hg._abi.HG_NEAREST)


def test_bilinear_batch_with_mixed_frame_sizes(ctx):
    """hg_warp_inverse_batch under HG_BILINEAR: frames of different sizes (one narrower than a quad) in one launch."""
    import torch
    rng = np.random.default_rng(7)
    W, H = 96, 70
    dev = torch.device("cuda", 0)
    imgs = [rng.integers(0, 256, (H, W, 4), dtype=np.uint8) for _ in range(3)]
    windows = [(-3, -2, 120, 80), (0, 0, 3, 40), (5, 7, 57, 33)]
    mats = np.stack([O.projective_from_squares(np.array([0, 0, 0, H, W, 0, W, H], np.float64) + rng.uniform(-6, 6, 8),
                                               np.array([0, 0, 0, H, W, 0, W, H], np.float64)) for _ in range(3)])
    srcs = [torch.from_numpy(i.reshape(-1).copy()).to(dev) for i in imgs]
    outs = [torch.zeros(w[2] * w[3] * 4, dtype=torch.uint8, device=dev) for w in windows]
    torch.cuda.synchronize()
    frames = [hg.HgFrame(s.data_ptr(), o.data_ptr(), W, H, *w) for s, o, w in zip(srcs, outs, windows)]
    ctx.set_sampling(hg._abi.HG_BILINEAR)
    try:
        ctx.warp_inverse_batch(1, mats, frames)
        ctx.synchronize()
    finally:
        ctx.set_sampling(hg._abi.HG_NEAREST)
    for img, o, w, m in zip(imgs, outs, windows, mats):
        want = O.warp_inverse_geometric_bilinear(img, W, H, m, *w).astype(np.int16)
        diff = np.abs(o.cpu().numpy().astype(np.int16) - want)
        assert diff.max() <= 1
