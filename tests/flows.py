"""Call sequences lifted from the reference's own test pages (test/test.js:34-350, test/nodeTest.js:5-13):
same points, same order of setter calls; the image is passed in its ImageData form.  Each flow takes a
Homography factory `mk(*ctor_args)`, an ImageData factory and the 400x400 RGBA test image, and returns the
warp result(s).  They run against the oracle restatement, against the product over the oracle stub (CPU)
and against the product over CUDA (GPU)."""
import math

import numpy as np

W = H = 400


def node_test(mk, img):  # test/nodeTest.js
    hm = mk()
    hm.setReferencePoints([[0, 0], [0, 1], [1, 0], [1, 1]], [[1 / 10, 1 / 2], [0, 1], [9 / 10, 1 / 2], [1, 1]])
    hm.setImage(img)
    return [hm.warp()], hm


def test1(mk, img):  # piecewise 2x upsample, 9 points (test.js:34)
    sq = [[0, 0], [0, 0.5], [0.5, 0], [0.5, 0.5], [0.5, 1], [1, 0.5], [1, 1], [0, 1], [1, 0]]
    sh = [[0, 0], [0, 1], [1, 0], [1, 1], [1, 2], [2, 1], [2, 2], [0, 2], [2, 0]]
    hm = mk("piecewiseaffine")
    hm.setReferencePoints(sq, sh)
    return [hm.warp(img)], hm


def test2(mk, img):  # piecewise, 4 normalised points (test.js:57)
    hm = mk("piecewiseaffine")
    hm.setReferencePoints([[0, 0], [0, 1], [1, 0], [1, 1]], [[1 / 5, 1 / 5], [0, 1 / 2], [1, 0], [6 / 8, 6 / 8]])
    return [hm.warp(img)], hm


def test3(mk, img):  # mixed ranges + overlap (test.js:83)
    hm = mk("piecewiseaffine")
    hm.setSourcePoints([[0, 0], [0, H], [W, 0], [W, H]], img)
    hm.setDestinyPoints([[0, 0], [0, 1 / 2], [1 / 2, 0], [1 / 6, 1 / 12]])
    return [hm.warp()], hm


def test4(mk, img):  # ctor width/height (test.js:109); ImageData input keeps its own 400x400
    nw, nh = W * 1.5, H / 1.5
    hm = mk("piecewiseaffine", nw, nh)
    hm.setSourcePoints([[0, 0], [0, nh], [nw, 0], [nw, nh]])
    hm.setImage(img)
    hm.setDestinyPoints([[0, 0], [0, nh], [nw, 0], [nw * 3 / 4, nh * 3 / 4]])
    return [hm.warp()], hm


def test5(mk, img):  # sinusoid, 21x11 grid (test.js:133)
    src, dst = [], []
    amplitude, n = 20, 8
    y = 0.0
    while y <= H:
        x = 0.0
        while x <= W:
            src.append([x, y])
            dst.append([x, amplitude + y + math.sin((x * n) / math.pi) * amplitude])
            x += W / 20
        y += H / 10
    hm = mk("piecewiseaffine", W, H)
    hm.setImage(img)
    hm.setSourcePoints(src)
    hm.setDestinyPoints(dst)
    return [hm.warp()], hm


def test6(mk, img):  # affine translation -> forward scatter (test.js:167)
    hm = mk("affine", W, H)
    hm.setSourcePoints([[0, 0], [0, H], [W, 0]])
    hm.setDestinyPoints([[100, 50], [100, H + 50], [W + 100, 50]])
    return [hm.warp(img)], hm


def test7(mk, img):  # affine rotation, normalised (test.js:195)
    hm = mk("affine")
    hm.setSourcePoints([[0, 0], [0, 1], [1, 0]], None)
    hm.setDestinyPoints([[0, 1 / 2], [1 / 2, 1], [1 / 2, 0]])
    return [hm.warp(img)], hm


def test8(mk, img):  # state consistency over repeats (test.js:220), 3 repeats
    hm = mk("affine")
    hm.setSourcePoints([[0, 0], [0, 1], [1, 0]])
    res = []
    for _ in range(3):
        hm.setDestinyPoints([[0, 0], [0, 1 / 1.25], [1 / 1.75, 0]])
        res.append(hm.warp(img))
        hm.setDestinyPoints([[0, 0], [0, 1.25], [1.75, 0]])
        res.append(hm.warp(img))
    return res, hm


def test9(mk, img):  # affine, mixed ranges (test.js:255)
    hm = mk("affine")
    hm.setSourcePoints([[0, 0], [0, 1], [1, 0]])
    hm.setImage(img)
    hm.setDestinyPoints([[0, 0], [W, H], [W, H / 5]])
    return [hm.warp()], hm


def test10(mk, img):  # projective identity (test.js:282)
    sq = [[0, 0], [0, H], [W, 0], [W, H]]
    hm = mk("projective", W, H)
    hm.setSourcePoints(sq)
    hm.setDestinyPoints(sq)
    return [hm.warp(img)], hm


def test11(mk, img):  # projective mirror (test.js:305)
    hm = mk("projective")
    hm.setSourcePoints([[0, 0], [0, 1], [1, 0], [1, 1]], None, W, H)
    hm.setDestinyPoints([[1 - 1 / 8, 0], [1 - 1 / 8, 1], [1 / 8, 0], [1 / 8, 1]])
    return [hm.warp(img)], hm


def test12(mk, img):  # opposite perspective (test.js:328)
    hm = mk("projective")
    hm.setSourcePoints([[0, 0], [0, 1], [1, 2 / 10], [1, 8 / 10]])
    hm.setDestinyPoints([[0, 2 / 10], [0, 8 / 10], [1, 0], [1, 1]])
    return [hm.warp(img)], hm


def test6_inverse(mk, img):  # translation forced through the inverse loop (applyAlwaysInverse)
    hm = mk("affine", W, H)
    hm.setSourcePoints([[0, 0], [0, H], [W, 0]])
    hm.setDestinyPoints([[100, 50], [100, H + 50], [W + 100, 50]])
    return [hm.warp(img, False, True)], hm


def test5_then_forward(mk, img):  # inverse piecewise warp, then a same-size forward one (map aliasing, Q8)
    res, hm = test5(mk, img)
    src = hm._srcPoints.copy()
    dst = src.copy()
    dst[0::2] += 3.0
    hm.setDestinyPoints(dst)
    res.append(hm.warp())
    return res, hm


ALL = [node_test, test1, test2, test3, test4, test5, test6, test7, test8, test9, test10, test11, test12,
       test6_inverse, test5_then_forward]
INVERSE_ONLY = [node_test, test1, test2, test3, test4, test5, test8, test9, test10, test11, test12, test6_inverse]


# ------------------------------------------------------------------ getTransformationMatrixAsCSS flows: mk -> string
sq4 = [[0, 0], [0, 1], [1, 0], [1, 1]]



def css1(mk):  # test/test.js:354-380 (testCSS1): element rect 300 x 150
    return mk("auto").getTransformationMatrixAsCSS([[0, 0], [0, 1], [1, 0]], [[0, 0], [1 / 2, 1], [1, 1 / 8]], 300.0, 150.0)


def css2(mk):  # test/test.js:383-425 (testCSS2): setters first, then the string, then the element's size
    hm = mk("projective")
    hm.setSourcePoints([list(p) for p in sq4])
    hm.setDestinyPoints([[0, 0], [0, 1], [1, 0.1], [1, 1]])
    return hm.getTransformationMatrixAsCSS() + " | " + hm.getTransformationMatrixAsCSS(None, None, 203.4, 150.0)


def bench_affine(mk):  # test/benchmark.js:358-434 shape: pixel-range points
    w = h = 400
    hm = mk("affine")
    hm.setSourcePoints([[0, 0], [0, h], [w, 0]])
    return hm.getTransformationMatrixAsCSS(None, [[0, h / 2], [w / 2, h * 0.8], [w / 2, 0]])


def bench_projective(mk):
    w = h = 400
    hm = mk("projective")
    hm.setSourcePoints([[0, 0], [0, h], [w, 0], [w, h]])
    return hm.getTransformationMatrixAsCSS(None, [[w / 10, 0], [w / 10, h], [w, h / 4], [w, 3 * h / 4]])


def random_ones(mk):
    rng = np.random.default_rng(91)
    out = []
    for k in range(40):
        n = 3 if k % 2 else 4
        src = rng.uniform(0, 900, (n, 2)).tolist()
        dst = rng.uniform(0, 900, (n, 2)).tolist()
        out.append(mk("auto").getTransformationMatrixAsCSS(src, dst, float(rng.integers(50, 900)), float(rng.integers(50, 900))))
    return "\n".join(out)


CSS = [css1, css2, bench_affine, bench_projective, random_ones]


# ------------------------------------------------------------------ BASELINE.json config 1: affine 3-point, 256 x 256
def config1(mk, mk_image):
    """image = default_rng(1) bytes; src [[0,0],[0,256],[256,0]] -> dst [[0,128],[128,204.8],[128,0]] (benchmark.js:204-205
    shape): pixel-range points, output size != input size -> the inverse loop."""
    w = h = 256
    data = np.random.default_rng(1).integers(0, 256, (h, w, 4), dtype=np.uint8).reshape(-1)
    hm = mk("affine")
    hm.setReferencePoints([[0, 0], [0, h], [w, 0]], [[0, h / 2], [w / 2, h * 0.8], [w / 2, 0]])
    return [hm.warp(mk_image(data.copy(), w, h))], hm
