"""homography.js_b200/js/homography_b200.mjs rendered into Python STATEMENT BY STATEMENT (line numbers of the .mjs in the
comments), so that the shim's own control flow — which differs from homography.py in places: every solve goes through
native.solveWithLimits, _induceObjective() re-solves, conversions through f64() / f32() — can be executed in an image that
has no JavaScript engine.  `native` is tests/napi_mock.Native: the real addon under the miniature Node-API runtime.

This file is a test aid and only as good as its fidelity to the .mjs; tests/test_js_binding_sources.py checks that every
native.* call and every public method of the .mjs appears here too.  JS semantics kept: null vs numbers in comparisons
(null <= 0 is true, null > 0 false), Math.round (ties up; Math.round(null) === 0), typed arrays as numpy arrays mutated in
place, bare-string throws as JsThrow."""
import math
from types import SimpleNamespace

import numpy as np

NORMALIZED_MAX = 8.0
MAX_CSS_DECIMAL = 5
KIND = {"affine": 0, "projective": 1}


class JsThrow(Exception):
    """`throw ('text')` of the shim."""


def _round(v):  # Math.round
    if v is None:
        return 0.0
    v = float(v)
    if v != v or v in (math.inf, -math.inf):
        return v
    return float(math.floor(v + 0.5))


def positive(v):  # (v) => v !== null && v > 0
    return v is not None and v > 0


def _le0(v):  # JS `v <= 0`: null coerces to 0
    return True if v is None else v <= 0


def _num(v):  # ToNumber for arithmetic: null -> 0
    return 0.0 if v is None else float(v)


def is_typed(a):
    return isinstance(a, np.ndarray)


def as_point_array(points):  # mjs:20
    return points if is_typed(points) else np.array(points, dtype=np.float64).reshape(-1).astype(np.float32)


def any_above(arr, limit):  # mjs:23
    return bool(np.any(np.asarray(arr, dtype=np.float64) > limit))


def scale_in_place(p, sx, sy, divide):  # mjs:27
    for i in range(p.size):
        s = sx if (i & 1) == 0 else sy
        p[i] = (float(p[i]) / s) if divide else (float(p[i]) * s)


def min_max_xy(p):  # mjs:33
    mnx = mny = math.inf
    mxx = mxy = -math.inf
    for i in range(p.size):
        v = float(p[i])
        if (i & 1) == 0:
            if v > mxx:
                mxx = v
            if v < mnx:
                mnx = v
        else:
            if v > mxy:
                mxy = v
            if v < mny:
                mny = v
    return mnx, mny, mxx, mxy


def select_transform(first, points):  # mjs:42
    n = points.size
    if first == "auto":
        if n == 6:
            return "affine"
        if n == 8:
            return "projective"
        if n > 8:
            return "piecewiseaffine"
        raise JsThrow(f"Transforms must contain at least 3 points but only {n / 2:g} were given")
    if first == "piecewiseaffine":
        if n < 6:
            raise JsThrow(f"A piecewise (or affine) transform needs to determine least three reference points but only {n / 2:g} were given")
        return first
    if first == "affine":
        if n != 6:
            raise JsThrow(f"An affine transform needs to determine exactly three reference points but {n / 2:g} were given")
        return first
    if first == "projective":
        if n != 8:
            raise JsThrow(f"A projective transform needs to determine exactly four reference points but {n / 2:g} were given")
        return first
    raise JsThrow(f'Transform "{first}" is unknown')


def f64(a):
    return a if (is_typed(a) and a.dtype == np.float64) else np.asarray(a, dtype=np.float64).copy()


def f32(a):
    return a if (is_typed(a) and a.dtype == np.float32) else np.asarray(a, dtype=np.float32).copy()


def _to_fixed(v, d):
    from homography_js_b200.homography import _js_to_fixed   # Number.prototype.toFixed (tested on its own)
    return _js_to_fixed(float(v), d)


class Homography:
    def __init__(self, native, transform="auto", width=None, height=None, device=0, triangulate=None):  # mjs:67
        self.native = native
        self._ctx = native.createContext(device)
        self._triangulate = triangulate
        self._width = None if width is None else _round(width)
        self._height = None if height is None else _round(height)
        self._objectiveWidth = self._objectiveHeight = None
        self._xOutputOffset = self._yOutputOffset = None
        self._srcPoints = self._dstPoints = None
        self.firstTransformSelected = self.transform = transform.lower()
        self._image = None
        self._minSrcX = self._minSrcY = self._maxSrcX = self._maxSrcY = None
        self._srcPointsAreNormalized = self._dstPointsAreNormalized = True
        self._mapState = None
        self._triangles = self._initialTriangles = None
        self._transformMatrix = self._piecewiseMatrices = None
        self._meshOnDevice = False
        self.last_path = None   # test aid (not in the .mjs): which loop warp() dispatched

    def setReferencePoints(self, srcPoints, dstPoints, image=None, width=None, height=None, srcNorm=None, dstNorm=None):  # mjs:88
        if srcPoints is None or dstPoints is None:   # typeof ... === 'undefined'
            raise JsThrow("Source and Destiny points must be defined when calling setReferencePoints().")
        self._dstPoints = None
        self.setSourcePoints(srcPoints, image, width, height, srcNorm)
        self.setDestinyPoints(dstPoints, dstNorm)

    def setSourcePoints(self, points, image=None, width=None, height=None, pointsAreNormalized=None):  # mjs:96
        pts = as_point_array(points)
        self._srcPoints = pts
        self._meshOnDevice = False
        self._srcPointsAreNormalized = (not any_above(pts, NORMALIZED_MAX)) if pointsAreNormalized is None else pointsAreNormalized
        self._transformMatrix = None
        self.transform = select_transform(self.firstTransformSelected, pts)
        self._objectiveWidth = self._objectiveHeight = None
        if image is not None:
            self.setImage(image, width, height)
        elif width is not None or height is not None:
            self._setSrcWidthHeight(width, height)
        if self._width is not None and self._height is not None and self._srcPointsAreNormalized:
            self._denormalizeSrc()
        if self._dstPoints is not None and self.transform != "piecewiseaffine":
            self._transformMatrix = self._solve(self._srcPoints, self._dstPoints)["matrix"]
        if self.transform == "piecewiseaffine" and self._mapState is None:
            self._triangles = self._initialTriangles
            self._piecewiseMatrices = None
            if (not self._srcPointsAreNormalized) or (positive(self._width) and positive(self._height)):
                self._setPiecewiseParameters()
            elif self._triangles is None:
                self._triangles = self._delaunay(self._srcPoints)

    def setImage(self, image, width=None, height=None):  # mjs:117
        if image is None or not is_typed(getattr(image, "data", None)):
            raise JsThrow("setImage() needs an ImageData-like object ({data, width, height})")
        self._image = image.data
        self.native.setImage(self._ctx, _clamped(image.data), image.width, image.height)
        self._setSrcWidthHeight(image.width, image.height)
        if self._srcPoints is not None and self.transform == "piecewiseaffine":
            self._setPiecewiseParameters()
        if self._dstPoints is not None and (_le0(self._objectiveWidth) or _le0(self._objectiveHeight)):
            self._induceObjective()

    def setDestinyPoints(self, points, pointsAreNormalized=None):  # mjs:126
        pts = as_point_array(points)
        if self._srcPoints is not None and pts.size != self._srcPoints.size:
            raise JsThrow(f"It must be the same amount of destiny points ({pts.size / 2:g}) than source points ({self._srcPoints.size / 2:g})")
        self._dstPoints = pts
        self._dstPointsAreNormalized = (not any_above(pts, NORMALIZED_MAX)) if pointsAreNormalized is None else pointsAreNormalized
        haveSize = positive(self._width) and positive(self._height)
        limitsDone = False
        if self.transform != "piecewiseaffine":
            if self._dstPointsAreNormalized and haveSize and self.transform == "projective":
                self._denormalizeDst()
            self._sameRange()
            r = self._solve(self._srcPoints, self._dstPoints, self._image is not None)
            self._transformMatrix = r["matrix"]
            if self._image is not None:
                self._setLimits(r["limits"])
                limitsDone = True
        else:
            self._piecewiseMatrices = None
        if not limitsDone and (self._image is not None or (self.transform == "piecewiseaffine" and haveSize)):
            self._induceObjective()
        if self.transform == "piecewiseaffine" and haveSize:
            if self._dstPointsAreNormalized:
                self._denormalizeDst()
            self._setPiecewiseParameters()

    def setTriangles(self, triangles):  # mjs:150
        self._triangles = triangles
        self._meshOnDevice = False
        if ((not self._srcPointsAreNormalized) or (positive(self._width) and positive(self._height))) and self._srcPoints is not None:
            self._setPiecewiseParameters()

    def warp(self, image=None, asHTMLPromise=False, applyAlwaysInverse=False):  # mjs:157
        if asHTMLPromise:
            raise JsThrow("asHTMLPromise needs a DOM; only ImageData results exist outside a browser")
        if image is not None:
            self.setImage(image)
        elif self._image is None:
            raise JsThrow("warp() must receive an image if it was not setted before through `setImage(img)` or  `setSourcePoints(points, img)`")
        oW, oH, W, H = (_num(v) for v in (self._objectiveWidth, self._objectiveHeight, self._width, self._height))
        area = oW * oH
        empty = (not (area >= 1)) or math.isnan(area)
        if self.transform == "piecewiseaffine":
            inverse = applyAlwaysInverse or (oW > W or oH > H or oW * 1.2 < W or oH * 1.2 < H)
            data = self._inversePiecewise(empty) if inverse else self._forwardPiecewise(empty)
        elif self.transform == "affine":
            inverse = applyAlwaysInverse or (oW != W or oH != H)
            data = self._inverseGeometric(empty) if inverse else self._forwardGeometric(empty)
        else:
            data = self._inverseGeometric(empty)
        if empty:
            return SimpleNamespace(data=np.zeros(4, np.uint8), width=1, height=1)
        return SimpleNamespace(data=data, width=int(oW), height=int(oH))

    def getTransformationMatrixAsCSS(self, srcPoints=None, dstPoints=None, width=None, height=None):  # mjs:181
        if width is not None or height is not None:
            self._setSrcWidthHeight(width, height)
        if srcPoints is not None:
            self.setSourcePoints(srcPoints, None, width, height)
        if dstPoints is not None:
            self.setDestinyPoints(dstPoints)
        if self._srcPoints is None:
            raise JsThrow("Impossible to calculate a transform when srcPoints are not set")
        elif self._dstPoints is None:
            raise JsThrow("Impossible to calculate a transform when dstPoints are not set")
        elif self._transformMatrix is None:
            raise JsThrow("Transform matrix can not be calculated")
        m, D = self._transformMatrix, MAX_CSS_DECIMAL
        if self.transform == "affine":
            return "matrix(" + ", ".join(_to_fixed(v, D) for v in m) + ")"
        if self.transform == "projective":
            cells, i = [], 0
            for dy in range(4):
                for dx in range(4):
                    if (dy == 2 and dx == 2) or (dy == 3 and dx == 3):
                        cells.append("1")
                    elif dy == 2 or dx == 2:
                        cells.append("0")
                    else:
                        cells.append(_to_fixed(m[(i * 3) % 8], D))
                        i += 1
            return "matrix3d(" + ", ".join(cells) + ")"
        raise JsThrow(f'Only "affine" or "projective" transforms can be applied on the CSS transform property, but {self.transform} selected')

    def transformHTMLElement(self, element, srcPoints=None, dstPoints=None):  # mjs:206
        rect = element.getBoundingClientRect()
        element.style.transform = self.getTransformationMatrixAsCSS(srcPoints, dstPoints, rect.width, rect.height)

    # ---------------------------------------------------------------- state plumbing
    def _solve(self, src, dst, withLimits=True):  # mjs:212
        kind = KIND.get(self.transform)
        if kind is None:
            raise JsThrow(f"{self.transform} transform does not exist")
        return self.native.solveWithLimits(self._ctx, kind, f64(src), f64(dst), self._width if withLimits else 1,
                                           self._height if withLimits else 1)

    def _setLimits(self, l):  # mjs:217
        self._xOutputOffset, self._yOutputOffset, self._objectiveWidth, self._objectiveHeight = (float(v) for v in l)

    def _denormalizeSrc(self):  # mjs:218
        scale_in_place(self._srcPoints, self._width, self._height, False)
        self._srcPointsAreNormalized = False
        self._meshOnDevice = False

    def _denormalizeDst(self):  # mjs:219
        scale_in_place(self._dstPoints, self._width, self._height, False)
        self._dstPointsAreNormalized = False

    def _delaunay(self, points):  # mjs:220
        return self._triangulate(points) if self._triangulate else self.native.delaunay(points)

    def _setSrcWidthHeight(self, width, height):  # mjs:226
        changed = self._width != width or self._height != height
        self._width, self._height = width, height
        if not changed:
            return
        self._width, self._height = _round(width), _round(height)
        self._mapState = None
        if self.transform == "projective":
            if self._srcPoints is not None and self._srcPointsAreNormalized:
                self._denormalizeSrc()
            if self._dstPoints is not None and self._dstPointsAreNormalized:
                self._denormalizeDst()
            if self._dstPoints is not None and self._srcPoints is not None:
                r = self._solve(self._srcPoints, self._dstPoints)
                self._transformMatrix = r["matrix"]
                self._setLimits(r["limits"])
        if self._srcPoints is not None and self.transform == "piecewiseaffine":
            self._setPiecewiseParameters()

    def _induceObjective(self):  # mjs:243
        if self.transform in ("affine", "projective"):
            if self._transformMatrix is None and self._srcPointsAreNormalized != self._dstPointsAreNormalized:
                self._sameRange()
            r = self._solve(self._srcPoints, self._dstPoints)
            if self._transformMatrix is None:
                self._transformMatrix = r["matrix"]
            self._setLimits(r["limits"])
        elif not self._dstPointsAreNormalized:
            a, b, c, d = min_max_xy(self._dstPoints)
            self._xOutputOffset, self._yOutputOffset = _round(a), _round(b)
            self._objectiveWidth = _round(c) - self._xOutputOffset
            self._objectiveHeight = _round(d) - self._yOutputOffset
        elif positive(self._width) and positive(self._height):
            a, b, c, d = min_max_xy(self._dstPoints)
            self._xOutputOffset, self._yOutputOffset = _round(a), _round(b)
            self._objectiveWidth = _round((c - a) * self._width)
            self._objectiveHeight = _round((d - b) * self._height)
        else:
            raise JsThrow("Trying to calculate a the output width and height of a Piecewise Affine transform but source width and height are not set")

    def _setPiecewiseParameters(self):  # mjs:264
        if self._srcPoints is None:
            raise JsThrow("Trying to set the Piecewise Affine Transform parameters before setting the Source Points.")
        if self._triangles is None:
            self._triangles = self._delaunay(self._srcPoints)
            self._meshOnDevice = False
        if self._srcPointsAreNormalized:
            if positive(self._width) and positive(self._height):
                self._denormalizeSrc()
            else:
                raise JsThrow("Trying to set the Piecewise Affine Transform parameters without knowing the source points ranges")
        if self._mapState is None:
            a, b, c, d = min_max_xy(self._srcPoints)
            self._minSrcX, self._minSrcY, self._maxSrcX, self._maxSrcY = _round(a), _round(b), _round(c), _round(d)
            self._mapState = "forward"
        if self._dstPoints is not None and self._piecewiseMatrices is None and self._triangles is not None:
            if self._dstPointsAreNormalized:
                self._denormalizeDst()
            if self._srcPointsAreNormalized != self._dstPointsAreNormalized:
                self._sameRange()
            self._uploadMesh()
            self._piecewiseMatrices = self.native.piecewiseMatrices(self._ctx, f32(self._dstPoints), len(self._triangles) / 3)

    def _uploadMesh(self):  # mjs:283
        if not self._meshOnDevice:
            self.native.setMesh(self._ctx, f32(self._srcPoints), np.asarray(self._triangles, dtype=np.uint32).copy())
            self._meshOnDevice = True

    def _sameRange(self):  # mjs:287
        if self._dstPointsAreNormalized == self._srcPointsAreNormalized:
            return
        haveSize = positive(self._width) and positive(self._height)
        if self._dstPointsAreNormalized and haveSize:
            scale_in_place(self._srcPoints, self._width, self._height, True)
            self._srcPointsAreNormalized = True
            self._meshOnDevice = False
        elif self._srcPointsAreNormalized and haveSize:
            self._denormalizeSrc()
        else:
            raise JsThrow("Impossible to put source and destiny points in the same range. Possible solutions: \n"
                          "1. Give a source width/height when calling setSrcPoints.\n2. Set the input image before.\n"
                          "3. Give Source and Destiny points in the same range (both normalized or both in image dimensions)")

    # ---------------------------------------------------------------- the four loops -> device
    def _inverseGeometric(self, empty):  # mjs:302
        self.last_path = "inverse_geometric"
        self._sameRange()
        if empty:
            return None
        return self.native.warpInversePoints(self._ctx, KIND[self.transform], f64(self._dstPoints), f64(self._srcPoints),
                                             self._xOutputOffset, self._yOutputOffset, self._objectiveWidth, self._objectiveHeight)

    def _forwardGeometric(self, empty):  # mjs:308
        self.last_path = "forward_geometric"
        if empty:
            return None
        return self.native.warpForwardMatrix(self._ctx, KIND[self.transform], self._transformMatrix, self._xOutputOffset,
                                             self._yOutputOffset, self._objectiveWidth, self._objectiveHeight)

    def _inversePiecewise(self, empty):  # mjs:313
        self.last_path = "inverse_piecewise"
        self._mapState = "inverse"
        if empty:
            return None
        self._uploadMesh()
        return self.native.warpPiecewiseInverse(self._ctx, f32(self._dstPoints), self._xOutputOffset, self._yOutputOffset,
                                                self._objectiveWidth, self._objectiveHeight, self._minSrcX, self._minSrcY)

    def _forwardPiecewise(self, empty):  # mjs:320
        self.last_path = "forward_piecewise"
        if empty:
            return None
        self._uploadMesh()
        return self.native.warpPiecewiseForward(self._ctx, f32(self._dstPoints), self._xOutputOffset, self._yOutputOffset,
                                                self._objectiveWidth, self._objectiveHeight, self._minSrcX, self._minSrcY,
                                                self._maxSrcX, self._maxSrcY, 1 if self._mapState == "inverse" else 0)


def _clamped(data):
    import napi_mock
    return np.ascontiguousarray(data, dtype=np.uint8).reshape(-1).view(napi_mock.Clamped)
