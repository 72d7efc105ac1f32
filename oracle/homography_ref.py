"""Literal CPU restatement of the reference's `class Homography` (H.js:38-1078).

TEST INFRASTRUCTURE ONLY (see oracle/hg_oracle.c header).  The state machine below follows the
reference statement by statement (line numbers cite /root/reference/Homography.js) and delegates
every arithmetic loop to the C oracle, so that `tests/` can compare the product's CUDA-backed
`Homography` with the reference's behaviour at the class surface.

Not restated here: the DOM branches (hidden canvas, HTMLImageElement in/out, `element.style`).  The CSS matrix
string (getTransformationMatrixAsCSS, H.js:548-586) is pure string work on the solved matrix and IS restated.  The
third-party `delaunator` triangulation (H.js:27, 1216) is restated in oracle/delaunator_ref.py from
the package's published algorithm — no reference fixture pins it (parity unpinned at that boundary);
pass triangles through setTriangles (H.js:517) to fix a mesh.
"""
from __future__ import annotations

import math

import numpy as np

from . import oracle as O

NORMALIZED_MAX = 8.0  # H.js:36
DIMS = 2              # H.js:34
MAX_CSS_DECIMAL = 5   # H.js:31


def js_to_fixed(x, digits):
    """Number.prototype.toFixed (ECMA-262 21.1.3.3) in exact rational arithmetic: n = the integer nearest to
    |x| * 10^digits (the larger one on a tie), printed with `digits` decimals; sign from `x < 0` (so -0 has none)."""
    from fractions import Fraction
    x = float(x)
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "Infinity" if x > 0 else "-Infinity"
    if abs(x) >= 1e21:
        return repr(x)  # ToString(x)
    scaled = Fraction(abs(x)) * 10 ** digits
    n = scaled.numerator // scaled.denominator
    if scaled - n >= Fraction(1, 2):
        n += 1
    text = str(n).rjust(digits + 1, "0")
    if digits:
        text = text[:-digits] + "." + text[-digits:]
    return ("-" if x < 0 else "") + text


class RefImageData:
    def __init__(self, data, width, height):
        self.data = data
        self.width = width
        self.height = height


def _is_view(a):  # ArrayBuffer.isView
    return isinstance(a, np.ndarray)


def _to_f32_flat(points):  # new Float32Array(points.flat())
    return np.array(points, dtype=np.float64).reshape(-1).astype(np.float32)


def _gt0(v):  # JS: v > 0 with null -> false
    return v is not None and v > 0


def _le0(v):  # JS: v <= 0 with null -> true (null coerces to 0)
    return True if v is None else v <= 0


def contains_value_greater_than(it, value):  # H.js:1539
    return bool(np.any(np.asarray(it, dtype=np.float64) > value))


def denormalize_points(p, w, h):  # H.js:1603 (in place, rounded to the array's own dtype)
    p[0::2] = (p[0::2].astype(np.float64) * float(w)).astype(p.dtype)
    p[1::2] = (p[1::2].astype(np.float64) * float(h)).astype(p.dtype)


def normalize_points(p, w, h):  # H.js:1621
    p[0::2] = (p[0::2].astype(np.float64) / float(w)).astype(p.dtype)
    p[1::2] = (p[1::2].astype(np.float64) / float(h)).astype(p.dtype)


def check_and_select_transform(transform, points):  # H.js:1444
    n = points.size
    if transform == "auto":
        if n == 3 * DIMS:
            return "affine"
        if n == 4 * DIMS:
            return "projective"
        if n > 4 * DIMS:
            return "piecewiseaffine"
        raise ValueError(f"Transforms must contain at least 3 points but only {n / DIMS:g} were given")
    if transform == "piecewiseaffine":
        if n < 3 * DIMS:
            raise ValueError("A piecewise (or affine) transform needs to determine least three reference points "
                             f"but only {n / DIMS:g} were given")
        return transform
    if transform == "affine":
        if n != 3 * DIMS:
            raise ValueError(f"An affine transform needs to determine exactly three reference points but {n / DIMS:g} were given")
        return transform
    if transform == "projective":
        if n != 4 * DIMS:
            raise ValueError(f"A projective transform needs to determine exactly four reference points but {n / DIMS:g} were given")
        return transform
    raise ValueError(f'Transform "{transform}" is unknown')


def stand_in_delaunay(points):
    """`new Delaunator(points).triangles` (H.js:1216): the oracle's own restatement of delaunator 5.0.0
    (oracle/delaunator_ref.py; the package is third-party and absent from the reference tree — parity unpinned)."""
    from oracle import delaunator_ref
    return np.asarray(delaunator_ref.triangles(np.asarray(points, dtype=np.float64).reshape(-1)), dtype=np.uint32)


class RefHomography:
    def __init__(self, transform="auto", width=None, height=None):  # H.js:78
        if width is not None:
            width = O.js_round(width)
        if height is not None:
            height = O.js_round(height)
        self._width = width
        self._height = height
        self._objectiveWidth = None
        self._objectiveHeight = None
        self._srcPoints = None
        self._dstPoints = None
        self.firstTransformSelected = transform.lower()
        self.transform = transform.lower()
        self._image = None
        self._maxSrcX = self._maxSrcY = self._minSrcX = self._minSrcY = None
        self._srcPointsAreNormalized = True
        self._dstPointsAreNormalized = True
        self._trianglesCorrespondencesMatrix = None
        self._triangles = None
        self._transformMatrix = None
        self._piecewiseMatrices = None
        self._initialTriangles = None
        self._xOutputOffset = None
        self._yOutputOffset = None
        self.last_path = None  # which of the four loops warp() ran (test aid, not in the reference)

    # ------------------------------------------------------------------ public setters
    def setReferencePoints(self, srcPoints, dstPoints, image=None, width=None, height=None,
                           srcPointsAreNormalized=None, dstPointsAreNormalized=None):  # H.js:173
        if srcPoints is None or dstPoints is None:
            raise ValueError("Source and Destiny points must be defined when calling setReferencePoints().")
        self._dstPoints = None
        self.setSourcePoints(srcPoints, image, width, height, srcPointsAreNormalized)
        self.setDestinyPoints(dstPoints, dstPointsAreNormalized)

    def setSourcePoints(self, points, image=None, width=None, height=None, pointsAreNormalized=None):  # H.js:218
        if not _is_view(points):
            points = _to_f32_flat(points)
        self._srcPoints = points
        self._srcPointsAreNormalized = (not contains_value_greater_than(points, NORMALIZED_MAX)
                                        if pointsAreNormalized is None else pointsAreNormalized)
        self._transformMatrix = None
        self.transform = check_and_select_transform(self.firstTransformSelected, self._srcPoints)
        self._objectiveWidth = None
        self._objectiveHeight = None
        if image is not None:
            self.setImage(image, width, height)
        elif width is not None or height is not None:
            self._setSrcWidthHeight(width, height)
        if self._width is not None and self._height is not None and self._srcPointsAreNormalized:
            denormalize_points(self._srcPoints, self._width, self._height)
            self._srcPointsAreNormalized = False
        if self._dstPoints is not None and self.transform != "piecewiseaffine":
            self._transformMatrix = O.calculate_transform_matrix(self.transform, self._srcPoints, self._dstPoints)
        if self.transform == "piecewiseaffine" and self._trianglesCorrespondencesMatrix is None:
            self._triangles = self._initialTriangles
            self._piecewiseMatrices = None
            if (not self._srcPointsAreNormalized) or (_gt0(self._width) and _gt0(self._height)):
                self._setPiecewiseAffineTransformParameters()
            elif self._triangles is None:
                self._triangles = stand_in_delaunay(self._srcPoints)

    def setImage(self, image, width=None, height=None):  # H.js:290 (ImageData form only)
        data = np.ascontiguousarray(image.data, dtype=np.uint8).reshape(-1)
        self._image = data
        self._setSrcWidthHeight(image.width, image.height)
        if self._srcPoints is not None and self.transform == "piecewiseaffine":
            self._setPiecewiseAffineTransformParameters()
        if self._dstPoints is not None and (_le0(self._objectiveWidth) or _le0(self._objectiveHeight)):
            self._induceBestObjectiveWidthAndHeight()

    def setDestinyPoints(self, points, pointsAreNormalized=None):  # H.js:337
        if not _is_view(points):
            points = _to_f32_flat(points)
        if self._srcPoints is not None and points.size != self._srcPoints.size:
            raise ValueError(f"It must be the same amount of destiny points ({points.size / DIMS:g}) "
                             f"than source points ({self._srcPoints.size / DIMS:g})")
        self._dstPoints = points
        self._dstPointsAreNormalized = (not contains_value_greater_than(points, NORMALIZED_MAX)
                                        if pointsAreNormalized is None else pointsAreNormalized)
        if self.transform != "piecewiseaffine":
            if (self._dstPointsAreNormalized and _gt0(self._width) and _gt0(self._height)
                    and self.transform == "projective"):
                denormalize_points(self._dstPoints, self._width, self._height)
                self._dstPointsAreNormalized = False
            self._putSrcAndDstPointsInSameRange()
            self._transformMatrix = O.calculate_transform_matrix(self.transform, self._srcPoints, self._dstPoints)
        else:
            self._piecewiseMatrices = None
        if self._image is not None or (self.transform == "piecewiseaffine" and _gt0(self._width) and _gt0(self._height)):
            self._induceBestObjectiveWidthAndHeight()
        if self.transform == "piecewiseaffine" and _gt0(self._width) and _gt0(self._height):
            if self._dstPointsAreNormalized:
                denormalize_points(self._dstPoints, self._width, self._height)
                self._dstPointsAreNormalized = False
            self._setPiecewiseAffineTransformParameters()

    def setTriangles(self, triangles):  # H.js:517
        self._triangles = np.ascontiguousarray(triangles, dtype=np.uint32).reshape(-1)
        if ((not self._srcPointsAreNormalized) or (_gt0(self._width) and _gt0(self._height))) and self._srcPoints is not None:
            self._setPiecewiseAffineTransformParameters()

    # ------------------------------------------------------------------ warp
    def warp(self, image=None, asHTMLPromise=False, applyAlwaysInverse=False):  # H.js:408
        if image is not None:
            self.setImage(image)
        elif self._image is None:
            raise ValueError("warp() must receive an image if it was not setted before through `setImage(img)` "
                             "or  `setSourcePoints(points, img)`")
        oW, oH, W, H = self._objectiveWidth, self._objectiveHeight, self._width, self._height
        if self.transform == "piecewiseaffine":
            if applyAlwaysInverse or (oW > W or oH > H or oW * 1.2 < W or oH * 1.2 < H):
                out = self._inversePiecewiseAffineWarp(self._image)
            else:
                out = self._piecewiseAffineWarp(self._image)
        elif self.transform == "affine":
            if applyAlwaysInverse or (oW != W or oH != H):
                out = self._inverseGeometricWarp(self._image)
            else:
                out = self._geometricWarp(self._image)
        else:  # projective
            out = self._inverseGeometricWarp(self._image)
        area = self._objectiveWidth * self._objectiveHeight
        if area >= 1 and not math.isnan(area):
            return RefImageData(out, int(self._objectiveWidth), int(self._objectiveHeight))
        return RefImageData(np.zeros(4, np.uint8), 1, 1)

    def getTransformationMatrixAsCSS(self, srcPoints=None, dstPoints=None, width=None, height=None):  # H.js:548
        if width is not None or height is not None:
            self._setSrcWidthHeight(width, height)
        if srcPoints is not None:
            self.setSourcePoints(srcPoints, None, width, height)
        if dstPoints is not None:
            self.setDestinyPoints(dstPoints)
        if self._srcPoints is None:
            raise ValueError("Impossible to calculate a transform when srcPoints are not set")
        elif self._dstPoints is None:
            raise ValueError("Impossible to calculate a transform when dstPoints are not set")
        elif self._transformMatrix is None:
            raise ValueError("Transform matrix can not be calculated")
        tm = self._transformMatrix
        if self.transform == "affine":  # H.js:560-567
            matrix = "matrix("
            for i in range(len(tm)):
                matrix += js_to_fixed(tm[i], MAX_CSS_DECIMAL)
                matrix += ", " if i < len(tm) - 1 else ")"
        elif self.transform == "projective":  # H.js:568-581
            matrix = "matrix3d("
            i = 0
            for dy in range(4):
                for dx in range(4):
                    if dy == 2 and dx == 2 or dy == 3 and dx == 3:
                        matrix += "1"
                    elif dy == 2 or dx == 2:
                        matrix += "0"
                    else:
                        matrix += js_to_fixed(tm[(i * 3) % 8], MAX_CSS_DECIMAL)
                        i += 1
                    matrix += ", " if dy * 4 + dx < 4 * 4 - 1 else ")"
        else:
            raise ValueError('Only "affine" or "projective" transforms can be applied on the CSS transform property, '
                             f"but {self.transform} selected")
        return matrix

    # ------------------------------------------------------------------ private plumbing
    def _setSrcWidthHeight(self, width, height):  # H.js:637
        last_w, last_h = self._width, self._height
        self._width, self._height = width, height
        if last_w != width or last_h != height:
            self._width = O.js_round(0.0 if width is None else width)    # Math.round(null) === 0
            self._height = O.js_round(0.0 if height is None else height)
            self._trianglesCorrespondencesMatrix = None
            if self.transform == "projective":
                if self._srcPoints is not None and self._srcPointsAreNormalized:
                    denormalize_points(self._srcPoints, self._width, self._height)
                    self._srcPointsAreNormalized = False
                if self._dstPoints is not None and self._dstPointsAreNormalized:
                    denormalize_points(self._dstPoints, self._width, self._height)
                    self._dstPointsAreNormalized = False
                if self._dstPoints is not None and self._srcPoints is not None:
                    self._transformMatrix = O.calculate_transform_matrix(self.transform, self._srcPoints, self._dstPoints)
                    self._induceBestObjectiveWidthAndHeight()
            if self._srcPoints is not None and self.transform == "piecewiseaffine":
                self._setPiecewiseAffineTransformParameters()

    def _induceBestObjectiveWidthAndHeight(self):  # H.js:693
        if self.transform in ("affine", "projective"):
            if self._transformMatrix is None:
                if self._srcPointsAreNormalized != self._dstPointsAreNormalized:
                    self._putSrcAndDstPointsInSameRange()
                self._transformMatrix = O.calculate_transform_matrix(self.transform, self._srcPoints, self._dstPoints)
            lim = O.transform_limits(self._transformMatrix, self._width, self._height)
            self._xOutputOffset, self._yOutputOffset, self._objectiveWidth, self._objectiveHeight = [float(v) for v in lim]
        elif not self._dstPointsAreNormalized:
            mm = O.minmax_xy(self._dstPoints)
            self._xOutputOffset, self._yOutputOffset = float(mm[0]), float(mm[1])
            self._objectiveWidth = float(mm[2]) - self._xOutputOffset
            self._objectiveHeight = float(mm[3]) - self._yOutputOffset
        elif _gt0(self._width) and _gt0(self._height):
            mn_x, mn_y, mx_x, mx_y = [float(v) for v in O.minmax_xy(self._dstPoints, False)]
            self._xOutputOffset = O.js_round(mn_x)
            self._yOutputOffset = O.js_round(mn_y)
            self._objectiveWidth = O.js_round((mx_x - mn_x) * self._width)
            self._objectiveHeight = O.js_round((mx_y - mn_y) * self._height)
        else:
            raise ValueError("Trying to calculate a the output width and height of a Piecewise Affine transform "
                             "but source width and height are not set")

    def _setPiecewiseAffineTransformParameters(self):  # H.js:738
        if self._srcPoints is None:
            raise ValueError("Trying to set the Piecewise Affine Transform parameters before setting the Source Points.")
        if self._triangles is None:
            self._triangles = stand_in_delaunay(self._srcPoints)
        if self._srcPointsAreNormalized:
            if _gt0(self._width) and _gt0(self._height):
                denormalize_points(self._srcPoints, self._width, self._height)
                self._srcPointsAreNormalized = False
            else:
                raise ValueError("Trying to set the Piecewise Affine Transform parameters without knowing the source points ranges")
        if (not self._srcPointsAreNormalized) and (self._triangles is None or self._trianglesCorrespondencesMatrix is None):
            mm = O.minmax_xy(self._srcPoints)
            self._minSrcX, self._minSrcY, self._maxSrcX, self._maxSrcY = [float(v) for v in mm]
            self._trianglesCorrespondencesMatrix = self._buildTrianglesCorrespondencesMatrix()
        if self._dstPoints is not None and self._piecewiseMatrices is None and self._triangles is not None:
            if self._dstPointsAreNormalized:
                denormalize_points(self._dstPoints, self._width, self._height)
                self._dstPointsAreNormalized = False
            self._piecewiseMatrices = self._calculatePiecewiseAffineTransformMatrices()

    def _calculatePiecewiseAffineTransformMatrices(self):  # H.js:785
        if self._srcPointsAreNormalized != self._dstPointsAreNormalized:
            self._putSrcAndDstPointsInSameRange()
        return O.piecewise_matrices(self._srcPoints, self._dstPoints, self._triangles)

    def _buildTrianglesCorrespondencesMatrix(self):  # H.js:817
        mw = self._maxSrcX - self._minSrcX
        length = mw * (self._maxSrcY - self._minSrcY)
        return O.build_index_map(self._srcPoints, self._triangles, mw, self._minSrcY, length)

    def _buildInverseTrianglesCorrespondencesMatrix(self):  # H.js:845
        length = self._objectiveWidth * self._objectiveHeight
        self._trianglesCorrespondencesMatrix = O.build_index_map(self._dstPoints, self._triangles, self._objectiveWidth,
                                                                 self._yOutputOffset, length)
        return self._trianglesCorrespondencesMatrix

    def _putSrcAndDstPointsInSameRange(self):  # H.js:876
        if self._dstPointsAreNormalized != self._srcPointsAreNormalized:
            if self._dstPointsAreNormalized and _gt0(self._width) and _gt0(self._height):
                normalize_points(self._srcPoints, self._width, self._height)
                self._srcPointsAreNormalized = True
            elif self._srcPointsAreNormalized and _gt0(self._width) and _gt0(self._height):
                denormalize_points(self._srcPoints, self._width, self._height)
                self._srcPointsAreNormalized = False
            else:
                raise ValueError("Impossible to put source and destiny points in the same range.")

    # ------------------------------------------------------------------ the four loops
    def _ints(self):
        return (int(self._width), int(self._height), int(self._xOutputOffset), int(self._yOutputOffset),
                int(self._objectiveWidth), int(self._objectiveHeight))

    def _geometricWarp(self, image):  # H.js:911
        self.last_path = "forward_geometric"
        W, H, xo, yo, oW, oH = self._ints()
        return O.warp_forward_geometric(image, W, H, self._transformMatrix, xo, yo, oW, oH)

    def _piecewiseAffineWarp(self, image):  # H.js:948
        self.last_path = "forward_piecewise"
        W, H, xo, yo, oW, oH = self._ints()
        return O.warp_forward_piecewise(image, W, H, self._trianglesCorrespondencesMatrix, self._piecewiseMatrices,
                                        xo, yo, oW, oH, int(self._minSrcX), int(self._minSrcY), int(self._maxSrcX),
                                        int(self._maxSrcY))

    def _inverseGeometricWarp(self, image):  # H.js:987
        self.last_path = "inverse_geometric"
        self._putSrcAndDstPointsInSameRange()
        inv = O.calculate_transform_matrix(self.transform, self._dstPoints, self._srcPoints)
        if math.isnan(self._objectiveWidth * self._objectiveHeight):
            return np.zeros(0, np.uint8)
        W, H, xo, yo, oW, oH = self._ints()
        return O.warp_inverse_geometric(image, W, H, inv, xo, yo, oW, oH)

    def _inversePiecewiseAffineWarp(self, image):  # H.js:1029
        self.last_path = "inverse_piecewise"
        imap = self._buildInverseTrianglesCorrespondencesMatrix()
        inv = O.inverse_matrices(self._piecewiseMatrices)
        W, H, xo, yo, oW, oH = self._ints()
        return O.warp_inverse_piecewise(image, W, H, imap, inv, xo, yo, oW, oH, int(self._minSrcX), int(self._minSrcY))
