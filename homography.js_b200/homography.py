"""Host-side mirror of the reference's `class Homography` (H.js:38-614) over the CUDA engine.

This is the Python twin of js/Homography.mjs: the same public surface
(`Homography(transform, width, height)`, `setReferencePoints`, `setSourcePoints`, `setDestinyPoints`,
`setImage`, `setTriangles`, `warp`, `getTransformationMatrixAsCSS`, `transformHTMLElement`) and the same state machine (normalised-range auto-detection,
in-place (de)normalisation of caller-owned point arrays, cache invalidation, forward/inverse dispatch
thresholds), with every arithmetic step executed by libhgwarp.so on the GPU:

    calculateTransformMatrix  (H.js:1237) -> hg_solve_affine / hg_solve_projective / hg_solve_with_limits
    calculateTransformLimits  (H.js:1503) -> hg_transform_limits
    _inverseGeometricWarp     (H.js:987)  -> hg_warp_inverse_points   (solve + pixel loop, one submission)
    _inversePiecewiseAffineWarp (H.js:1029) -> hg_warp_piecewise_inverse
    _geometricWarp / _piecewiseAffineWarp (H.js:911/948) -> hg_warp_forward_matrix / hg_warp_piecewise_forward

Errors the reference raises as bare strings are raised as HomographyError with the same text.
There is no CPU fallback: without the CUDA library / a GPU the constructor raises.
"""
from __future__ import annotations

import decimal
import math

import numpy as np

from . import _abi

NORMALIZED_MAX = 8.0  # H.js:36
DIMS = 2              # H.js:34
MAX_CSS_DECIMAL = 5   # H.js:31


class HomographyError(Exception):
    """The reference throws bare strings; the text is preserved."""


class ImageData:
    """ImageData-like result of warp(): RGBA8 bytes, width, height (H.js:437)."""

    def __init__(self, data: np.ndarray, width: int, height: int):
        self.data = data
        self.width = width
        self.height = height

    def as_array(self) -> np.ndarray:
        return self.data.reshape(self.height, self.width, 4)

    @classmethod
    def from_png(cls, png_bytes: bytes) -> "ImageData":
        """What `loadImage()` + canvas `getImageData` give the reference (test/nodeTest.js:11, H.js:1071-1076):
        the RGBA8 pixels of a PNG file (hg_png_decode)."""
        a = _abi.png_decode(png_bytes)
        return cls(a.reshape(-1), a.shape[1], a.shape[0])

    @classmethod
    def from_jpeg(cls, jpeg_bytes: bytes) -> "ImageData":
        """The RGBA8 pixels of a JPEG file, as a browser's canvas returns them (hg_jpeg_decode)."""
        a = _abi.jpeg_decode(jpeg_bytes)
        return cls(a.reshape(-1), a.shape[1], a.shape[0])

    def to_png(self) -> bytes:
        """The result as PNG file bytes (the reference's HTMLImageElementFromImageData / toDataURL, H.js:467-496)."""
        return _abi.png_encode(np.asarray(self.data, dtype=np.uint8).reshape(self.height, self.width, 4))


def _js_round(x: float) -> float:
    """Math.round: nearest integer, ties toward +inf (host-side scalar bookkeeping only)."""
    if x is None:   # Math.round(null) === 0
        return 0.0
    if x != x or x in (math.inf, -math.inf) or abs(x) >= 4503599627370496.0:
        return x
    r = math.floor(x)
    return float(r + 1) if x - r >= 0.5 else float(r)


def _js_to_fixed(x: float, digits: int) -> str:
    """Number.prototype.toFixed: the decimal string of the EXACT binary value, rounded to `digits` places with ties
    away from zero ("let n be an integer for which n / 10^f - x is as close to zero as possible; if there are two
    such n, pick the larger n", applied to |x|); -0 prints without a sign, a negative that rounds to zero keeps it."""
    x = float(x)
    if x != x:
        return "NaN"
    if x in (math.inf, -math.inf):
        return "Infinity" if x > 0 else "-Infinity"
    if abs(x) >= 1e21:
        return repr(x)  # ToString(x): shortest round-trip digits, "1e+21" form — the same text in both languages
    with decimal.localcontext() as ctx:
        ctx.prec = 80
        q = decimal.Decimal(abs(x)).quantize(decimal.Decimal(1).scaleb(-digits), rounding=decimal.ROUND_HALF_UP)
    return ("-" if x < 0 else "") + format(q, "f")


def _positive(v) -> bool:      # JS `v > 0` (null -> false)
    return v is not None and v > 0


def _not_positive(v) -> bool:  # JS `v <= 0` (null coerces to 0 -> true; NaN -> false)
    return True if v is None else v <= 0


def _as_point_array(points):
    """`if(!ArrayBuffer.isView(points)) points = new Float32Array(points.flat())` (H.js:220, 339).
    numpy arrays play the role of typed arrays: they are kept (and later mutated in place)."""
    if isinstance(points, np.ndarray) and points.dtype in (np.float32, np.float64) and points.ndim == 1:
        return points
    return np.asarray(points, dtype=np.float64).reshape(-1).astype(np.float32)


def _scale_in_place(p: np.ndarray, sx: float, sy: float, divide: bool):
    """denormalizePoints / normalizePoints (H.js:1603/1621): double math, stored in the array's dtype."""
    x = p[0::2].astype(np.float64)
    y = p[1::2].astype(np.float64)
    if divide:
        x, y = x / sx, y / sy
    else:
        x, y = x * sx, y * sy
    p[0::2] = x.astype(p.dtype)
    p[1::2] = y.astype(p.dtype)


def _minmax_xy(p: np.ndarray):
    """minmaxXYofArray (H.js:1558), unrounded: (minX, minY, maxX, maxY); strict compares skip NaN."""
    x = p[0::2].astype(np.float64)
    y = p[1::2].astype(np.float64)
    x, y = x[~np.isnan(x)], y[~np.isnan(y)]
    mnx = float(x.min()) if x.size else math.inf
    mxx = float(x.max()) if x.size else -math.inf
    mny = float(y.min()) if y.size else math.inf
    mxy = float(y.max()) if y.size else -math.inf
    return mnx, mny, mxx, mxy


def _select_transform(first: str, points: np.ndarray) -> str:
    """checkAndSelectTransform (H.js:1444)."""
    n = points.size
    if first == "auto":
        if n == 3 * DIMS:
            return "affine"
        if n == 4 * DIMS:
            return "projective"
        if n > 4 * DIMS:
            return "piecewiseaffine"
        raise HomographyError(f"Transforms must contain at least 3 points but only {n / DIMS:g} were given")
    if first == "piecewiseaffine":
        if n < 3 * DIMS:
            raise HomographyError("A piecewise (or affine) transform needs to determine least three reference points "
                                  f"but only {n / DIMS:g} were given")
        return first
    if first == "affine":
        if n != 3 * DIMS:
            raise HomographyError(f"An affine transform needs to determine exactly three reference points but {n / DIMS:g} were given")
        return first
    if first == "projective":
        if n != 4 * DIMS:
            raise HomographyError(f"A projective transform needs to determine exactly four reference points but {n / DIMS:g} were given")
        return first
    raise HomographyError(f'Transform "{first}" is unknown')


def default_triangulation(points: np.ndarray) -> np.ndarray:
    """`new Delaunator(points).triangles` (H.js:1216): the library's host-side restatement of delaunator 5.0.0
    (hg_delaunay, csrc/delaunay_host.cuh).  The package is third-party and absent from the reference tree, so the
    triangle order is by construction, not pinned by any reference fixture; pass triangles explicitly with
    setTriangles() (the reference's own hook, H.js:517) to reproduce a given mesh."""
    return _abi.delaunay(points)


class Homography:
    _KIND = {"affine": _abi.HG_AFFINE, "projective": _abi.HG_PROJECTIVE}

    def __init__(self, transform: str = "auto", width=None, height=None, device: int = 0, context=None,
                 sampling: str = "nearest", pinned_output: bool = False):
        """`sampling="bilinear"` is an EXTENSION (the reference only has Math.round sampling): it applies to the
        inverse affine / projective loop and is accurate to <= 1 LSB per channel (see csrc/bilinear.cuh).
        `pinned_output=True`: the inverse warps return their ImageData over page-locked memory owned by this object (two
        buffers used alternately, so a result stays valid until the warp after next) — the copy engine then writes the
        result at link speed instead of staging it through the driver (in Node: an external ArrayBuffer)."""
        self._ctx = context if context is not None else _abi.Context(device)
        self._pinned_output = bool(pinned_output)
        self._out_ring = [None, None]
        self._out_next = 0
        if sampling not in ("nearest", "bilinear"):
            raise HomographyError(f'sampling "{sampling}" is unknown')
        self._sampling = _abi.HG_BILINEAR if sampling == "bilinear" else _abi.HG_NEAREST
        self._width = None if width is None else _js_round(width)
        self._height = None if height is None else _js_round(height)
        self._objectiveWidth = None
        self._objectiveHeight = None
        self._xOutputOffset = None
        self._yOutputOffset = None
        self._srcPoints = None
        self._dstPoints = None
        self.firstTransformSelected = transform.lower()
        self.transform = transform.lower()
        self._image = None            # host view of the RGBA bytes (the device copy lives in the context)
        self._image_on_device = False
        self._minSrcX = self._minSrcY = self._maxSrcX = self._maxSrcY = None
        self._srcPointsAreNormalized = True
        self._dstPointsAreNormalized = True
        # which index map the reference would currently hold in _trianglesCorrespondencesMatrix:
        # None | "forward" | "inverse"  (H.js:115, 759, 847 — one field shared by both maps)
        self._map_state = None
        self._triangles = None
        self._initialTriangles = None
        self._transformMatrix = None
        self._piecewiseMatrices = None
        self._mesh_on_device = False
        self.last_path = None

    # ------------------------------------------------------------------ public API
    def setReferencePoints(self, srcPoints, dstPoints, image=None, width=None, height=None,
                           srcPointsAreNormalized=None, dstPointsAreNormalized=None):
        """H.js:173."""
        if srcPoints is None or dstPoints is None:
            raise HomographyError("Source and Destiny points must be defined when calling setReferencePoints().")
        self._dstPoints = None
        self.setSourcePoints(srcPoints, image, width, height, srcPointsAreNormalized)
        self.setDestinyPoints(dstPoints, dstPointsAreNormalized)

    def setSourcePoints(self, points, image=None, width=None, height=None, pointsAreNormalized=None):
        """H.js:218."""
        pts = _as_point_array(points)
        self._srcPoints = pts
        self._mesh_on_device = False
        self._srcPointsAreNormalized = (not bool(np.any(pts > NORMALIZED_MAX))) if pointsAreNormalized is None \
            else pointsAreNormalized
        self._transformMatrix = None
        self.transform = _select_transform(self.firstTransformSelected, pts)
        self._objectiveWidth = None
        self._objectiveHeight = None
        if image is not None:
            self.setImage(image, width, height)
        elif width is not None or height is not None:
            self._setSrcWidthHeight(width, height)
        if self._width is not None and self._height is not None and self._srcPointsAreNormalized:
            self._denormalize_src()
        if self._dstPoints is not None and self.transform != "piecewiseaffine":
            self._transformMatrix = self._solve(self._srcPoints, self._dstPoints)
        if self.transform == "piecewiseaffine" and self._map_state is None:
            self._triangles = self._initialTriangles
            self._piecewiseMatrices = None
            if (not self._srcPointsAreNormalized) or (_positive(self._width) and _positive(self._height)):
                self._setPiecewiseAffineTransformParameters()
            elif self._triangles is None:
                self._triangles = default_triangulation(self._srcPoints)

    def setImage(self, image, width=None, height=None):
        """H.js:290 — the Node / ImageData form: any object with .data (RGBA8), .width, .height."""
        data = getattr(image, "data", None)
        if data is None and isinstance(image, np.ndarray) and image.ndim == 3 and image.shape[2] == 4:
            image = ImageData(np.ascontiguousarray(image, dtype=np.uint8).reshape(-1), image.shape[1], image.shape[0])
            data = image.data
        if data is None:
            raise HomographyError("setImage() needs an ImageData-like object ({data, width, height}); "
                                  "HTMLImageElement inputs exist only in a browser")
        w, h = int(image.width), int(image.height)
        self._image = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
        self._ctx.image_set(self._image, w, h)   # device-resident like this._image
        self._image_on_device = True
        self._setSrcWidthHeight(w, h)
        if self._srcPoints is not None and self.transform == "piecewiseaffine":
            self._setPiecewiseAffineTransformParameters()
        if self._dstPoints is not None and (_not_positive(self._objectiveWidth) or _not_positive(self._objectiveHeight)):
            self._induceBestObjectiveWidthAndHeight()

    def setDestinyPoints(self, points, pointsAreNormalized=None):
        """H.js:337."""
        pts = _as_point_array(points)
        if self._srcPoints is not None and pts.size != self._srcPoints.size:
            raise HomographyError(f"It must be the same amount of destiny points ({pts.size / DIMS:g}) "
                                  f"than source points ({self._srcPoints.size / DIMS:g})")
        self._dstPoints = pts
        self._dstPointsAreNormalized = (not bool(np.any(pts > NORMALIZED_MAX))) if pointsAreNormalized is None \
            else pointsAreNormalized
        have_size = _positive(self._width) and _positive(self._height)
        limits_done = False
        if self.transform != "piecewiseaffine":
            if self._dstPointsAreNormalized and have_size and self.transform == "projective":
                self._denormalize_dst()
            self._putSrcAndDstPointsInSameRange()
            if self._image is not None:
                # matrix + output extent in one device submission (H.js:357 + 365-366)
                self._transformMatrix, lim = self._ctx.solve_with_limits(
                    self._KIND[self.transform], self._srcPoints, self._dstPoints, self._width, self._height)
                self._set_limits(lim)
                limits_done = True
            else:
                self._transformMatrix = self._solve(self._srcPoints, self._dstPoints)
        else:
            self._piecewiseMatrices = None
        if not limits_done and (self._image is not None or (self.transform == "piecewiseaffine" and have_size)):
            self._induceBestObjectiveWidthAndHeight()
        if self.transform == "piecewiseaffine" and have_size:
            if self._dstPointsAreNormalized:
                self._denormalize_dst()
            self._setPiecewiseAffineTransformParameters()

    def setTriangles(self, triangles):
        """H.js:517."""
        self._triangles = np.ascontiguousarray(triangles, dtype=np.uint32).reshape(-1)
        self._mesh_on_device = False
        if ((not self._srcPointsAreNormalized) or (_positive(self._width) and _positive(self._height))) \
                and self._srcPoints is not None:
            self._setPiecewiseAffineTransformParameters()

    def warp(self, image=None, asHTMLPromise=False, applyAlwaysInverse=False):
        """H.js:408.  Returns an ImageData-like object synchronously."""
        if asHTMLPromise:
            raise HomographyError("asHTMLPromise needs a DOM; only ImageData results exist outside a browser")
        if image is not None:
            self.setImage(image)
        elif self._image is None:
            raise HomographyError("warp() must receive an image if it was not setted before through `setImage(img)` "
                                  "or  `setSourcePoints(points, img)`")
        oW, oH, W, H = self._objectiveWidth, self._objectiveHeight, self._width, self._height
        area = oW * oH
        empty = not (area >= 1) or math.isnan(area)   # H.js:436-441: 1x1 transparent fallback
        if self.transform == "piecewiseaffine":
            inverse = applyAlwaysInverse or (oW > W or oH > H or oW * 1.2 < W or oH * 1.2 < H)
            run = self._inversePiecewiseAffineWarp if inverse else self._piecewiseAffineWarp
        elif self.transform == "affine":
            inverse = applyAlwaysInverse or (oW != W or oH != H)
            run = self._inverseGeometricWarp if inverse else self._geometricWarp
        else:
            run = self._inverseGeometricWarp
        out = run(empty)
        if empty:
            return ImageData(np.zeros(4, np.uint8), 1, 1)
        return ImageData(out, int(oW), int(oH))

    def getTransformationMatrixAsCSS(self, srcPoints=None, dstPoints=None, width=None, height=None) -> str:
        """H.js:548-586: the current affine / projective matrix as the value of the CSS `transform` property —
        `matrix(a, b, c, d, e, f)` from the six float32 coefficients, `matrix3d(...)` from the eight doubles (the 3x3
        transposed into a column-major 4x4 with the identity's z row / column), every number through toFixed(5)."""
        if width is not None or height is not None:
            self._setSrcWidthHeight(width, height)
        if srcPoints is not None:
            self.setSourcePoints(srcPoints, None, width, height)
        if dstPoints is not None:
            self.setDestinyPoints(dstPoints)
        if self._srcPoints is None:
            raise HomographyError("Impossible to calculate a transform when srcPoints are not set")
        if self._dstPoints is None:
            raise HomographyError("Impossible to calculate a transform when dstPoints are not set")
        if self._transformMatrix is None:
            raise HomographyError("Transform matrix can not be calculated")
        m = [float(v) for v in self._transformMatrix]
        if self.transform == "affine":
            return "matrix(" + ", ".join(_js_to_fixed(v, MAX_CSS_DECIMAL) for v in m) + ")"
        if self.transform == "projective":
            cells, i = [], 0
            for dy in range(4):
                for dx in range(4):
                    if (dy == 2 and dx == 2) or (dy == 3 and dx == 3):
                        cells.append("1")
                    elif dy == 2 or dx == 2:
                        cells.append("0")
                    else:
                        cells.append(_js_to_fixed(m[(i * 3) % 8], MAX_CSS_DECIMAL))
                        i += 1
            return "matrix3d(" + ", ".join(cells) + ")"
        raise HomographyError('Only "affine" or "projective" transforms can be applied on the CSS transform property, '
                              f"but {self.transform} selected")

    def transformHTMLElement(self, element, srcPoints=None, dstPoints=None):
        """H.js:611 — duck-typed outside a browser: `element.getBoundingClientRect()` gives .width / .height and the
        string lands in `element.style.transform`."""
        rect = element.getBoundingClientRect()
        element.style.transform = self.getTransformationMatrixAsCSS(srcPoints, dstPoints, rect.width, rect.height)

    # ------------------------------------------------------------------ state plumbing (H.js:637-896)
    def _solve(self, src, dst):
        if self.transform == "affine":
            return self._ctx.solve_affine(src, dst)
        if self.transform == "projective":
            return self._ctx.solve_projective(src, dst)
        raise HomographyError(f"{self.transform} transform does not exist")

    def _denormalize_src(self):
        _scale_in_place(self._srcPoints, self._width, self._height, divide=False)
        self._srcPointsAreNormalized = False
        self._mesh_on_device = False

    def _denormalize_dst(self):
        _scale_in_place(self._dstPoints, self._width, self._height, divide=False)
        self._dstPointsAreNormalized = False

    def _set_limits(self, lim):
        self._xOutputOffset, self._yOutputOffset, self._objectiveWidth, self._objectiveHeight = (float(v) for v in lim)

    def _setSrcWidthHeight(self, width, height):
        """H.js:637."""
        changed = (self._width != width) or (self._height != height)
        self._width, self._height = width, height
        if not changed:
            return
        self._width = _js_round(width)
        self._height = _js_round(height)
        self._map_state = None
        if self.transform == "projective":
            if self._srcPoints is not None and self._srcPointsAreNormalized:
                self._denormalize_src()
            if self._dstPoints is not None and self._dstPointsAreNormalized:
                self._denormalize_dst()
            if self._dstPoints is not None and self._srcPoints is not None:
                self._transformMatrix, lim = self._ctx.solve_with_limits(
                    _abi.HG_PROJECTIVE, self._srcPoints, self._dstPoints, self._width, self._height)
                self._set_limits(lim)
        if self._srcPoints is not None and self.transform == "piecewiseaffine":
            self._setPiecewiseAffineTransformParameters()

    def _induceBestObjectiveWidthAndHeight(self):
        """H.js:693."""
        if self.transform in ("affine", "projective"):
            if self._transformMatrix is None:
                if self._srcPointsAreNormalized != self._dstPointsAreNormalized:
                    self._putSrcAndDstPointsInSameRange()
                self._transformMatrix = self._solve(self._srcPoints, self._dstPoints)
            self._set_limits(self._ctx.transform_limits(self._transformMatrix, self._width, self._height))
        elif not self._dstPointsAreNormalized:
            if self._dstPoints.dtype == np.float32 and hasattr(self._ctx, "piecewise_extents"):
                # difference of ROUNDED extrema (H.js:707-710), on the device
                lim = self._ctx.piecewise_extents(self._dstPoints.reshape(1, -1, 2))[0]
                self._set_limits(lim)
            else:  # Float64Array points (kept as given, H.js:220): same arithmetic on the host
                mnx, mny, mxx, mxy = _minmax_xy(self._dstPoints)
                self._xOutputOffset, self._yOutputOffset = _js_round(mnx), _js_round(mny)
                self._objectiveWidth = _js_round(mxx) - self._xOutputOffset
                self._objectiveHeight = _js_round(mxy) - self._yOutputOffset
        elif _positive(self._width) and _positive(self._height):
            mnx, mny, mxx, mxy = _minmax_xy(self._dstPoints)
            self._xOutputOffset, self._yOutputOffset = _js_round(mnx), _js_round(mny)
            self._objectiveWidth = _js_round((mxx - mnx) * self._width)
            self._objectiveHeight = _js_round((mxy - mny) * self._height)
        else:
            raise HomographyError("Trying to calculate a the output width and height of a Piecewise Affine transform "
                                  "but source width and height are not set")

    def _setPiecewiseAffineTransformParameters(self):
        """H.js:738."""
        if self._srcPoints is None:
            raise HomographyError("Trying to set the Piecewise Affine Transform parameters before setting the Source Points.")
        if self._triangles is None:
            self._triangles = default_triangulation(self._srcPoints)
            self._mesh_on_device = False
        if self._srcPointsAreNormalized:
            if _positive(self._width) and _positive(self._height):
                self._denormalize_src()
            else:
                raise HomographyError("Trying to set the Piecewise Affine Transform parameters without knowing the source points ranges")
        if self._map_state is None:
            mnx, mny, mxx, mxy = _minmax_xy(self._srcPoints)
            self._minSrcX, self._minSrcY = _js_round(mnx), _js_round(mny)
            self._maxSrcX, self._maxSrcY = _js_round(mxx), _js_round(mxy)
            # the reference rasterises the forward map here (H.js:759); the engine builds it lazily,
            # only when a forward warp is actually dispatched — the state flag is what matters
            self._map_state = "forward"
        if self._dstPoints is not None and self._piecewiseMatrices is None and self._triangles is not None:
            if self._dstPointsAreNormalized:
                self._denormalize_dst()
            if self._srcPointsAreNormalized != self._dstPointsAreNormalized:
                self._putSrcAndDstPointsInSameRange()
            self._upload_mesh()
            self._piecewiseMatrices = self._ctx.piecewise_matrices(self._dstPoints)

    def _upload_mesh(self):
        if not self._mesh_on_device:
            self._ctx.piecewise_set_mesh(self._srcPoints.astype(np.float32), self._triangles)
            self._mesh_on_device = True

    def _putSrcAndDstPointsInSameRange(self):
        """H.js:876."""
        if self._dstPointsAreNormalized == self._srcPointsAreNormalized:
            return
        have_size = _positive(self._width) and _positive(self._height)
        if self._dstPointsAreNormalized and have_size:
            _scale_in_place(self._srcPoints, self._width, self._height, divide=True)
            self._srcPointsAreNormalized = True
            self._mesh_on_device = False
        elif self._srcPointsAreNormalized and have_size:
            self._denormalize_src()
        else:
            raise HomographyError(
                "Impossible to put source and destiny points in the same range. Possible solutions: \n"
                "1. Give a source width/height when calling setSrcPoints.\n"
                "2. Set the input image before.\n"
                "3. Give Source and Destiny points in the same range (both normalized or both in image dimensions)")

    # ------------------------------------------------------------------ the four loops -> device
    def _pinned_out(self, nbytes: int):
        """The next page-locked result buffer (None when the object returns ordinary arrays)."""
        if not self._pinned_output:
            return None
        i = self._out_next
        self._out_next ^= 1
        if self._out_ring[i] is None or self._out_ring[i].size < nbytes:
            self._out_ring[i] = self._ctx.pinned_array(nbytes + nbytes // 8)
        return self._out_ring[i][:nbytes]

    def _window(self):
        return (int(self._xOutputOffset), int(self._yOutputOffset), int(self._objectiveWidth), int(self._objectiveHeight))

    def _inverseGeometricWarp(self, empty):
        """H.js:987."""
        self.last_path = "inverse_geometric"
        self._putSrcAndDstPointsInSameRange()
        if empty:
            return None
        xo, yo, oW, oH = self._window()
        if self._sampling != _abi.HG_NEAREST:
            self._ctx.set_sampling(self._sampling)
        try:
            buf = self._pinned_out(oW * oH * 4)
            if buf is not None:
                self._ctx.warp_inverse_points(self._KIND[self.transform], self._dstPoints, self._srcPoints, xo, yo, oW, oH,
                                              out_host_ptr=buf.ctypes.data)
                return buf
            return self._ctx.warp_inverse_points(self._KIND[self.transform], self._dstPoints, self._srcPoints, xo, yo, oW, oH)
        finally:
            if self._sampling != _abi.HG_NEAREST:
                self._ctx.set_sampling(_abi.HG_NEAREST)

    def _geometricWarp(self, empty):
        """H.js:911."""
        self.last_path = "forward_geometric"
        if empty:
            return None
        xo, yo, oW, oH = self._window()
        return self._ctx.warp_forward_matrix(self._transformMatrix, xo, yo, oW, oH)

    def _inversePiecewiseAffineWarp(self, empty):
        """H.js:1029."""
        self.last_path = "inverse_piecewise"
        self._map_state = "inverse"   # the shared map field now holds the inverse map (H.js:847-850)
        if empty:
            return None
        xo, yo, oW, oH = self._window()
        self._upload_mesh()
        buf = self._pinned_out(oW * oH * 4)
        if buf is not None:
            self._ctx.warp_piecewise_inverse(self._dstPoints, xo, yo, oW, oH, int(self._minSrcX), int(self._minSrcY),
                                             out_host_ptr=buf.ctypes.data)
            return buf
        return self._ctx.warp_piecewise_inverse(self._dstPoints, xo, yo, oW, oH, int(self._minSrcX), int(self._minSrcY))

    def _piecewiseAffineWarp(self, empty):
        """H.js:948."""
        self.last_path = "forward_piecewise"
        if empty:
            return None
        xo, yo, oW, oH = self._window()
        self._upload_mesh()
        return self._ctx.warp_piecewise_forward(self._dstPoints, xo, yo, oW, oH, int(self._minSrcX), int(self._minSrcY),
                                                int(self._maxSrcX), int(self._maxSrcY),
                                                use_inverse_map=(self._map_state == "inverse"))
