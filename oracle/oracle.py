"""ctypes face of the CPU oracle (oracle/hg_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
Every function cites the reference line it restates (H.js = /root/reference/Homography.js).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libhgoracle.so")
_lib = None

_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_i16p = C.POINTER(C.c_int16)


def build(force: bool = False) -> str:
    """Compile the oracle with the recipe in oracle/Makefile."""
    src = os.path.join(_HERE, "hg_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], env={k: v for k, v in os.environ.items() if k != "CC"})
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.orc_js_round.restype = C.c_double
        L.orc_js_round.argtypes = [C.c_double]
        L.orc_js_toint32.restype = C.c_int32
        L.orc_js_toint32.argtypes = [C.c_double]
        L.orc_affine_from_triangles.argtypes = [_f64p, _f64p, _f32p]
        L.orc_inverse_affine.argtypes = [_f32p, _f32p]
        L.orc_projective_from_squares.argtypes = [_f64p, _f64p, _f64p]
        L.orc_apply_affine.argtypes = [_f32p, C.c_double, C.c_double, _f64p]
        L.orc_apply_projective.argtypes = [_f64p, C.c_double, C.c_double, _f64p]
        L.orc_transform_limits.argtypes = [C.c_int, C.c_void_p, C.c_double, C.c_double, _f64p]
        L.orc_minmax_xy.argtypes = [_f64p, C.c_int64, C.c_int, _f64p]
        L.orc_fill_triangle.argtypes = [_f32p, C.c_int32, C.c_double, C.c_double, _i16p, C.c_int64]
        L.orc_build_index_map.argtypes = [_f32p, _u32p, C.c_int32, C.c_double, C.c_double, _i16p, C.c_int64]
        L.orc_piecewise_matrices.argtypes = [_f32p, _f32p, _u32p, C.c_int32, _f32p]
        L.orc_warp_inverse_geometric.argtypes = [C.c_int, _u8p, C.c_int32, C.c_int32, C.c_void_p] + [C.c_int32] * 4 + [_u8p, C.c_int]
        L.orc_warp_inverse_geometric_bilinear.argtypes = [C.c_int, _u8p, C.c_int32, C.c_int32, C.c_void_p] + [C.c_int32] * 4 + [_u8p, C.c_int]
        L.orc_warp_forward_geometric.argtypes = [C.c_int, _u8p, C.c_int32, C.c_int32, C.c_void_p] + [C.c_int32] * 4 + [_u8p]
        L.orc_warp_inverse_piecewise.argtypes = [_u8p, C.c_int32, C.c_int32, _i16p, C.c_int64, _f32p, C.c_int32] + [C.c_int32] * 6 + [_u8p, C.c_int]
        L.orc_warp_forward_piecewise.argtypes = [_u8p, C.c_int32, C.c_int32, _i16p, C.c_int64, _f32p, C.c_int32] + [C.c_int32] * 8 + [_u8p]
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


def _img(image):
    a = np.ascontiguousarray(image, dtype=np.uint8)
    return a.reshape(-1)


# ---------------------------------------------------------------- JS number helpers
def js_round(x: float) -> float:
    """Math.round (ties toward +inf)."""
    return lib().orc_js_round(float(x))


def js_toint32(x: float) -> int:
    """~~x."""
    return lib().orc_js_toint32(float(x))


# ---------------------------------------------------------------- solves
def affine_from_triangles(src, dst) -> np.ndarray:
    """affineMatrixFromTriangles, H.js:1265 -> Float32Array(6)."""
    s = np.ascontiguousarray(src, dtype=np.float64).reshape(-1)
    d = np.ascontiguousarray(dst, dtype=np.float64).reshape(-1)
    out = np.empty(6, np.float32)
    lib().orc_affine_from_triangles(_p(s, _f64p), _p(d, _f64p), _p(out, _f32p))
    return out


def inverse_affine(m) -> np.ndarray:
    """inverseAffineMatrix, H.js:1345 -> Float32Array(6)."""
    a = np.ascontiguousarray(m, dtype=np.float32).reshape(-1)
    out = np.empty(6, np.float32)
    lib().orc_inverse_affine(_p(a, _f32p), _p(out, _f32p))
    return out


def projective_from_squares(src, dst) -> np.ndarray:
    """projectiveMatrixFromSquares + numeric.js solve, H.js:1320/1650 -> Array(8) of f64."""
    s = np.ascontiguousarray(src, dtype=np.float64).reshape(-1)
    d = np.ascontiguousarray(dst, dtype=np.float64).reshape(-1)
    out = np.empty(8, np.float64)
    lib().orc_projective_from_squares(_p(s, _f64p), _p(d, _f64p), _p(out, _f64p))
    return out


def calculate_transform_matrix(transform: str, src, dst) -> np.ndarray:
    """calculateTransformMatrix, H.js:1237."""
    if transform == "affine":
        return affine_from_triangles(src, dst)
    if transform == "projective":
        return projective_from_squares(src, dst)
    raise ValueError(f"{transform} transform does not exist")


def transform_limits(matrix, width, height):
    """calculateTransformLimits, H.js:1503 -> [xOff, yOff, oW, oH] (floats; may be NaN)."""
    m = np.ascontiguousarray(matrix)
    out = np.empty(4, np.float64)
    if m.size == 6:
        m = m.astype(np.float32)
        lib().orc_transform_limits(0, m.ctypes.data, float(width), float(height), _p(out, _f64p))
    elif m.size == 8:
        m = m.astype(np.float64)
        lib().orc_transform_limits(1, m.ctypes.data, float(width), float(height), _p(out, _f64p))
    else:
        raise ValueError(f"Transform matrix have an incorrect shape --> {m.size}")
    return out


def minmax_xy(points, rounded=True):
    """minmaxXYofArray, H.js:1558 -> [minX, minY, maxX, maxY]."""
    a = np.ascontiguousarray(points, dtype=np.float64).reshape(-1)
    out = np.empty(4, np.float64)
    lib().orc_minmax_xy(_p(a, _f64p), a.size, int(bool(rounded)), _p(out, _f64p))
    return out


# ---------------------------------------------------------------- index map
def build_index_map(points, triangles, matrix_width, y_offset, length) -> np.ndarray:
    """_build(Inverse)TrianglesCorrespondencesMatrix + fillTriangle, H.js:817/845/1111 -> Int16Array."""
    p = np.ascontiguousarray(points, dtype=np.float32).reshape(-1)
    t = np.ascontiguousarray(triangles, dtype=np.uint32).reshape(-1)
    m = np.empty(max(int(length), 0), np.int16)
    lib().orc_build_index_map(_p(p, _f32p), _p(t, _u32p), t.size // 3, float(matrix_width), float(y_offset),
                              _p(m, _i16p), m.size)
    return m


def piecewise_matrices(src_points, dst_points, triangles) -> np.ndarray:
    """_calculatePiecewiseAffineTransformMatrices, H.js:785 -> (T,6) float32."""
    s = np.ascontiguousarray(src_points, dtype=np.float32).reshape(-1)
    d = np.ascontiguousarray(dst_points, dtype=np.float32).reshape(-1)
    t = np.ascontiguousarray(triangles, dtype=np.uint32).reshape(-1)
    out = np.empty((t.size // 3, 6), np.float32)
    lib().orc_piecewise_matrices(_p(s, _f32p), _p(d, _f32p), _p(t, _u32p), t.size // 3, _p(out, _f32p))
    return out


def inverse_matrices(mats) -> np.ndarray:
    m = np.ascontiguousarray(mats, dtype=np.float32).reshape(-1, 6)
    return np.stack([inverse_affine(r) for r in m]) if len(m) else m.copy()


# ---------------------------------------------------------------- warp loops
def warp_inverse_geometric(image, W, H, inv, xOff, yOff, oW, oH, threads=1) -> np.ndarray:
    """_inverseGeometricWarp pixel loop, H.js:987 (inv = the dst->src matrix)."""
    img = _img(image)
    inv = np.ascontiguousarray(inv)
    kind = 0 if inv.size == 6 else 1
    inv = inv.astype(np.float32 if kind == 0 else np.float64)
    out = np.zeros(max(oW, 0) * max(oH, 0) * 4, np.uint8)
    lib().orc_warp_inverse_geometric(kind, _p(img, _u8p), W, H, inv.ctypes.data, xOff, yOff, oW, oH,
                                     _p(out, _u8p), threads)
    return out


def warp_inverse_geometric_bilinear(image, W, H, inv, xOff, yOff, oW, oH, threads=1) -> np.ndarray:
    """EXTENSION (no counterpart in the reference): bilinear sampling, see hg_oracle.c."""
    img = _img(image)
    inv = np.ascontiguousarray(inv)
    kind = 0 if inv.size == 6 else 1
    inv = inv.astype(np.float32 if kind == 0 else np.float64)
    out = np.zeros(max(oW, 0) * max(oH, 0) * 4, np.uint8)
    lib().orc_warp_inverse_geometric_bilinear(kind, _p(img, _u8p), W, H, inv.ctypes.data, xOff, yOff, oW, oH,
                                              _p(out, _u8p), threads)
    return out


def warp_forward_geometric(image, W, H, fwd, xOff, yOff, oW, oH) -> np.ndarray:
    """_geometricWarp, H.js:911."""
    img = _img(image)
    fwd = np.ascontiguousarray(fwd)
    kind = 0 if fwd.size == 6 else 1
    fwd = fwd.astype(np.float32 if kind == 0 else np.float64)
    out = np.zeros(max(oW, 0) * max(oH, 0) * 4, np.uint8)
    lib().orc_warp_forward_geometric(kind, _p(img, _u8p), W, H, fwd.ctypes.data, xOff, yOff, oW, oH,
                                     _p(out, _u8p))
    return out


def warp_inverse_piecewise(image, W, H, index_map, inv_mats, xOff, yOff, oW, oH, minSrcX, minSrcY,
                           threads=1) -> np.ndarray:
    """_inversePiecewiseAffineWarp pixel loop, H.js:1042."""
    img = _img(image)
    m = np.ascontiguousarray(index_map, dtype=np.int16)
    inv = np.ascontiguousarray(inv_mats, dtype=np.float32).reshape(-1)
    out = np.zeros(max(oW, 0) * max(oH, 0) * 4, np.uint8)
    lib().orc_warp_inverse_piecewise(_p(img, _u8p), W, H, _p(m, _i16p), m.size, _p(inv, _f32p), inv.size // 6,
                                     xOff, yOff, oW, oH, minSrcX, minSrcY, _p(out, _u8p), threads)
    return out


def warp_forward_piecewise(image, W, H, index_map, fwd_mats, xOff, yOff, oW, oH, minSrcX, minSrcY, maxSrcX,
                           maxSrcY) -> np.ndarray:
    """_piecewiseAffineWarp, H.js:948."""
    img = _img(image)
    m = np.ascontiguousarray(index_map, dtype=np.int16)
    fwd = np.ascontiguousarray(fwd_mats, dtype=np.float32).reshape(-1)
    out = np.zeros(max(oW, 0) * max(oH, 0) * 4, np.uint8)
    lib().orc_warp_forward_piecewise(_p(img, _u8p), W, H, _p(m, _i16p), m.size, _p(fwd, _f32p), fwd.size // 6,
                                     xOff, yOff, oW, oH, minSrcX, minSrcY, maxSrcX, maxSrcY, _p(out, _u8p))
    return out


def max_threads() -> int:
    return lib().orc_max_threads()
