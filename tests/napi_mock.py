"""Runs the Node-API addon (homography.js_b200/js/hgwarp_napi.c) WITHOUT Node.js: the addon is linked with a miniature
Node-API runtime (tests/napi_mock/napi_mock.c) into one shared library, and `Native` below plays the role of
`require('./hgwarp.node')` — same export names, JS numbers and typed arrays in, typed arrays / objects out.

`NapiContext` adapts that to the engine interface the Python twin of the class (homography.py) drives, making the call
sequence  class surface -> native.* -> N-API marshalling -> C ABI -> CUDA  executable end to end.  Test infrastructure only."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "homography.js_b200")

# napi_typedarray_type (js/napi_min.h, same order as node_api.h)
TYPES = {np.dtype(np.int8): 0, np.dtype(np.uint8): 1, np.dtype(np.int16): 3, np.dtype(np.uint16): 4, np.dtype(np.int32): 5,
         np.dtype(np.uint32): 6, np.dtype(np.float32): 7, np.dtype(np.float64): 8}
DTYPES = {0: np.int8, 1: np.uint8, 2: np.uint8, 3: np.int16, 4: np.uint16, 5: np.int32, 6: np.uint32, 7: np.float32, 8: np.float64}
V_UNDEFINED, V_NUMBER, V_EXTERNAL, V_OBJECT, V_ARRAYBUFFER, V_TYPEDARRAY, V_FUNCTION = range(7)

_LIBS = {}


class JsError(Exception):
    """A JavaScript exception thrown by the addon (napi_throw_error)."""


class Clamped(np.ndarray):
    """Marks a uint8 array as a Uint8ClampedArray (what ImageData.data is)."""


def build(cpu_double: bool = False):
    """gcc: addon + mock runtime -> one shared library in a temp dir, linked against the in-tree libhgwarp.so.
    cpu_double=True additionally compiles tests/napi_mock/hgwarp_cpu_double.c INTO the library: the GPU entry points the addon
    binds are then answered by the CPU oracle (bound locally with -Bsymbolic), so the marshalling can be checked without a GPU;
    the host-only entry points (hg_delaunay, hg_png_*) still come from libhgwarp.so."""
    if cpu_double in _LIBS:
        return _LIBS[cpu_double]
    out = os.path.join(tempfile.mkdtemp(prefix="hgnapi_"), "libhgnapi_mock.so")
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    extra = []
    if cpu_double:
        from oracle import oracle as O
        O.build()
        odir = os.path.join(ROOT, "oracle")
        extra = [os.path.join(ROOT, "tests", "napi_mock", "hgwarp_cpu_double.c"), "-Wl,-Bsymbolic", "-L" + odir, "-lhgoracle",
                 "-Wl,-rpath," + odir]
    subprocess.check_call([gcc, "-std=c11", "-O1", "-g", "-Wall", "-Wextra", "-Werror", "-fPIC", "-shared", "-fvisibility=hidden",
                           "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "napi_mock", "napi_mock.c"),
                           os.path.join(PKG, "js", "hgwarp_napi.c")] + extra +
                          ["-L" + PKG, "-lhgwarp", "-Wl,-rpath," + PKG, "-lm", "-o", out], env=env)
    L = C.CDLL(out)
    vp = C.c_void_p
    L.mock_env_create.restype = vp
    L.mock_env_destroy.argtypes = [vp]
    L.mock_module_name.restype = C.c_char_p
    L.mock_export_count.argtypes = [vp]
    L.mock_export_name.argtypes = [vp, C.c_int]
    L.mock_export_name.restype = C.c_char_p
    L.mock_number.argtypes = [vp, C.c_double]
    L.mock_number.restype = vp
    L.mock_undefined.argtypes = [vp]
    L.mock_undefined.restype = vp
    L.mock_typedarray.argtypes = [vp, C.c_int, vp, C.c_size_t]
    L.mock_typedarray.restype = vp
    L.mock_call.argtypes = [vp, C.c_char_p, C.c_int, C.POINTER(vp)]
    L.mock_call.restype = vp
    L.mock_exception.argtypes = [vp]
    L.mock_exception.restype = C.c_char_p
    L.mock_kind.argtypes = [vp]
    L.mock_get_number.argtypes = [vp]
    L.mock_get_number.restype = C.c_double
    L.mock_typedarray_get.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.POINTER(vp)]
    L.mock_get_property.argtypes = [vp, C.c_char_p]
    L.mock_get_property.restype = vp
    _LIBS[cpu_double] = L
    return L


class External:
    def __init__(self, handle):
        self.handle = handle


class Native:
    """`const native = require('./hgwarp.node')` under the mock runtime: native.<export>(...args)."""

    def __init__(self, cpu_double: bool = False):
        self.L = build(cpu_double)
        self.env = self.L.mock_env_create()
        assert self.env, "the addon did not register a module"
        self.exports = [self.L.mock_export_name(self.env, i).decode() for i in range(self.L.mock_export_count(self.env))]

    def close(self):
        if self.env:
            self.L.mock_env_destroy(self.env)   # runs the externals' finalizers (hg_ctx_destroy), like the GC would
            self.env = None

    def _to_js(self, a, keep):
        L = self.L
        if a is None:
            return L.mock_undefined(self.env)
        if isinstance(a, External):
            return a.handle
        if isinstance(a, (bool, int, float, np.integer, np.floating)):
            return L.mock_number(self.env, float(a))
        if isinstance(a, np.ndarray):
            arr = np.ascontiguousarray(a)
            keep.append(arr)
            t = 2 if isinstance(a, Clamped) else TYPES[arr.dtype]
            return L.mock_typedarray(self.env, t, arr.ctypes.data, arr.size)
        raise TypeError(f"no JS value for {type(a)}")

    def _from_js(self, v):
        L = self.L
        k = L.mock_kind(v)
        if k == V_UNDEFINED:
            return None
        if k == V_NUMBER:
            return L.mock_get_number(v)
        if k == V_EXTERNAL:
            return External(v)
        if k == V_TYPEDARRAY:
            t, n, p = C.c_int(), C.c_size_t(), C.c_void_p()
            assert L.mock_typedarray_get(v, C.byref(t), C.byref(n), C.byref(p)) == 0
            dt = np.dtype(DTYPES[t.value])
            if n.value == 0:
                return np.empty(0, dt)
            return np.frombuffer(C.string_at(p.value, n.value * dt.itemsize), dtype=dt).copy()
        if k == V_OBJECT:
            out = {}
            for name in ("data", "width", "height", "matrix", "limits"):
                q = L.mock_get_property(v, name.encode())
                if q:
                    out[name] = self._from_js(q)
            return out
        raise TypeError(f"unexpected JS value kind {k}")

    def call(self, name, *args):
        keep = []
        argv = (C.c_void_p * max(len(args), 1))(*[self._to_js(a, keep) for a in args])
        r = self.L.mock_call(self.env, name.encode(), len(args), argv)
        if not r:
            raise JsError((self.L.mock_exception(self.env) or b"?").decode())
        return self._from_js(r)

    def __getattr__(self, name):
        if name.startswith("_") or name in ("L", "env", "exports"):
            raise AttributeError(name)
        return lambda *a: self.call(name, *a)


class NapiContext:
    """The engine interface homography.py drives (_abi.Context's subset), answered through native.* — the calls
    js/homography_b200.mjs makes, with the same argument conversions (f64() / f32() / Uint32Array.from)."""

    def __init__(self, native: Native, device: int = 0):
        self.n = native
        self.ctx = native.createContext(device)
        self._last_solve = None

    def close(self):
        pass   # the external's finalizer destroys the context when the environment goes away

    @staticmethod
    def _f64(a):
        return np.ascontiguousarray(a, dtype=np.float64).reshape(-1)

    @staticmethod
    def _f32(a):
        return np.ascontiguousarray(a, dtype=np.float32).reshape(-1)

    def image_set(self, rgba, w, h):
        self.n.setImage(self.ctx, np.ascontiguousarray(rgba, dtype=np.uint8).reshape(-1).view(Clamped), w, h)

    def solve_with_limits(self, kind, src, dst, w, h):
        s, d = self._f64(src).copy(), self._f64(dst).copy()
        self._last_solve = (kind, s, d)
        r = self.n.solveWithLimits(self.ctx, kind, s, d, w, h)
        return r["matrix"], r["limits"]

    def solve_affine(self, src, dst):      # the shim's _solve(src, dst, withLimits = false): size 1 x 1, limits unused
        return self.solve_with_limits(0, src, dst, 1, 1)[0]

    def solve_projective(self, src, dst):
        return self.solve_with_limits(1, src, dst, 1, 1)[0]

    def transform_limits(self, matrix, w, h):
        """The shim has no separate limits call: _induceObjective() solves again with the current size (same points ->
        same matrix).  The points of the last solve are replayed; the matrix must come out identical."""
        kind, s, d = self._last_solve
        r = self.n.solveWithLimits(self.ctx, kind, s, d, w, h)
        m = np.asarray(matrix)
        assert np.array_equal(r["matrix"].view(np.uint8), m.astype(r["matrix"].dtype).view(np.uint8))
        return r["limits"]

    def warp_inverse_points(self, kind, dst_pts, src_pts, x_off, y_off, o_w, o_h, **kw):
        return self.n.warpInversePoints(self.ctx, kind, self._f64(dst_pts), self._f64(src_pts), x_off, y_off, o_w, o_h)

    def warp_forward_matrix(self, fwd, x_off, y_off, o_w, o_h, **kw):
        m = np.asarray(fwd)
        kind = 0 if m.size == 6 else 1
        return self.n.warpForwardMatrix(self.ctx, kind, self._f32(m) if kind == 0 else self._f64(m), x_off, y_off, o_w, o_h)

    def piecewise_set_mesh(self, src_pts, tris):
        self._n_tris = int(np.asarray(tris).size // 3)
        self.n.setMesh(self.ctx, self._f32(src_pts), np.ascontiguousarray(tris, dtype=np.uint32).reshape(-1))

    def piecewise_matrices(self, dst_pts, want_inverse=False):
        assert not want_inverse
        return self.n.piecewiseMatrices(self.ctx, self._f32(dst_pts), self._n_tris).reshape(-1, 6)

    def warp_piecewise_inverse(self, dst_pts, x_off, y_off, o_w, o_h, min_src_x, min_src_y, **kw):
        return self.n.warpPiecewiseInverse(self.ctx, self._f32(dst_pts), x_off, y_off, o_w, o_h, min_src_x, min_src_y)

    def warp_piecewise_forward(self, dst_pts, x_off, y_off, o_w, o_h, min_src_x, min_src_y, max_src_x, max_src_y,
                               use_inverse_map=False, **kw):
        return self.n.warpPiecewiseForward(self.ctx, self._f32(dst_pts), x_off, y_off, o_w, o_h, min_src_x, min_src_y,
                                           max_src_x, max_src_y, 1 if use_inverse_map else 0)
