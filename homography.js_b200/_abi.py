"""ctypes binding of libhgwarp.so — the same symbols an N-API addon binds (include/hgwarp.h).

There is deliberately NO fallback: if the CUDA library is missing or no GPU is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HGWARP_LIB") or os.path.join(_HERE, "libhgwarp.so")  # env override: A/B builds

HG_OK, HG_ERR_INVALID, HG_ERR_CUDA, HG_ERR_NOMEM, HG_ERR_UNSUPPORTED, HG_ERR_STATE = range(6)
HG_AFFINE, HG_PROJECTIVE = 0, 1
HG_NEAREST, HG_BILINEAR = 0, 1

# every symbol include/hgwarp.h declares (tests check the .so exports exactly these)
SYMBOLS = [
    "hg_abi_version", "hg_device_count", "hg_ctx_create", "hg_ctx_destroy", "hg_last_error",
    "hg_ctx_synchronize", "hg_ctx_stream", "hg_timer_start", "hg_timer_stop", "hg_ctx_set_sampling", "hg_launch_count",
    "hg_profile_enable", "hg_profile_read",
    "hg_image_set", "hg_image_set_device",
    "hg_solve_affine", "hg_solve_projective", "hg_inverse_affine", "hg_transform_limits", "hg_solve_with_limits",
    "hg_warp_inverse_matrix", "hg_warp_inverse_points", "hg_warp_forward_matrix",
    "hg_delaunay", "hg_png_decode", "hg_jpeg_decode", "hg_png_encode", "hg_png_encode_bound",
    "hg_piecewise_set_mesh", "hg_piecewise_mesh_size", "hg_piecewise_matrices", "hg_piecewise_extents", "hg_build_index_map",
    "hg_warp_piecewise_inverse", "hg_warp_piecewise_forward",
    "hg_warp_inverse_batch", "hg_warp_piecewise_inverse_batch",
    "hg_warp_inverse_points_batch", "hg_warp_forward_batch", "hg_warp_piecewise_forward_batch",
    "hg_warp_piecewise_stream", "hg_stream_slot_bytes", "hg_checksum_frames",
    "hg_pipe_create", "hg_pipe_submit", "hg_pipe_wait", "hg_pipe_flush", "hg_pipe_destroy",
    "hg_pipe_create_piecewise", "hg_pipe_submit_piecewise",
    "hg_host_alloc_pinned_ex", "hg_pcie_probe",
    "hg_debug_rcp_max_error", "hg_debug_quotient_at_least", "hg_debug_force_general", "hg_debug_piecewise_binning", "hg_debug_piecewise_stats",
    "hg_dev_alloc", "hg_dev_free", "hg_host_alloc_pinned", "hg_host_free_pinned",
    "hg_memcpy_h2d", "hg_memcpy_d2h", "hg_output_device",
]


class HgFrame(C.Structure):
    _fields_ = [("src_dev", C.c_void_p), ("out_dev", C.c_void_p),
                ("src_w", C.c_int32), ("src_h", C.c_int32),
                ("x_off", C.c_int32), ("y_off", C.c_int32), ("o_w", C.c_int32), ("o_h", C.c_int32)]


class HgStreamInfo(C.Structure):
    _fields_ = [("x_off", C.c_int32), ("y_off", C.c_int32), ("o_w", C.c_int32), ("o_h", C.c_int32),
                ("slot", C.c_int32), ("status", C.c_int32)]


CHECKSUM_MUL = 2654435761
CHECKSUM_LEN_MUL = 0x9E3779B97F4A7C15


def checksum_reference(rgba_bytes: np.ndarray) -> int:
    """The checksum hg_checksum_frames computes, in numpy (for comparing device frames with oracle frames):
    sum_i pixel_i * (((i * 2654435761) mod 2^32) | 1) + n * 0x9E3779B97F4A7C15 (mod 2^64), pixel_i = little-endian RGBA8 word."""
    px = np.ascontiguousarray(rgba_bytes, dtype=np.uint8).reshape(-1).view("<u4").astype(np.uint64)
    n = px.size
    w = ((np.arange(n, dtype=np.uint64) * np.uint64(CHECKSUM_MUL)) & np.uint64(0xFFFFFFFF)) | np.uint64(1)
    with np.errstate(over="ignore"):
        return (int((px * w).sum(dtype=np.uint64)) + n * CHECKSUM_LEN_MUL) & 0xFFFFFFFFFFFFFFFF


class HgError(RuntimeError):
    def __init__(self, status, text):
        super().__init__(f"hgwarp status {status}: {text}")
        self.status = status
        self.text = text


_lib = None


def load():
    """dlopen libhgwarp.so; raises if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` "
                           "(nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i, d = C.c_void_p, C.c_int, C.c_double
    L.hg_abi_version.restype = i
    L.hg_device_count.argtypes = [C.POINTER(i)]
    L.hg_ctx_create.argtypes = [i, C.POINTER(vp)]
    L.hg_ctx_destroy.argtypes = [vp]
    L.hg_last_error.argtypes = [vp]
    L.hg_last_error.restype = C.c_char_p
    L.hg_ctx_synchronize.argtypes = [vp]
    L.hg_ctx_stream.argtypes = [vp, C.POINTER(vp)]
    L.hg_timer_start.argtypes = [vp]
    L.hg_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
    L.hg_ctx_set_sampling.argtypes = [vp, i]
    L.hg_launch_count.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.hg_profile_enable.argtypes = [vp, i]
    L.hg_profile_read.argtypes = [vp, C.POINTER(d), C.POINTER(C.c_uint64)]
    L.hg_image_set.argtypes = [vp, vp, i, i]
    L.hg_image_set_device.argtypes = [vp, vp, i, i]
    L.hg_solve_affine.argtypes = [vp, vp, vp, vp]
    L.hg_solve_projective.argtypes = [vp, vp, vp, vp]
    L.hg_inverse_affine.argtypes = [vp, vp, vp]
    L.hg_transform_limits.argtypes = [vp, i, vp, d, d, vp]
    L.hg_solve_with_limits.argtypes = [vp, i, vp, vp, d, d, vp, vp]
    L.hg_warp_inverse_matrix.argtypes = [vp, i, vp, i, i, i, i, vp, vp]
    L.hg_warp_inverse_points.argtypes = [vp, i, vp, vp, i, i, i, i, vp, vp]
    L.hg_warp_forward_matrix.argtypes = [vp, i, vp, i, i, i, i, vp, vp]
    L.hg_piecewise_set_mesh.argtypes = [vp, vp, i, vp, i]
    L.hg_piecewise_mesh_size.argtypes = [vp, C.POINTER(i), C.POINTER(i)]
    L.hg_piecewise_matrices.argtypes = [vp, vp, vp, vp]
    L.hg_piecewise_extents.argtypes = [vp, vp, i, i, vp]
    L.hg_build_index_map.argtypes = [vp, vp, d, d, C.c_int64, vp]
    L.hg_warp_piecewise_inverse.argtypes = [vp, vp, i, i, i, i, i, i, vp, vp]
    L.hg_warp_piecewise_forward.argtypes = [vp, vp, i, i, i, i, i, i, i, i, i, vp, vp]
    L.hg_warp_inverse_batch.argtypes = [vp, i, vp, C.POINTER(HgFrame), i]
    L.hg_warp_piecewise_inverse_batch.argtypes = [vp, vp, C.POINTER(HgFrame), i, i, i]
    L.hg_warp_inverse_points_batch.argtypes = [vp, i, vp, vp, C.POINTER(HgFrame), i]
    L.hg_warp_forward_batch.argtypes = [vp, i, vp, C.POINTER(HgFrame), i]
    L.hg_warp_piecewise_forward_batch.argtypes = [vp, vp, C.POINTER(HgFrame), i, i, i, i, i]
    L.hg_warp_piecewise_stream.argtypes = [vp, vp, i, C.c_int64, i, i, vp, i, i, i, vp, i, i, i, C.POINTER(HgStreamInfo)]
    L.hg_stream_slot_bytes.argtypes = [i, i]
    L.hg_stream_slot_bytes.restype = C.c_size_t
    L.hg_checksum_frames.argtypes = [vp, C.POINTER(HgFrame), i, vp]
    L.hg_pipe_create_piecewise.argtypes = [vp, i, i, i, i, i, C.POINTER(vp)]
    L.hg_pipe_submit_piecewise.argtypes = [vp, vp, vp, i, i, vp, vp, C.POINTER(C.c_uint64)]
    L.hg_host_alloc_pinned_ex.argtypes = [vp, C.c_size_t, i, C.POINTER(vp)]
    L.hg_pcie_probe.argtypes = [vp, C.c_size_t, i, C.POINTER(d), C.POINTER(d), C.POINTER(d)]
    L.hg_pipe_create.argtypes = [vp, i, i, i, i, i, i, C.POINTER(vp)]
    L.hg_pipe_submit.argtypes = [vp, vp, vp, vp, i, i, i, i, vp, C.POINTER(C.c_uint64)]
    L.hg_pipe_wait.argtypes = [vp, C.c_uint64]
    L.hg_pipe_flush.argtypes = [vp]
    L.hg_pipe_destroy.argtypes = [vp]
    L.hg_debug_force_general.argtypes = [vp, i]
    L.hg_debug_piecewise_binning.argtypes = [vp, i]
    L.hg_debug_piecewise_stats.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.hg_debug_rcp_max_error.argtypes = [vp, i, i, C.POINTER(d)]
    L.hg_debug_quotient_at_least.argtypes = [vp, vp, vp, vp, i, vp]
    L.hg_dev_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    L.hg_dev_free.argtypes = [vp, vp]
    L.hg_host_alloc_pinned.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    L.hg_host_free_pinned.argtypes = [vp, vp]
    L.hg_memcpy_h2d.argtypes = [vp, vp, vp, C.c_size_t]
    L.hg_memcpy_d2h.argtypes = [vp, vp, vp, C.c_size_t]
    L.hg_output_device.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.hg_delaunay.argtypes = [vp, i, vp, i, C.POINTER(i)]
    L.hg_png_decode.argtypes = [vp, C.c_size_t, vp, C.c_size_t, C.POINTER(i), C.POINTER(i)]
    L.hg_jpeg_decode.argtypes = [vp, C.c_size_t, vp, C.c_size_t, C.POINTER(i), C.POINTER(i)]
    L.hg_png_encode_bound.argtypes = [i, i]
    L.hg_png_encode_bound.restype = C.c_size_t
    L.hg_png_encode.argtypes = [vp, i, i, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    _lib = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data


class Context:
    """One GPU + one CUDA stream (hg_ctx)."""

    def __init__(self, device: int = 0):
        self.L = load()
        h = C.c_void_p()
        st = self.L.hg_ctx_create(device, C.byref(h))
        if st != HG_OK:
            raise HgError(st, (self.L.hg_last_error(None) or b"").decode())
        self.h = h
        self.device = device
        self._pinned = []

    def close(self):
        if getattr(self, "h", None):
            for ptr in getattr(self, "_pinned", []):
                self.L.hg_host_free_pinned(self.h, ptr)
            self._pinned = []
            self.L.hg_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st):
        if st != HG_OK:
            raise HgError(st, (self.L.hg_last_error(self.h) or b"").decode())

    # ---- plumbing
    def synchronize(self):
        self._ck(self.L.hg_ctx_synchronize(self.h))

    def stream(self) -> int:
        s = C.c_void_p()
        self._ck(self.L.hg_ctx_stream(self.h, C.byref(s)))
        return s.value or 0

    def timer_start(self):
        self._ck(self.L.hg_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        self._ck(self.L.hg_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def set_sampling(self, sampling: int):
        """HG_NEAREST (the reference's Math.round sampling) or HG_BILINEAR (extension, <= 1 LSB vs its oracle)."""
        self._ck(self.L.hg_ctx_set_sampling(self.h, sampling))

    def launch_count(self) -> int:
        n = C.c_uint64()
        self._ck(self.L.hg_launch_count(self.h, C.byref(n)))
        return n.value

    def profile_enable(self, on: bool = True):
        self._ck(self.L.hg_profile_enable(self.h, int(on)))

    def profile_read(self):
        ms, n = C.c_double(), C.c_uint64()
        self._ck(self.L.hg_profile_read(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def debug_rcp_max_error(self, biased_exponent: int = 1023, negative: bool = False) -> float:
        e = C.c_double()
        self._ck(self.L.hg_debug_rcp_max_error(self.h, biased_exponent, int(negative), C.byref(e)))
        return e.value

    def debug_quotient_at_least(self, N, D, b) -> np.ndarray:
        N, D, b = (np.ascontiguousarray(v, dtype=np.float64) for v in (N, D, b))
        out = np.zeros(N.size, np.int32)
        self._ck(self.L.hg_debug_quotient_at_least(self.h, _ptr(N), _ptr(D), _ptr(b), int(N.size), _ptr(out)))
        return out.astype(bool)

    def debug_piecewise_stats(self):
        a, b = C.c_uint64(), C.c_uint64()
        self._ck(self.L.hg_debug_piecewise_stats(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def debug_force_general(self, on: bool):
        self._ck(self.L.hg_debug_force_general(self.h, int(on)))

    def debug_piecewise_binning(self, mode: int):
        self._ck(self.L.hg_debug_piecewise_binning(self.h, int(mode)))

    def dev_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._ck(self.L.hg_dev_alloc(self.h, nbytes, C.byref(p)))
        return p.value

    def dev_free(self, p: int):
        self._ck(self.L.hg_dev_free(self.h, p))

    def host_alloc_pinned(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._ck(self.L.hg_host_alloc_pinned(self.h, nbytes, C.byref(p)))
        return p.value

    def host_alloc_pinned_ex(self, nbytes: int, write_combined: bool = False) -> int:
        p = C.c_void_p()
        self._ck(self.L.hg_host_alloc_pinned_ex(self.h, nbytes, int(write_combined), C.byref(p)))
        return p.value

    def pinned_array(self, nbytes: int, write_combined: bool = False) -> np.ndarray:
        """A uint8 numpy view over page-locked host memory owned by the library (freed by the context's close())."""
        ptr = self.host_alloc_pinned_ex(nbytes, write_combined)
        self._pinned.append(ptr)
        return np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(ptr))

    def pcie_probe(self, nbytes: int = 64 << 20, iters: int = 8):
        """(h2d, d2h, bidirectional) GB/s of raw pinned copies on this GPU's host link (hg_pcie_probe)."""
        a, b, cc = C.c_double(), C.c_double(), C.c_double()
        self._ck(self.L.hg_pcie_probe(self.h, nbytes, iters, C.byref(a), C.byref(b), C.byref(cc)))
        return a.value, b.value, cc.value

    def host_free_pinned(self, p: int):
        self._ck(self.L.hg_host_free_pinned(self.h, p))

    def memcpy_h2d(self, dst_dev: int, src_host: int, nbytes: int):
        self._ck(self.L.hg_memcpy_h2d(self.h, dst_dev, src_host, nbytes))

    def memcpy_d2h(self, dst_host: int, src_dev: int, nbytes: int):
        self._ck(self.L.hg_memcpy_d2h(self.h, dst_host, src_dev, nbytes))

    def output_device(self):
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self.L.hg_output_device(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    # ---- image
    def image_set(self, rgba: np.ndarray, w: int, h: int):
        a = np.ascontiguousarray(rgba, dtype=np.uint8).reshape(-1)
        if a.size != w * h * 4:
            raise ValueError(f"image buffer has {a.size} bytes, expected {w}x{h}x4")
        self._ck(self.L.hg_image_set(self.h, a.ctypes.data, w, h))

    def image_set_host_ptr(self, host_ptr: int, w: int, h: int):
        self._ck(self.L.hg_image_set(self.h, host_ptr, w, h))

    def image_set_device(self, dev_ptr: int, w: int, h: int):
        self._ck(self.L.hg_image_set_device(self.h, dev_ptr, w, h))

    # ---- solves
    def solve_affine(self, src, dst) -> np.ndarray:
        s = np.ascontiguousarray(src, dtype=np.float64).reshape(-1)
        d = np.ascontiguousarray(dst, dtype=np.float64).reshape(-1)
        assert s.size == 6 and d.size == 6
        out = np.empty(6, np.float32)
        self._ck(self.L.hg_solve_affine(self.h, _ptr(s), _ptr(d), _ptr(out)))
        return out

    def solve_projective(self, src, dst) -> np.ndarray:
        s = np.ascontiguousarray(src, dtype=np.float64).reshape(-1)
        d = np.ascontiguousarray(dst, dtype=np.float64).reshape(-1)
        assert s.size == 8 and d.size == 8
        out = np.empty(8, np.float64)
        self._ck(self.L.hg_solve_projective(self.h, _ptr(s), _ptr(d), _ptr(out)))
        return out

    def inverse_affine(self, m) -> np.ndarray:
        a = np.ascontiguousarray(m, dtype=np.float32).reshape(-1)
        assert a.size == 6
        out = np.empty(6, np.float32)
        self._ck(self.L.hg_inverse_affine(self.h, _ptr(a), _ptr(out)))
        return out

    def transform_limits(self, matrix, w, h) -> np.ndarray:
        m = np.ascontiguousarray(matrix)
        kind = HG_AFFINE if m.size == 6 else HG_PROJECTIVE
        m = m.astype(np.float32 if kind == HG_AFFINE else np.float64).reshape(-1)
        out = np.empty(4, np.float64)
        self._ck(self.L.hg_transform_limits(self.h, kind, _ptr(m), float(w), float(h), _ptr(out)))
        return out

    def solve_with_limits(self, kind, src, dst, w, h):
        s = np.ascontiguousarray(src, dtype=np.float64).reshape(-1)
        d = np.ascontiguousarray(dst, dtype=np.float64).reshape(-1)
        m = np.empty(6, np.float32) if kind == HG_AFFINE else np.empty(8, np.float64)
        lim = np.empty(4, np.float64)
        self._ck(self.L.hg_solve_with_limits(self.h, kind, _ptr(s), _ptr(d), float(w), float(h), _ptr(m), _ptr(lim)))
        return m, lim

    # ---- warps
    def warp_inverse_matrix(self, inv, x_off, y_off, o_w, o_h, to_host=True, out_dev=None):
        m = np.ascontiguousarray(inv)
        kind = HG_AFFINE if m.size == 6 else HG_PROJECTIVE
        m = m.astype(np.float32 if kind == HG_AFFINE else np.float64).reshape(-1)
        out = np.empty(o_w * o_h * 4, np.uint8) if to_host else None
        self._ck(self.L.hg_warp_inverse_matrix(self.h, kind, _ptr(m), x_off, y_off, o_w, o_h, _ptr(out), out_dev))
        return out

    def warp_inverse_points(self, kind, dst_pts, src_pts, x_off, y_off, o_w, o_h, to_host=True, out_dev=None,
                            out_host_ptr=None):
        d = np.ascontiguousarray(dst_pts, dtype=np.float64).reshape(-1)
        s = np.ascontiguousarray(src_pts, dtype=np.float64).reshape(-1)
        out = None
        hp = out_host_ptr
        if hp is None and to_host:
            out = np.empty(o_w * o_h * 4, np.uint8)
            hp = out.ctypes.data
        self._ck(self.L.hg_warp_inverse_points(self.h, kind, _ptr(d), _ptr(s), x_off, y_off, o_w, o_h, hp, out_dev))
        return out

    def warp_forward_matrix(self, fwd, x_off, y_off, o_w, o_h, to_host=True, out_dev=None):
        m = np.ascontiguousarray(fwd)
        kind = HG_AFFINE if m.size == 6 else HG_PROJECTIVE
        m = m.astype(np.float32 if kind == HG_AFFINE else np.float64).reshape(-1)
        out = np.empty(o_w * o_h * 4, np.uint8) if to_host else None
        self._ck(self.L.hg_warp_forward_matrix(self.h, kind, _ptr(m), x_off, y_off, o_w, o_h, _ptr(out), out_dev))
        return out

    def warp_inverse_batch(self, kind, inv_matrices: np.ndarray, frames):
        m = np.ascontiguousarray(inv_matrices, dtype=np.float32 if kind == HG_AFFINE else np.float64)
        arr = (HgFrame * len(frames))(*frames)
        self._ck(self.L.hg_warp_inverse_batch(self.h, kind, _ptr(m), arr, len(frames)))

    def warp_inverse_points_batch(self, kind, dst_pts: np.ndarray, src_pts: np.ndarray, frames):
        """Per-frame solve + pixel loop for a batch (hg_warp_inverse_points_batch); points: (n_frames, 6 | 8) doubles."""
        dd = np.ascontiguousarray(dst_pts, dtype=np.float64)
        ss = np.ascontiguousarray(src_pts, dtype=np.float64)
        arr = frames if isinstance(frames, C.Array) else (HgFrame * len(frames))(*frames)
        self._ck(self.L.hg_warp_inverse_points_batch(self.h, kind, _ptr(dd), _ptr(ss), arr, len(arr)))

    def warp_forward_batch(self, kind, fwd_matrices: np.ndarray, frames):
        m = np.ascontiguousarray(fwd_matrices, dtype=np.float32 if kind == HG_AFFINE else np.float64)
        arr = frames if isinstance(frames, C.Array) else (HgFrame * len(frames))(*frames)
        self._ck(self.L.hg_warp_forward_batch(self.h, kind, _ptr(m), arr, len(arr)))

    def warp_piecewise_forward_batch(self, dst_pts, frames, min_src_x, min_src_y, max_src_x, max_src_y):
        dd = np.ascontiguousarray(dst_pts, dtype=np.float32).reshape(-1)
        arr = frames if isinstance(frames, C.Array) else (HgFrame * len(frames))(*frames)
        self._ck(self.L.hg_warp_piecewise_forward_batch(self.h, _ptr(dd), arr, len(arr), min_src_x, min_src_y, max_src_x, max_src_y))

    def checksum_frames(self, frames) -> np.ndarray:
        arr = frames if isinstance(frames, C.Array) else (HgFrame * len(frames))(*frames)
        out = np.zeros(len(arr), np.uint64)
        self._ck(self.L.hg_checksum_frames(self.h, arr, len(arr), _ptr(out)))
        return out

    def stream_slot_bytes(self, max_out_w: int, max_out_h: int) -> int:
        return int(self.L.hg_stream_slot_bytes(max_out_w, max_out_h))

    def warp_piecewise_stream(self, dst_pts, first_frame, min_src_x, min_src_y, out_ring_dev, n_slots, max_out_w, max_out_h,
                              src_ring_dev=None, n_src=0, src_w=0, src_h=0, info=None):
        """hg_warp_piecewise_stream: dst_pts (n_frames, n_pts, 2) float32; returns the HgStreamInfo array."""
        dd = np.ascontiguousarray(dst_pts, dtype=np.float32)
        n = dd.shape[0]
        if info is None:
            info = (HgStreamInfo * n)()
        self._ck(self.L.hg_warp_piecewise_stream(self.h, _ptr(dd), n, int(first_frame), min_src_x, min_src_y, src_ring_dev, n_src,
                                                 src_w, src_h, out_ring_dev, n_slots, max_out_w, max_out_h, info))
        return info

    # ---- piecewise
    def piecewise_set_mesh(self, src_pts, tris):
        p = np.ascontiguousarray(src_pts, dtype=np.float32).reshape(-1)
        t = np.ascontiguousarray(tris, dtype=np.uint32).reshape(-1)
        self._n_tris = t.size // 3
        self._n_pts = p.size // 2
        self._ck(self.L.hg_piecewise_set_mesh(self.h, _ptr(p), p.size // 2, _ptr(t), t.size // 3))

    def piecewise_matrices(self, dst_pts, want_inverse=False):
        d = np.ascontiguousarray(dst_pts, dtype=np.float32).reshape(-1)
        fwd = np.empty((self._n_tris, 6), np.float32)
        inv = np.empty((self._n_tris, 6), np.float32) if want_inverse else None
        self._ck(self.L.hg_piecewise_matrices(self.h, _ptr(d), _ptr(fwd), _ptr(inv)))
        return (fwd, inv) if want_inverse else fwd

    def piecewise_extents(self, dst_pts) -> np.ndarray:
        """dst_pts: (n_frames, n_pts, 2) float32 -> (n_frames, 4) [xOff, yOff, oW, oH] (H.js:706-710)."""
        d = np.ascontiguousarray(dst_pts, dtype=np.float32)
        if d.ndim == 2:
            d = d[None]
        out = np.empty((d.shape[0], 4), np.float64)
        self._ck(self.L.hg_piecewise_extents(self.h, _ptr(d), d.shape[1], d.shape[0], _ptr(out)))
        return out

    def build_index_map(self, pts, map_width, y_offset, map_len) -> np.ndarray:
        p = np.ascontiguousarray(pts, dtype=np.float32).reshape(-1)
        out = np.empty(max(int(map_len), 0), np.int16)
        self._ck(self.L.hg_build_index_map(self.h, _ptr(p), float(map_width), float(y_offset), int(map_len), _ptr(out)))
        return out

    def warp_piecewise_inverse(self, dst_pts, x_off, y_off, o_w, o_h, min_src_x, min_src_y, to_host=True,
                               out_dev=None, out_host_ptr=None):
        d = np.ascontiguousarray(dst_pts, dtype=np.float32).reshape(-1)
        out = None
        hp = out_host_ptr
        if hp is None and to_host:
            out = np.empty(o_w * o_h * 4, np.uint8)
            hp = out.ctypes.data
        self._ck(self.L.hg_warp_piecewise_inverse(self.h, _ptr(d), x_off, y_off, o_w, o_h, min_src_x, min_src_y,
                                                  hp, out_dev))
        return out

    def warp_piecewise_forward(self, dst_pts, x_off, y_off, o_w, o_h, min_src_x, min_src_y, max_src_x, max_src_y,
                               use_inverse_map=False, to_host=True, out_dev=None):
        d = np.ascontiguousarray(dst_pts, dtype=np.float32).reshape(-1)
        out = np.empty(o_w * o_h * 4, np.uint8) if to_host else None
        self._ck(self.L.hg_warp_piecewise_forward(self.h, _ptr(d), x_off, y_off, o_w, o_h, min_src_x, min_src_y,
                                                  max_src_x, max_src_y, int(use_inverse_map), _ptr(out), out_dev))
        return out

    def warp_piecewise_inverse_batch(self, dst_pts, frames, min_src_x, min_src_y):
        d = np.ascontiguousarray(dst_pts, dtype=np.float32).reshape(-1)
        arr = frames if isinstance(frames, C.Array) else (HgFrame * len(frames))(*frames)
        self._ck(self.L.hg_warp_piecewise_inverse_batch(self.h, _ptr(d), arr, len(frames), min_src_x, min_src_y))


class Pipe:
    """hg_pipe: pipelined host-to-host stream of independent inverse warps (video use case)."""

    def __init__(self, ctx: Context, kind: int, src_w: int, src_h: int, max_out_w: int, max_out_h: int, depth: int = 3):
        self.ctx, self.L = ctx, ctx.L
        h = C.c_void_p()
        ctx._ck(self.L.hg_pipe_create(ctx.h, kind, src_w, src_h, max_out_w, max_out_h, depth, C.byref(h)))
        self.h = h
        self.npts = 6 if kind == HG_AFFINE else 8

    def submit(self, rgba_host_ptr: int, dst_pts, src_pts, x_off, y_off, o_w, o_h, out_host_ptr: int) -> int:
        d = np.ascontiguousarray(dst_pts, dtype=np.float64).reshape(-1)
        s = np.ascontiguousarray(src_pts, dtype=np.float64).reshape(-1)
        assert d.size == self.npts and s.size == self.npts
        t = C.c_uint64()
        self.ctx._ck(self.L.hg_pipe_submit(self.h, rgba_host_ptr, d.ctypes.data, s.ctypes.data, x_off, y_off, o_w, o_h,
                                           out_host_ptr, C.byref(t)))
        return t.value

    def wait(self, ticket: int):
        self.ctx._ck(self.L.hg_pipe_wait(self.h, ticket))

    @classmethod
    def piecewise(cls, ctx: "Context", src_w: int, src_h: int, max_out_w: int, max_out_h: int, depth: int = 3) -> "Pipe":
        """hg_pipe_create_piecewise: the same pipeline for inverse piecewise frames of the context mesh."""
        self = cls.__new__(cls)
        self.ctx, self.L = ctx, ctx.L
        h = C.c_void_p()
        ctx._ck(self.L.hg_pipe_create_piecewise(ctx.h, src_w, src_h, max_out_w, max_out_h, depth, C.byref(h)))
        self.h = h
        self.npts = 0
        return self

    def submit_piecewise(self, rgba_host_ptr, dst_pts, min_src_x, min_src_y, out_host_ptr: int):
        """Returns (ticket, (x_off, y_off, o_w, o_h)); rgba_host_ptr None = the context image."""
        d = np.ascontiguousarray(dst_pts, dtype=np.float32).reshape(-1)
        t = C.c_uint64()
        win = (C.c_int32 * 4)()
        self.ctx._ck(self.L.hg_pipe_submit_piecewise(self.h, rgba_host_ptr, d.ctypes.data, min_src_x, min_src_y, out_host_ptr,
                                                      win, C.byref(t)))
        return t.value, tuple(win)

    def flush(self):
        self.ctx._ck(self.L.hg_pipe_flush(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.L.hg_pipe_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def delaunay(points) -> np.ndarray:
    """hg_delaunay: `new Delaunator(points).triangles` (H.js:1216) — host only, needs no GPU.  points: (n, 2) or flat
    [x0, y0, ...]; returns a flat uint32 array, three vertex ids per triangle."""
    pts = np.ascontiguousarray(np.asarray(points, dtype=np.float64).reshape(-1))
    n = pts.size // 2
    out = np.zeros(3 * max(2 * n - 5, 0), np.uint32)
    cnt = C.c_int()
    st = load().hg_delaunay(_ptr(pts), n, _ptr(out) if out.size else None, out.size // 3, C.byref(cnt))
    if st:
        raise HgError(st, "hg_delaunay failed")
    return out[:3 * cnt.value].copy()


def png_decode(data: bytes) -> np.ndarray:
    """hg_png_decode: PNG file bytes -> (h, w, 4) uint8 RGBA, the layout of ImageData.data.  Host only."""
    buf = np.frombuffer(data, dtype=np.uint8)
    w, h = C.c_int(), C.c_int()
    L = load()
    if L.hg_png_decode(_ptr(buf), buf.size, None, 0, C.byref(w), C.byref(h)):
        raise HgError(1, "hg_png_decode: not a PNG this decoder supports")
    out = np.empty((h.value, w.value, 4), np.uint8)
    if L.hg_png_decode(_ptr(buf), buf.size, _ptr(out), out.size, C.byref(w), C.byref(h)):
        raise HgError(1, "hg_png_decode: malformed image data")
    return out


def jpeg_decode(data: bytes) -> np.ndarray:
    """hg_jpeg_decode: JPEG file bytes (baseline or progressive) -> (h, w, 4) uint8 RGBA (alpha 255), the bytes getImageData returns.  Host only.
    Raises HgError with status HG_ERR_UNSUPPORTED for CMYK / arithmetic-coded / lossless files."""
    buf = np.frombuffer(data, dtype=np.uint8)
    w, h = C.c_int(), C.c_int()
    L = load()
    st = L.hg_jpeg_decode(_ptr(buf), buf.size, None, 0, C.byref(w), C.byref(h))
    if st:
        raise HgError(st, "hg_jpeg_decode: not a JPEG this decoder supports")
    out = np.empty((h.value, w.value, 4), np.uint8)
    st = L.hg_jpeg_decode(_ptr(buf), buf.size, _ptr(out), out.size, C.byref(w), C.byref(h))
    if st:
        raise HgError(st, "hg_jpeg_decode: malformed or unsupported image data")
    return out


def png_encode(rgba) -> bytes:
    """hg_png_encode: (h, w, 4) uint8 RGBA -> PNG file bytes.  Host only."""
    a = np.ascontiguousarray(rgba, dtype=np.uint8)
    if a.ndim != 3 or a.shape[2] != 4:
        raise ValueError("png_encode expects an (h, w, 4) uint8 array")
    L = load()
    cap = L.hg_png_encode_bound(a.shape[1], a.shape[0])
    out = np.empty(cap, np.uint8)
    n = C.c_size_t()
    if L.hg_png_encode(_ptr(a), a.shape[1], a.shape[0], _ptr(out), cap, C.byref(n)):
        raise HgError(1, "hg_png_encode failed")
    return out[:n.value].tobytes()


def device_count() -> int:
    n = C.c_int()
    load().hg_device_count(C.byref(n))
    return n.value
