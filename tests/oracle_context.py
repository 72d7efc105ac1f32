"""A stand-in for homography_js_b200._abi.Context that answers every engine call with the CPU oracle.

Used by the CPU-only tests to exercise the PRODUCT's host-side state machine (homography.py) without a
GPU, and by nothing else: the product never imports it.
"""
import numpy as np

from oracle import oracle as O


class OracleContext:
    def __init__(self):
        self.img = None
        self.W = self.H = 0
        self.src_pts = self.tris = None
        self.last_map = None
        self.calls = []

    def close(self):
        pass

    def image_set(self, rgba, w, h):
        self.calls.append("image_set")
        self.img, self.W, self.H = np.array(rgba, dtype=np.uint8).reshape(-1), w, h

    def solve_affine(self, src, dst):
        self.calls.append("solve_affine")
        return O.affine_from_triangles(src, dst)

    def solve_projective(self, src, dst):
        self.calls.append("solve_projective")
        return O.projective_from_squares(src, dst)

    def inverse_affine(self, m):
        return O.inverse_affine(m)

    def transform_limits(self, matrix, w, h):
        self.calls.append("transform_limits")
        return O.transform_limits(matrix, w, h)

    def solve_with_limits(self, kind, src, dst, w, h):
        self.calls.append("solve_with_limits")
        m = O.affine_from_triangles(src, dst) if kind == 0 else O.projective_from_squares(src, dst)
        return m, O.transform_limits(m, w, h)

    def warp_inverse_points(self, kind, dst_pts, src_pts, x_off, y_off, o_w, o_h, **kw):
        self.calls.append("warp_inverse_points")
        inv = O.affine_from_triangles(dst_pts, src_pts) if kind == 0 else O.projective_from_squares(dst_pts, src_pts)
        return O.warp_inverse_geometric(self.img, self.W, self.H, inv, x_off, y_off, o_w, o_h)

    def warp_inverse_matrix(self, inv, x_off, y_off, o_w, o_h, **kw):
        return O.warp_inverse_geometric(self.img, self.W, self.H, inv, x_off, y_off, o_w, o_h)

    def warp_forward_matrix(self, fwd, x_off, y_off, o_w, o_h, **kw):
        self.calls.append("warp_forward_matrix")
        return O.warp_forward_geometric(self.img, self.W, self.H, fwd, x_off, y_off, o_w, o_h)

    def piecewise_set_mesh(self, src_pts, tris):
        self.calls.append("piecewise_set_mesh")
        self.src_pts = np.array(src_pts, dtype=np.float32).reshape(-1)
        self.tris = np.array(tris, dtype=np.uint32).reshape(-1)

    def piecewise_matrices(self, dst_pts, want_inverse=False):
        fwd = O.piecewise_matrices(self.src_pts, dst_pts, self.tris)
        return (fwd, O.inverse_matrices(fwd)) if want_inverse else fwd

    def piecewise_extents(self, dst_pts):
        d = np.asarray(dst_pts, dtype=np.float32)
        out = np.empty((d.shape[0], 4), np.float64)
        for f in range(d.shape[0]):
            mm = O.minmax_xy(d[f])
            out[f] = [mm[0], mm[1], mm[2] - mm[0], mm[3] - mm[1]]
        return out

    def build_index_map(self, pts, map_width, y_offset, map_len):
        self.last_map = O.build_index_map(pts, self.tris, map_width, y_offset, map_len)
        return self.last_map

    def warp_piecewise_inverse(self, dst_pts, x_off, y_off, o_w, o_h, min_src_x, min_src_y, **kw):
        self.calls.append("warp_piecewise_inverse")
        fwd = O.piecewise_matrices(self.src_pts, dst_pts, self.tris)
        self.last_map = O.build_index_map(dst_pts, self.tris, o_w, y_off, o_w * o_h)
        return O.warp_inverse_piecewise(self.img, self.W, self.H, self.last_map, O.inverse_matrices(fwd), x_off, y_off,
                                        o_w, o_h, min_src_x, min_src_y)

    def warp_piecewise_forward(self, dst_pts, x_off, y_off, o_w, o_h, min_src_x, min_src_y, max_src_x, max_src_y,
                               use_inverse_map=False, **kw):
        self.calls.append("warp_piecewise_forward")
        fwd = O.piecewise_matrices(self.src_pts, dst_pts, self.tris)
        if use_inverse_map:
            m = self.last_map
        else:
            mw = max_src_x - min_src_x
            m = O.build_index_map(self.src_pts, self.tris, mw, min_src_y, mw * (max_src_y - min_src_y))
        return O.warp_forward_piecewise(self.img, self.W, self.H, m, fwd, x_off, y_off, o_w, o_h, min_src_x, min_src_y,
                                        max_src_x, max_src_y)
