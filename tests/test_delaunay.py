"""hg_delaunay (host-side restatement of delaunator 5.0.0, the third-party package behind H.js:1216) against the
oracle's independent pure-Python restatement — bit for bit, order included — and against scipy's Delaunay as a SET of
triangles.  Host only: runs without a GPU (the library just has to load)."""
import numpy as np
import pytest

import homography_js_b200 as hg
from oracle import delaunator_ref


def _cxx(points):
    return hg._abi.delaunay(np.asarray(points, dtype=np.float64))


def _py(points):
    return np.asarray(delaunator_ref.triangles(np.asarray(points, dtype=np.float64).reshape(-1)), dtype=np.uint32)


def _tri_set(t):
    return {tuple(sorted(map(int, t[i:i + 3]))) for i in range(0, len(t), 3)}


@pytest.mark.parametrize("seed", range(20))
def test_random_points_match_python_restatement_and_scipy(seed):
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    n = int(rng.integers(3, 400))
    pts = rng.uniform(0, 1000, (n, 2))
    if seed % 2:
        pts = pts.astype(np.float32).astype(np.float64)  # Float32Array source points, as the reference stores them
    a, b = _cxx(pts), _py(pts)
    assert np.array_equal(a, b)
    assert len(a) % 3 == 0 and len(a) > 0
    assert _tri_set(a) == _tri_set(Delaunay(pts).simplices.reshape(-1))
    # every triangle keeps the seed's winding (the seed is swapped until orient2d >= 0)
    for i in range(0, len(a), 3):
        p, q, r = pts[a[i]], pts[a[i + 1]], pts[a[i + 2]]
        assert delaunator_ref.orient2d(p[0], p[1], q[0], q[1], r[0], r[1]) > 0


def test_degenerate_inputs_agree():
    """Regular grids (co-circular quadruples, ties in the distance sort), duplicates, collinear points, tiny sets:
    whatever the algorithm does there, both restatements do the same, element for element."""
    gx, gy = np.meshgrid(np.arange(10) * 426.0, np.arange(10) * 240.0)
    cases = [
        np.stack([gx.ravel(), gy.ravel()], 1),                                   # the 10 x 10 grid of config 3
        np.stack([np.arange(21).repeat(11) * 20.0, np.tile(np.arange(11), 21) * 20.0], 1),  # test.js test5 shape
        np.array([[0, 0], [0, 400], [400, 0], [400, 400]], float),              # a square: one co-circular flip decision
        np.array([[0, 0], [0, 200], [200, 0], [200, 200], [100, 100], [100, 100], [0, 0]], float),  # duplicates
        np.array([[0, 0], [1, 1], [2, 2], [3, 3]], float),                       # collinear -> no triangles
        np.array([[5, 5], [5, 5], [5, 5]], float),                               # coincident -> no triangles
        np.array([[0, 0], [1, 0]], float),                                       # fewer than three points
        np.array([[0, 0], [1e-7, 1e-7], [4, 0], [0, 4], [4, 4.000000001]], float),
    ]
    for pts in cases:
        a, b = _cxx(pts), _py(pts)
        assert np.array_equal(a, b), pts.tolist()
    assert len(_cxx(cases[4])) == 0 and len(_cxx(cases[5])) == 0 and len(_cxx(cases[6])) == 0
    assert len(_cxx(cases[0])) == 3 * 162  # a 10 x 10 grid triangulates into 2 * 9 * 9 triangles
    assert len(_cxx(cases[2])) == 6


def test_exact_orientation_on_nearly_collinear_points():
    """The orientation predicate must be sign-exact where the floating-point determinant is not: points on a line with
    one ulp of offset, at large coordinates."""
    rng = np.random.default_rng(3)
    base = rng.uniform(1e5, 2e5, (40, 1)) * np.array([[1.0, 0.5]])
    base[:, 1] = base[:, 0] * 0.5
    pts = base.copy()
    pts[::3, 1] = np.nextafter(pts[::3, 1], np.inf)
    pts[1::3, 1] = np.nextafter(pts[1::3, 1], -np.inf)
    pts = np.concatenate([pts, [[0.0, 1e5], [3e5, -1e4]]])
    assert np.array_equal(_cxx(pts), _py(pts))


def test_homography_default_triangulation_uses_the_restatement():
    pts = np.array([[0, 0], [0, 0.5], [0.5, 0], [0.5, 0.5], [0.5, 1], [1, 0.5], [1, 1], [0, 1], [1, 0]], np.float32)
    from homography_js_b200.homography import default_triangulation
    t = default_triangulation(pts)
    assert np.array_equal(t, _py(pts)) and len(t) == 3 * 8


def test_capacity_and_argument_errors():
    import ctypes as C
    L = hg._abi.load()
    pts = np.array([[0, 0], [1, 0], [0, 1], [1, 1.5]], np.float64)
    out = np.zeros(3, np.uint32)
    cnt = C.c_int(-1)
    assert L.hg_delaunay(pts.ctypes.data, 4, out.ctypes.data, 1, C.byref(cnt)) == 1  # HG_ERR_INVALID: needs 2 triangles
    assert cnt.value == 0
    assert L.hg_delaunay(None, 4, out.ctypes.data, 1, C.byref(cnt)) == 1


def test_benchmark_sized_mesh_is_fast_and_complete():
    """The reference's largest benchmark mesh: 160 x 80 points -> 23,000+ triangles (test/benchmark.js:48, README.md:317).
    A jittered grid of that size triangulates in well under a second on the host and is a complete triangulation of its
    convex hull (Euler: T = 2n - 2 - hull vertices)."""
    import time
    from scipy.spatial import ConvexHull
    rng = np.random.default_rng(12)
    gx, gy = np.meshgrid(np.arange(160) * 5.0, np.arange(80) * 5.0)
    pts = (np.stack([gx.ravel(), gy.ravel()], 1) + rng.uniform(-1.5, 1.5, (160 * 80, 2))).astype(np.float32).astype(np.float64)
    t0 = time.perf_counter()
    t = _cxx(pts)
    dt = time.perf_counter() - t0
    n, hull = len(pts), len(ConvexHull(pts).vertices)
    assert len(t) // 3 == 2 * n - 2 - hull
    assert dt < 2.0, dt
    # every point is used
    assert len(np.unique(t)) == n


def test_medium_mesh_matches_python_restatement():
    rng = np.random.default_rng(13)
    pts = rng.uniform(0, 4000, (2500, 2)).astype(np.float32).astype(np.float64)
    assert np.array_equal(_cxx(pts), _py(pts))


def test_hostile_point_sets_under_sanitizers(tmp_path):
    """hg_delaunay takes caller data as is: duplicates, collinear sets, 1e300, denormals, NaN and Inf run through the C++
    under AddressSanitizer + UBSan with a time limit (no out-of-bounds access, no endless hull walk)."""
    import os
    import shutil
    import subprocess
    from conftest import ROOT
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if not gxx:
        pytest.skip("no g++")
    exe = tmp_path / "delaunay_fuzz"
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    build = subprocess.run([gxx, "-std=c++17", "-O1", "-g", "-ffp-contract=off", "-fsanitize=address,undefined",
                            "-fno-sanitize-recover=all", os.path.join(ROOT, "tests", "delaunay_fuzz_harness.cpp"), "-o", str(exe)],
                           capture_output=True, text=True, env=env)
    if build.returncode != 0 and "sanitize" in build.stderr:
        pytest.skip("sanitizer runtime not available")
    assert build.returncode == 0, build.stderr[-3000:]
    run = subprocess.run([str(exe), "1500"], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, (run.stdout[-500:], run.stderr[-3000:])
    assert run.stdout.startswith("triangles ")
