"""The Node-API addon (js/hgwarp_napi.c) EXECUTED without Node.js, under the miniature Node-API runtime of
tests/napi_mock/ (see tests/napi_mock.py):

  CPU  host-only exports (delaunay, pngDecode, pngEncode) against the C ABI called directly; argument checking and the
       thrown JS errors; and the whole class surface driven through native.* with the GPU entry points answered by a CPU
       test double (tests/napi_mock/hgwarp_cpu_double.c -> oracle), which pins every argument the addon marshals;
  GPU  the same flows with the addon bound to the real libhgwarp.so (class surface -> native.* -> N-API -> C ABI -> CUDA)."""
import numpy as np
import pytest

import flows
import homography_js_b200 as hg
import napi_mock as M
from oracle.homography_ref import RefHomography, RefImageData


@pytest.fixture(scope="module")
def native_cpu():
    n = M.Native(cpu_double=True)
    yield n
    n.close()


def test_addon_registers_the_exports_the_shim_calls(native_cpu):
    assert M.build(True).mock_module_name() == b"hgwarp"
    assert set(native_cpu.exports) == {"createContext", "setImage", "solveWithLimits", "warpInversePoints", "warpForwardMatrix",
                                       "setMesh", "delaunay", "pngDecode", "jpegDecode", "pngEncode", "piecewiseMatrices",
                                       "warpPiecewiseInverse", "warpPiecewiseForward"}


def test_host_only_exports_match_the_c_abi(native_cpu):
    rng = np.random.default_rng(0)
    for dt in (np.float64, np.float32):
        pts = rng.uniform(0, 100, (40, 2)).astype(dt)
        got = native_cpu.delaunay(pts.reshape(-1))
        assert got.dtype == np.uint32 and np.array_equal(got, hg._abi.delaunay(pts.astype(np.float64)))
    assert native_cpu.delaunay(np.zeros(2, np.float64)).size == 0          # one point: no triangles
    img = rng.integers(0, 256, (9, 13, 4), dtype=np.uint8)
    png = native_cpu.pngEncode(img.reshape(-1).view(M.Clamped), 13, 9)
    assert bytes(png) == hg._abi.png_encode(img)
    back = native_cpu.pngDecode(png)
    assert (back["width"], back["height"]) == (13, 9) and np.array_equal(back["data"].reshape(9, 13, 4), img)
    PIL = pytest.importorskip("PIL.Image")
    import io
    b = io.BytesIO()
    PIL.fromarray(img[..., :3].copy(), "RGB").save(b, format="JPEG", quality=85)
    jpg = np.frombuffer(b.getvalue(), np.uint8)
    back = native_cpu.jpegDecode(jpg)
    assert (back["width"], back["height"]) == (13, 9)
    assert np.array_equal(back["data"].reshape(9, 13, 4), hg._abi.jpeg_decode(b.getvalue()))
    with pytest.raises(M.JsError, match="not a JPEG"):
        native_cpu.jpegDecode(np.arange(64, dtype=np.uint8))


def test_argument_errors_become_js_exceptions(native_cpu):
    with pytest.raises(M.JsError, match="wrong number of arguments"):
        native_cpu.pngEncode(np.zeros(16, np.uint8), 2)
    with pytest.raises(M.JsError, match=r"delaunay\(Float32Array \| Float64Array\)"):
        native_cpu.delaunay(np.zeros(6, np.int32))
    with pytest.raises(M.JsError, match="not a PNG"):
        native_cpu.pngDecode(np.arange(64, dtype=np.uint8))
    with pytest.raises(M.JsError, match="pngEncode"):
        native_cpu.pngEncode(np.zeros(15, np.uint8).view(M.Clamped), 2, 2)   # shorter than w*h*4
    ctx = native_cpu.createContext(0)
    with pytest.raises(M.JsError, match="setImage"):
        native_cpu.setImage(ctx, np.zeros(10, np.uint8).view(M.Clamped), 4, 4)
    with pytest.raises(M.JsError, match="solveWithLimits"):
        native_cpu.solveWithLimits(ctx, 1, np.zeros(6), np.zeros(8), 10, 10)     # projective needs 8 values
    with pytest.raises(M.JsError, match="status 5: no image set"):
        native_cpu.warpInversePoints(ctx, 0, np.array([0, 0, 0, 1, 1, 0.]), np.array([0, 0, 0, 1, 1, 0.]), 0, 0, 4, 4)
    with pytest.raises(M.JsError, match="createContext|status"):
        native_cpu.createContext(7)                                            # the double only knows device 0


@pytest.mark.parametrize("flow", flows.ALL, ids=lambda f: f.__name__)
def test_class_surface_through_the_addon_cpu_double(native_cpu, flow, golden):
    ref_res, ref = flow(lambda *a: RefHomography(*a), RefImageData(golden["src"].reshape(-1).copy(), 400, 400))
    got_res, got = flow(lambda *a: hg.Homography(*a, context=M.NapiContext(native_cpu)),
                        hg.ImageData(golden["src"].reshape(-1).copy(), 400, 400))
    assert got.last_path == ref.last_path
    for r, g in zip(ref_res, got_res):
        assert (g.width, g.height) == (r.width, r.height)
        assert g.data.dtype == np.uint8 and np.array_equal(g.data, r.data)


def test_node_golden_through_the_addon_cpu_double(native_cpu, golden):
    res, _ = flows.node_test(lambda *a: hg.Homography(*a, context=M.NapiContext(native_cpu)),
                             hg.ImageData(golden["src"].reshape(-1).copy(), 400, 400))
    assert np.array_equal(res[0].as_array(), golden["out"])


# ------------------------------------------------------------------ the real library (B200 box)
@pytest.fixture(scope="module")
def native_gpu():
    n = M.Native(cpu_double=False)
    yield n
    n.close()


@pytest.mark.gpu
@pytest.mark.parametrize("flow", [flows.node_test, flows.test1, flows.test5, flows.test6, flows.test10, flows.test5_then_forward],
                         ids=lambda f: f.__name__)
def test_class_surface_through_the_addon_over_cuda(native_gpu, flow, golden):
    ref_res, ref = flow(lambda *a: RefHomography(*a), RefImageData(golden["src"].reshape(-1).copy(), 400, 400))
    got_res, got = flow(lambda *a: hg.Homography(*a, context=M.NapiContext(native_gpu)),
                        hg.ImageData(golden["src"].reshape(-1).copy(), 400, 400))
    assert got.last_path == ref.last_path
    for r, g in zip(ref_res, got_res):
        assert (g.width, g.height) == (r.width, r.height)
        assert np.array_equal(g.data, r.data)


# ------------------------------------------------------------------ the JS shim's own control flow (transliterated)
def _shim_factory(native):
    import js_shim_transliteration as T
    return lambda *a: T.Homography(native, *a)


@pytest.mark.parametrize("flow", flows.ALL, ids=lambda f: f.__name__)
def test_js_shim_control_flow_through_the_addon_cpu_double(native_cpu, flow, golden):
    """js/homography_b200.mjs rendered statement by statement into Python (tests/js_shim_transliteration.py), driving the real
    addon: the shim's call protocol (solveWithLimits for every solve, re-solve in _induceObjective, f64 / f32 conversions)
    gives the reference's bytes for every flow of the reference's test page."""
    ref_res, ref = flow(lambda *a: RefHomography(*a), RefImageData(golden["src"].reshape(-1).copy(), 400, 400))
    got_res, got = flow(_shim_factory(native_cpu), hg.ImageData(golden["src"].reshape(-1).copy(), 400, 400))
    assert got.transform == ref.transform and got.last_path == ref.last_path
    for r, g in zip(ref_res, got_res):
        assert (g.width, g.height) == (r.width, r.height)
        assert np.array_equal(g.data, r.data)
    assert np.array_equal(got._srcPoints, ref._srcPoints) and np.array_equal(got._dstPoints, ref._dstPoints)
    for name in ("_xOutputOffset", "_yOutputOffset", "_objectiveWidth", "_objectiveHeight", "_width", "_height"):
        assert getattr(got, name) == getattr(ref, name), name


@pytest.mark.parametrize("flow", flows.CSS, ids=lambda f: f.__name__)
def test_js_shim_css_strings_through_the_addon_cpu_double(native_cpu, flow):
    assert flow(_shim_factory(native_cpu)) == flow(lambda *a: RefHomography(*a))


def test_transliteration_covers_the_shim():
    """Every method of the .mjs class and every native.* call it makes exists in the transliteration, and vice versa."""
    import os
    import re
    from conftest import ROOT
    mjs = open(os.path.join(ROOT, "homography.js_b200", "js", "homography_b200.mjs")).read()
    py = open(os.path.join(ROOT, "tests", "js_shim_transliteration.py")).read()
    js_methods = set(re.findall(r"\n  (_?[A-Za-z]+)\(", mjs)) - {"constructor", "if", "for", "switch", "return"}
    py_methods = set(re.findall(r"\n    def (_?[A-Za-z]+)\(", py)) - {"__init__"}
    assert js_methods == py_methods, js_methods ^ py_methods
    js_calls = set(re.findall(r"\bnative\.([A-Za-z]+)\(", mjs)) - {"pngDecode", "jpegDecode", "pngEncode"}   # module-level helpers
    py_calls = set(re.findall(r"\bnative\.([A-Za-z]+)\(", py))
    assert js_calls == py_calls, js_calls ^ py_calls
    # the same reference-defined throw texts
    for text in re.findall(r"throw \('([^'$]{30,})'\)", mjs):
        assert text in py, text
