"""The PRODUCT's host-side state machine (homography.js_b200/homography.py) against the literal
restatement of the reference class, with the engine calls answered by the CPU oracle (tests/oracle_context.py).
Covers the setter order / normalisation / dispatch logic without a GPU."""
import numpy as np
import pytest

import flows
import homography_js_b200 as hg
from oracle.homography_ref import RefHomography, RefImageData
from oracle_context import OracleContext


def _run(flow, golden):
    ref_res, ref = flow(lambda *a: RefHomography(*a), RefImageData(golden["src"].reshape(-1).copy(), 400, 400))
    got_res, got = flow(lambda *a: hg.Homography(*a, context=OracleContext()),
                        hg.ImageData(golden["src"].reshape(-1).copy(), 400, 400))
    return ref_res, ref, got_res, got


@pytest.mark.parametrize("flow", flows.ALL, ids=lambda f: f.__name__)
def test_flow_matches_reference_restatement(flow, golden):
    ref_res, ref, got_res, got = _run(flow, golden)
    assert got.transform == ref.transform
    assert got.last_path == ref.last_path
    assert len(ref_res) == len(got_res)
    for r, g in zip(ref_res, got_res):
        assert (g.width, g.height) == (r.width, r.height)
        assert np.array_equal(g.data, r.data)
    # point arrays are mutated in place exactly like the reference's typed arrays
    assert np.array_equal(got._srcPoints, ref._srcPoints)
    assert np.array_equal(got._dstPoints, ref._dstPoints)
    assert got._srcPointsAreNormalized == ref._srcPointsAreNormalized
    assert got._dstPointsAreNormalized == ref._dstPointsAreNormalized
    for name in ("_xOutputOffset", "_yOutputOffset", "_objectiveWidth", "_objectiveHeight", "_width", "_height"):
        assert getattr(got, name) == getattr(ref, name), name


def test_node_flow_reproduces_reference_golden(golden):
    res, _ = flows.node_test(lambda *a: hg.Homography(*a, context=OracleContext()),
                             hg.ImageData(golden["src"].reshape(-1).copy(), 400, 400))
    assert np.array_equal(res[0].as_array(), golden["out"])


def test_caller_owned_typed_array_is_mutated_in_place(golden):
    """T1: a Float32Array handed to a setter is denormalised in place (H.js:244)."""
    src = np.array([0, 0, 0, 1, 1, 0, 1, 1], np.float32)
    hm = hg.Homography("projective", context=OracleContext())
    hm.setSourcePoints(src, None, 400, 400)
    assert list(src) == [0, 0, 0, 400, 400, 0, 400, 400]


def test_error_texts_match_reference():
    hm = hg.Homography("affine", context=OracleContext())
    with pytest.raises(hg.HomographyError, match="exactly three reference points but 4 were given"):
        hm.setSourcePoints([[0, 0], [0, 1], [1, 0], [1, 1]])
    hm = hg.Homography(context=OracleContext())
    with pytest.raises(hg.HomographyError, match="at least 3 points but only 2 were given"):
        hm.setSourcePoints([[0, 0], [0, 1]])
    hm = hg.Homography("projective", context=OracleContext())
    hm.setSourcePoints([[0, 0], [0, 1], [1, 0], [1, 1]])
    with pytest.raises(hg.HomographyError, match=r"same amount of destiny points \(3\) than source points \(4\)"):
        hm.setDestinyPoints([[0, 0], [0, 1], [1, 0]])
    with pytest.raises(hg.HomographyError, match="warp\\(\\) must receive an image"):
        hm.warp()
    with pytest.raises(hg.HomographyError, match='Transform "bogus" is unknown'):
        hg.Homography("bogus", context=OracleContext()).setSourcePoints([[0, 0], [0, 1], [1, 0]])


def test_mixed_ranges_without_size_raises_like_reference():
    hm = hg.Homography("affine", context=OracleContext())
    hm.setSourcePoints([[0, 0], [0, 1], [1, 0]])
    with pytest.raises(hg.HomographyError, match="Impossible to put source and destiny points in the same range"):
        hm.setDestinyPoints([[0, 0], [400, 400], [400, 80]])


def test_dispatch_thresholds(golden):
    """H.js:421/426: forward loop iff the output size equals (affine) / is within [1/1.2, 1] of (piecewise) the input."""
    img = hg.ImageData(golden["src"].reshape(-1).copy(), 400, 400)
    hm = hg.Homography("affine", context=OracleContext())
    hm.setSourcePoints([[0, 0], [0, 400], [400, 0]])
    hm.setDestinyPoints([[7, 9], [7, 409], [407, 9]])
    hm.warp(img)
    assert hm.last_path == "forward_geometric"
    hm.warp(None, False, True)
    assert hm.last_path == "inverse_geometric"
    hm.setDestinyPoints([[0, 0], [0, 401], [400, 0]])
    hm.warp()
    assert hm.last_path == "inverse_geometric"


# ------------------------------------------------------------------ getTransformationMatrixAsCSS / transformHTMLElement
@pytest.mark.parametrize("flow", flows.CSS, ids=lambda f: f.__name__)
def test_css_matrix_matches_reference_restatement(flow):
    want = flow(lambda *a: RefHomography(*a))
    got = flow(lambda *a: hg.Homography(*a, context=OracleContext()))
    assert got == want
    assert got.startswith("matrix")


def test_css_errors_and_element_form():
    with pytest.raises(hg.HomographyError, match="srcPoints are not set"):
        hg.Homography("affine", context=OracleContext()).getTransformationMatrixAsCSS()
    hm = hg.Homography("affine", context=OracleContext())
    hm.setSourcePoints([[0, 0], [0, 1], [1, 0]])
    with pytest.raises(hg.HomographyError, match="dstPoints are not set"):
        hm.getTransformationMatrixAsCSS()
    pw = hg.Homography("piecewiseaffine", 10, 10, context=OracleContext())
    pw.setReferencePoints([[0, 0], [0, 1], [1, 0], [1, 1], [0.5, 0.5]], [[0, 0], [0, 1], [1, 0], [1, 1], [0.4, 0.6]])
    pw._transformMatrix = np.zeros(6, np.float32)   # the reference checks the matrix field before the transform kind
    with pytest.raises(hg.HomographyError, match='Only "affine" or "projective" transforms'):
        pw.getTransformationMatrixAsCSS()

    class Style:
        transform = None

    class Element:  # what transformHTMLElement touches of a DOM element (H.js:611-614)
        style = Style()

        def getBoundingClientRect(self):
            class R:
                width, height = 320.0, 200.0
            return R()

    el = Element()
    hg.Homography("auto", context=OracleContext()).transformHTMLElement(el, [[0, 0], [0, 1], [1, 0]], [[0, 0], [1 / 2, 1], [1, 1 / 8]])
    assert el.style.transform == "matrix(1.00000, 0.12500, 0.50000, 1.00000, 0.00000, 0.00000)"


def test_config1_affine_256_plumbing():
    """BASELINE.json configs[0]: the class surface end to end on the CPU (engine calls answered by the oracle)."""
    ref_res, ref = flows.config1(lambda *a: RefHomography(*a), RefImageData)
    got_res, got = flows.config1(lambda *a: hg.Homography(*a, context=OracleContext()), hg.ImageData)
    assert got.last_path == ref.last_path == "inverse_geometric"
    assert (got_res[0].width, got_res[0].height) == (ref_res[0].width, ref_res[0].height) == (256, 205)
    assert np.array_equal(got_res[0].data, ref_res[0].data)
    assert ref_res[0].data.reshape(-1, 4)[:, 3].any()   # not an empty image
