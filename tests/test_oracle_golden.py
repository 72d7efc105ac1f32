"""Pins the CPU oracle to the reference: its one pixel-exact golden (test/nodeTest.js ->
test/transformedImage.png) and known-answer cases that hold by construction for the inputs of
test/test.js.  CPU only."""
import numpy as np

import flows
from oracle import oracle as O
from oracle.homography_ref import RefHomography, RefImageData


def _img(golden):
    return RefImageData(golden["src"].reshape(-1).copy(), 400, 400)


def test_node_golden_via_class(golden):
    res, hm = flows.node_test(lambda *a: RefHomography(*a), _img(golden))
    r = res[0]
    assert (r.width, r.height) == (400, 200)
    assert hm.transform == "projective" and hm.last_path == "inverse_geometric"
    assert (hm._xOutputOffset, hm._yOutputOffset) == (0.0, 200.0)
    got = r.data.reshape(200, 400, 4)
    assert np.array_equal(got, golden["out"]), f"{(got != golden['out']).any(axis=2).sum()} pixels differ"


def test_node_golden_via_functions(golden):
    """Same golden through the bare functions: DLT solve (both directions), limits, inverse loop."""
    src = golden["src_points"].reshape(-1) * 400.0
    dst = golden["dst_points"].reshape(-1) * 400.0
    src32, dst32 = src.astype(np.float32), dst.astype(np.float32)  # Float32Array storage (H.js:220)
    fwd = O.projective_from_squares(src32, dst32)
    lim = O.transform_limits(fwd, 400, 400)
    assert list(lim) == [0.0, 200.0, 400.0, 200.0]
    inv = O.projective_from_squares(dst32, src32)
    out = O.warp_inverse_geometric(golden["src"], 400, 400, inv, 0, 200, 400, 200)
    assert np.array_equal(out.reshape(200, 400, 4), golden["out"])
    # multi-threaded variant of the loop (used as the CPU baseline) is bit-identical
    out_mt = O.warp_inverse_geometric(golden["src"], 400, 400, inv, 0, 200, 400, 200, threads=4)
    assert np.array_equal(out, out_mt)


def test_projective_identity_is_identity(golden):  # test.js:282 (test10)
    res, hm = flows.test10(lambda *a: RefHomography(*a), _img(golden))
    assert hm.last_path == "inverse_geometric"
    assert np.array_equal(res[0].data.reshape(400, 400, 4), golden["src"])


def test_affine_translation_forward_is_identity(golden):  # test.js:167 (test6): forward scatter path
    res, hm = flows.test6(lambda *a: RefHomography(*a), _img(golden))
    assert hm.last_path == "forward_geometric"
    assert list(hm._transformMatrix) == [1, 0, 0, 1, 100, 50]
    assert (hm._xOutputOffset, hm._yOutputOffset, hm._objectiveWidth, hm._objectiveHeight) == (100, 50, 400, 400)
    assert np.array_equal(res[0].data.reshape(400, 400, 4), golden["src"])


def test_affine_translation_inverse_is_identity(golden):
    res, hm = flows.test6_inverse(lambda *a: RefHomography(*a), _img(golden))
    assert hm.last_path == "inverse_geometric"
    assert np.array_equal(res[0].data.reshape(400, 400, 4), golden["src"])


def test_projective_mirror(golden):
    """x' = W - x: column x of the output is source column W - x (x >= 1); Q2: column 0 maps to sx = W."""
    hm = RefHomography("projective")
    hm.setSourcePoints([[0, 0], [0, 1], [1, 0], [1, 1]], None, 400, 400)
    hm.setDestinyPoints([[1, 0], [1, 1], [0, 0], [0, 1]])
    r = hm.warp(_img(golden))
    assert (r.width, r.height) == (400, 400)
    got = r.data.reshape(400, 400, 4)
    src = golden["src"]
    assert np.array_equal(got[:, 1:], src[:, ::-1][:, :-1])


def test_piecewise_2x_upsample(golden):  # test.js:34 (test1)
    res, hm = flows.test1(lambda *a: RefHomography(*a), _img(golden))
    r = res[0]
    assert (r.width, r.height) == (800, 800) and hm.last_path == "inverse_piecewise"
    got = r.data.reshape(800, 800, 4)
    src = golden["src"]
    # every inverse matrix is exactly [0.5,0,0,0.5,0,0]; Math.round(0.5*x) = (x+1)//2
    iy = (np.arange(799) + 1) // 2
    ix = (np.arange(799) + 1) // 2
    assert np.array_equal(got[:799, :799], src[iy][:, ix])
    # Q2: x = 799 -> round(399.5) = 400 -> flat index runs into the next row's first pixel
    assert np.array_equal(got[:797, 799], src[(np.arange(797) + 1) // 2 + 1, 0])
    # last output row -> source row 400 -> past the end of the image -> transparent
    assert not got[799].any()


def test_state_consistency_loop(golden):  # test.js:220 (test8)
    res, _ = flows.test8(lambda *a: RefHomography(*a), _img(golden))
    for k in (2, 4):
        assert np.array_equal(res[k].data, res[0].data)
        assert np.array_equal(res[k + 1].data, res[1].data)


def test_zero_area_output_gives_1x1(golden):  # H.js:436-441
    hm = RefHomography("affine")
    hm.setSourcePoints([[0, 0], [0, 400], [400, 0]])
    hm.setDestinyPoints([[10, 10], [10, 10], [10, 10]])  # collapses everything to one point
    r = hm.warp(_img(golden))
    assert (r.width, r.height) == (1, 1) and not r.data.any()
