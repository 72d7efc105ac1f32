"""hg_png_decode / hg_png_encode (host-side image ingest / egress) against Pillow and against the
reference's own fixtures.  Host only: runs without a GPU."""
import io
import os

import numpy as np
import pytest

import homography_js_b200 as hg

PIL = pytest.importorskip("PIL.Image")
REF = "/root/reference/test"


def _png_bytes(img, **kw):
    b = io.BytesIO()
    img.save(b, format="PNG", **kw)
    return b.getvalue()


def _pillow_rgba(data):
    return np.asarray(PIL.open(io.BytesIO(data)).convert("RGBA"))


@pytest.mark.parametrize("mode", ["RGBA", "RGB", "L", "LA", "P", "1", "I;16"])
def test_decode_matches_pillow_for_every_colour_type(mode):
    rng = np.random.default_rng(11)
    w, h = 61, 37  # odd width: sub-byte rows end mid-byte
    if mode == "RGBA":
        img = PIL.fromarray(rng.integers(0, 256, (h, w, 4), dtype=np.uint8), "RGBA")
    elif mode == "RGB":
        img = PIL.fromarray(rng.integers(0, 256, (h, w, 3), dtype=np.uint8), "RGB")
    elif mode == "L":
        img = PIL.fromarray(rng.integers(0, 256, (h, w), dtype=np.uint8), "L")
    elif mode == "LA":
        img = PIL.fromarray(rng.integers(0, 256, (h, w, 2), dtype=np.uint8), "LA")
    elif mode == "P":
        img = PIL.fromarray(rng.integers(0, 256, (h, w, 3), dtype=np.uint8), "RGB").quantize(37)
    elif mode == "1":
        img = PIL.fromarray((rng.integers(0, 2, (h, w)) * 255).astype(np.uint8), "L").convert("1")
    else:
        img = PIL.fromarray(rng.integers(0, 65536, (h, w), dtype=np.uint16))
    data = _png_bytes(img)
    got = hg._abi.png_decode(data)
    if mode == "I;16":
        want = np.asarray(PIL.open(io.BytesIO(data))).astype(np.uint16)
        assert np.array_equal(got[..., 0], (want >> 8).astype(np.uint8)) and (got[..., 3] == 255).all()
        assert np.array_equal(got[..., 0], got[..., 1]) and np.array_equal(got[..., 0], got[..., 2])
    else:
        assert np.array_equal(got, _pillow_rgba(data))


def test_every_filter_type_and_transparency_chunks():
    """Pillow picks filters adaptively: a smooth gradient + noise image exercises Sub / Up / Average / Paeth; palette
    transparency (tRNS) and an RGB colour key are honoured."""
    y, x = np.mgrid[0:96, 0:128]
    rng = np.random.default_rng(5)
    a = np.stack([(x * 2) % 256, (y * 3) % 256, (x + y) % 256, 255 - (x % 256)], -1).astype(np.uint8)
    a[40:60] = rng.integers(0, 256, (20, 128, 4), dtype=np.uint8)
    for lvl in (1, 9):
        data = _png_bytes(PIL.fromarray(a, "RGBA"), compress_level=lvl)
        assert np.array_equal(hg._abi.png_decode(data), a)
    p = PIL.fromarray(a[..., :3], "RGB").quantize(16)
    data = _png_bytes(p, transparency=3)
    assert np.array_equal(hg._abi.png_decode(data), _pillow_rgba(data))
    rgb = a[..., :3].copy()
    rgb[10:20, 10:20] = (1, 2, 3)
    data = _png_bytes(PIL.fromarray(rgb, "RGB"), transparency=(1, 2, 3))
    got = hg._abi.png_decode(data)
    assert np.array_equal(got, _pillow_rgba(data)) and (got[10:20, 10:20, 3] == 0).all()


def test_encode_round_trips_and_pillow_reads_it():
    rng = np.random.default_rng(6)
    for w, h in ((1, 1), (3, 5), (400, 200), (257, 63)):
        a = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        if w > 100:
            a[:, : w // 2] = a[:1, : w // 2]  # vertically constant half: the Up filter wins there
        data = hg._abi.png_encode(a)
        assert np.array_equal(hg._abi.png_decode(data), a)
        assert np.array_equal(_pillow_rgba(data), a)


def test_malformed_files_are_rejected():
    a = np.zeros((4, 4, 4), np.uint8)
    good = hg._abi.png_encode(a)
    for bad in (b"", b"not a png", good[:20], good[:-1], good[:40] + bytes([good[40] ^ 1]) + good[41:]):
        with pytest.raises(hg.HgError):
            hg._abi.png_decode(bad)
    inter = _png_bytes(PIL.fromarray(a, "RGBA"))  # sanity: Pillow's own file decodes
    assert hg._abi.png_decode(inter).shape == (4, 4, 4)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
def test_reference_fixtures_decode_to_the_committed_golden_arrays(golden):
    """The reference's own PNGs (test/testImgLogoBlack.png -> test/transformedImage.png, test/nodeTest.js) decode to
    exactly the arrays the golden fixture holds (decoded independently when the fixture was made)."""
    src = hg._abi.png_decode(open(os.path.join(REF, "testImgLogoBlack.png"), "rb").read())
    out = hg._abi.png_decode(open(os.path.join(REF, "transformedImage.png"), "rb").read())
    assert src.shape == (400, 400, 4) and out.shape == (200, 400, 4)
    assert np.array_equal(src.reshape(-1), np.asarray(golden["src"]).reshape(-1))
    assert np.array_equal(out.reshape(-1), np.asarray(golden["out"]).reshape(-1))
    for name in ("testImg.png", "testImgLogoWhite.png"):
        data = open(os.path.join(REF, name), "rb").read()
        assert np.array_equal(hg._abi.png_decode(data), _pillow_rgba(data))


def _interlaced_png(a):
    """An Adam7-interlaced 8-bit RGBA PNG of `a`, built by hand (Pillow cannot write interlaced files)."""
    import struct
    import zlib
    h, w = a.shape[:2]
    raw = b""
    for x0, y0, dx, dy in ((0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)):
        sub = a[y0::dy, x0::dx]
        if sub.size == 0:
            continue
        prev = np.zeros(sub.shape[1] * 4, np.uint8)
        for y, row in enumerate(sub.reshape(sub.shape[0], -1)):
            ft = (0, 2, 1)[y % 3]  # None, Up, Sub in turn
            if ft == 0:
                f = row
            elif ft == 2:
                f = (row.astype(np.int16) - prev).astype(np.uint8)
            else:
                left = np.concatenate([np.zeros(4, np.uint8), row[:-4]])
                f = (row.astype(np.int16) - left).astype(np.uint8)
            raw += bytes([ft]) + f.tobytes()
            prev = row

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)

    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 1)) +
            chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b""))


def test_adam7_interlaced_files_decode():
    rng = np.random.default_rng(8)
    for w, h in ((1, 1), (3, 2), (8, 8), (37, 21), (64, 5)):  # small sizes leave some passes empty
        a = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        data = _interlaced_png(a)
        assert np.array_equal(_pillow_rgba(data), a)          # the hand-built file is a valid interlaced PNG
        assert np.array_equal(hg._abi.png_decode(data), a)
