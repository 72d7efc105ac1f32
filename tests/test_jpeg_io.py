"""hg_jpeg_decode (host-side JPEG ingest, baseline and progressive) against libjpeg-turbo through Pillow —
byte for byte: the
bytes a browser's getImageData returns for the file.  Host only: runs without a GPU."""
import io
import os
import shutil
import subprocess

import numpy as np
import pytest

import homography_js_b200 as hg
from conftest import ROOT

PIL = pytest.importorskip("PIL.Image")


def _picture(h, w, seed=0):
    """Smooth gradients with a block of noise: every frequency, saturated chroma at the edges."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    a = np.stack([(np.sin(x / 7.0) + 1) * 127, (np.cos(y / 5.0) + 1) * 127, (x + y) % 256], -1).astype(np.uint8)
    if h > 6 and w > 8:
        a[h // 3:h // 2, w // 4:w // 2] = rng.integers(0, 256, (h // 2 - h // 3, w // 2 - w // 4, 3), dtype=np.uint8)
    return a


def _jpeg(img, **kw):
    b = io.BytesIO()
    img.save(b, format="JPEG", **kw)
    return b.getvalue()


def _same_as_pillow(data):
    want = np.asarray(PIL.open(io.BytesIO(data)).convert("RGB"))
    got = hg._abi.jpeg_decode(data)
    assert got.shape == want.shape[:2] + (4,)
    assert (got[..., 3] == 255).all()
    assert np.array_equal(got[..., :3], want)


@pytest.mark.parametrize("progressive", [False, True], ids=["baseline", "progressive"])
@pytest.mark.parametrize("subsampling", [0, 1, 2], ids=["444", "422", "420"])
def test_decode_matches_libjpeg_turbo_for_every_size_and_quality(subsampling, progressive):
    """Odd sizes (partial MCUs), components one or two samples wide (no triangle filter there), whole-MCU sizes; progressive
    files exercise all four scan kinds (DC / AC, first / refinement) and end-of-band runs."""
    for h, w in [(64, 64), (37, 53), (1, 1), (8, 9), (17, 3), (2, 2), (100, 255), (131, 97), (5, 4), (16, 5), (3, 6)]:
        for q in (1, 30, 75, 95, 100):
            data = _jpeg(PIL.fromarray(_picture(h, w), "RGB"), quality=q, subsampling=subsampling, progressive=progressive,
                         optimize=progressive and q == 75)
            assert (b"\xff\xc2" in data) == progressive
            _same_as_pillow(data)


def test_grey_custom_huffman_restart_intervals_rgb_files_and_noise():
    img = PIL.fromarray(_picture(97, 133), "RGB")
    _same_as_pillow(_jpeg(img.convert("L"), quality=80))
    _same_as_pillow(_jpeg(img.convert("L"), quality=80, progressive=True))
    _same_as_pillow(_jpeg(img, quality=80, optimize=True, subsampling=2))            # per-image Huffman tables
    for kw in ({"restart_marker_blocks": 3}, {"restart_marker_rows": 1}, {"restart_marker_blocks": 1, "subsampling": 2},
               {"restart_marker_blocks": 2, "subsampling": 2, "progressive": True}):
        try:
            data = _jpeg(img, quality=70, **kw)
        except TypeError:
            continue   # an older Pillow without the restart options
        assert b"\xff\xdd" in data
        _same_as_pillow(data)
    try:
        _same_as_pillow(_jpeg(img, quality=90, keep_rgb=True, subsampling=0))        # RGB components, Adobe transform 0
    except TypeError:
        pass
    rng = np.random.default_rng(4)
    noise = PIL.fromarray(rng.integers(0, 256, (63, 65, 3), dtype=np.uint8), "RGB")
    _same_as_pillow(_jpeg(noise, quality=100, subsampling=2))
    _same_as_pillow(_jpeg(noise, quality=10, subsampling=1))


def test_full_hd_frame():
    _same_as_pillow(_jpeg(PIL.fromarray(_picture(1080, 1920), "RGB"), quality=85))
    _same_as_pillow(_jpeg(PIL.fromarray(_picture(1080, 1920), "RGB"), quality=85, progressive=True))


def test_unsupported_and_malformed_files_are_refused_not_guessed():
    img = PIL.fromarray(_picture(40, 40), "RGB")
    base = _jpeg(img, quality=80)
    i = base.index(b"\xff\xc0")
    for data in (_jpeg(img.convert("CMYK"), quality=80),          # four components
                 base[:i + 1] + b"\xc9" + base[i + 2:],            # the same frame declared arithmetic-coded (SOF9)
                 base[:i + 1] + b"\xc3" + base[i + 2:],            # lossless (SOF3)
                 base[:i + 4] + b"\x0c" + base[i + 5:]):           # 12-bit samples
        with pytest.raises(hg.HgError) as e:
            hg._abi.jpeg_decode(data)
        assert e.value.status == hg._abi.HG_ERR_UNSUPPORTED
    good = _jpeg(img, quality=80)
    for data in (b"", b"\xff\xd8", good[:40], b"\x89PNG\r\n\x1a\n" + good, good[:200]):
        with pytest.raises(hg.HgError):
            hg._abi.jpeg_decode(data)
    # without its JFIF segment the file is still a JPEG (component ids 1, 2, 3 -> YCbCr): same pixels
    assert good[2:4] == b"\xff\xe0"
    assert np.array_equal(hg._abi.jpeg_decode(good[:2] + good[20:]), hg._abi.jpeg_decode(good))
    # a frame header that promises 65535 x 65535 pixels in a 700-byte file: refused before anything is allocated
    i = good.index(b"\xff\xc0")
    bomb = good[:i + 5] + b"\xff\xff\xff\xff" + good[i + 9:]
    import ctypes as C
    w, h = C.c_int(), C.c_int()
    assert hg._abi.load().hg_jpeg_decode(bomb, len(bomb), None, 0, C.byref(w), C.byref(h)) == hg._abi.HG_ERR_UNSUPPORTED  # > 2^28 pixels
    big = np.empty(1, np.uint8)
    assert hg._abi.load().hg_jpeg_decode(bomb, len(bomb), big.ctypes.data, 65535 * 65535 * 4, C.byref(w), C.byref(h)) != 0
    # within the pixel ceiling the size query answers, and the decode refuses a frame its bytes cannot pay for
    bomb = good[:i + 5] + b"\x3f\xff\x3f\xff" + good[i + 9:]
    assert hg._abi.load().hg_jpeg_decode(bomb, len(bomb), None, 0, C.byref(w), C.byref(h)) == 0 and w.value == 0x3fff
    assert hg._abi.load().hg_jpeg_decode(bomb, len(bomb), big.ctypes.data, 0x3fff * 0x3fff * 4, C.byref(w), C.byref(h)) != 0


def test_truncated_file_without_eoi_is_decoded_when_complete():
    good = _jpeg(PIL.fromarray(_picture(24, 24), "RGB"), quality=90)
    assert good.endswith(b"\xff\xd9")
    assert np.array_equal(hg._abi.jpeg_decode(good[:-2]), hg._abi.jpeg_decode(good))


def test_mutated_jpegs_under_sanitizers(tmp_path):
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if not gxx:
        pytest.skip("no g++")
    exe = tmp_path / "jpeg_fuzz"
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    build = subprocess.run([gxx, "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
                            os.path.join(ROOT, "tests", "jpeg_fuzz_harness.cpp"), "-o", str(exe)], capture_output=True, text=True, env=env)
    if build.returncode != 0 and "sanitize" in build.stderr:
        pytest.skip("sanitizer runtime not available")
    assert build.returncode == 0, build.stderr[-3000:]
    files = []
    img = PIL.fromarray(_picture(29, 43), "RGB")
    for name, kw in (("a", dict(quality=75, subsampling=2)), ("b", dict(quality=90, subsampling=1, optimize=True)),
                     ("c", dict(quality=60, subsampling=0)), ("d", dict(quality=80)),
                     ("e", dict(quality=75, subsampling=2, progressive=True)), ("f", dict(quality=50, subsampling=0, progressive=True))):
        p = tmp_path / (name + ".jpg")
        p.write_bytes(_jpeg(img.convert("L") if name == "d" else img, **kw))
        files.append(str(p))
    try:
        p = tmp_path / "r.jpg"
        p.write_bytes(_jpeg(img, quality=70, restart_marker_blocks=2, subsampling=2))
        files.append(str(p))
    except TypeError:
        pass
    p = tmp_path / "x.jpg"
    p.write_bytes(_with_exif_orientation(_jpeg(img, quality=70), 6, True))
    files.append(str(p))
    run = subprocess.run([str(exe), "1500"] + files, capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, (run.stdout[-500:], run.stderr[-3000:])
    decoded, rejected = [int(t.rstrip(",")) for t in run.stdout.split() if t.rstrip(",").isdigit()]
    assert decoded > 100 and rejected > 100


def _with_exif_orientation(jpeg: bytes, orientation: int, big_endian: bool) -> bytes:
    """Splice a minimal APP1 Exif segment (TIFF header + IFD0 with the Orientation entry) in after SOI."""
    import struct
    e = ">" if big_endian else "<"
    tiff = (b"MM" if big_endian else b"II") + struct.pack(e + "HI", 42, 8) + struct.pack(e + "H", 1) + \
        struct.pack(e + "HHIHH", 0x0112, 3, 1, orientation, 0) + struct.pack(e + "I", 0)
    body = b"Exif\0\0" + tiff
    return jpeg[:2] + b"\xff\xe1" + struct.pack(">H", len(body) + 2) + body + jpeg[2:]


@pytest.mark.parametrize("big_endian", [False, True], ids=["II", "MM"])
def test_exif_orientation_is_applied_like_a_browser_canvas(big_endian):
    """Browsers draw a JPEG the way its Exif Orientation says (image-orientation: from-image): getImageData returns the turned
    picture.  Checker: Pillow's exif_transpose on libjpeg-turbo's pixels."""
    ops = pytest.importorskip("PIL.ImageOps")
    base = _jpeg(PIL.fromarray(_picture(37, 53), "RGB"), quality=90, subsampling=2)
    upright = hg._abi.jpeg_decode(base)
    for o in range(1, 9):
        data = _with_exif_orientation(base, o, big_endian)
        im = PIL.open(io.BytesIO(data))
        assert im.getexif().get(0x0112) == o
        want = np.asarray(ops.exif_transpose(im).convert("RGB"))
        got = hg._abi.jpeg_decode(data)
        assert got.shape[:2] == ((53, 37) if o >= 5 else (37, 53))
        assert np.array_equal(got[..., :3], want), o
    # out-of-range values and damaged segments leave the picture as stored
    for data in (_with_exif_orientation(base, 9, big_endian), _with_exif_orientation(base, 0, big_endian),
                 _with_exif_orientation(base, 6, big_endian)[:2] + b"\xff\xe1\x00\x10Exif\0\0II*\0\xff\xff\xff\x7f" + base[2:]):
        assert np.array_equal(hg._abi.jpeg_decode(data), upright)


# ------------------------------------------------------------------ one scan per component (non-interleaved sequential)
def _reencode_non_interleaved(jpeg: bytes) -> bytes:
    """Pillow only writes interleaved baseline scans.  This transcodes one losslessly into the other legal sequential layout —
    one scan per component over the component's own block raster — by Huffman-decoding the coefficients and re-encoding them
    with the file's own tables (T.81 F.1.2 / F.2.2)."""
    import struct
    segs, pos = [], 2
    while True:
        assert jpeg[pos] == 0xFF
        m = jpeg[pos + 1]
        ln = struct.unpack(">H", jpeg[pos + 2:pos + 4])[0]
        segs.append((m, jpeg[pos + 4:pos + 2 + ln]))
        pos += 2 + ln
        if m == 0xDA:
            break
    end = jpeg.rindex(b"\xff\xd9")
    ecs = jpeg[pos:end].replace(b"\xff\x00", b"\xff")
    sof = next(b for m, b in segs if m == 0xC0)
    H, W, nc = struct.unpack(">HHB", sof[1:6])
    comps = [(sof[6 + 3 * i], sof[7 + 3 * i] >> 4, sof[7 + 3 * i] & 15) for i in range(nc)]
    max_h, max_v = max(c[1] for c in comps), max(c[2] for c in comps)
    dec, enc = {}, {}
    for m, b in segs:
        if m != 0xC4:
            continue
        o = 0
        while o < len(b):
            tc_th, counts = b[o], b[o + 1:o + 17]
            vals = b[o + 17:o + 17 + sum(counts)]
            o += 17 + sum(counts)
            code, k, d, e = 0, 0, {}, {}
            for ln in range(1, 17):
                for _ in range(counts[ln - 1]):
                    d[(ln, code)] = vals[k]
                    e[vals[k]] = (code, ln)
                    code += 1
                    k += 1
                code <<= 1
            dec[tc_th], enc[tc_th] = d, e
    sos = segs[-1][1]
    sel = {sos[1 + 2 * i]: (sos[2 + 2 * i] >> 4, sos[2 + 2 * i] & 15) for i in range(sos[0])}
    bits = "".join(f"{byte:08b}" for byte in ecs)
    cur = 0

    def sym(table):
        nonlocal cur
        code = ln = 0
        while True:
            code = (code << 1) | int(bits[cur])
            cur += 1
            ln += 1
            if (ln, code) in table:
                return table[(ln, code)]

    def receive(s):
        nonlocal cur
        if s == 0:
            return 0
        v = int(bits[cur:cur + s], 2)
        cur += s
        return v if v >= (1 << (s - 1)) else v - (1 << s) + 1

    mcus_x, mcus_y = -(-W // (8 * max_h)), -(-H // (8 * max_v))
    blocks = {cid: {} for cid, _, _ in comps}
    pred = {cid: 0 for cid, _, _ in comps}
    for my in range(mcus_y):
        for mx in range(mcus_x):
            for cid, hs, vs in comps:
                td, ta = sel[cid]
                for by in range(vs):
                    for bx in range(hs):
                        zz = [0] * 64
                        pred[cid] += receive(sym(dec[td]))
                        zz[0] = pred[cid]
                        k = 1
                        while k < 64:
                            rs = sym(dec[0x10 | ta])
                            r, s = rs >> 4, rs & 15
                            if s == 0:
                                if r != 15:
                                    break
                                k += 16
                                continue
                            k += r
                            zz[k] = receive(s)
                            k += 1
                        blocks[cid][(mx * hs + bx, my * vs + by)] = zz
    out = bytearray(b"\xff\xd8")
    for m, b in segs[:-1]:
        out += bytes([0xFF, m]) + struct.pack(">H", len(b) + 2) + b

    def put_bits(acc, value, n):
        acc.append(f"{value:0{n}b}" if n else "")

    def category(v):
        return 0 if v == 0 else abs(v).bit_length()

    for cid, hs, vs in comps:
        td, ta = sel[cid]
        acc, last = [], 0
        bw, bh = -(-(-(-W * hs // max_h)) // 8), -(-(-(-H * vs // max_v)) // 8)   # ceil(ceil(W hs / max_h) / 8)
        for Y in range(bh):
            for X in range(bw):
                zz = blocks[cid][(X, Y)]
                diff, last = zz[0] - last, zz[0]
                s = category(diff)
                put_bits(acc, *enc[td][s])
                put_bits(acc, diff if diff >= 0 else diff + (1 << s) - 1, s)
                run = 0
                for k in range(1, 64):
                    if zz[k] == 0:
                        run += 1
                        continue
                    while run > 15:
                        put_bits(acc, *enc[0x10 | ta][0xF0])
                        run -= 16
                    s = category(zz[k])
                    put_bits(acc, *enc[0x10 | ta][(run << 4) | s])
                    put_bits(acc, zz[k] if zz[k] >= 0 else zz[k] + (1 << s) - 1, s)
                    run = 0
                if run:
                    put_bits(acc, *enc[0x10 | ta][0x00])
        stream = "".join(acc)
        stream += "1" * (-len(stream) % 8)
        data = bytes(int(stream[i:i + 8], 2) for i in range(0, len(stream), 8)).replace(b"\xff", b"\xff\x00")
        out += b"\xff\xda" + struct.pack(">HB", 8, 1) + bytes([cid, (td << 4) | ta, 0, 63, 0]) + data
    return bytes(out + b"\xff\xd9")


@pytest.mark.parametrize("subsampling", [0, 1, 2], ids=["444", "422", "420"])
def test_one_scan_per_component_layout(subsampling):
    src = _jpeg(PIL.fromarray(_picture(45, 61), "RGB"), quality=80, subsampling=subsampling)
    multi = _reencode_non_interleaved(src)
    assert multi.count(b"\xff\xda") == 3
    want = np.asarray(PIL.open(io.BytesIO(src)).convert("RGB"))
    assert np.array_equal(np.asarray(PIL.open(io.BytesIO(multi)).convert("RGB")), want)   # the transcoding is lossless
    assert np.array_equal(hg._abi.jpeg_decode(multi)[..., :3], want)


# ------------------------------------------------------------------ hostile files: work bounded by the file size
def _segments(jpeg: bytes):
    """(marker, start, end) of every marker segment up to and including the first SOS header."""
    out, pos = [], 2
    while pos + 4 <= len(jpeg):
        assert jpeg[pos] == 0xFF
        m = jpeg[pos + 1]
        ln = (jpeg[pos + 2] << 8) | jpeg[pos + 3]
        out.append((m, pos, pos + 2 + ln))
        pos += 2 + ln
        if m == 0xDA:
            break
    return out


def test_crafted_progressive_file_cannot_buy_minutes_of_decoding():
    """The advisor's denial-of-service file: a frame header claiming 4096 x 4096 followed by hundreds of AC scans with no
    entropy-coded data.  Every scan used to walk all 262,144 blocks (5 KB -> 4.7 s, 70 KB -> more than 10 minutes); now the
    scan count and the blocks walked are bounded by the file size and the file is refused at once."""
    import struct
    import time
    base = _jpeg(PIL.fromarray(_picture(64, 64)[..., 0], "L"), quality=80, progressive=True)
    segs = _segments(base)
    sof = next(s for s in segs if s[0] == 0xC2)
    sos = segs[-1]
    data = bytearray(base)
    data[sof[1] + 5: sof[1] + 9] = struct.pack(">HH", 4096, 4096)          # height, width
    # an AC-first scan header of the single component (Ss=1, Se=63), repeated with no data behind it
    empty_scan = b"\xff\xda" + struct.pack(">HBBBBBB", 8, 1, data[sos[1] + 5], 0x00, 1, 63, 0)
    pad = b"\xff\xfe" + struct.pack(">H", 4098) + bytes(4096)               # a comment, so the size guard of the header passes
    crafted = bytes(data[:sof[1]]) + pad + bytes(data[sof[1]:-2]) + empty_scan * 500 + b"\xff\xd9"
    t0 = time.perf_counter()
    with pytest.raises(hg.HgError):
        hg._abi.jpeg_decode(crafted)
    assert time.perf_counter() - t0 < 2.0
    # a header claiming more pixels than the decoder supports is refused before anything is allocated
    huge = bytearray(base)
    huge[sof[1] + 5: sof[1] + 9] = struct.pack(">HH", 65535, 65535)
    t0 = time.perf_counter()
    with pytest.raises(hg.HgError) as e:
        hg._abi.jpeg_decode(bytes(huge))
    assert e.value.status == hg._abi.HG_ERR_UNSUPPORTED and time.perf_counter() - t0 < 0.5
    # honest progressive files are untouched by the bounds
    _same_as_pillow(_jpeg(PIL.fromarray(_picture(300, 200), "RGB"), quality=85, progressive=True))


def test_exif_segment_after_the_frame_header_still_turns_the_picture():
    """An APP1 Exif segment between SOF and SOS: the size query must report the turned size (it used to stop at SOF and
    report 30 x 20 for pixels laid out 20 x 30)."""
    ops = pytest.importorskip("PIL.ImageOps")
    base = _jpeg(PIL.fromarray(_picture(20, 30), "RGB"), quality=90)
    exif = _with_exif_orientation(base, 6, False)
    app1_end = 4 + ((exif[4] << 8) | exif[5])
    app1 = exif[2:app1_end]
    segs = _segments(base)
    sos = segs[-1]
    moved = base[:sos[1]] + app1 + base[sos[1]:]          # Exif right in front of SOS, after SOF / DHT
    got = hg._abi.jpeg_decode(moved)
    assert got.shape[:2] == (30, 20)
    assert np.array_equal(got[..., :3], np.asarray(ops.exif_transpose(PIL.open(io.BytesIO(exif))).convert("RGB")))
