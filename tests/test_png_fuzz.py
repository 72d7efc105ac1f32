"""The PNG decoder reads untrusted bytes: mutated files (chunk CRCs re-sealed so the damage gets past the first check) run
through it under AddressSanitizer + UBSan.  Host only."""
import io
import os
import shutil
import subprocess

import numpy as np
import pytest

import homography_js_b200 as hg
from conftest import ROOT

PIL = pytest.importorskip("PIL.Image")


def _bases(tmp_path):
    rng = np.random.default_rng(3)
    w, h = 23, 17
    rgba = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    files = []

    def save(name, img, **kw):
        b = io.BytesIO()
        img.save(b, format="PNG", **kw)
        p = tmp_path / name
        p.write_bytes(b.getvalue())
        files.append(str(p))

    save("rgba.png", PIL.fromarray(rgba, "RGBA"))
    save("rgb.png", PIL.fromarray(rgba[..., :3].copy(), "RGB"))
    save("pal.png", PIL.fromarray(rgba[..., :3].copy(), "RGB").quantize(13))
    save("bit.png", PIL.fromarray((rgba[..., 0] > 127).astype(np.uint8) * 255, "L").convert("1"))
    save("g16.png", PIL.fromarray(rng.integers(0, 65536, (h, w), dtype=np.uint16)))
    save("la.png", PIL.fromarray(rgba[..., :2].copy(), "LA"))
    own = tmp_path / "own.png"
    own.write_bytes(hg._abi.png_encode(rgba))     # the product's own encoder output
    files.append(str(own))
    return files


def test_mutated_pngs_under_sanitizers(tmp_path):
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if not gxx:
        pytest.skip("no g++")
    exe = tmp_path / "png_fuzz"
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    build = subprocess.run([gxx, "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
                            os.path.join(ROOT, "tests", "png_fuzz_harness.cpp"), "-o", str(exe), "-lz"],
                           capture_output=True, text=True, env=env)
    if build.returncode != 0 and "sanitize" in build.stderr:
        pytest.skip("sanitizer runtime not available")
    assert build.returncode == 0, build.stderr[-3000:]
    run = subprocess.run([str(exe), "1500"] + _bases(tmp_path), capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, (run.stdout[-500:], run.stderr[-3000:])
    decoded, rejected = [int(t.rstrip(",")) for t in run.stdout.split() if t.rstrip(",").isdigit()]
    assert decoded > 100 and rejected > 100   # both outcomes were exercised


def test_decompression_bomb_header_is_rejected_without_allocating():
    """A tiny file whose IHDR claims 65536 x 65536: rejected (deflate cannot expand that far), not a 17 GB allocation."""
    import struct
    import zlib

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))

    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", 65536, 65536, 8, 6, 0, 0, 0)) + \
        chunk(b"IDAT", zlib.compress(b"\0" * 64)) + chunk(b"IEND", b"")
    import ctypes as C
    L = hg._abi.load()
    w, h = C.c_int(0), C.c_int(0)
    buf = (C.c_uint8 * 16)()
    # header query succeeds (the chunks are well-formed); a decode into a too-small buffer is refused by the capacity check
    assert L.hg_png_decode(png, len(png), None, 0, C.byref(w), C.byref(h)) == 0 and (w.value, h.value) == (65536, 65536)
    assert L.hg_png_decode(png, len(png), buf, 16, C.byref(w), C.byref(h)) != 0
