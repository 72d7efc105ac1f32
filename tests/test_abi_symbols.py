"""The C-ABI library builds for sm_100a, loads, and exports exactly the symbols include/hgwarp.h declares.
No compute calls (no GPU here)."""
import ctypes
import os
import re
import subprocess

import homography_js_b200 as hg
from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "hgwarp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_match_binding_table():
    assert _declared() == sorted(hg._abi.SYMBOLS)


def test_library_builds_and_exports_every_declared_symbol():
    import importlib.util
    spec = importlib.util.spec_from_file_location("hg_build", os.path.join(ROOT, "homography.js_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    lib_path = mod.build()
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib_path], text=True)
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("hg_"))
    assert exported == _declared()
    L = ctypes.CDLL(lib_path)
    for name in _declared():
        assert hasattr(L, name), name
    assert hg._abi.load().hg_abi_version() == 1


def test_library_is_compiled_for_sm_100a():
    out = subprocess.run(["cuobjdump", "-lelf", hg._abi.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        import pytest
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_no_device_fails_loudly_not_silently():
    """Without a GPU, creating a context raises: there is no CPU fallback in the product."""
    import pytest
    if hg.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(hg.HgError):
        hg.Context(0)


def test_product_does_not_import_the_oracle():
    """The shipped package may MENTION the oracle in comments (where an extension's definition lives), but must never
    import, load, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "homography.js_b200")
    forbidden = re.compile(r"(^|\s)(import\s+oracle|from\s+oracle|from\s+\.\.?oracle)|libhgoracle|orc_[a-z_]+\s*\(|"
                           r"#include\s+[\"<][^\">]*oracle|sys\.path[^\n]*oracle", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".mjs", ".js", ".c", ".cc")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                m = forbidden.search(text)
                assert m is None, (os.path.join(dirpath, f), m.group(0))
