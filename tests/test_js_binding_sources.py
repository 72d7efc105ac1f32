"""The Node.js face cannot run here (no JavaScript engine, no node_api.h in the image): what CAN be checked is that
the N-API addon source compiles against the hand-declared Node-API subset, binds only symbols the C ABI declares,
and that the JS shim calls exactly the functions the addon registers."""
import os
import re
import subprocess

import homography_js_b200 as hg
from conftest import ROOT

JS = os.path.join(ROOT, "homography.js_b200", "js")


def test_napi_addon_compiles():
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    subprocess.check_call([gcc, "-std=c11", "-Wall", "-Wextra", "-Werror", "-fsyntax-only",
                           "-I" + os.path.join(ROOT, "include"), os.path.join(JS, "hgwarp_napi.c")], env=env)


def test_addon_binds_only_declared_abi_symbols():
    src = open(os.path.join(JS, "hgwarp_napi.c")).read()
    used = set(re.findall(r"\b(hg_[a-z0-9_]+)\s*\(", src))
    assert used and used <= set(hg._abi.SYMBOLS), used - set(hg._abi.SYMBOLS)


def test_js_shim_calls_match_addon_exports():
    addon = open(os.path.join(JS, "hgwarp_napi.c")).read()
    exported = set(re.findall(r'\{"([A-Za-z]+)", 0, [A-Za-z]+, 0, 0, 0, napi_default, 0\}', addon))
    shim = open(os.path.join(JS, "homography_b200.mjs")).read()
    called = set(re.findall(r"\bnative\.([A-Za-z]+)\(", shim))
    assert called == exported, (called, exported)
    for name in ("setReferencePoints", "setSourcePoints", "setDestinyPoints", "setImage", "setTriangles", "warp",
                 "getTransformationMatrixAsCSS", "transformHTMLElement"):
        assert re.search(r"\n  %s\(" % name, shim), name
    assert "export { Homography" in shim


def test_js_shim_lexes_cleanly_and_brackets_balance():
    """No JavaScript engine here: at least the shim must tokenize without error tokens (pygments' ECMAScript lexer) and its
    brackets must balance outside strings, template literals and comments."""
    pygments = __import__("pytest").importorskip("pygments")
    from pygments.lexers import JavascriptLexer
    from pygments.token import Error, Punctuation
    src = open(os.path.join(JS, "homography_b200.mjs")).read()
    stack, pairs = [], {")": "(", "]": "[", "}": "{"}
    for tok, text in JavascriptLexer().get_tokens(src):
        assert tok is not Error, repr(text)
        if tok in Punctuation:
            for ch in text:
                if ch in "([{":
                    stack.append(ch)
                elif ch in ")]}":
                    assert stack and stack.pop() == pairs[ch], "unbalanced " + ch
    # template literals `${...}` are lexed as string parts with their own interpolation tokens; whatever remains must be closed
    assert not stack, stack
