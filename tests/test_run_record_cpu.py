"""The run-record builder of the fused piecewise path (pwf_make_record_from, csrc/piecewise_fused.cuh) compiled for the HOST
from the kernel's own source text and compared, record for record, with a column-by-column sweep of the bin.

The builder resolves the overlaps of a 64-column bin of the triangle map (H.js:1111-1126: later triangles overwrite earlier ones,
i.e. the highest id covering a column wins; H.js:848: the map is an Int16Array) and run-length encodes it.  The sweep below is
the definition: per column the highest covering id, the Int16 wrap, then runs of equal ids."""
import ctypes
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "homography.js_b200", "csrc", "piecewise_fused.cuh")

SHIM = r"""
#include <algorithm>
#include <cstdint>
#include <cstddef>
struct uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
#define __device__
#define __forceinline__ inline
using std::min;
using std::max;
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline void atomicOr(int *p, int v) { *p |= v; }
constexpr int PW_BIN_W = 64;
constexpr int PW_BIN_CAP = 8;
"""

DRIVER = r"""
extern "C" void make_records(const unsigned *ent, const unsigned *cnt, const int *full, int n_tris, int n, unsigned *rec, int *status)
{
    for (int i = 0; i < n; ++i) {
        alignas(16) unsigned e[PW_BIN_CAP];
        for (int k = 0; k < PW_BIN_CAP; ++k) e[k] = ent[(size_t)i * PW_BIN_CAP + k];
        uint4 r0, r1;
        status[i] = 0;
        pwf_make_record_from(e, cnt[i], full[i], n_tris, status + i, r0, r1);
        unsigned *o = rec + (size_t)i * 8;
        o[0] = r0.x; o[1] = r0.y; o[2] = r0.z; o[3] = r0.w; o[4] = r1.x; o[5] = r1.y; o[6] = r1.z; o[7] = r1.w;
    }
}
"""


def _extract(text: str, start_pat: str) -> str:
    """The function that starts at `start_pat`, up to its closing brace at column 0."""
    a = text.index(start_pat)
    b = text.index("\n}\n", a) + 3
    return text[a:b]


@pytest.fixture(scope="module")
def builder():
    text = open(SRC).read()
    body = (_extract(text, "__device__ __forceinline__ int pwf_map_id(") + "constexpr int PW_RUN_CAP = 8;\n" +
            _extract(text, "__device__ __forceinline__ void pwf_make_record_from("))
    assert re.search(r"reinterpret_cast<const uint4 \*>\(ent_p\)", body)
    with tempfile.TemporaryDirectory() as d:
        cpp, so = os.path.join(d, "rec.cpp"), os.path.join(d, "librec.so")
        open(cpp, "w").write(SHIM + body + DRIVER)
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-o", so, cpp])
        lib = ctypes.CDLL(so)
        lib.make_records.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        yield lib


def _sweep(ent, cnt, full, n_tris):
    """(mask, eight int16 ids, overflow) by a column-by-column sweep."""
    cols = np.full(64, full, np.int64)
    for e in ent[:min(cnt, 8)]:
        lo, hi, t = int(e) & 127, (int(e) >> 7) & 127, int(e) >> 14
        cols[lo:hi] = np.maximum(cols[lo:hi], t)
    ids16 = []
    for raw in cols:
        t = -1 if raw < 0 else int(np.int16(np.uint16(int(raw) & 0xFFFF)))
        ids16.append(t if 0 <= t < n_tris else -1)
    mask, ids, prev = 0, [], None
    for c, t in enumerate(ids16):
        if t != prev:
            if len(ids) == 8:
                return mask, ids, True
            mask |= 1 << c
            ids.append(t)
            prev = t
    return mask, ids + [-1] * (8 - len(ids)), False


def _cases(rng, n, n_tris, max_cnt, id_hi):
    ent = np.zeros((n, 8), np.uint32)
    cnt = rng.integers(0, max_cnt + 1, n).astype(np.uint32)
    full = np.where(rng.random(n) < 0.5, -1, rng.integers(0, id_hi, n)).astype(np.int32)
    for i in range(n):
        for k in range(int(cnt[i])):
            kind = rng.integers(0, 4)
            lo = 0 if kind == 0 else int(rng.integers(0, 64))
            hi = 64 if kind == 1 else int(rng.integers(lo + 1, 65))
            if kind == 3 and k:   # abut the previous entry
                lo = min(int((ent[i, k - 1] >> 7) & 127), 63)
                hi = int(rng.integers(lo + 1, 65))
            t = int(ent[i, k - 1] >> 14) if (k and rng.random() < 0.1) else int(rng.integers(0, id_hi))
            ent[i, k] = (t << 14) | (hi << 7) | lo
    return ent, cnt, full


@pytest.mark.parametrize("n_tris,max_cnt,id_hi", [(40, 3, 40), (7938, 4, 7938), (100, 8, 120), (1 << 17, 8, 1 << 17), (5, 8, 6)])
def test_run_records_equal_a_column_sweep(builder, n_tris, max_cnt, id_hi):
    rng = np.random.default_rng(n_tris + max_cnt)
    n = 6000
    ent, cnt, full = _cases(rng, n, n_tris, max_cnt, id_hi)
    rec = np.zeros((n, 8), np.uint32)
    status = np.zeros(n, np.int32)
    builder.make_records(ent.ctypes.data, cnt.ctypes.data, full.ctypes.data, n_tris, n, rec.ctypes.data, status.ctypes.data)
    overflows = 0
    for i in range(n):
        mask, ids, over = _sweep(ent[i], int(cnt[i]), int(full[i]), n_tris)
        assert bool(status[i]) == over, (i, ent[i], cnt[i], full[i])
        if over:
            overflows += 1
            continue
        got_mask = int(rec[i, 0]) | (int(rec[i, 1]) << 32)
        got_ids = [int(np.int16(np.uint16((int(rec[i, 4 + k // 2]) >> (16 * (k & 1))) & 0xFFFF))) for k in range(8)]
        assert (got_mask, got_ids) == (mask, ids), (i, [hex(int(v)) for v in ent[i]], cnt[i], full[i])
    if max_cnt == 8:
        assert overflows > 0   # the more-than-eight-runs flag is exercised
