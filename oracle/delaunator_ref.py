"""TEST INFRASTRUCTURE (oracle) — never imported by the product.

Pure-Python restatement of the triangulation the reference obtains from a third-party package:

    import Delaunator from 'https://cdn.skypack.dev/delaunator@5.0.0'          Homography.js:27
    function Delaunay(points) { return new Delaunator(points).triangles }       Homography.js:1216-1218

delaunator 5.0.0 (-> robust-predicates 3.0.1; package.json:10-12, package-lock.json:17-24,33-37) is NOT under
/root/reference, so this follows the algorithm the package publishes: seed triangle near the bounding-box centre,
points sorted by distance from its circumcentre with the package's own quicksort, sweep over an advancing convex hull
found through an angular hash, edge flips by the in-circle test with a 512-entry stack.  Python floats are IEEE doubles
evaluated without contraction, i.e. JS Numbers.  The orientation predicate is exact: here through `fractions.Fraction`
(an independent route from the C++ product code, which uses floating-point expansions).

PARITY UNPINNED: the reference holds no triangulation fixture; this file and csrc/delaunay_host.cuh are two independent
restatements that must agree bit for bit (tests/test_delaunay.py), and both must give scipy's Delaunay triangle SET on
points in general position.
"""
import math
from fractions import Fraction

EPSILON = 2.0 ** -52


def orient2d(ax, ay, bx, by, cx, cy):
    """Sign-exact (ay-cy)(bx-cx) - (ax-cx)(by-cy), the convention of robust-predicates' orient2d."""
    detleft = (ay - cy) * (bx - cx)
    detright = (ax - cx) * (by - cy)
    det = detleft - detright
    if abs(det) >= 3.3306690738754716e-16 * abs(detleft + detright):
        return det
    if not all(math.isfinite(v) for v in (ax, ay, bx, by, cx, cy)):
        return det
    f = Fraction
    exact = (f(ay) - f(cy)) * (f(bx) - f(cx)) - (f(ax) - f(cx)) * (f(by) - f(cy))
    return 1.0 if exact > 0 else (-1.0 if exact < 0 else 0.0)


def _dist(ax, ay, bx, by):
    dx = ax - bx
    dy = ay - by
    return dx * dx + dy * dy


def _in_circle(ax, ay, bx, by, cx, cy, px, py):
    dx = ax - px; dy = ay - py; ex = bx - px; ey = by - py; fx = cx - px; fy = cy - py
    ap = dx * dx + dy * dy; bp = ex * ex + ey * ey; cp = fx * fx + fy * fy
    return dx * (ey * cp - bp * fy) - dy * (ex * cp - bp * fx) + ap * (ex * fy - ey * fx) < 0


def _div(a, b):
    """JS division: x / 0 is +-Infinity or NaN, never an exception."""
    if b == 0:
        if a == 0 or a != a:
            return math.nan
        return math.copysign(math.inf, a) * math.copysign(1.0, b)
    return a / b


def _circum(ax, ay, bx, by, cx, cy):
    dx = bx - ax; dy = by - ay; ex = cx - ax; ey = cy - ay
    bl = dx * dx + dy * dy; cl = ex * ex + ey * ey
    d = _div(0.5, dx * ey - dy * ex)
    return (ey * bl - dy * cl) * d, (dx * cl - ex * bl) * d


def _circumradius(ax, ay, bx, by, cx, cy):
    x, y = _circum(ax, ay, bx, by, cx, cy)
    return x * x + y * y


def _pseudo_angle(dx, dy):
    p = _div(dx, abs(dx) + abs(dy))
    return (3 - p if dy > 0 else 1 + p) / 4


def _quicksort(ids, dists, left, right):
    if right - left <= 20:
        for i in range(left + 1, right + 1):
            temp = ids[i]
            td = dists[temp]
            j = i - 1
            while j >= left and dists[ids[j]] > td:
                ids[j + 1] = ids[j]
                j -= 1
            ids[j + 1] = temp
        return
    median = (left + right) >> 1
    i = left + 1
    j = right
    ids[median], ids[i] = ids[i], ids[median]
    if dists[ids[left]] > dists[ids[right]]:
        ids[left], ids[right] = ids[right], ids[left]
    if dists[ids[i]] > dists[ids[right]]:
        ids[i], ids[right] = ids[right], ids[i]
    if dists[ids[left]] > dists[ids[i]]:
        ids[left], ids[i] = ids[i], ids[left]
    temp = ids[i]
    td = dists[temp]
    while True:
        i += 1
        while dists[ids[i]] < td:
            i += 1
        j -= 1
        while dists[ids[j]] > td:
            j -= 1
        if j < i:
            break
        ids[i], ids[j] = ids[j], ids[i]
    ids[left + 1] = ids[j]
    ids[j] = temp
    if right - i + 1 >= j - left:
        _quicksort(ids, dists, i, right)
        _quicksort(ids, dists, left, j - 1)
    else:
        _quicksort(ids, dists, left, j - 1)
        _quicksort(ids, dists, i, right)


def triangles(points):
    """`new Delaunator(points).triangles` as a flat list of vertex ids (three per triangle)."""
    c = [float(v) for v in points]
    n = len(c) >> 1
    if n < 3:
        return []
    max_tri = max(2 * n - 5, 0)
    tri = [0] * (max_tri * 3)
    half = [-1] * (max_tri * 3)
    hash_size = math.ceil(math.sqrt(n))
    hull_prev = [0] * n
    hull_next = [0] * n
    hull_tri = [0] * n
    hull_hash = [-1] * hash_size
    ids = list(range(n))
    dists = [0.0] * n
    xs = c[0::2]
    ys = c[1::2]
    min_x = math.inf; min_y = math.inf; max_x = -math.inf; max_y = -math.inf
    for i in range(n):
        if xs[i] < min_x: min_x = xs[i]
        if ys[i] < min_y: min_y = ys[i]
        if xs[i] > max_x: max_x = xs[i]
        if ys[i] > max_y: max_y = ys[i]
    cx = (min_x + max_x) / 2
    cy = (min_y + max_y) / 2

    i0 = i1 = i2 = None
    best = math.inf
    for i in range(n):
        d = _dist(cx, cy, xs[i], ys[i])
        if d < best:
            i0, best = i, d
    if i0 is None:
        return []
    best = math.inf
    for i in range(n):
        if i == i0:
            continue
        d = _dist(xs[i0], ys[i0], xs[i], ys[i])
        if d < best and d > 0:
            i1, best = i, d
    if i1 is None:
        return []
    min_r = math.inf
    for i in range(n):
        if i == i0 or i == i1:
            continue
        r = _circumradius(xs[i0], ys[i0], xs[i1], ys[i1], xs[i], ys[i])
        if r < min_r:
            i2, min_r = i, r
    if min_r == math.inf:
        return []  # collinear input: a hull, no triangles
    if orient2d(xs[i0], ys[i0], xs[i1], ys[i1], xs[i2], ys[i2]) < 0:
        i1, i2 = i2, i1
    ox, oy = _circum(xs[i0], ys[i0], xs[i1], ys[i1], xs[i2], ys[i2])
    ccx = xs[i0] + ox
    ccy = ys[i0] + oy
    for i in range(n):
        dists[i] = _dist(xs[i], ys[i], ccx, ccy)
    _quicksort(ids, dists, 0, n - 1)

    def hash_key(x, y):
        a = _pseudo_angle(x - ccx, y - ccy) * hash_size
        if a != a:
            return 0
        return math.floor(a) % hash_size

    state = {"len": 0, "start": i0}

    def link(a, b):
        half[a] = b
        if b != -1:
            half[b] = a

    def add_triangle(a0, a1, a2, a, b, cc):
        t = state["len"]
        tri[t] = a0; tri[t + 1] = a1; tri[t + 2] = a2
        link(t, a); link(t + 1, b); link(t + 2, cc)
        state["len"] += 3
        return t

    stack = []

    def legalize(a):
        ar = 0
        while True:
            b = half[a]
            a0 = a - a % 3
            ar = a0 + (a + 2) % 3
            if b == -1:
                if not stack:
                    break
                a = stack.pop()
                continue
            b0 = b - b % 3
            al = a0 + (a + 1) % 3
            bl = b0 + (b + 2) % 3
            p0 = tri[ar]; pr = tri[a]; pl = tri[al]; p1 = tri[bl]
            if _in_circle(xs[p0], ys[p0], xs[pr], ys[pr], xs[pl], ys[pl], xs[p1], ys[p1]):
                tri[a] = p1
                tri[b] = p0
                hbl = half[bl]
                if hbl == -1:
                    e = state["start"]
                    while True:
                        if hull_tri[e] == bl:
                            hull_tri[e] = a
                            break
                        e = hull_prev[e]
                        if e == state["start"]:
                            break
                link(a, hbl)
                link(b, half[ar])
                link(ar, bl)
                br = b0 + (b + 1) % 3
                if len(stack) < 512:
                    stack.append(br)
            else:
                if not stack:
                    break
                a = stack.pop()
        return ar

    hull_next[i0] = hull_prev[i2] = i1
    hull_next[i1] = hull_prev[i0] = i2
    hull_next[i2] = hull_prev[i1] = i0
    hull_tri[i0] = 0; hull_tri[i1] = 1; hull_tri[i2] = 2
    hull_hash[hash_key(xs[i0], ys[i0])] = i0
    hull_hash[hash_key(xs[i1], ys[i1])] = i1
    hull_hash[hash_key(xs[i2], ys[i2])] = i2
    add_triangle(i0, i1, i2, -1, -1, -1)

    xp = yp = 0.0
    for k in range(n):
        i = ids[k]
        x = xs[i]; y = ys[i]
        if k > 0 and abs(x - xp) <= EPSILON and abs(y - yp) <= EPSILON:
            continue
        xp = x; yp = y
        if i == i0 or i == i1 or i == i2:
            continue
        start = 0
        key = hash_key(x, y)
        for j in range(hash_size):
            start = hull_hash[(key + j) % hash_size]
            if start != -1 and start != hull_next[start]:
                break
        start = hull_prev[start]
        e = start
        while True:
            q = hull_next[e]
            if not (orient2d(x, y, xs[e], ys[e], xs[q], ys[q]) >= 0):
                break
            e = q
            if e == start:
                e = -1
                break
        if e == -1:
            continue
        t = add_triangle(e, i, hull_next[e], -1, -1, hull_tri[e])
        hull_tri[i] = legalize(t + 2)
        hull_tri[e] = t
        nn = hull_next[e]
        while True:
            q = hull_next[nn]
            if not (orient2d(x, y, xs[nn], ys[nn], xs[q], ys[q]) < 0):
                break
            t = add_triangle(nn, i, q, hull_tri[i], -1, hull_tri[nn])
            hull_tri[i] = legalize(t + 2)
            hull_next[nn] = nn
            nn = q
        if e == start:
            while True:
                q = hull_prev[e]
                if not (orient2d(x, y, xs[q], ys[q], xs[e], ys[e]) < 0):
                    break
                t = add_triangle(q, i, e, -1, hull_tri[e], hull_tri[q])
                legalize(t + 2)
                hull_tri[q] = t
                hull_next[e] = e
                e = q
        state["start"] = hull_prev[i] = e
        hull_next[e] = hull_prev[nn] = i
        hull_next[i] = nn
        hull_hash[hash_key(x, y)] = i
        hull_hash[hash_key(xs[e], ys[e])] = e
    return tri[:state["len"]]
