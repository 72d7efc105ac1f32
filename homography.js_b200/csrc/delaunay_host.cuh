// delaunay_host.cuh — host-side Delaunay triangulation behind hg_delaunay (include/hgwarp.h).
//
// The reference triangulates its source points with a third-party package that is NOT part of its tree:
//     import Delaunator from 'https://cdn.skypack.dev/delaunator@5.0.0';                       H.js:27
//     function Delaunay(points) { return new Delaunator(points).triangles; }                    H.js:1216-1218
// (package.json:10-12 pins delaunator ^5.0.0, package-lock.json:17-24 resolves 5.0.0 -> robust-predicates 3.0.1).
// Which triangles exist, their ORDER (it decides who wins shared edges and overlaps in the index map, Q7) and their
// VERTEX order (it decides the pivot of the affine solve, A6) all come from that package, so this file restates the
// algorithm delaunator 5.0.0 publishes, step for step, in IEEE double without contraction (JS Numbers):
//   1. seed: the point nearest the bounding-box centre, its nearest neighbour, and the third point giving the smallest
//      circumcircle; made counter-clockwise by the robust orientation test;
//   2. all points sorted by squared distance from the seed triangle's circumcentre (the package's own quicksort:
//      insertion sort below 21 elements, median-of-three partition above — ties keep ITS order, so it is restated too);
//   3. sweep: each point is joined to the edges of the advancing convex hull it can see (found through an angular hash
//      of the hull), walking forward and backward, and every new triangle is legalised by edge flips (in-circle test,
//      explicit stack of 512 edges);
//   4. triangles are emitted in creation order, three vertex ids each.
// The orientation predicate is exact (robust-predicates' orient2d returns the sign of the exact determinant): a
// floating-point filter first, then an exact expansion of the six products.  The in-circle test is the package's plain
// floating-point formula.  NO fixture of the reference pins any triangulation: order-compatibility with the
// real package is by construction only — "parity unpinned" — and tests/ checks this file against an independent
// Python restatement (oracle/delaunator_ref.py) and, as a set of triangles, against scipy's Delaunay.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

namespace hg_delaunay_detail {

// ---- exact sign of a sum of doubles (Shewchuk-style grow-expansion with zero elimination)
inline void two_sum(double a, double b, double &s, double &e)
{
    s = a + b;
    const double bb = s - a;
    e = (a - (s - bb)) + (b - bb);
}

inline void two_product(double a, double b, double &p, double &e)
{
    p = a * b;
    e = std::fma(a, b, -p);
}

// adds b to the non-overlapping expansion e[0..n) (increasing magnitude), in place
inline int grow_expansion(double *e, int n, double b)
{
    double q = b;
    int m = 0;
    for (int i = 0; i < n; ++i) {
        double s, err;
        two_sum(q, e[i], s, err);
        q = s;
        if (err != 0.0) e[m++] = err;
    }
    if (q != 0.0 || m == 0) e[m++] = q;
    return m;
}

// sign-exact orientation determinant (ay-cy)(bx-cx) - (ax-cx)(by-cy), the convention of robust-predicates' orient2d
// (negative when a, b, c turn counter-clockwise in a y-up frame)
inline double orient2d(double ax, double ay, double bx, double by, double cx, double cy)
{
    const double detleft = (ay - cy) * (bx - cx);
    const double detright = (ax - cx) * (by - cy);
    const double det = detleft - detright;
    const double detsum = std::fabs(detleft + detright);
    if (std::fabs(det) >= 3.3306690738754716e-16 * detsum) return det;
    // exact: det = ay*bx - ay*cx - cy*bx - ax*by + ax*cy + cx*by   (the cx*cy terms cancel)
    const double f[6][2] = {{ay, bx}, {-ay, cx}, {-cy, bx}, {-ax, by}, {ax, cy}, {cx, by}};
    double e[16];
    int n = 0;
    for (int i = 0; i < 6; ++i) {
        double p, t;
        two_product(f[i][0], f[i][1], p, t);
        n = grow_expansion(e, n, t);
        n = grow_expansion(e, n, p);
    }
    return e[n - 1];  // the most significant component carries the sign of the exact sum
}

inline double dist2(double ax, double ay, double bx, double by)
{
    const double dx = ax - bx, dy = ay - by;
    return dx * dx + dy * dy;
}

inline bool in_circle(double ax, double ay, double bx, double by, double cx, double cy, double px, double py)
{
    const double dx = ax - px, dy = ay - py, ex = bx - px, ey = by - py, fx = cx - px, fy = cy - py;
    const double ap = dx * dx + dy * dy, bp = ex * ex + ey * ey, cp = fx * fx + fy * fy;
    return dx * (ey * cp - bp * fy) - dy * (ex * cp - bp * fx) + ap * (ex * fy - ey * fx) < 0;
}

inline double circumradius2(double ax, double ay, double bx, double by, double cx, double cy)
{
    const double dx = bx - ax, dy = by - ay, ex = cx - ax, ey = cy - ay;
    const double bl = dx * dx + dy * dy, cl = ex * ex + ey * ey, d = 0.5 / (dx * ey - dy * ex);
    const double x = (ey * bl - dy * cl) * d, y = (dx * cl - ex * bl) * d;
    return x * x + y * y;
}

inline void circumcenter(double ax, double ay, double bx, double by, double cx, double cy, double &ox, double &oy)
{
    const double dx = bx - ax, dy = by - ay, ex = cx - ax, ey = cy - ay;
    const double bl = dx * dx + dy * dy, cl = ex * ex + ey * ey, d = 0.5 / (dx * ey - dy * ex);
    ox = ax + (ey * bl - dy * cl) * d;
    oy = ay + (dx * cl - ex * bl) * d;
}

// monotone in the true angle, no trigonometry; in [0, 1]
inline double pseudo_angle(double dx, double dy)
{
    const double p = dx / (std::fabs(dx) + std::fabs(dy));
    return (dy > 0 ? 3 - p : 1 + p) / 4;
}

// the package's sort of ids by dists[id]: its tie behaviour is part of the triangle order
inline void sort_ids(std::vector<uint32_t> &ids, const std::vector<double> &d, long left, long right)
{
    auto swp = [&](long i, long j) {
        const uint32_t t = ids[(size_t)i];
        ids[(size_t)i] = ids[(size_t)j];
        ids[(size_t)j] = t;
    };
    if (right - left <= 20) {
        for (long i = left + 1; i <= right; ++i) {
            const uint32_t tmp = ids[(size_t)i];
            const double td = d[tmp];
            long j = i - 1;
            while (j >= left && d[ids[(size_t)j]] > td) {
                ids[(size_t)j + 1] = ids[(size_t)j];
                --j;
            }
            ids[(size_t)(j + 1)] = tmp;
        }
        return;
    }
    const long median = (left + right) >> 1;
    long i = left + 1, j = right;
    swp(median, i);
    if (d[ids[(size_t)left]] > d[ids[(size_t)right]]) swp(left, right);
    if (d[ids[(size_t)i]] > d[ids[(size_t)right]]) swp(i, right);
    if (d[ids[(size_t)left]] > d[ids[(size_t)i]]) swp(left, i);
    const uint32_t tmp = ids[(size_t)i];
    const double td = d[tmp];
    for (;;) {
        do ++i; while (d[ids[(size_t)i]] < td);
        do --j; while (d[ids[(size_t)j]] > td);
        if (j < i) break;
        swp(i, j);
    }
    ids[(size_t)left + 1] = ids[(size_t)j];
    ids[(size_t)j] = tmp;
    if (right - i + 1 >= j - left) {
        sort_ids(ids, d, i, right);
        sort_ids(ids, d, left, j - 1);
    } else {
        sort_ids(ids, d, left, j - 1);
        sort_ids(ids, d, i, right);
    }
}

struct Sweep {
    const double *xy;
    size_t n;
    std::vector<uint32_t> tri;
    std::vector<int32_t> half;
    std::vector<uint32_t> hull_prev, hull_next, hull_tri;
    std::vector<int32_t> hull_hash;
    size_t hash_size = 0, tri_len = 0;
    uint32_t hull_start = 0;
    double ccx = 0, ccy = 0;
    uint32_t edge_stack[512];

    size_t hash_key(double x, double y) const
    {
        const double a = pseudo_angle(x - ccx, y - ccy) * (double)hash_size;
        if (!(a == a)) return 0;  // the point IS the circumcentre (0 / 0): any bucket will do
        return (size_t)((long long)std::floor(a) % (long long)hash_size);
    }
    void link(long a, long b)
    {
        half[(size_t)a] = (int32_t)b;
        if (b != -1) half[(size_t)b] = (int32_t)a;
    }
    size_t add_triangle(uint32_t i0, uint32_t i1, uint32_t i2, long a, long b, long c)
    {
        const size_t t = tri_len;
        tri[t] = i0;
        tri[t + 1] = i1;
        tri[t + 2] = i2;
        link((long)t, a);
        link((long)t + 1, b);
        link((long)t + 2, c);
        tri_len += 3;
        return t;
    }
    // flips edges around half-edge a until the Delaunay condition holds; returns the half-edge to keep on the hull
    uint32_t legalize(size_t a)
    {
        size_t depth = 0, ar = 0;
        for (;;) {
            const long b = half[a];
            const size_t a0 = a - a % 3;
            ar = a0 + (a + 2) % 3;
            if (b == -1) {  // hull edge
                if (depth == 0) break;
                a = edge_stack[--depth];
                continue;
            }
            const size_t b0 = (size_t)b - (size_t)b % 3;
            const size_t al = a0 + (a + 1) % 3, bl = b0 + ((size_t)b + 2) % 3;
            const uint32_t p0 = tri[ar], pr = tri[a], pl = tri[al], p1 = tri[bl];
            if (in_circle(xy[2 * p0], xy[2 * p0 + 1], xy[2 * pr], xy[2 * pr + 1], xy[2 * pl], xy[2 * pl + 1], xy[2 * p1], xy[2 * p1 + 1])) {
                tri[a] = p1;
                tri[(size_t)b] = p0;
                const long hbl = half[bl];
                if (hbl == -1) {  // the flipped edge was on the hull: repoint the hull's triangle reference
                    uint32_t e = hull_start;
                    do {
                        if (hull_tri[e] == bl) {
                            hull_tri[e] = (uint32_t)a;
                            break;
                        }
                        e = hull_prev[e];
                    } while (e != hull_start);
                }
                link((long)a, hbl);
                link(b, half[ar]);
                link((long)ar, (long)bl);
                const size_t br = b0 + ((size_t)b + 1) % 3;
                if (depth < 512) edge_stack[depth++] = (uint32_t)br;
            } else {
                if (depth == 0) break;
                a = edge_stack[--depth];
            }
        }
        return (uint32_t)ar;
    }
};

// triangles (vertex ids, three per triangle, in creation order) of the points xy[0..2n); empty for n < 3 or collinear input
inline std::vector<uint32_t> triangulate(const double *xy, size_t n)
{
    std::vector<uint32_t> none;
    if (n < 3) return none;
    const double inf = std::numeric_limits<double>::infinity();
    Sweep S;
    S.xy = xy;
    S.n = n;
    const size_t max_tri = 2 * n >= 5 ? 2 * n - 5 : 0;
    S.tri.assign(max_tri * 3, 0);
    S.half.assign(max_tri * 3, -1);
    S.hash_size = (size_t)std::ceil(std::sqrt((double)n));
    S.hull_prev.assign(n, 0);
    S.hull_next.assign(n, 0);
    S.hull_tri.assign(n, 0);
    S.hull_hash.assign(S.hash_size, -1);
    std::vector<uint32_t> ids(n);
    std::vector<double> dists(n);

    double minx = inf, miny = inf, maxx = -inf, maxy = -inf;
    for (size_t i = 0; i < n; ++i) {
        const double x = xy[2 * i], y = xy[2 * i + 1];
        if (x < minx) minx = x;
        if (y < miny) miny = y;
        if (x > maxx) maxx = x;
        if (y > maxy) maxy = y;
        ids[i] = (uint32_t)i;
    }
    const double cx = (minx + maxx) / 2, cy = (miny + maxy) / 2;

    size_t i0 = 0, i1 = 0, i2 = 0;
    bool have = false;
    double best = inf;
    for (size_t i = 0; i < n; ++i) {
        const double d = dist2(cx, cy, xy[2 * i], xy[2 * i + 1]);
        if (d < best) { i0 = i; best = d; have = true; }
    }
    if (!have) return none;  // NaN coordinates
    const double i0x = xy[2 * i0], i0y = xy[2 * i0 + 1];
    best = inf;
    have = false;
    for (size_t i = 0; i < n; ++i) {
        if (i == i0) continue;
        const double d = dist2(i0x, i0y, xy[2 * i], xy[2 * i + 1]);
        if (d < best && d > 0) { i1 = i; best = d; have = true; }
    }
    if (!have) return none;  // all points coincide
    double i1x = xy[2 * i1], i1y = xy[2 * i1 + 1];
    double min_r = inf;
    for (size_t i = 0; i < n; ++i) {
        if (i == i0 || i == i1) continue;
        const double r = circumradius2(i0x, i0y, i1x, i1y, xy[2 * i], xy[2 * i + 1]);
        if (r < min_r) { i2 = i; min_r = r; }
    }
    if (min_r == inf) return none;  // collinear: the package returns only a hull, no triangles
    double i2x = xy[2 * i2], i2y = xy[2 * i2 + 1];
    if (orient2d(i0x, i0y, i1x, i1y, i2x, i2y) < 0) {
        const size_t t = i1;
        const double tx = i1x, ty = i1y;
        i1 = i2; i1x = i2x; i1y = i2y;
        i2 = t; i2x = tx; i2y = ty;
    }
    circumcenter(i0x, i0y, i1x, i1y, i2x, i2y, S.ccx, S.ccy);
    for (size_t i = 0; i < n; ++i) dists[i] = dist2(xy[2 * i], xy[2 * i + 1], S.ccx, S.ccy);
    sort_ids(ids, dists, 0, (long)n - 1);

    S.hull_start = (uint32_t)i0;
    S.hull_next[i0] = S.hull_prev[i2] = (uint32_t)i1;
    S.hull_next[i1] = S.hull_prev[i0] = (uint32_t)i2;
    S.hull_next[i2] = S.hull_prev[i1] = (uint32_t)i0;
    S.hull_tri[i0] = 0;
    S.hull_tri[i1] = 1;
    S.hull_tri[i2] = 2;
    S.hull_hash[S.hash_key(i0x, i0y)] = (int32_t)i0;
    S.hull_hash[S.hash_key(i1x, i1y)] = (int32_t)i1;
    S.hull_hash[S.hash_key(i2x, i2y)] = (int32_t)i2;
    S.add_triangle((uint32_t)i0, (uint32_t)i1, (uint32_t)i2, -1, -1, -1);

    const double eps = 2.220446049250313e-16;  // 2^-52
    double xp = 0, yp = 0;
    for (size_t k = 0; k < n; ++k) {
        const uint32_t i = ids[k];
        const double x = xy[2 * (size_t)i], y = xy[2 * (size_t)i + 1];
        if (k > 0 && std::fabs(x - xp) <= eps && std::fabs(y - yp) <= eps) continue;  // near-duplicate
        xp = x;
        yp = y;
        if (i == i0 || i == i1 || i == i2) continue;

        // a hull vertex near the point's direction, through the angular hash
        long start = 0;
        const size_t key = S.hash_key(x, y);
        for (size_t j = 0; j < S.hash_size; ++j) {
            start = S.hull_hash[(key + j) % S.hash_size];
            if (start != -1 && (uint32_t)start != S.hull_next[(size_t)start]) break;
        }
        start = S.hull_prev[(size_t)start];
        long e = start;
        uint32_t q;
        for (;;) {
            q = S.hull_next[(size_t)e];
            if (!(orient2d(x, y, xy[2 * (size_t)e], xy[2 * (size_t)e + 1], xy[2 * (size_t)q], xy[2 * (size_t)q + 1]) >= 0)) break;
            e = q;
            if (e == start) {
                e = -1;
                break;
            }
        }
        if (e == -1) continue;  // no visible edge: a near-duplicate

        size_t t = S.add_triangle((uint32_t)e, i, S.hull_next[(size_t)e], -1, -1, (long)S.hull_tri[(size_t)e]);
        S.hull_tri[i] = S.legalize(t + 2);
        S.hull_tri[(size_t)e] = (uint32_t)t;

        // forward along the hull
        uint32_t nn = S.hull_next[(size_t)e];
        for (;;) {
            q = S.hull_next[nn];
            if (!(orient2d(x, y, xy[2 * (size_t)nn], xy[2 * (size_t)nn + 1], xy[2 * (size_t)q], xy[2 * (size_t)q + 1]) < 0)) break;
            t = S.add_triangle(nn, i, q, (long)S.hull_tri[i], -1, (long)S.hull_tri[nn]);
            S.hull_tri[i] = S.legalize(t + 2);
            S.hull_next[nn] = nn;  // removed from the hull
            nn = q;
        }
        // backward from the other side
        if (e == start) {
            for (;;) {
                q = S.hull_prev[(size_t)e];
                if (!(orient2d(x, y, xy[2 * (size_t)q], xy[2 * (size_t)q + 1], xy[2 * (size_t)e], xy[2 * (size_t)e + 1]) < 0)) break;
                t = S.add_triangle(q, i, (uint32_t)e, -1, (long)S.hull_tri[(size_t)e], (long)S.hull_tri[q]);
                S.legalize(t + 2);
                S.hull_tri[q] = (uint32_t)t;
                S.hull_next[(size_t)e] = (uint32_t)e;  // removed from the hull
                e = q;
            }
        }
        S.hull_start = S.hull_prev[i] = (uint32_t)e;
        S.hull_next[(size_t)e] = S.hull_prev[nn] = i;
        S.hull_next[i] = nn;
        S.hull_hash[S.hash_key(x, y)] = (int32_t)i;
        S.hull_hash[S.hash_key(xy[2 * (size_t)e], xy[2 * (size_t)e + 1])] = (int32_t)e;
    }
    S.tri.resize(S.tri_len);
    return S.tri;
}

}  // namespace hg_delaunay_detail
