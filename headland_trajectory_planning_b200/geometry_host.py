"""Host-side construction of the static geometry that is uploaded to HBM.

These are the polygons the reference builds with shapely in its constructors
(``orchard_geometry_environment.py:277-353``, ``reference_line_heuristic.py:65-67``)
restated as vertex arrays: the GEOS buffer of a two-point line with flat caps is
an exact rectangle, ``Point.buffer(r, cap_style="square")`` is an axis-aligned
square, and the round-cap buffer (``quad_segs=16``) is a 66-vertex convex polygon.
Only construction happens here (once per environment); every per-pose predicate
runs on the GPU.
"""
import math

import numpy as np

QUAD_SEGS = 16
LANE_RADIUS = 6.0
CAPSULE_VERTS = 66
_R_IN = LANE_RADIUS * math.cos(math.pi / (4 * QUAD_SEGS)) - 1e-6
_R_OUT = LANE_RADIUS + 1e-6


class _Ring:
    def __init__(self, pts):
        self._pts = pts

    @property
    def xy(self):
        return (list(self._pts[:, 0]), list(self._pts[:, 1]))

    @property
    def coords(self):
        return [tuple(p) for p in self._pts]


class Poly:
    """Minimal stand-in for ``shapely.geometry.Polygon`` for the unchanged OBCA
    consumer, which reads ``.exterior.xy`` / ``.exterior.coords`` / ``.bounds``
    (``obca_py/optimizer.py:143-178``)."""

    def __init__(self, pts):
        pts = np.asarray(pts, dtype=np.float64)
        if not (pts[0] == pts[-1]).all():
            pts = np.vstack([pts, pts[:1]])
        self._pts = pts
        self.exterior = _Ring(pts)

    @property
    def bounds(self):
        return (self._pts[:, 0].min(), self._pts[:, 1].min(), self._pts[:, 0].max(), self._pts[:, 1].max())

    def __repr__(self):
        return "POLYGON ((" + ", ".join("%.3f %.3f" % tuple(p) for p in self._pts) + "))"


def open_ccw(pts):
    p = np.asarray(pts, dtype=np.float64)
    if len(p) > 1 and (p[0] == p[-1]).all():
        p = p[:-1]
    twice_area = np.sum(p[:, 0] * np.roll(p[:, 1], -1) - np.roll(p[:, 0], -1) * p[:, 1])
    return np.ascontiguousarray(p[::-1] if twice_area < 0 else p)


def flat_line_buffer(p0, p1, dist):
    """Tree-row rectangle, ``LineString(row).buffer(w/2, cap_style=2)``."""
    dx, dy = p1[0] - p0[0], p1[1] - p0[1]
    ln = math.sqrt(dx * dx + dy * dy)
    ux, uy = dist * dx / ln, dist * dy / ln
    return open_ccw([[p0[0] - uy, p0[1] + ux], [p1[0] - uy, p1[1] + ux],
                     [p1[0] + uy, p1[1] - ux], [p0[0] + uy, p0[1] - ux]])


def square_point_buffer(x, y, r):
    """Obstacle square, ``Point(x, y).buffer(r, cap_style="square")``."""
    return open_ccw([[x + r, y + r], [x + r, y - r], [x - r, y - r], [x - r, y + r]])


def capsule_polygon(p0, p1, r=LANE_RADIUS, quad_segs=QUAD_SEGS):
    """Lane capsule, ``LineString([p0, p1]).buffer(r, cap_style=1, join_style=3)``:
    offset segment left, fillet round p1, offset segment right, fillet round p0;
    fillet vertices at ``angle + pi/2 - k*pi/(2*quad_segs)``, k = 1..2*quad_segs-1."""
    dx, dy = p1[0] - p0[0], p1[1] - p0[1]
    ln = math.sqrt(dx * dx + dy * dy)
    ux, uy = r * dx / ln, r * dy / ln
    quantum = math.pi / 2.0 / quad_segs
    ring = [(p0[0] - uy, p0[1] + ux), (p1[0] - uy, p1[1] + ux)]
    for centre, ang in ((p1, math.atan2(dy, dx)), (p0, math.atan2(-dy, -dx))):
        start, end = ang + math.pi / 2, ang - math.pi / 2
        total = abs(start - end)
        nseg = int(total / quantum + 0.5)
        inc = total / nseg
        for i in range(1, nseg):
            a = start - i * inc
            ring.append((centre[0] + r * math.cos(a), centre[1] + r * math.sin(a)))
        if centre is p1:
            ring += [(p1[0] + uy, p1[1] - ux), (p0[0] + uy, p0[1] - ux)]
    out = open_ccw(ring)
    assert len(out) == CAPSULE_VERTS
    return out


def _strictly_inside_capsule(seg, poly, q):
    ax, ay, bx, by = seg
    ex, ey = bx - ax, by - ay
    t = ((q[0] - ax) * ex + (q[1] - ay) * ey) / (ex * ex + ey * ey)
    t = min(max(t, 0.0), 1.0)
    d = math.hypot(q[0] - (ax + t * ex), q[1] - (ay + t * ey))
    if d <= _R_IN:
        return True
    if d > _R_OUT:
        return False
    nxt = np.roll(poly, -1, axis=0)
    g = (nxt[:, 1] - poly[:, 1]) * (q[0] - poly[:, 0]) - (nxt[:, 0] - poly[:, 0]) * (q[1] - poly[:, 1])
    return bool((g < 0).all())


def lane_critical_points(segs, polys):
    """Vertices of the boundary of the union of capsules that are not vertices of a
    single capsule: crossings of two capsule boundaries not strictly inside a third.
    A footprint holding such a point in its interior is not inside the lane even
    when all its edges are covered."""
    out = []
    S = len(polys)
    for i in range(S):
        a0 = polys[i]
        a1 = np.roll(a0, -1, axis=0)
        for j in range(i + 1, S):
            b0 = polys[j]
            b1 = np.roll(b0, -1, axis=0)
            r = (a1 - a0)[:, None, :]
            s = (b1 - b0)[None, :, :]
            qp = b0[None, :, :] - a0[:, None, :]
            rxs = r[..., 0] * s[..., 1] - r[..., 1] * s[..., 0]
            with np.errstate(divide="ignore", invalid="ignore"):
                t = (qp[..., 0] * s[..., 1] - qp[..., 1] * s[..., 0]) / rxs
                u = (qp[..., 0] * r[..., 1] - qp[..., 1] * r[..., 0]) / rxs
            hit = (np.abs(rxs) > 1e-12) & (t >= 0) & (t <= 1) & (u >= 0) & (u <= 1)
            for a, b in zip(*np.nonzero(hit)):
                q = a0[a] + t[a, b] * (a1[a] - a0[a])
                if not any(_strictly_inside_capsule(segs[k], polys[k], q)
                           for k in range(S) if k not in (i, j)):
                    out.append(q)
    return np.array(out, dtype=np.float64).reshape(-1, 2)
