"""Exact float64 footprint predicates -- the CPU statement of what shapely/GEOS
computes for the reference on the warm-start path.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  **parity unpinned** at the GEOS
boundary: shapely is not installable here, so these predicates DEFINE parity for
the CUDA kernels.  They restate, for the shapes the reference actually builds
(SURVEY.md section 8a-8..a-11):

  * ``CarModel.get_path_poly`` + ``nearest.intersects(path_poly)``
    (car_model.py:39-73, orchard_geometry_environment.py:414-431): the swept body
    is a union of per-pose rectangles and the obstacles are closed convex
    polygons, so "any pose rectangle meets any obstacle (closed sets)" is the
    same boolean.  Per pair: separating-axis test with STRICT separation.
  * ``field_range_poly.contains(path_poly)`` (orchard_geometry_environment.py:374-378):
    every pose rectangle inside the closed, possibly non-convex field polygon:
    4 corners inside-or-on, and no polygon edge meets the rectangle's open
    interior.
  * ``guided_lane.contains(path_poly)`` (reference_line_heuristic.py:65-67,81,105-108):
    every pose rectangle inside the UNION of the per-segment capsule polygons
    (``LineString.buffer(6, cap_style=round, join_style=bevel)``, shapely-2 default
    ``quad_segs=16`` -> 66-vertex convex polygons, restated from GEOS
    ``OffsetSegmentGenerator``): corners covered, rectangle edges covered by the
    union of clip intervals, and no boundary vertex of the union (pairwise
    boundary crossing not interior to a third capsule) strictly inside the
    rectangle.

Operation order inside each predicate is part of the contract: the device's
float64 path evaluates the same expressions without FMA contraction.
"""
import math

import numpy as np

QUAD_SEGS = 16                       # shapely >= 2.0 default for buffer()
LANE_RADIUS = 6.0                    # reference_line_heuristic.py:66
# radius of the circle inscribed in the polygonal round cap: every point closer
# than this to the segment is inside the capsule POLYGON; every point farther
# than the radius is outside.  Only the sliver in between needs the half-planes.
LANE_INSCRIBED = LANE_RADIUS * math.cos(math.pi / (4 * QUAD_SEGS))
_BAND_IN = LANE_INSCRIBED - 1e-6
_BAND_OUT = LANE_RADIUS + 1e-6


# ----------------------------------------------------------------- footprints
def body_extent(axle_to_back, axle_to_front, width):
    """Local-frame extent [x0, x1, y0, y1] of the body rectangle, car_model.py:102-119."""
    return (-axle_to_back, axle_to_front, -(width / 2), width / 2)


def aux_extent(feature):
    """Local extent of one implement rectangle ``[[x_lt, y_lt], height, width]``,
    car_model.py:146-160: corners (x,y),(x+w,y),(x+w,y-h),(x,y-h)."""
    x, y = feature[0][0], feature[0][1]
    h, w = feature[1], feature[2]
    return (x, x + w, y - h, y)


def rect_corners(poses, ext):
    """World corners of the local rectangle ``ext`` at every pose: (P,4,2).
    Vertex order (x0,y1),(x0,y0),(x1,y0),(x1,y1) (car_model.py:102-119);
    X = (c*lx - s*ly) + x, Y = (s*lx + c*ly) + y with c,s = libm cos/sin(yaw)
    (car_model.py:46-50)."""
    poses = np.asarray(poses, dtype=np.float64).reshape(-1, 3)
    x0, x1, y0, y1 = ext
    lx = np.array([x0, x0, x1, x1])
    ly = np.array([y1, y0, y0, y1])
    c = np.cos(poses[:, 2])[:, None]
    s = np.sin(poses[:, 2])[:, None]
    X = (c * lx - s * ly) + poses[:, 0:1]
    Y = (s * lx + c * ly) + poses[:, 1:2]
    return np.stack([X, Y], axis=-1)


def to_local(poses, pts):
    """Points (M,2) in every pose frame: u = c*dx + s*dy, w = c*dy - s*dx. -> (P,M),(P,M)"""
    poses = np.asarray(poses, dtype=np.float64).reshape(-1, 3)
    c = np.cos(poses[:, 2])[:, None]
    s = np.sin(poses[:, 2])[:, None]
    dx = pts[None, :, 0] - poses[:, 0:1]
    dy = pts[None, :, 1] - poses[:, 1:2]
    return c * dx + s * dy, c * dy - s * dx


# ------------------------------------------------------------ polygon helpers
def ccw(poly):
    """Return the ring (N,2) without closing point, counter-clockwise."""
    p = np.asarray(poly, dtype=np.float64)
    if len(p) > 1 and p[0, 0] == p[-1, 0] and p[0, 1] == p[-1, 1]:
        p = p[:-1]
    area2 = np.sum(p[:, 0] * np.roll(p[:, 1], -1) - np.roll(p[:, 0], -1) * p[:, 1])
    return p[::-1].copy() if area2 < 0 else p.copy()


def halfplanes(poly_ccw):
    """Outward normals (unnormalised) of a CCW convex ring: n_i = (e_y, -e_x)."""
    e = np.roll(poly_ccw, -1, axis=0) - poly_ccw
    return np.stack([e[:, 1], -e[:, 0]], axis=1)


def line_flat_buffer(p0, p1, dist):
    """``LineString([p0,p1]).buffer(dist, cap_style=flat)`` -> the 4-vertex rectangle
    (orchard_geometry_environment.py:277-286), GEOS ``computeOffsetSegment``:
    u = dist*d/len; left offset = (x - uy, y + ux)."""
    dx, dy = p1[0] - p0[0], p1[1] - p0[1]
    ln = math.sqrt(dx * dx + dy * dy)
    ux, uy = dist * dx / ln, dist * dy / ln
    return ccw(np.array([[p0[0] - uy, p0[1] + ux], [p1[0] - uy, p1[1] + ux],
                         [p1[0] + uy, p1[1] - ux], [p0[0] + uy, p0[1] - ux]]))


def point_square_buffer(x, y, r):
    """``Point(x,y).buffer(r, cap_style='square')`` (orchard_geometry_environment.py:350):
    axis-aligned square of half-side r."""
    return ccw(np.array([[x + r, y + r], [x + r, y - r], [x - r, y - r], [x - r, y + r]]))


def capsule_polygon(p0, p1, r=LANE_RADIUS, quad_segs=QUAD_SEGS):
    """``LineString([p0,p1]).buffer(r, cap_style=round, join_style=bevel)`` restated
    from GEOS ``OffsetCurveBuilder::computeLineBufferCurve`` /
    ``OffsetSegmentGenerator::addLineEndCap`` / ``addDirectedFillet``: left offset
    segment, 31 interior fillet points clockwise round p1 at ``angle + pi/2 -
    k*pi/(2*quad_segs)``, right offset segment, 31 fillet points round p0.
    66 vertices, all on or inside the true capsule.  Returned CCW."""
    dx, dy = p1[0] - p0[0], p1[1] - p0[1]
    ln = math.sqrt(dx * dx + dy * dy)
    ux, uy = r * dx / ln, r * dy / ln
    quantum = math.pi / 2.0 / quad_segs
    pts = [(p0[0] - uy, p0[1] + ux), (p1[0] - uy, p1[1] + ux)]

    def fillet(p, ang):
        start, end = ang + math.pi / 2, ang - math.pi / 2
        total = abs(start - end)
        nseg = int(total / quantum + 0.5)
        inc = total / nseg
        for i in range(1, nseg):
            a = start - i * inc
            pts.append((p[0] + r * math.cos(a), p[1] + r * math.sin(a)))

    fillet(p1, math.atan2(dy, dx))
    pts += [(p1[0] + uy, p1[1] - ux), (p0[0] + uy, p0[1] - ux)]
    fillet(p0, math.atan2(-dy, -dx))
    return ccw(np.array(pts))


# ----------------------------------------------------- obstacle SAT (closed)
def rects_hit_convex(poses, ext, poly_ccw, corners=None):
    """(P,) bool: pose rectangle meets the closed convex polygon.  Separated iff
    some edge normal of either shape STRICTLY separates (touching = collision,
    GEOS ``intersects``)."""
    if corners is None:
        corners = rect_corners(poses, ext)
    n = halfplanes(poly_ccw)                                  # (m,2)
    # obstacle axes: d[p,i,k] = nx_i*(rx_k - vx_i) + ny_i*(ry_k - vy_i)
    ddx = corners[:, None, :, 0] - poly_ccw[None, :, None, 0]
    ddy = corners[:, None, :, 1] - poly_ccw[None, :, None, 1]
    d = n[None, :, 0, None] * ddx + n[None, :, 1, None] * ddy
    sep = (d.min(axis=2) > 0).any(axis=1)
    # rectangle axes in the pose frame
    u, w = to_local(poses, poly_ccw)
    x0, x1, y0, y1 = ext
    sep |= (u.min(axis=1) > x1) | (u.max(axis=1) < x0) | (w.min(axis=1) > y1) | (w.max(axis=1) < y0)
    return ~sep


# ------------------------------------------------ field polygon containment
def points_in_closed_polygon(px, py, poly):
    """Even-odd crossing test with explicit on-edge acceptance.  px,py any shape."""
    px = np.asarray(px, dtype=np.float64)
    py = np.asarray(py, dtype=np.float64)
    inside = np.zeros(px.shape, dtype=bool)
    onedge = np.zeros(px.shape, dtype=bool)
    n = len(poly)
    for i in range(n):
        ax, ay = poly[i]
        bx, by = poly[(i + 1) % n]
        cross = (bx - ax) * (py - ay) - (by - ay) * (px - ax)
        onedge |= ((cross == 0) & (px >= min(ax, bx)) & (px <= max(ax, bx))
                   & (py >= min(ay, by)) & (py <= max(ay, by)))
        straddle = (ay > py) != (by > py)
        if by != ay:
            xint = (bx - ax) * (py - ay) / (by - ay) + ax
            inside ^= straddle & (px < xint)
    return inside | onedge


def segments_meet_open_rect(ua, wa, ub, wb, ext):
    """Liang-Barsky with strict inequalities: does the segment (ua,wa)->(ub,wb)
    (pose-frame coordinates, arrays) meet the OPEN rectangle ``ext``?"""
    x0, x1, y0, y1 = ext
    t0 = np.zeros(ua.shape)
    t1 = np.ones(ua.shape)
    dead = np.zeros(ua.shape, dtype=bool)
    du = ub - ua
    dw = wb - wa
    for a, d, lo, hi in ((ua, du, x0, x1), (wa, dw, y0, y1)):
        with np.errstate(divide="ignore", invalid="ignore"):
            tlo = (lo - a) / d
            thi = (hi - a) / d
        zero = d == 0
        dead |= zero & ((a <= lo) | (a >= hi))
        pos = d > 0
        neg = d < 0
        t0 = np.where(pos, np.maximum(t0, tlo), t0)
        t1 = np.where(pos, np.minimum(t1, thi), t1)
        t0 = np.where(neg, np.maximum(t0, thi), t0)
        t1 = np.where(neg, np.minimum(t1, tlo), t1)
    return (~dead) & (t0 < t1)


def rects_inside_polygon(poses, ext, poly, corners=None):
    """(P,) bool: pose rectangle subset of the closed simple polygon (GEOS ``contains``)."""
    if corners is None:
        corners = rect_corners(poses, ext)
    ok = points_in_closed_polygon(corners[..., 0], corners[..., 1], poly).all(axis=1)
    u, w = to_local(poses, np.asarray(poly, dtype=np.float64))
    ub, wb = np.roll(u, -1, axis=1), np.roll(w, -1, axis=1)
    cut = segments_meet_open_rect(u, w, ub, wb, ext).any(axis=1)
    return ok & ~cut


# ------------------------------------------------------------ lane (capsules)
class Lane:
    """Union of capsule polygons (reference_line_heuristic.py:50-82)."""

    def __init__(self, waypoints):
        wp = np.asarray(waypoints, dtype=np.float64)
        self.seg_p0 = wp[:-1].copy()
        self.seg_p1 = wp[1:].copy()
        self.polys = [capsule_polygon(a, b) for a, b in zip(self.seg_p0, self.seg_p1)]
        self.normals = [halfplanes(p) for p in self.polys]
        self.critical = self._critical_points()

    # distance from points (..,) to segment i
    def _seg_dist(self, i, px, py):
        ax, ay = self.seg_p0[i]
        bx, by = self.seg_p1[i]
        ex, ey = bx - ax, by - ay
        ee = ex * ex + ey * ey
        t = ((px - ax) * ex + (py - ay) * ey) / ee
        t = np.clip(t, 0.0, 1.0)
        qx = ax + t * ex
        qy = ay + t * ey
        return np.hypot(px - qx, py - qy)

    def _halfplane_vals(self, i, px, py):
        """g[..., j] = nx_j*(px - vx_j) + ny_j*(py - vy_j)"""
        v = self.polys[i]
        n = self.normals[i]
        return n[:, 0] * (px[..., None] - v[:, 0]) + n[:, 1] * (py[..., None] - v[:, 1])

    def points_in(self, i, px, py, strict=False):
        """Points in capsule polygon i (closed, or strict interior)."""
        px = np.asarray(px, dtype=np.float64)
        py = np.asarray(py, dtype=np.float64)
        d = self._seg_dist(i, px, py)
        res = d <= _BAND_IN
        band = (~res) & (d <= _BAND_OUT)
        if band.any():
            g = self._halfplane_vals(i, px[band], py[band])
            res[band] = (g < 0).all(axis=-1) if strict else (g <= 0).all(axis=-1)
        return res

    def _critical_points(self):
        """Vertices of the union's boundary that are not vertices of a single
        capsule: proper crossings of two capsule boundaries that are not strictly
        inside a third capsule."""
        out = []
        S = len(self.polys)
        for i in range(S):
            A0 = self.polys[i]
            A1 = np.roll(A0, -1, axis=0)
            for j in range(i + 1, S):
                B0 = self.polys[j]
                B1 = np.roll(B0, -1, axis=0)
                r = (A1 - A0)[:, None, :]
                s = (B1 - B0)[None, :, :]
                qp = B0[None, :, :] - A0[:, None, :]
                rxs = r[..., 0] * s[..., 1] - r[..., 1] * s[..., 0]
                with np.errstate(divide="ignore", invalid="ignore"):
                    t = (qp[..., 0] * s[..., 1] - qp[..., 1] * s[..., 0]) / rxs
                    u = (qp[..., 0] * r[..., 1] - qp[..., 1] * r[..., 0]) / rxs
                hit = (np.abs(rxs) > 1e-12) & (t >= 0) & (t <= 1) & (u >= 0) & (u <= 1)
                ii, jj = np.nonzero(hit)
                for a, b in zip(ii, jj):
                    q = A0[a] + t[a, b] * (A1[a] - A0[a])
                    covered = False
                    for k in range(S):
                        if k in (i, j):
                            continue
                        if self.points_in(k, np.array([q[0]]), np.array([q[1]]), strict=True)[0]:
                            covered = True
                            break
                    if not covered:
                        out.append(q)
        return np.array(out, dtype=np.float64).reshape(-1, 2)

    def search_segment(self, x, y):
        """Index of the LAST capsule whose interior holds the point, else -1
        (reference_line_heuristic.py:120-129)."""
        last = -1
        for i in range(len(self.polys)):
            if self.points_in(i, np.array([x]), np.array([y]), strict=True)[0]:
                last = i
        return last

    def rects_inside(self, poses, ext, corners=None):
        """(P,) bool: pose rectangle inside the union of capsule polygons."""
        poses = np.asarray(poses, dtype=np.float64).reshape(-1, 3)
        if corners is None:
            corners = rect_corners(poses, ext)
        P = len(poses)
        S = len(self.polys)
        cin = np.zeros((S, P, 4), dtype=bool)
        for i in range(S):
            cin[i] = self.points_in(i, corners[..., 0], corners[..., 1])
        inside = cin.all(axis=2).any(axis=0)                  # step 0: one capsule holds all 4
        anyc = cin.any(axis=0).all(axis=1)                    # step 1: every corner covered
        todo = np.nonzero(anyc & ~inside)[0]
        for p in todo:
            inside[p] = self._rect_inside_slow(poses[p], ext, corners[p])
        return inside

    def _rect_inside_slow(self, pose, ext, cor):
        S = len(self.polys)
        for k in range(4):                                    # step 2: edge coverage
            a = cor[k]
            b = cor[(k + 1) % 4]
            ivs = []
            for i in range(S):
                g0 = self._halfplane_vals(i, a[0:1], a[1:2])[0]
                g1 = self._halfplane_vals(i, b[0:1], b[1:2])[0]
                lo, hi = 0.0, 1.0
                ok = True
                for x0, x1 in zip(g0, g1):
                    if x0 <= 0 and x1 <= 0:
                        continue
                    if x0 > 0 and x1 > 0:
                        ok = False
                        break
                    tc = x0 / (x0 - x1)
                    if x0 > 0:
                        lo = max(lo, tc)
                    else:
                        hi = min(hi, tc)
                if ok and lo <= hi:
                    ivs.append((lo, hi))
            ivs.sort()
            cover = 0.0
            for lo, hi in ivs:
                if lo > cover:
                    return False
                cover = max(cover, hi)
            if cover < 1.0:
                return False
        if len(self.critical):                                # step 3: union-boundary vertices
            u, w = to_local(pose[None, :], self.critical)
            x0, x1, y0, y1 = ext
            if ((u > x0) & (u < x1) & (w > y0) & (w < y1)).any():
                return False
        return True


# ------------------------------------------------- union of footprint rectangles: boundary vertices
def union_boundary_vertices(poses, ext, tol=1e-12):
    """Vertices of ``unary_union`` of the rectangles ``ext`` at ``poses`` (car_model.py:39-53): rectangle corners that
    no other rectangle covers (strictly) plus proper crossings of two rectangles' edges that no third rectangle
    covers.  GEOS is unavailable: this DEFINES the vertex set for the oracle (parity unpinned); vertices on holes of
    the union, which ``.exterior`` would skip, are included."""
    poses = np.asarray(poses, dtype=np.float64).reshape(-1, 3)
    R = len(poses)
    corners = rect_corners(poses, ext)                      # (R,4,2)
    c, s = np.cos(poses[:, 2]), np.sin(poses[:, 2])
    x0, x1, y0, y1 = ext

    def covered(pts, skip):
        """pts (M,2); skip: list of rectangle indices per point not to test -> (M,) bool"""
        dx = pts[:, None, 0] - poses[None, :, 0]
        dy = pts[:, None, 1] - poses[None, :, 1]
        u = c[None, :] * dx + s[None, :] * dy
        w = c[None, :] * dy - s[None, :] * dx
        ins = (u > x0 + tol) & (u < x1 - tol) & (w > y0 + tol) & (w < y1 - tol)
        for m, sk in enumerate(skip):
            ins[m, list(sk)] = False
        return ins.any(axis=1)

    pts = corners.reshape(-1, 2)
    keep = ~covered(pts, [(i // 4,) for i in range(4 * R)])
    out = [pts[keep]]
    cross, skips = [], []
    for i in range(R):
        A0 = corners[i]
        A1 = np.roll(A0, -1, axis=0)
        for j in range(i + 1, R):
            B0 = corners[j]
            B1 = np.roll(B0, -1, axis=0)
            r = (A1 - A0)[:, None, :]
            sv = (B1 - B0)[None, :, :]
            qp = B0[None, :, :] - A0[:, None, :]
            den = r[..., 0] * sv[..., 1] - r[..., 1] * sv[..., 0]
            with np.errstate(divide="ignore", invalid="ignore"):
                t = (qp[..., 0] * sv[..., 1] - qp[..., 1] * sv[..., 0]) / den
                uu = (qp[..., 0] * r[..., 1] - qp[..., 1] * r[..., 0]) / den
            hit = (np.abs(den) >= 1e-14) & (t >= 0) & (t <= 1) & (uu >= 0) & (uu <= 1)
            for a, b in zip(*np.nonzero(hit)):
                cross.append(A0[a] + t[a, b] * (A1[a] - A0[a]))
                skips.append((i, j))
    if cross:
        cross = np.array(cross)
        out.append(cross[~covered(cross, skips)])
    return np.concatenate(out)


def signed_distance_to_ring(pts, poly):
    """Distance of every point to the ring of ``poly``, negative unless the point is strictly inside."""
    pts = np.asarray(pts, dtype=np.float64).reshape(-1, 2)
    best = np.full(len(pts), np.inf)
    n = len(poly)
    for i in range(n):
        a, b = poly[i], poly[(i + 1) % n]
        e = b - a
        t = np.clip(((pts[:, 0] - a[0]) * e[0] + (pts[:, 1] - a[1]) * e[1]) / (e[0] * e[0] + e[1] * e[1]), 0.0, 1.0)
        best = np.minimum(best, np.hypot(pts[:, 0] - (a[0] + t * e[0]), pts[:, 1] - (a[1] + t * e[1])))
    inside = points_in_closed_polygon(pts[:, 0], pts[:, 1], poly) & (best > 0)
    return np.where(inside, best, -best)
