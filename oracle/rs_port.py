"""Table-driven float64 restatement of the reference's Reeds-Shepp module,
``path_planner/utils/reeds_shepp.py`` (byte-identical copy at
``utils/reeds_shepp.py``).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  PINNED: ``oracle/gen_golden.py``
runs this port against the reference's own module (which imports in the build
container once ``matplotlib`` is stubbed) and requires bit-for-bit equality of
word sets, order, lengths, sample counts and sampled states; the vectors are
committed as ``tests/golden/rs_*.npz``.

Layout differs from the reference on purpose: the 46 candidate words are one
table (solver, mirror signs, "backwards" flag, length pattern, letters) that the
CUDA evaluator mirrors one thread per row; the arithmetic inside every solver
keeps the reference's operation order so validity flags and lengths are
bit-identical.

Reference map:
  word solvers SLS/LSL/LSR/LRL/LRLRn/LRLRp/LRSR/LRSL/LRSLR  reeds_shepp.py:90-160,257-283,322-350,425-440
  family order SCS,CSC,CCC,CCCC,CCSC,CCSCC                    reeds_shepp.py:565-582
  dedup / length cap (``set_path``)                           reeds_shepp.py:68-87
  sampler (``generate_local_course`` + ``interpolate``)       reeds_shepp.py:471-562
  world transform (``calc_all_paths``)                        reeds_shepp.py:39-65
  ``M`` / ``R`` / ``pi_2_pi``                                  reeds_shepp.py:586-617
"""
import math

PI = math.pi
MAX_LENGTH = 1000.0          # reeds_shepp.py:7
HALF_PI = 0.5 * PI


def mod2pi(theta):
    """``M`` (reeds_shepp.py:606-617): Python floored ``%`` then fold to (-pi, pi]."""
    phi = theta % (2.0 * PI)
    if phi < -PI:
        phi += 2.0 * PI
    if phi > PI:
        phi -= 2.0 * PI
    return phi


def polar(x, y):
    """``R`` (reeds_shepp.py:596-603)."""
    return math.hypot(x, y), math.atan2(y, x)


def pi_2_pi(theta):
    """reeds_shepp.py:586-593."""
    while theta > PI:
        theta -= 2.0 * PI
    while theta < -PI:
        theta += 2.0 * PI
    return theta


# ---------------------------------------------------------------- word solvers
def _sls(x, y, phi):          # reeds_shepp.py:144-160
    phi = mod2pi(phi)
    if y > 0.0 and 0.0 < phi < PI * 0.99:
        xd = -y / math.tan(phi) + x
        t = xd - math.tan(phi / 2.0)
        u = phi
        v = math.sqrt((x - xd) ** 2 + y ** 2) - math.tan(phi / 2.0)
        return True, t, u, v
    elif y < 0.0 and 0.0 < phi < PI * 0.99:
        xd = -y / math.tan(phi) + x
        t = xd - math.tan(phi / 2.0)
        u = phi
        v = -math.sqrt((x - xd) ** 2 + y ** 2) - math.tan(phi / 2.0)
        return True, t, u, v
    return False, 0.0, 0.0, 0.0


def _lsl(x, y, phi):          # reeds_shepp.py:90-98
    u, t = polar(x - math.sin(phi), y - 1.0 + math.cos(phi))
    if t >= 0.0:
        v = mod2pi(phi - t)
        if v >= 0.0:
            return True, t, u, v
    return False, 0.0, 0.0, 0.0


def _lsr(x, y, phi):          # reeds_shepp.py:101-114
    u1, t1 = polar(x + math.sin(phi), y - 1.0 - math.cos(phi))
    u1 = u1 ** 2
    if u1 >= 4.0:
        u = math.sqrt(u1 - 4.0)
        theta = math.atan2(2.0, u)
        t = mod2pi(t1 + theta)
        v = mod2pi(t - phi)
        if t >= 0.0 and v >= 0.0:
            return True, t, u, v
    return False, 0.0, 0.0, 0.0


def _lrl(x, y, phi):          # reeds_shepp.py:117-128
    u1, t1 = polar(x - math.sin(phi), y - 1.0 + math.cos(phi))
    if u1 <= 4.0:
        u = -2.0 * math.asin(0.25 * u1)
        t = mod2pi(t1 + 0.5 * u + PI)
        v = mod2pi(phi - t + u)
        if t >= 0.0 and u <= 0.0:
            return True, t, u, v
    return False, 0.0, 0.0, 0.0


def _tau_omega(u, v, xi, eta, phi):   # reeds_shepp.py:239-254
    delta = mod2pi(u - v)
    A = math.sin(u) - math.sin(delta)
    B = math.cos(u) - math.cos(delta) - 1.0
    t1 = math.atan2(eta * A - xi * B, xi * A + eta * B)
    t2 = 2.0 * (math.cos(delta) - math.cos(v) - math.cos(u)) + 3.0
    if t2 < 0:
        tau = mod2pi(t1 + PI)
    else:
        tau = mod2pi(t1)
    omega = mod2pi(tau - u + v - phi)
    return tau, omega


def _lrlrn(x, y, phi):        # reeds_shepp.py:257-268
    xi = x + math.sin(phi)
    eta = y - 1.0 - math.cos(phi)
    rho = 0.25 * (2.0 + math.sqrt(xi * xi + eta * eta))
    if rho <= 1.0:
        u = math.acos(rho)
        t, v = _tau_omega(u, -u, xi, eta, phi)
        if t >= 0.0 and v <= 0.0:
            return True, t, u, v
    return False, 0.0, 0.0, 0.0


def _lrlrp(x, y, phi):        # reeds_shepp.py:271-283
    xi = x + math.sin(phi)
    eta = y - 1.0 - math.cos(phi)
    rho = (20.0 - xi * xi - eta * eta) / 16.0
    if 0.0 <= rho <= 1.0:
        u = -math.acos(rho)
        if u >= -0.5 * PI:
            t, v = _tau_omega(u, u, xi, eta, phi)
            if t >= 0.0 and v >= 0.0:
                return True, t, u, v
    return False, 0.0, 0.0, 0.0


def _lrsr(x, y, phi):         # reeds_shepp.py:322-334
    xi = x + math.sin(phi)
    eta = y - 1.0 - math.cos(phi)
    rho, theta = polar(-eta, xi)
    if rho >= 2.0:
        t = theta
        u = 2.0 - rho
        v = mod2pi(t + 0.5 * PI - phi)
        if t >= 0.0 and u <= 0.0 and v <= 0.0:
            return True, t, u, v
    return False, 0.0, 0.0, 0.0


def _lrsl(x, y, phi):         # reeds_shepp.py:337-350
    xi = x - math.sin(phi)
    eta = y - 1.0 + math.cos(phi)
    rho, theta = polar(xi, eta)
    if rho >= 2.0:
        r = math.sqrt(rho * rho - 4.0)
        u = 2.0 - r
        t = mod2pi(theta + math.atan2(r, -2.0))
        v = mod2pi(phi - 0.5 * PI - t)
        if t >= 0.0 and u <= 0.0 and v <= 0.0:
            return True, t, u, v
    return False, 0.0, 0.0, 0.0


def _lrslr(x, y, phi):        # reeds_shepp.py:425-440
    xi = x + math.sin(phi)
    eta = y - 1.0 - math.cos(phi)
    rho, theta = polar(xi, eta)
    if rho >= 2.0:
        u = 4.0 - math.sqrt(rho * rho - 4.0)
        if u <= 0.0:
            t = mod2pi(math.atan2((4.0 - u) * xi - 2.0 * eta, -2.0 * xi + (u - 4.0) * eta))
            v = mod2pi(t - phi)
            if t >= 0.0 and v >= 0.0:
                return True, t, u, v
    return False, 0.0, 0.0, 0.0


SOLVERS = (_sls, _lsl, _lsr, _lrl, _lrlrn, _lrlrp, _lrsr, _lrsl, _lrslr)
S_SLS, S_LSL, S_LSR, S_LRL, S_LRLRN, S_LRLRP, S_LRSR, S_LRSL, S_LRSLR = range(9)

# length patterns: how (t, u, v) map to the signed segment lengths before the
# mirror negation.  'H' is the fixed -pi/2 arc of the CCSC / CCSCC families.
P_TUV, P_VUT, P_TUnUV, P_TUUV, P_THUV, P_VUHT, P_THUHV = range(7)


def _pattern(p, t, u, v):
    if p == P_TUV:
        return [t, u, v]
    if p == P_VUT:
        return [v, u, t]
    if p == P_TUnUV:
        return [t, u, -u, v]
    if p == P_TUUV:
        return [t, u, u, v]
    if p == P_THUV:
        return [t, -0.5 * PI, u, v]
    if p == P_VUHT:
        return [v, u, -0.5 * PI, t]
    return [t, -0.5 * PI, u, -0.5 * PI, v]


def _build_table():
    """46 rows in the reference's evaluation order (reeds_shepp.py:131-141,
    163-236, 286-319, 353-422, 443-468).  Row = (solver, sx, sy, backwards,
    pattern, letters).  Mirror variants always come in the order
    (+,+), (-,+), (+,-), (-,-); the phi sign is sx*sy, lengths are negated when
    sx < 0 and the letters are L<->R swapped when sy < 0."""
    rows = []

    def quad(solver, backwards, pattern, letters_pos, letters_neg):
        for sx, sy in ((1, 1), (-1, 1), (1, -1), (-1, -1)):
            rows.append((solver, sx, sy, backwards, pattern,
                         letters_pos if sy > 0 else letters_neg))

    rows.append((S_SLS, 1, 1, False, P_TUV, "SLS"))
    rows.append((S_SLS, 1, -1, False, P_TUV, "SRS"))
    quad(S_LSL, False, P_TUV, "LSL", "RSR")
    quad(S_LSR, False, P_TUV, "LSR", "RSL")
    quad(S_LRL, False, P_TUV, "LRL", "RLR")
    quad(S_LRL, True, P_VUT, "LRL", "RLR")
    quad(S_LRLRN, False, P_TUnUV, "LRLR", "RLRL")
    quad(S_LRLRP, False, P_TUUV, "LRLR", "RLRL")
    quad(S_LRSL, False, P_THUV, "LRSL", "RLSR")
    quad(S_LRSR, False, P_THUV, "LRSR", "RLSL")
    quad(S_LRSL, True, P_VUHT, "LSRL", "RSLR")
    quad(S_LRSR, True, P_VUHT, "RSRL", "LSLR")
    quad(S_LRSLR, False, P_THUHV, "LRSLR", "RLSRL")
    assert len(rows) == 46
    return tuple(rows)


WORD_TABLE = _build_table()


class RSPath:
    """Mirror of ``PATH`` (reeds_shepp.py:12-23)."""
    __slots__ = ("lengths", "ctypes", "L", "x", "y", "yaw", "cs", "directions", "cand")

    def __init__(self, lengths, ctypes, L, cand):
        self.lengths = lengths
        self.ctypes = ctypes
        self.L = L
        self.cand = cand          # row of WORD_TABLE that produced it (port-only)
        self.x = []
        self.y = []
        self.yaw = []
        self.cs = []
        self.directions = []


def eval_candidates(x, y, phi):
    """All 46 rows -> list of (flag, lengths or None).  No dedup."""
    out = []
    xb = x * math.cos(phi) + y * math.sin(phi)       # reeds_shepp.py:217-218, 387-388
    yb = x * math.sin(phi) - y * math.cos(phi)
    for solver, sx, sy, backwards, pattern, letters in WORD_TABLE:
        ax, ay = (xb, yb) if backwards else (x, y)
        ax = ax if sx > 0 else -ax
        ay = ay if sy > 0 else -ay
        aphi = phi if sx * sy > 0 else -phi
        flag, t, u, v = SOLVERS[solver](ax, ay, aphi)
        if not flag:
            out.append((False, None))
            continue
        lens = _pattern(pattern, t, u, v)
        if sx < 0:
            lens = [-l for l in lens]
        out.append((True, lens))
    return out


def generate_path(q0, q1, maxc):
    """reeds_shepp.py:565-582 + ``set_path`` (:68-87): normalise, evaluate the 46
    candidates in order, signed-sum dedup against earlier accepted paths with
    identical letters, drop L >= 1000, assert L >= 0.01."""
    dx = q1[0] - q0[0]
    dy = q1[1] - q0[1]
    dth = q1[2] - q0[2]
    c = math.cos(q0[2])
    s = math.sin(q0[2])
    x = (c * dx + s * dy) * maxc
    y = (-s * dx + c * dy) * maxc

    paths = []
    for cand, (flag, lens) in enumerate(eval_candidates(x, y, dth)):
        if not flag:
            continue
        letters = WORD_TABLE[cand][5]
        dup = False
        for pe in paths:
            if pe.ctypes == list(letters):
                if sum([a - b for a, b in zip(pe.lengths, lens)]) <= 0.01:
                    dup = True
                    break
        if dup:
            continue
        L = sum([abs(i) for i in lens])
        if L >= MAX_LENGTH:
            continue
        assert L >= 0.01
        paths.append(RSPath(lens, list(letters), L, cand))
    return paths


def _interp(ind, l, m, maxc, ox, oy, oyaw, px, py, pyaw, cs, directions):
    """``interpolate`` (reeds_shepp.py:533-562)."""
    if m == "S":
        px[ind] = ox + l / maxc * math.cos(oyaw)
        py[ind] = oy + l / maxc * math.sin(oyaw)
        pyaw[ind] = oyaw
        cs[ind] = 0
    else:
        ldx = math.sin(l) / maxc
        if m == "L":
            ldy = (1.0 - math.cos(l)) / maxc
            cs[ind] = maxc
        else:
            ldy = (1.0 - math.cos(l)) / (-maxc)
            cs[ind] = -maxc
        gdx = math.cos(-oyaw) * ldx + math.sin(-oyaw) * ldy
        gdy = -math.sin(-oyaw) * ldx + math.cos(-oyaw) * ldy
        px[ind] = ox + gdx
        py[ind] = oy + gdy
        pyaw[ind] = oyaw + l if m == "L" else oyaw - l
    directions[ind] = 1 if l > 0.0 else -1


def generate_local_course(L, lengths, mode, maxc, step_size):
    """reeds_shepp.py:471-530.  Quirks kept: the first sample of a segment
    overwrites the previous segment's end point (cusps are not emitted), the
    carry-over ``pd = +-d - ll`` can start a segment with an offset of the wrong
    sign (sample tagged with the wrong direction), trailing entries are popped
    while ``px[-1] == 0.0``."""
    point_num = int(L / step_size) + len(lengths) + 3
    px = [0.0] * point_num
    py = [0.0] * point_num
    pyaw = [0.0] * point_num
    directions = [0] * point_num
    cs = [0] * point_num
    ind = 1
    directions[0] = 1 if lengths[0] > 0.0 else -1
    ll = 0.0
    for i, (m, l) in enumerate(zip(mode, lengths)):
        d = step_size if l > 0.0 else -step_size
        ox, oy, oyaw = px[ind], py[ind], pyaw[ind]
        ind -= 1
        if i >= 1 and (lengths[i - 1] * lengths[i]) > 0:
            pd = -d - ll
        else:
            pd = d - ll
        while abs(pd) <= abs(l):
            ind += 1
            _interp(ind, pd, m, maxc, ox, oy, oyaw, px, py, pyaw, cs, directions)
            pd += d
        ll = l - pd - d
        ind += 1
        _interp(ind, l, m, maxc, ox, oy, oyaw, px, py, pyaw, cs, directions)
    while px[-1] == 0.0:
        px.pop()
        py.pop()
        pyaw.pop()
        directions.pop()
        cs.pop()
    return px, py, pyaw, cs, directions


def calc_all_paths(sx, sy, syaw, gx, gy, gyaw, maxc, step_size=0.2):
    """reeds_shepp.py:39-65."""
    q0 = [sx, sy, syaw]
    q1 = [gx, gy, gyaw]
    paths = generate_path(q0, q1, maxc)
    for path in paths:
        x, y, yaw, cs, directions = generate_local_course(
            path.L, path.lengths, path.ctypes, maxc, step_size * maxc)
        path.x = [math.cos(-q0[2]) * ix + math.sin(-q0[2]) * iy + q0[0] for (ix, iy) in zip(x, y)]
        path.y = [-math.sin(-q0[2]) * ix + math.cos(-q0[2]) * iy + q0[1] for (ix, iy) in zip(x, y)]
        path.yaw = [pi_2_pi(iyaw + q0[2]) for iyaw in yaw]
        path.directions = directions
        path.cs = cs
        path.lengths = [l / maxc for l in path.lengths]
        path.L = path.L / maxc
    return paths
