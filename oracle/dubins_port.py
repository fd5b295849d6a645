"""Restatement of the ``dubins`` package the reference imports (``requirements.txt:14``:
``git+https://github.com/AgRoboticsResearch/pydubins.git``, UN-PINNED and un-vendored -- a Cython wrapper of Andrew
Walker's ``dubins.c``).  TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  **parity unpinned**: neither the package
nor its source is in ``/root/reference`` and it cannot be installed here, so this file restates the PUBLISHED
algorithm of dubins.c 1.x (Walker 2008-2018; Shkel & Lumelsky 2001 classification) and is anchored on the reference's
call sites only (``hybrid_a_star_search.py:294-295``, ``utils/navigation_utils.py:206-215``):

    path = dubins.shortest_path(q0, q1, turning_radius)      # six words LSL LSR RSL RSR RLR LRL, first minimum wins
    configurations, distances = path.sample_many(step)       # t = 0, step, 2 step, ... < length

Every formula below is dubins.c's, in its operation order (float64)."""
import math

EDUBOK, EDUBNOPATH = 0, 4
LSL, LSR, RSL, RSR, RLR, LRL = range(6)
L_SEG, S_SEG, R_SEG = 0, 1, 2
DIRDATA = ((L_SEG, S_SEG, L_SEG), (L_SEG, S_SEG, R_SEG), (R_SEG, S_SEG, L_SEG), (R_SEG, S_SEG, R_SEG),
           (R_SEG, L_SEG, R_SEG), (L_SEG, R_SEG, L_SEG))
WORD_NAMES = ("LSL", "LSR", "RSL", "RSR", "RLR", "LRL")


def fmodr(x, y):
    return x - y * math.floor(x / y)


def mod2pi(theta):
    return fmodr(theta, 2 * math.pi)


class _Inter:
    pass


def intermediate_results(q0, q1, rho):
    if rho <= 0.0:
        raise ValueError("rho must be positive")
    dx, dy = q1[0] - q0[0], q1[1] - q0[1]
    D = math.sqrt(dx * dx + dy * dy)
    d = D / rho
    theta = 0.0
    if d > 0:
        theta = mod2pi(math.atan2(dy, dx))
    alpha = mod2pi(q0[2] - theta)
    beta = mod2pi(q1[2] - theta)
    r = _Inter()
    r.alpha, r.beta, r.d = alpha, beta, d
    r.sa, r.sb, r.ca, r.cb = math.sin(alpha), math.sin(beta), math.cos(alpha), math.cos(beta)
    r.c_ab = math.cos(alpha - beta)
    r.d_sq = d * d
    return r


def word(i, t):
    """(t, p, q) of word ``t`` or None (dubins_LSL .. dubins_LRL)."""
    if t == LSL:
        tmp0 = i.d + i.sa - i.sb
        p_sq = 2 + i.d_sq - (2 * i.c_ab) + (2 * i.d * (i.sa - i.sb))
        if p_sq >= 0:
            tmp1 = math.atan2((i.cb - i.ca), tmp0)
            return (mod2pi(tmp1 - i.alpha), math.sqrt(p_sq), mod2pi(i.beta - tmp1))
    elif t == RSR:
        tmp0 = i.d - i.sa + i.sb
        p_sq = 2 + i.d_sq - (2 * i.c_ab) + (2 * i.d * (i.sb - i.sa))
        if p_sq >= 0:
            tmp1 = math.atan2((i.ca - i.cb), tmp0)
            return (mod2pi(i.alpha - tmp1), math.sqrt(p_sq), mod2pi(tmp1 - i.beta))
    elif t == LSR:
        p_sq = -2 + (i.d_sq) + (2 * i.c_ab) + (2 * i.d * (i.sa + i.sb))
        if p_sq >= 0:
            p = math.sqrt(p_sq)
            tmp0 = math.atan2((-i.ca - i.cb), (i.d + i.sa + i.sb)) - math.atan2(-2.0, p)
            return (mod2pi(tmp0 - i.alpha), p, mod2pi(tmp0 - mod2pi(i.beta)))
    elif t == RSL:
        p_sq = -2 + i.d_sq + (2 * i.c_ab) - (2 * i.d * (i.sa + i.sb))
        if p_sq >= 0:
            p = math.sqrt(p_sq)
            tmp0 = math.atan2((i.ca + i.cb), (i.d - i.sa - i.sb)) - math.atan2(2.0, p)
            return (mod2pi(i.alpha - tmp0), p, mod2pi(i.beta - tmp0))
    elif t == RLR:
        tmp0 = (6. - i.d_sq + 2 * i.c_ab + 2 * i.d * (i.sa - i.sb)) / 8.
        phi = math.atan2(i.ca - i.cb, i.d - i.sa + i.sb)
        if abs(tmp0) <= 1:
            p = mod2pi((2 * math.pi) - math.acos(tmp0))
            tt = mod2pi(i.alpha - phi + mod2pi(p / 2.))
            return (tt, p, mod2pi(i.alpha - i.beta - tt + mod2pi(p)))
    elif t == LRL:
        tmp0 = (6. - i.d_sq + 2 * i.c_ab + 2 * i.d * (i.sb - i.sa)) / 8.
        phi = math.atan2(i.ca - i.cb, i.d + i.sa - i.sb)
        if abs(tmp0) <= 1:
            p = mod2pi(2 * math.pi - math.acos(tmp0))
            tt = mod2pi(-i.alpha - phi + p / 2.)
            return (tt, p, mod2pi(mod2pi(i.beta) - i.alpha - tt + mod2pi(p)))
    return None


def _segment(t, qi, kind):
    st, ct = math.sin(qi[2]), math.cos(qi[2])
    if kind == L_SEG:
        qt = [+math.sin(qi[2] + t) - st, -math.cos(qi[2] + t) + ct, t]
    elif kind == R_SEG:
        qt = [-math.sin(qi[2] - t) + st, +math.cos(qi[2] - t) - ct, -t]
    else:
        qt = [ct * t, st * t, 0.0]
    return [qt[0] + qi[0], qt[1] + qi[1], qt[2] + qi[2]]


class DubinsPath:
    def __init__(self, qi, param, rho, kind):
        self.qi, self.param, self.rho, self.type = tuple(qi), tuple(param), rho, kind

    def path_length(self):
        return (self.param[0] + self.param[1] + self.param[2]) * self.rho

    def path_type(self):
        return self.type

    def sample(self, t):
        tprime = t / self.rho
        types = DIRDATA[self.type]
        qi = [0.0, 0.0, self.qi[2]]
        p1, p2 = self.param[0], self.param[1]
        q1 = _segment(p1, qi, types[0])
        q2 = _segment(p2, q1, types[1])
        if tprime < p1:
            q = _segment(tprime, qi, types[0])
        elif tprime < (p1 + p2):
            q = _segment(tprime - p1, q1, types[1])
        else:
            q = _segment(tprime - p1 - p2, q2, types[2])
        return (q[0] * self.rho + self.qi[0], q[1] * self.rho + self.qi[1], mod2pi(q[2]))

    def sample_many(self, step_size):
        qs, ts = [], []
        x, length = 0.0, self.path_length()
        while x < length:
            qs.append(self.sample(x))
            ts.append(x)
            x += step_size
        return qs, ts


def shortest_path(q0, q1, rho):
    inter = intermediate_results(q0, q1, rho)
    best, best_cost, best_word = None, math.inf, -1
    for w in range(6):
        params = word(inter, w)
        if params is not None:
            cost = params[0] + params[1] + params[2]
            if cost < best_cost:
                best_word, best_cost, best = w, cost, params
    if best_word < 0:
        raise RuntimeError("no Dubins path")
    return DubinsPath(q0, best, rho, best_word)
