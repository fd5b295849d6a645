"""cr_sincos (csrc/hl_crmath.cuh) must return the CORRECTLY ROUNDED sin / cos: the host build of the same header
is checked against a 70-digit Taylor evaluation rounded once (Fraction -> float), on random arguments, headings,
arc lengths at the planner's step sizes, values next to multiples of pi/2 and the exact doubles math.pi & co."""
import math
import os
import struct
import subprocess
import tempfile
from decimal import Decimal, getcontext
from fractions import Fraction

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = r'''
#include <stdio.h>
#include "hl_crmath.cuh"
int main(void) {
  double x;
  while (fread(&x, sizeof x, 1, stdin) == 1) { double r[2]; cr_sincos(x, &r[0], &r[1]); fwrite(r, sizeof(double), 2, stdout); }
  return 0; }'''

getcontext().prec = 90
PI = Decimal("3.14159265358979323846264338327950288419716939937510582097494459230781640628620899862803482534211706798")


def _exact_sincos(x):
    """Correctly rounded (sin x, cos x) of the double x."""
    d = Decimal(x)                                   # exact
    k = int((d / (PI / 2)).to_integral_value())
    r = d - Decimal(k) * PI / 2
    term, s, n = r, r, 1
    while abs(term) > Decimal(10) ** -80:
        term = -term * r * r / ((2 * n) * (2 * n + 1)); s += term; n += 1
    term, c, n = Decimal(1), Decimal(1), 1
    while abs(term) > Decimal(10) ** -80:
        term = -term * r * r / ((2 * n - 1) * (2 * n)); c += term; n += 1
    q = k % 4
    sv, cv = [(s, c), (c, -s), (-s, -c), (-c, s)][q]
    return float(Fraction(sv)), float(Fraction(cv))


def test_cr_sincos_is_correctly_rounded():
    rng = np.random.default_rng(0)
    maxc = math.tan(0.55) / 1.9
    xs = np.concatenate([
        rng.uniform(-10, 10, 3000), rng.uniform(-1000, 1000, 500), rng.uniform(-1e-3, 1e-3, 200),
        np.arange(-60, 61) * 0.1 * maxc, np.arange(-60, 61) * 0.2 * maxc,
        np.array([math.pi, -math.pi, math.pi / 2, -math.pi / 2, math.pi / 4, 3 * math.pi / 4, 2 * math.pi, 1e-300, 0.0,
                  math.radians(10), math.radians(45), 0.55, 1.5707963267948966, 1.5707963267948968, 3.1415926535897936]),
        (np.arange(1, 200) * (math.pi / 2)) + rng.uniform(-1e-9, 1e-9, 199),
    ])
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.cpp")
        open(c, "w").write(SRC)
        exe = os.path.join(d, "t")
        subprocess.check_call(["g++", "-O2", "-I", os.path.join(ROOT, "headland_trajectory_planning_b200", "csrc"), c, "-o", exe])
        out = subprocess.run([exe], input=xs.astype("<f8").tobytes(), capture_output=True, check=True).stdout
    got = np.frombuffer(out, dtype="<f8").reshape(-1, 2)
    bad, libm_bad = 0, 0
    for x, (s, c) in zip(xs, got):
        es, ec = _exact_sincos(float(x))
        bad += (s != es) + (c != ec)
        libm_bad += (math.sin(x) != es) + (math.cos(x) != ec)
    assert bad == 0, f"{bad} of {2 * len(xs)} values are not correctly rounded"
    assert libm_bad <= 2 * len(xs) // 500            # glibc itself is the correctly rounded value (almost) always
