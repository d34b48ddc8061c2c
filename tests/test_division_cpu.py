"""The streaming kernels divide by a launch-uniform sigma with `div_rn` (csrc/update.cu): q0 = RN(a r), two FMA residual
corrections, r = RN(1/b) from the host.  This replays that recurrence in exact rational arithmetic (Fraction -> float is
correctly rounded) and checks it against the correctly rounded quotient on random and adversarial operands -- the
claim behind "bit-identical to __ddiv_rn" (the GPU test compares the kernels with torch's fp64 division)."""
import random
from fractions import Fraction

import numpy as np


def _rn(x: Fraction) -> float:
    return float(x)


def _fma(a: float, b: float, c: float) -> float:
    return _rn(Fraction(a) * Fraction(b) + Fraction(c))


def _div_rn(a: float, b: float) -> float:
    r = 1.0 / b
    q0 = a * r
    q1 = _fma(_fma(-q0, b, a), r, q0)
    return _fma(_fma(-q1, b, a), r, q1)


def test_two_step_markstein_division_is_correctly_rounded():
    rng = random.Random(1)
    divisors = [80.0, 0.002, 1.0, 3.0, 0.1, 57.58598425, 1.9999999999999998, 1.0000000000000002, float(np.nextafter(4.0, 0.0))]
    divisors += [rng.uniform(0.002, 80.0) for _ in range(40)]
    divisors += [float(np.float64(2.0 ** rng.randint(-9, 6)) * (2.0 - 2.0 ** -52)) for _ in range(6)]      # all-ones significands
    bad = 0
    for b in divisors:
        for _ in range(1500):
            kind = rng.random()
            if kind < 0.5:
                a = rng.uniform(-100.0, 100.0)
            elif kind < 0.8:                     # fp32-representable numerators scaled by a step size, as in the kernels
                a = float(np.float32(rng.gauss(0.0, 1.0))) * rng.uniform(-80.0, 80.0)
            else:                                # quotients next to a rounding boundary: a = RN(q b) for a "round" q
                q = float(np.float32(rng.uniform(-4.0, 4.0)))
                a = float(np.nextafter(q * b, rng.choice([-np.inf, np.inf])))
            want = _rn(Fraction(a) / Fraction(b))
            got = _div_rn(a, b)
            bad += got != want
    assert bad == 0
