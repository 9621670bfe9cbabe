"""Closed-form ground-state data of the periodic transverse-field Ising chain at finite N.

Callers' side of the path: the reference's drivers compare their AD results with these
Jordan-Wigner expressions (examples/TFIM/E0.py:9-23, examples/TFIM_vumps/analytic.py:8-12).  They are the
parity targets at sizes no CPU run can reach (N = 24 ... 30) and what `bench.py` and
the multi-GPU self-check assert against — the product path never imports `oracle/`.

Momenta are the Neveu-Schwarz set k_m = (2m + 1) pi / N, m = 0..N-1, which for even N is the grid of
E0.py:15-18; E0 = -1/2 sum_m eps_m with eps_m = 2 sqrt(g^2 - 2 g cos k_m + 1) is exact for the even-parity
ground state at any finite N.  chi_F = 1/4 sum_{k_m > 0} sin^2 k_m / (1 + g^2 - 2 g cos k_m)^2.
"""
from __future__ import annotations

import math
from typing import NamedTuple


class TFIMExact(NamedTuple):
    E0: float
    dE0: float
    d2E0: float
    chiF: float


def tfim_exact(N: int, g: float) -> TFIMExact:
    E0 = dE0 = d2E0 = chiF = 0.0
    for m in range(N):
        k = (2 * m + 1) * math.pi / N
        c, s = math.cos(k), math.sin(k)
        w2 = g * g - 2.0 * g * c + 1.0
        eps = 2.0 * math.sqrt(w2)
        E0 -= 0.5 * eps
        dE0 -= 0.5 * 4.0 * (g - c) / eps                    # E0.py:19
        d2E0 -= 0.5 * 16.0 * s * s / eps ** 3               # E0.py:20
        if 0.0 < k < math.pi:
            chiF += 0.25 * s * s / (w2 * w2)
    return TFIMExact(E0, dE0, d2E0, chiF)
