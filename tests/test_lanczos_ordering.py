"""CPU (numpy): why the Lanczos step keeps the reference's order of operations.

libdsea forms r0 = u - alpha q_i - beta q_{i-1} BEFORE its single Gram-Schmidt sweep (Lanczos.py:61-66).
Sweeping u itself (r = u - Q Q^T u) is algebraically the same and 24 n bytes cheaper per step, but its
orthogonality defect obeys the bare three-term recurrence and explodes once beta gets small.  This test
pins that decision with the case that exposed it: TFIM N=8 with k = dim = 256 (Krylov space exhausted),
the shape of the reference's own examples (N=10, k=300 > number of distinct eigenvalues)."""
import numpy as np

from oracle import dsea_oracle as orc


def _lanczos(H, k, rng, recurrence_first):
    n = H.shape[0]
    Q = np.zeros((n, k))
    a, b = np.zeros(k), np.zeros(k)
    q = rng.standard_normal(n)
    Q[:, 0] = q / np.linalg.norm(q)
    for i in range(k):
        u = H @ Q[:, i]
        a[i] = Q[:, i] @ u
        if i == k - 1:
            break
        if recurrence_first:
            r = u - a[i] * Q[:, i] - (b[i - 1] * Q[:, i - 1] if i else 0.0)
            r = r - Q[:, :i + 1] @ (Q[:, :i + 1].T @ r)
        else:
            r = u - Q[:, :i + 1] @ (Q[:, :i + 1].T @ u)
        b[i] = np.linalg.norm(r)
        Q[:, i + 1] = r / b[i]
    T = np.diag(a) + np.diag(b[:k - 1], 1) + np.diag(b[:k - 1], -1)
    return np.linalg.eigvalsh(T)[0]


def test_recurrence_before_sweep_is_what_keeps_k_equal_dim_stable():
    N, g = 8, 0.5
    H = orc.TFIMOracle(N, g).dense().numpy()
    exact = np.linalg.eigvalsh(H)[0]
    good = _lanczos(H, 256, np.random.default_rng(0), recurrence_first=True)
    bad = _lanczos(H, 256, np.random.default_rng(0), recurrence_first=False)
    assert abs(good - exact) < 1e-10 * abs(exact)
    assert abs(bad - exact) > 1e-3 * abs(exact)          # the cheaper variant returns a Ritz value outside the spectrum
    # and the oracle (the reference's algorithm) agrees with the stable ordering
    E, _ = orc.extreme_eigpair(orc.TFIMOracle(N, g).H, 256, 256, orc.SeededDraws(1), "min")
    assert abs(E.item() - exact) < 1e-10 * abs(exact)
