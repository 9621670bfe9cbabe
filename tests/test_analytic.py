"""CPU: the closed forms the product path asserts against (`dominantsparseeigenad_b200.analytic`) agree with the
oracle's independent implementation, with BASELINE.md section 3 and with dense exact diagonalisation at small N."""
import numpy as np
import pytest
import torch

from dominantsparseeigenad_b200.analytic import tfim_exact
from oracle import dsea_oracle as orc

BASELINE_TABLE = [   # N, g, E0/N, (dE0/dg)/N, (d2E0/dg2)/N, chiF   
    (20, 1.005050539970398, -1.277782264486651, -0.642889422229133, -1.106829271033164, 11.735098717835),
    (24, 1.0, -1.274149024889876, -0.637074512444938, -1.174367973012996, 17.25),
    (24, 1.5, -1.671926552941317, -0.877323292575251, -0.183237299245418, 0.534180712080),
    (28, 1.0, -1.273907645616076, -0.636953822808038, -1.223485960519997, 23.625),
    (30, 1.5, -1.671926242197181, -0.877327825904887, -0.183168835493561, 0.666785743369),
]


@pytest.mark.parametrize("row", BASELINE_TABLE)
def test_closed_forms_match_baseline_table(row):
    N, g, e, de, d2e, chi = row
    ex = tfim_exact(N, g)
    assert abs(ex.E0 / N - e) < 1e-13 and abs(ex.dE0 / N - de) < 1e-13 and abs(ex.d2E0 / N - d2e) < 1e-12
    assert abs(ex.chiF - chi) < 1e-9 * max(1.0, chi)


@pytest.mark.parametrize("N", [5, 8, 9, 12])
def test_closed_forms_match_oracle_and_dense_diagonalisation(N):
    for g in (0.7, 1.0, 1.3):
        ex = tfim_exact(N, g)
        o = orc.tfim_analytic(N, g)
        assert np.allclose([ex.E0, ex.dE0, ex.d2E0, ex.chiF], [float(x) for x in o], rtol=1e-12, atol=1e-12)
        if N <= 9:
            H = orc.TFIMOracle(N, g).dense()
            assert abs(torch.linalg.eigvalsh(H)[0].item() - ex.E0) < 1e-10 * N
