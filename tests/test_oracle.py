"""CPU: the oracle restatement against the committed golden vectors (reference outputs) and the
upstream result files.  The oracle itself was pinned to the live reference by oracle/gen_golden.py."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import dsea_oracle as orc
from conftest import f32round


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("N", [3, 4, 10, 12, 16])
def test_integer_tables_bit_exact(golden, N):
    t = golden("tfim_tables.npz")
    flips, diag = orc.tfim_flip_table(N), orc.tfim_diagonal(N)
    assert _sha(flips) == str(t[f"flips_sha256_N{N}"])
    assert _sha(diag) == str(t[f"diag_sha256_N{N}"])
    if N <= 10:
        assert np.array_equal(flips, t[f"flips_N{N}"]) and flips.dtype == np.int64
        assert np.array_equal(diag, t[f"diag_N{N}"])
    assert np.array_equal(diag, orc.tfim_diagonal_closed_form(N, np.arange(1 << N)))


@pytest.mark.parametrize("N", [4, 10, 12])
def test_matvec_kat(golden, N):
    k = golden("tfim_matvec_kat.npz")
    m = orc.TFIMOracle(N, float(k[f"g_N{N}"]))
    v, w = torch.from_numpy(k[f"v_N{N}"]), torch.from_numpy(k[f"w_N{N}"])
    assert np.array_equal(m.H(v).numpy(), k[f"Hv_N{N}"])
    assert np.array_equal(m.pHpg(v).numpy(), k[f"pHpg_v_N{N}"])
    assert np.array_equal(m.Hadjoint_to_gadjoint(w, v).numpy(), k[f"adj_N{N}"])


def test_lanczos_and_cg_kat(golden):
    k = golden("lanczos_cg_kat.npz")
    A = torch.from_numpy(k["A"])
    q0 = torch.from_numpy(k["q0"])
    lo, vlo, hi, vhi = orc.extreme_eigpair(lambda v: A.matmul(v), A.shape[0], int(k["k"]),
                                           orc.ListDraws([q0, q0]), "both")
    assert abs(lo.item() - float(k["eval_min"])) < 1e-12 and abs(hi.item() - float(k["eval_max"])) < 1e-12
    assert np.allclose(vlo.numpy(), k["evec_min"], atol=1e-10) and np.allclose(vhi.numpy(), k["evec_max"], atol=1e-10)
    B = torch.from_numpy(k["cg_A"])
    x = orc.cg_solve(lambda v: B.matmul(v), torch.from_numpy(k["cg_b"]), torch.from_numpy(k["cg_x0"]))
    assert np.allclose(x.numpy(), k["cg_x"], atol=1e-12, rtol=0)


@pytest.mark.parametrize("tag", ["N10_g1", "N10_g3", "N12_g1"])
def test_tfim_derivatives_golden(golden, tag):
    d = golden("tfim_derivatives.npz")
    N = int(tag[1:3])
    g, k, seed = float(d[tag + "_g"]), int(d[tag + "_k"]), int(d[tag + "_seed"])
    ref = d[tag + "_ref"]
    m = orc.TFIMOracle(N, torch.tensor([g], dtype=torch.float64, requires_grad=True))
    E0, dE0, d2E0, _ = orc.tfim_energy_derivatives(m, k, orc.SeededDraws(seed))
    assert abs(E0 - ref[0]) < 1e-12 * abs(ref[0])
    assert abs(dE0 - ref[1]) < 1e-9 * abs(ref[1])
    assert abs(d2E0 - ref[2]) < 1e-7 * abs(ref[2])
    m = orc.TFIMOracle(N, torch.tensor([g], dtype=torch.float64, requires_grad=True))
    _, _, chif = orc.tfim_fidelity_susceptibility(m, k, orc.SeededDraws(seed + 1))
    assert abs(chif - ref[3]) < 1e-6 * abs(ref[3])
    an = d[tag + "_analytic"]
    assert abs(E0 - an[0]) < 1e-12 * abs(an[0]) and abs(dE0 - an[1]) < 1e-6 * abs(an[1])


@pytest.mark.parametrize("N", [10, 16, 20])
def test_analytic_matches_upstream_result_files(golden, N):
    """examples/TFIM/datas/*.npz hold per-site values on g = float64(float32(linspace(.5,1.5,100)))."""
    up = golden("upstream_tfim.npz")
    gs = up[f"gs_N{N}"]
    for i in (0, 25, 50, 75, 99):
        g = f32round(gs[i])
        E0, dE0, d2E0, chif = orc.tfim_analytic(N, g)
        assert abs(E0 / N - up[f"E0s_N{N}"][i]) < 1e-12 * abs(E0 / N)
        assert abs(dE0 / N - up[f"dE0s_N{N}"][i]) < 1e-6 * abs(dE0 / N)
        if g >= 1.0:      # g<1: near-degenerate ground state, reference itself scatters
            assert abs(d2E0 / N - up[f"d2E0s_N{N}"][i]) < 1e-5 * abs(d2E0 / N)
            assert abs(chif - up[f"chiFs_N{N}"][i]) < 1e-5 * abs(chif)


def test_oracle_reproduces_upstream_N10_point(golden):
    up = golden("upstream_tfim.npz")
    i, N, k = 60, 10, 300
    g = f32round(up["gs_N10"][i])
    m = orc.TFIMOracle(N, torch.tensor([g], dtype=torch.float64, requires_grad=True))
    E0, dE0, d2E0, _ = orc.tfim_energy_derivatives(m, k, orc.SeededDraws(3))
    assert abs(E0 / N - up["E0s_N10"][i]) < 1e-12 * abs(E0 / N)
    assert abs(dE0 / N - up["dE0s_N10"][i]) < 1e-6 * abs(dE0 / N)
    assert abs(d2E0 / N - up["d2E0s_N10"][i]) < 1e-6 * abs(d2E0 / N)


def test_schrodinger_config1(golden):
    s = golden("schrodinger1d.npz")
    for variant in ("matrixAD", "sparseAD"):
        mdl = orc.Schrodinger1DOracle(300)
        loss = (mdl.loss_dense if variant == "matrixAD" else mdl.loss_sparse)(300, orc.SeededDraws(31337))
        grad, = torch.autograd.grad(loss, mdl.potential)
        assert abs(loss.item() - float(s[variant + "_loss"])) < 1e-10
        assert np.allclose(grad.numpy(), s[variant + "_grad"], atol=1e-9, rtol=1e-6)
    assert abs(float(s["matrixAD_loss"]) - 0.099454537767) < 1e-11      # SURVEY section 6 KAT
