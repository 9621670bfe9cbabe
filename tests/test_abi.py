"""CPU: libdsea.so builds/loads and exports exactly the C ABI that include/dsea.h declares; the
host-callable bit maps are bit-exact against the reference tables.  No GPU compute here."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from dominantsparseeigenad_b200 import _build, _lib


@pytest.fixture(scope="module")
def lib():
    _build.build()
    return _lib.load()


def _header_functions():
    src = open(os.path.join(ROOT, "include", "dsea.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dsea_[A-Za-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = _header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in dsea.h but not exported by libdsea.so"
        assert n in _lib.PROTOTYPES, f"{n} declared in dsea.h but has no ctypes prototype"
    assert sorted(_lib.PROTOTYPES) == names


def test_library_is_sm100a_only(lib):
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out


def test_version_and_error_string(lib):
    assert lib.dsea_version() == 100
    assert lib.dsea_ctx_set_option(None, b"x", 0) != 0
    assert b"NULL" in lib.dsea_last_error()


@pytest.mark.parametrize("N", [3, 4, 10])
def test_host_bit_maps_bit_exact(lib, golden, N):
    t = golden("tfim_tables.npz")
    flips, diag = t[f"flips_N{N}"], t[f"diag_N{N}"]
    for s in range(1 << N):
        assert lib.dsea_tfim_diag(N, s) == diag[s]
        for i in range(N):
            assert lib.dsea_tfim_flip_index(N, s, i) == flips[s, i]


def test_host_bit_maps_large_N(lib):
    from oracle import dsea_oracle as orc
    rng = np.random.default_rng(0)
    for N in (20, 24, 28, 30, 36):
        s = rng.integers(0, 1 << N, size=2000, dtype=np.int64)
        want = orc.tfim_diagonal_closed_form(N, s)
        got = np.array([lib.dsea_tfim_diag(N, int(x)) for x in s])
        assert np.array_equal(want, got)
    # closed form == the reference's table construction (oracle.tfim_diagonal restates TFIM.py:39-46)
    assert np.array_equal(orc.tfim_diagonal(14), np.array([lib.dsea_tfim_diag(14, s) for s in range(1 << 14)]))


def test_col_stride(lib):
    for n in (1, 15, 16, 17, 300, 1000, 1 << 20):
        ld = lib.dsea_col_stride(n)
        assert ld >= n and ld % 16 == 0 and ld - n < 16


def test_no_cuda_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dominantsparseeigenad_b200 import runtime
    with pytest.raises(_lib.DseaError):
        runtime.context()


def test_reference_import_alias_resolves_to_the_same_modules():
    """`import DominantSparseEigenAD.symeig` must be the very module object of the implementation, so that the
    reference's module-global rebinding protocol (symeig.py:66,87) works through the alias."""
    import DominantSparseEigenAD.symeig as alias_symeig
    from DominantSparseEigenAD.CG import CG_torch, CGSubspace          # noqa: F401
    from DominantSparseEigenAD.Lanczos import Lanczos, symeigLanczos    # noqa: F401
    from DominantSparseEigenAD.eig import DominantEig                   # noqa: F401
    from DominantSparseEigenAD.symeig import DominantSymeig             # noqa: F401
    import dominantsparseeigenad_b200 as impl
    assert alias_symeig is impl.symeig
    assert hasattr(alias_symeig, "setDominantSparseSymeig")
