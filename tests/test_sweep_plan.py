"""CPU: the TFIM sweep schedule and tile addressing (the index logic of tfim.cu), emulated on the host.

For every (local bits, tile bits, run bits) the plan must (1) make every local spin bit the responsibility of
exactly one sweep, and (2) within each sweep, map (tile, element) -> global index bijectively, with
in-tile flips of the handled bits equal to global flips of the corresponding spin bit.  A numpy emulation of
the sweeps driven by the library's own plan must reproduce the oracle's H v bit-for-bit in structure
(to rounding in value)."""
import ctypes as C

import numpy as np
import pytest
import torch

from dominantsparseeigenad_b200 import _build, _lib


@pytest.fixture(scope="module")
def lib():
    _build.build()
    return _lib.load()


def plan(lib, L, T, run=0, direct=1):
    buf = (C.c_int * 200)()
    n = lib.dsea_tfim_plan(L, T, run, direct, buf)
    assert n > 0, (L, T, run)
    return [tuple(buf[5 * j:5 * j + 5]) for j in range(n)]


def tile_to_global(L, sw):
    """Global indices of every (tile, element) of one sweep, as an array [ntiles, 2^T].  Mirrors tile_base() in
    tfim.cu: a sweep with direct bits orders its tiles with the UPPER address bits fastest."""
    T, c, hshift, b0, nd = sw
    ntiles = 1 << (L - T)
    t = np.arange(ntiles, dtype=np.int64)[:, None]
    e = np.arange(1 << T, dtype=np.int64)[None, :]
    mid = hshift - c
    up = L - (hshift + T - c)
    if nd > 0:
        assert nd == up
        t_up, t_mid = t & ((1 << up) - 1), t >> up
    else:
        t_mid, t_up = t & ((1 << mid) - 1), t >> mid
    base = (t_mid << c) | (t_up << (hshift + T - c))
    return base | (e & ((1 << c) - 1)) | ((e >> c) << hshift)


@pytest.mark.parametrize("T", [3, 5, 8, 12, 13, 14])
def test_every_bit_handled_once_and_tiles_are_bijective(lib, T):
    for L in range(1, 31):
        for run, direct in ((0, 1), (0, 0), (2, 1), (4, 0)):
            sweeps = plan(lib, L, T, run, direct)
            handled = []
            for (Ts, c, hshift, b0, nd) in sweeps:
                assert 1 <= Ts <= max(T, 1) or L < T
                assert 0 < c <= Ts and b0 in (0, c)
                assert 0 <= nd <= 4 and (nd == 0 or (direct and run == 0))
                for b in range(b0, Ts):
                    handled.append(b if b < c else hshift + (b - c))
                handled += [hshift + (Ts - c) + d for d in range(nd)]          # direct bits sit right above the tile
            assert sorted(handled) == list(range(L)), (L, T, run, sweeps)
            if run == 0 and T >= 10:
                assert all(c >= 4 for (Ts, c, _, _, _) in sweeps if Ts == T), (L, T, sweeps)   # >= 128-byte runs
            if L <= 16:
                for sw in sweeps:
                    g = tile_to_global(L, sw)
                    assert np.array_equal(np.sort(g.ravel()), np.arange(1 << L)), (L, T, run, sw)
                    Ts, c, hshift, b0, nd = sw
                    e = np.arange(1 << Ts)
                    for b in range(b0, Ts):
                        spin_bit = b if b < c else hshift + (b - c)
                        assert np.array_equal(g[:, e ^ (1 << b)], g ^ (1 << spin_bit))


def test_production_plans():
    """The plans the benchmark sizes get with 13-bit tiles: two sweeps up to 26 local bits (the last one with
    128-byte runs and <= 4 direct bits), three sweeps beyond; partner tiles of the direct bits are adjacent."""
    lib = _lib.load()
    assert plan(lib, 20, 13) == [(13, 13, 13, 0, 0), (13, 6, 13, 6, 0)]
    assert plan(lib, 22, 13) == [(13, 13, 13, 0, 0), (13, 4, 13, 4, 0)]
    assert plan(lib, 24, 13) == [(13, 13, 13, 0, 0), (13, 4, 13, 4, 2)]
    assert plan(lib, 26, 13) == [(13, 13, 13, 0, 0), (13, 4, 13, 4, 4)]
    assert plan(lib, 27, 13) == [(13, 13, 13, 0, 0), (13, 6, 13, 6, 0), (13, 6, 20, 6, 0)]
    assert [s[4] for s in plan(lib, 28, 13)] == [0, 0, 0] and len(plan(lib, 24, 13, direct=0)) == 3
    assert plan(lib, 23, 13) == [(13, 13, 13, 0, 0), (13, 4, 13, 4, 1)]
    g = tile_to_global(23, (13, 4, 13, 4, 1))[:, 0]            # element 0 of consecutive tiles
    assert g[1] == g[0] ^ (1 << 22) and g[2] == g[0] ^ (1 << 4)


@pytest.mark.parametrize("N,T,run", [(10, 13, 0), (12, 5, 0), (14, 6, 2), (15, 7, 3), (16, 10, 0), (17, 11, 0)])
def test_emulated_sweeps_reproduce_oracle_matvec(lib, N, T, run):
    from oracle import dsea_oracle as orc
    g = 1.3
    rng = np.random.default_rng(N)
    v = rng.standard_normal(1 << N)
    u = None
    for j, sw in enumerate(plan(lib, N, T, run)):
        Ts, c, hshift, b0, nd = sw
        gi = tile_to_global(N, sw)
        tile = v[gi]
        e = np.arange(1 << Ts)
        acc = np.zeros_like(tile)
        for b in range(b0, Ts):
            acc += tile[:, e ^ (1 << b)]
        for d in range(nd):                                    # direct bits: partner elements straight from the vector
            acc += v[gi ^ (1 << (hshift + Ts - c + d))]
        if j == 0:
            diag = np.array([lib.dsea_tfim_diag(N, int(s)) for s in range(1 << N)])
            u = np.empty_like(v)
            u[gi] = diag[gi] * tile - g * acc
        else:
            u[gi] = u[gi] - g * acc
    want = orc.TFIMOracle(N, g).H(torch.from_numpy(v)).numpy()
    assert np.abs(u - want).max() <= 1e-13 * np.abs(want).max()
