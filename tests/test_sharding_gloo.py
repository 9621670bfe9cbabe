"""CPU, world_size 2 and 4 over gloo: the host-side logic of the sharded path.

Each rank rebuilds its shard of H v from (a) local flips inside the shard, (b) the diagonal evaluated on
GLOBAL indices with libdsea's host-callable bit map, and (c) whole-shard swaps with `partners()` — the
exact decomposition the CUDA kernels + NCCL implement — and compares it with the oracle's full matvec.
"""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT  # noqa: F401


def _worker(rank, world, port, N, g, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from dominantsparseeigenad_b200 import _lib
        from dominantsparseeigenad_b200.sharding import ShardLayout, broadcast_bytes, dist_dot_reference
        from oracle import dsea_oracle as orc
        lib = _lib.load()
        lay = ShardLayout(N, world, rank)
        L, n_loc = lay.local_bits, lay.n_loc
        rng = np.random.default_rng(42)
        v, w = rng.standard_normal(1 << N), rng.standard_normal(1 << N)
        v_loc = torch.from_numpy(v[lay.offset:lay.offset + n_loc].copy())
        w_loc = torch.from_numpy(w[lay.offset:lay.offset + n_loc].copy())
        # (a) local flips
        s_loc = np.arange(n_loc)
        acc = np.zeros(n_loc)
        for i in range(L):
            acc += v_loc.numpy()[s_loc ^ (1 << i)]
        # (c) one whole-shard swap per top bit
        for j, peer in enumerate(lay.partners()):
            buf = torch.empty(n_loc, dtype=torch.float64)
            if rank < peer:
                dist.send(v_loc, peer); dist.recv(buf, peer)
            else:
                dist.recv(buf, peer); dist.send(v_loc, peer)
            acc += buf.numpy()
        # (b) diagonal on the global index
        diag = np.array([lib.dsea_tfim_diag(N, lay.global_index(int(s))) for s in s_loc])
        u_loc = diag * v_loc.numpy() - g * acc
        want = orc.TFIMOracle(N, g).H(torch.from_numpy(v)).numpy()[lay.offset:lay.offset + n_loc]
        assert np.abs(u_loc - want).max() <= 1e-13 * np.abs(want).max()
        # flips that leave the shard land on the partner, at the same local index
        for j, peer in enumerate(lay.partners()):
            s = lay.global_index(5)
            t = lib.dsea_tfim_flip_index(N, s, L + j)
            assert lay.owner(t) == peer and (t & (n_loc - 1)) == 5
        # distributed dot
        d = dist_dot_reference(v_loc, w_loc)
        assert abs(d.item() - float(np.dot(v, w))) <= 1e-12 * abs(float(np.dot(v, w))) + 1e-12
        # unique-id style broadcast
        payload = bytes(range(128)) if rank == 0 else b""
        assert broadcast_bytes(payload, 128, 0) == bytes(range(128))
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as exc:  # surface the failure in the parent
        import traceback
        q.put((rank, traceback.format_exc() + repr(exc)))


@pytest.mark.parametrize("world,N", [(2, 10), (4, 11)])
def test_sharded_decomposition_over_gloo(world, N):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, 1.3, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in results), results


def test_shard_layout_arithmetic():
    from dominantsparseeigenad_b200.sharding import ShardLayout
    lay = ShardLayout(30, 8, 5)
    assert lay.top_bits == 3 and lay.local_bits == 27 and lay.n_loc == 1 << 27
    assert lay.offset == 5 << 27 and lay.partners() == [4, 7, 1]
    assert lay.owner(lay.global_index(123)) == 5
    with pytest.raises(ValueError):
        ShardLayout(10, 3, 0)
    with pytest.raises(ValueError):
        ShardLayout(2, 4, 0)
