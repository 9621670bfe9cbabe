"""Host-side description of how the 2^N state vector is sharded across ranks.

Rank r of P = 2^p owns the contiguous global indices s = (r << L) | s_loc with L = N - p, i.e. the top p
spin bits ARE the rank.  Consequences the CUDA path relies on (and tests/test_sharding_gloo.py checks
with two CPU processes over gloo):
  * flips of bits i < L stay inside the shard;
  * a flip of top bit L + j pairs s_loc on rank r with the SAME s_loc on rank r ^ (1 << j): a whole-shard
    swap with one partner per top bit (NVSwitch makes all partners equidistant);
  * the diagonal needs the GLOBAL index because the periodic bond couples bit N-1 with bit 0 and bit L
    with bit L-1;
  * dot products are local sums followed by an allreduce.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import torch


@dataclass(frozen=True)
class ShardLayout:
    N: int
    world: int
    rank: int

    def __post_init__(self):
        if self.world < 1 or self.world & (self.world - 1):
            raise ValueError("world must be a power of two")
        if not 0 <= self.rank < self.world:
            raise ValueError("rank out of range")
        if self.N - self.top_bits < 1:
            raise ValueError("too few spins for this many ranks")

    @property
    def top_bits(self) -> int:
        return self.world.bit_length() - 1

    @property
    def local_bits(self) -> int:
        return self.N - self.top_bits

    @property
    def n_loc(self) -> int:
        return 1 << self.local_bits

    @property
    def offset(self) -> int:
        return self.rank << self.local_bits

    def partners(self) -> List[int]:
        """partners()[j] holds the amplitudes reached by flipping spin bit L + j."""
        return [self.rank ^ (1 << j) for j in range(self.top_bits)]

    def global_index(self, s_loc: int) -> int:
        return self.offset | s_loc

    def owner(self, s: int) -> int:
        return s >> self.local_bits


def broadcast_bytes(payload: bytes, nbytes: int, src: int = 0, device=None) -> bytes:
    """Broadcasts a fixed-size byte string (the ncclUniqueId) over the default process group."""
    import torch.distributed as dist
    t = torch.zeros(nbytes, dtype=torch.uint8)
    if dist.get_rank() == src:
        assert len(payload) == nbytes
        t = torch.tensor(list(payload), dtype=torch.uint8)
    if device is not None and dist.get_backend() == "nccl":
        t = t.to(device)
    dist.broadcast(t, src=src)
    return bytes(t.cpu().tolist())


def dist_dot_reference(a_loc: torch.Tensor, b_loc: torch.Tensor) -> torch.Tensor:
    """What `dominantsparseeigenad_b200.dot` computes on sharded vectors, spelled with torch.distributed
    (used by the CPU tests; the product path does this inside libdsea with NCCL)."""
    import torch.distributed as dist
    s = torch.dot(a_loc, b_loc).reshape(1)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(s)
    return s[0]
