"""Process-level runtime: the libdsea context (one per process = one per GPU), pointer/stream
plumbing between torch tensors and the C ABI, and the start-vector source.

torch is used here only for device memory, streams and (when launched under torchrun) for the
rendezvous that distributes the ncclUniqueId.  No compute goes through torch on the hot path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, Optional

import torch

from . import _lib

F64 = torch.float64

_ctx = None
_start_vector_hook: Optional[Callable[[int, str], Optional[torch.Tensor]]] = None
_draw_counter = 0
_draw_seed = None
cg_start = os.environ.get("DSEA_CG_START", "random")      # "random" (reference, CG.py:58,121) | "zero" (opt-in)
# Opt-in: draw the random start vectors in the EVEN sector of the global spin flip F = prod_i sx_i,
# v <- v + F v with (F v)[s] = v[~s].  The TFIM Hamiltonian commutes with F and its finite-N ground state is even, but
# for g < 1 the odd partner is only ~g^N above it (2e-7 at N=20, g=0.5), so a generic start vector makes the Lanczos
# vector and every CG solve pick up seed-dependent odd-sector contamination: the reference's own d2E0 / chi_F scatter
# by 1e-4 there.  Starting inside the even sector removes that (rounding re-seeds the odd sector only at
# 1e-16).  Only meaningful for operators that commute with F (TFIM); default off = the reference's plain randn.
parity_sector = os.environ.get("DSEA_PARITY_SECTOR", "none")   # "none" | "even"
stats = {"cg_iters": [], "lanczos_calls": 0, "cg_calls": 0}


def record_cg_iterations(n: int, keep: int = 4096) -> None:
    """Work accounting for bench.py / tests; bounded so long optimisation loops do not grow it forever."""
    log = stats["cg_iters"]
    log.append(int(n))
    if len(log) > keep:
        del log[:-keep]


class Context:
    """Owns a dsea_ctx*.  rank/world follow torch.distributed when it is initialised."""

    def __init__(self):
        if not torch.cuda.is_available():
            raise _lib.DseaError("dominantsparseeigenad_b200 needs a CUDA device (B200, sm_100a); "
                                 "there is no CPU fallback.")
        self.lib = _lib.load()
        rank, world = 0, 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            rank, world = torch.distributed.get_rank(), torch.distributed.get_world_size()
        if world > 1:
            self.device_index = int(os.environ.get("LOCAL_RANK", rank % max(torch.cuda.device_count(), 1)))
        else:
            self.device_index = torch.cuda.current_device()
        torch.cuda.set_device(self.device_index)
        self.device = torch.device("cuda", self.device_index)
        id_buf = None
        if world > 1:
            from .sharding import broadcast_bytes
            raw = (C.c_ubyte * 128)()
            if rank == 0:
                _lib.check(self.lib.dsea_nccl_unique_id(raw))
            id_buf = (C.c_ubyte * 128)(*broadcast_bytes(bytes(raw), 128, 0, self.device))
        handle = C.c_void_p()
        _lib.check(self.lib.dsea_ctx_create(self.device_index, rank, world, id_buf, C.byref(handle)))
        self.handle = handle
        self.rank, self.world = rank, world
        self.log2world = world.bit_length() - 1

    def set_option(self, key: str, value: int) -> None:
        _lib.check(self.lib.dsea_ctx_set_option(self.handle, key.encode(), int(value)))

    def p2p_enabled(self) -> bool:
        return bool(self.lib.dsea_ctx_p2p(self.handle))

    def launch_count(self) -> int:
        return int(self.lib.dsea_launch_count(self.handle))

    PROFILE_KINDS = ("matvec", "reorth_dots", "reorth_update", "ritz", "cg_update", "normalise", "tridiag",
                     "adjoint", "noop", "exchange")

    def profile_enable(self, on: bool = True) -> None:
        _lib.check(self.lib.dsea_profile_enable(self.handle, 1 if on else 0))

    def profile_collect(self) -> dict:
        """{kind: {"ms", "bytes", "launches"}} accumulated since the previous collect (synchronises)."""
        n = len(self.PROFILE_KINDS)
        ms, by, cnt = (C.c_double * n)(), (C.c_double * n)(), (C.c_int64 * n)()
        _lib.check(self.lib.dsea_profile_collect(self.handle, n, ms, by, cnt))
        return {k: {"ms": ms[i], "bytes": by[i], "launches": int(cnt[i])} for i, k in enumerate(self.PROFILE_KINDS)}

    def col_stride(self, n: int) -> int:
        return int(self.lib.dsea_col_stride(n))


def context() -> Context:
    global _ctx
    if _ctx is None:
        _ctx = Context()
        for key, env in (("tfim_tile_bits", "DSEA_TFIM_TILE_BITS"), ("tfim_run_bits", "DSEA_TFIM_RUN_BITS"),
                         ("cg_check_every", "DSEA_CG_CHECK_EVERY"), ("reorth_ctas_per_sm", "DSEA_REORTH_CTAS"),
                         ("p2p", "DSEA_P2P"), ("tfim_pipeline", "DSEA_TFIM_PIPELINE"),
                         ("mailbox", "DSEA_MAILBOX"), ("tfim_tma", "DSEA_TFIM_TMA"), ("basis_fp32", "DSEA_BASIS_FP32"),
                         ("tfim_direct", "DSEA_TFIM_DIRECT"), ("tfim_fuse_scale", "DSEA_TFIM_FUSE_SCALE"),
                         ("cg_fuse_push", "DSEA_CG_FUSE_PUSH"), ("tfim_pipe_threads", "DSEA_TFIM_PIPE_THREADS"),
                         ("tfim_generic_min_operands", "DSEA_TFIM_GENERIC_MIN_OPERANDS"), ("pdl", "DSEA_PDL"), ("pdl_staged", "DSEA_PDL_STAGED"), ("tfim_stage", "DSEA_TFIM_STAGE"),
                         ("fuse_small", "DSEA_FUSE_SMALL")):
            if os.environ.get(env):
                _ctx.set_option(key, int(os.environ[env]))
    return _ctx


def set_basis_precision(kind: str) -> None:
    """"fp64" (default; the reference's precision, Lanczos.py:43,49) or "fp32": opt-in shadow basis for native TFIM
    operators — the Lanczos vectors are stored rounded to fp32 (half the HBM traffic of the re-orthogonalisation and
    half the footprint), every accumulation stays fp64, and the eigenpair is polished in fp64 by one
    Jacobi-Davidson step, so E0 / psi0 / gradients keep the tolerances of the fp64 path (tests/test_gpu_parity.py)."""
    if kind not in ("fp64", "fp32"):
        raise ValueError("basis precision must be 'fp64' or 'fp32'")
    context().set_option("basis_fp32", 1 if kind == "fp32" else 0)


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a contiguous fp64 CUDA tensor (None passes NULL)."""
    if t is None:
        return None
    assert t.is_cuda and t.dtype == F64 and t.is_contiguous(), (t.device, t.dtype, t.is_contiguous())
    return t.data_ptr()


def dev_vec(t: torch.Tensor, device: torch.device) -> torch.Tensor:
    """fp64, contiguous, 16-byte aligned copy-if-needed of `t` on `device` (detached)."""
    t = t.detach()
    if t.device != device or t.dtype != F64:
        t = t.to(device=device, dtype=F64)
    if not t.is_contiguous():
        t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone()
    return t


def empty(n: int, device: torch.device) -> torch.Tensor:
    return torch.empty(int(n), dtype=F64, device=device)


# ---- start vectors -------------------------------------------------------------------------------
def set_start_vector_hook(fn: Optional[Callable[[int, str], Optional[torch.Tensor]]]) -> None:
    """Tests use this to inject q0 / x0.  `fn(n_loc, kind)` with kind in {"lanczos", "cg"} returns a
    tensor (any device) or None to fall through to the device Philox generator."""
    global _start_vector_hook
    _start_vector_hook = fn


def reset_draw_counter() -> None:
    """Restarts the start-vector stream (the next draw is draw 1 of the current torch seed)."""
    global _draw_counter
    _draw_counter = 0


def start_vector(n_loc: int, kind: str, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Standard-normal start vector on the GPU.  Reference: torch.randn at Lanczos.py:52, CG.py:58,121.

    Reproducible under torch.manual_seed: the Philox key is torch's initial seed and the stream id a draw
    counter that restarts whenever that seed changes (so `manual_seed(s)` followed by the same sequence of
    solves draws the same vectors; re-seeding with the SAME value does not restart it — call
    `reset_draw_counter()` for that).  The Philox counter is the global element index, so sharded runs draw
    the same vector as a single GPU.
    """
    global _draw_counter, _draw_seed
    ctx = context()
    if out is None:
        out = empty(n_loc, ctx.device)
    if _start_vector_hook is not None:
        v = _start_vector_hook(n_loc, kind)
        if v is not None:
            out.copy_(v.to(device=ctx.device, dtype=F64))
            return out
    seed = torch.initial_seed() & 0xFFFFFFFFFFFFFFFF
    if seed != _draw_seed:
        _draw_seed, _draw_counter = seed, 0
    _draw_counter += 1
    _lib.check(ctx.lib.dsea_randn(ctx.handle, n_loc, seed, _draw_counter, ptr(out), stream_ptr()))
    if parity_sector == "even":
        out.add_(spin_flip(out))
    return out


def spin_flip(v: torch.Tensor) -> torch.Tensor:
    """(F v)[s] = v[~s] for the global index s: on one GPU the reversed vector; sharded, the reversed shard of the
    mirror rank P-1-r (one pairwise exchange through torch.distributed)."""
    ctx = context()
    mine = v.flip(0).contiguous()
    if ctx.world == 1:
        return mine
    import torch.distributed as dist
    peer = ctx.world - 1 - ctx.rank
    got = torch.empty_like(mine)
    ops = [dist.P2POp(dist.isend, mine, peer), dist.P2POp(dist.irecv, got, peer)]
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    return got
