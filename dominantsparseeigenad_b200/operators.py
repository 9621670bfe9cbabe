"""Linear operators for the dominant-eigenpair path.

Three native (device-resident) operators wrap a `dsea_op*` descriptor of libdsea:

* `TFIM(N, device)`          mirrors examples/TFIM/TFIM.py:5-101 (attributes N, dim, g; callables H,
                             pHpg, Hadjoint_to_gadjoint) but holds NO (2^N, N) index table: flips and the
                             diagonal come from bit arithmetic inside the kernels.
* `SparseMatrixOperator`     A(p) = CSR + diag(p): the "explicit sparse matrix + trainable diagonal" case of
                             examples/schrodinger1D.py:18-34 (Hsparse / Hadjoint_to_padjoint).
* `DenseOperator`            symmetric dense matrix (Lanczos.py:48, CG.py:23).

`CallbackOperator` adapts an arbitrary Python callable `A(v)` (the reference's "A is a function"
contract): the operator application happens in user code, everything else (re-orthogonalisation,
tridiagonal eigensolve, Ritz GEMV, CG vector updates, projections) still runs in libdsea.

All of them expose the same solver surface used by symeig.py / CG.py / Lanczos.py:
    lanczos(param, k, which) -> (evals[2], evec_min, evec_max, info)
    cg(param, shift, b, x0)  -> x
    adjoint(v1, v2)          -> gradient w.r.t. the parameter (torch-differentiable in v1, v2)
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional

import torch

from . import _lib, runtime
from .runtime import F64, context, dev_vec, empty, ptr, stream_ptr

CG_EPS = 1e-7          # CG.py:25


# =================================================================================================
# differentiable level-1 pieces (global reductions go through libdsea so they are shard-aware)
# =================================================================================================
class Dot(torch.autograd.Function):
    """a . b as a 0-dim tensor, summed over all ranks.  Replaces torch.matmul(v, w) on vectors."""

    @staticmethod
    def forward(ctx, a, b):
        rt = context()
        a_, b_ = dev_vec(a, rt.device), dev_vec(b, rt.device)
        out = torch.empty((), dtype=F64, device=rt.device)
        _lib.check(rt.lib.dsea_dot(rt.handle, a_.numel(), ptr(a_), ptr(b_), out.data_ptr(), stream_ptr()))
        ctx.save_for_backward(a, b)
        return out

    @staticmethod
    def backward(ctx, go):
        a, b = ctx.saved_tensors
        return scale(go, b), scale(go, a)


def dot(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return Dot.apply(a, b)


class Scale(torch.autograd.Function):
    """s * v for a REPLICATED scalar s (0-dim or one element) and a SHARDED vector v.

    A plain `s * v` would let torch reduce the gradient of s with a rank-local sum; here it is the
    global inner product (allreduced inside libdsea), so replicated scalars stay replicated in every
    order of differentiation.  With one rank this is exactly torch's broadcasting multiply.
    """

    @staticmethod
    def forward(ctx, s, v):
        ctx.save_for_backward(s, v)
        return s.reshape(()).to(v.device) * v

    @staticmethod
    def backward(ctx, go):
        s, v = ctx.saved_tensors
        grad_s = Dot.apply(go, v).reshape(s.shape).to(s.device) if ctx.needs_input_grad[0] else None
        grad_v = Scale.apply(s, go) if ctx.needs_input_grad[1] else None
        return grad_s, grad_v


def scale(s, v: torch.Tensor) -> torch.Tensor:
    """Shard-safe `s * v`; python floats and non-differentiable scalars take the cheap path."""
    if not isinstance(s, torch.Tensor):
        return s * v
    if s.numel() != 1:
        raise ValueError("scale() expects a single-element scalar")
    return Scale.apply(s, v)


class Project(torch.autograd.Function):
    """b - (psi . b) psi   (symeig.py:27,80; CG.py:59,67,122,132) as one dot pass + one fused axpy pass."""

    @staticmethod
    def forward(ctx, psi, b):
        rt = context()
        psi_, b_ = dev_vec(psi, rt.device), dev_vec(b, rt.device)
        out = empty(b_.numel(), rt.device)
        _lib.check(rt.lib.dsea_project(rt.handle, b_.numel(), ptr(psi_), ptr(b_), ptr(out), stream_ptr()))
        ctx.save_for_backward(psi, b)
        return out.to(b.device)             # CPU tensors in => CPU tensors out, like Scale and Dot's callers

    @staticmethod
    def backward(ctx, go):
        psi, b = ctx.saved_tensors
        grad_b = Project.apply(psi, go).to(b.device) if ctx.needs_input_grad[1] else None
        grad_psi = None
        if ctx.needs_input_grad[0]:
            go_p, b_p = go.to(psi.device), b.to(psi.device)
            grad_psi = -scale(dot(psi, b_p).to(psi.device), go_p) - scale(dot(go_p, psi).to(psi.device), b_p)
        return grad_psi, grad_b


def project(psi: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return Project.apply(psi, b)


# =================================================================================================
# native operators
# =================================================================================================
class NativeOperator:
    """Base of the operators libdsea applies itself."""

    _dsea_native = True

    def __init__(self):
        self.rt = context()
        self.device = self.rt.device
        self.handle = C.c_void_p()
        self.n_loc = 0

    def __del__(self):
        try:
            if getattr(self, "handle", None) and self.handle.value:
                self.rt.lib.dsea_op_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass

    # ---- parameter handling (overridden) ----
    def param_ptr(self, param: Optional[torch.Tensor]):
        raise NotImplementedError

    def _work(self) -> Optional[torch.Tensor]:
        nw = int(self.rt.lib.dsea_op_work_doubles(self.handle))
        return empty(nw, self.device) if nw else None

    # ---- raw (non-differentiable) applications ----
    def matvec_raw(self, param, v: torch.Tensor, shift: Optional[torch.Tensor] = None) -> torch.Tensor:
        v_ = dev_vec(v, self.device)
        u = empty(self.n_loc, self.device)
        work = self._work()
        keep = self.param_ptr(param)
        _lib.check(self.rt.lib.dsea_matvec(self.rt.handle, self.handle, keep[0], ptr(shift), ptr(v_), ptr(u), None,
                                           ptr(work), stream_ptr()))
        return u

    def lanczos(self, param, k: int, which: int, want_info: bool = False):
        """Device-resident k-step Lanczos + tridiagonal eigensolve + Ritz vectors (Lanczos.py:3-105)."""
        rt, lib = self.rt, self.rt.lib
        n, ldq = self.n_loc, self.rt.col_stride(self.n_loc)
        # k * ldq doubles, or k float columns under the opt-in fp32 shadow basis (runtime.set_basis_precision)
        Q = empty(int(lib.dsea_lanczos_basis_doubles(self.handle, k)), self.device)
        runtime.start_vector(n, "lanczos", out=Q[:n])
        work = empty(int(lib.dsea_lanczos_work_doubles(self.handle)), self.device)
        alpha, beta = empty(k, self.device), empty(k, self.device)
        evals = empty(2, self.device)
        vmin = empty(n, self.device) if which in (_lib.DSEA_MIN, _lib.DSEA_BOTH) else None
        vmax = empty(n, self.device) if which in (_lib.DSEA_MAX, _lib.DSEA_BOTH) else None
        info = (C.c_int64 * 2)() if want_info else None
        keep = self.param_ptr(param)
        _lib.check(lib.dsea_lanczos(rt.handle, self.handle, keep[0], k, which, ptr(Q), ptr(work), ptr(alpha),
                                    ptr(beta), ptr(evals), ptr(vmin), ptr(vmax), info, stream_ptr()))
        runtime.stats["lanczos_calls"] += 1
        fp32 = Q.numel() != k * ldq
        return evals, vmin, vmax, {"Q": Q, "ldq": ldq, "alpha": alpha, "beta": beta, "basis": "fp32" if fp32 else "fp64",
                                   "k_eff": int(info[0]) if info else None}

    def cg(self, param, shift: Optional[torch.Tensor], b: torch.Tensor, x0: torch.Tensor,
           eps: float = CG_EPS, maxit: int = 0) -> torch.Tensor:
        """Solves (A(param) - shift) x = b from x0 (CG.py:3-41); returns a new tensor."""
        rt, lib = self.rt, self.rt.lib
        b_ = dev_vec(b, self.device)
        x = dev_vec(x0, self.device).clone()
        shift_ = None if shift is None else dev_vec(shift.reshape(-1), self.device)
        work = empty(int(lib.dsea_cg_work_doubles(self.handle)), self.device)
        iters = C.c_int64(0)
        keep = self.param_ptr(param)
        if maxit <= 0:
            maxit = int(getattr(self, "dim", self.n_loc))      # CG.py:32 iterates at most n (GLOBAL dimension) times
        _lib.check(lib.dsea_cg(rt.handle, self.handle, keep[0], ptr(shift_), ptr(b_), ptr(x), ptr(work), eps, maxit,
                               C.byref(iters), stream_ptr()), allow_noconv=True)
        runtime.stats["cg_calls"] += 1
        runtime.record_cg_iterations(int(iters.value))
        return x


class _PHpg(torch.autograd.Function):
    """(dH/dg) v = -sum_i v[s ^ (1<<i)]  (TFIM.py:58-65).  Linear and symmetric: its own adjoint."""

    @staticmethod
    def forward(ctx, model, v):
        v_ = dev_vec(v, model.device)
        u = empty(model.n_loc, model.device)
        work = model._work()
        _lib.check(model.rt.lib.dsea_tfim_dHdg(model.rt.handle, model.handle, ptr(v_), ptr(u), ptr(work), stream_ptr()))
        ctx.model = model
        return u

    @staticmethod
    def backward(ctx, go):
        return None, _PHpg.apply(ctx.model, go)


class _TFIMAdjoint(torch.autograd.Function):
    """v1^T (dH/dg) v2 as a shape-[1] tensor (TFIM.py:100-101), fused gather + dot, no output vector."""

    @staticmethod
    def forward(ctx, model, v1, v2):
        a, b = dev_vec(v1, model.device), dev_vec(v2, model.device)
        out = torch.empty(1, dtype=F64, device=model.device)
        work = model._work()
        _lib.check(model.rt.lib.dsea_adjoint(model.rt.handle, model.handle, ptr(a), ptr(b), out.data_ptr(), ptr(work),
                                             stream_ptr()))
        ctx.model = model
        ctx.save_for_backward(v1, v2)
        return out

    @staticmethod
    def backward(ctx, go):
        v1, v2 = ctx.saved_tensors
        m = ctx.model
        return None, scale(go, _PHpg.apply(m, v2)), scale(go, _PHpg.apply(m, v1))


class _TFIMMatvec(torch.autograd.Function):
    """H(g) v  (TFIM.py:91-98), differentiable in g and v."""

    @staticmethod
    def forward(ctx, model, g, v):
        ctx.model = model
        ctx.save_for_backward(g, v)
        return model.matvec_raw(g, v)

    @staticmethod
    def backward(ctx, go):
        g, v = ctx.saved_tensors
        m = ctx.model
        grad_g = _TFIMAdjoint.apply(m, go, v).reshape(g.shape) if ctx.needs_input_grad[1] else None
        grad_v = _TFIMMatvec.apply(m, g, go) if ctx.needs_input_grad[2] else None
        return None, grad_g, grad_v


class _BoundCallable:
    """A callable attribute that carries its native operator (so set*() can detect the fast path)."""

    def __init__(self, op, fn, name):
        self._dsea_operator = op
        self._fn = fn
        self.__name__ = name

    def __call__(self, *args):
        return self._fn(*args)


class TFIM(NativeOperator):
    """1-D transverse-field Ising chain, H = -sum_i (g sx_i + sz_i sz_{i+1}), periodic.

    Drop-in for examples/TFIM/TFIM.py: same attribute names (`N`, `dim`, `g`, `device`) and callables
    (`H`, `pHpg`, `Hadjoint_to_gadjoint`).  Under torchrun with world > 1 every vector is the rank-local
    shard of `n_loc = 2^N / world` amplitudes (top log2(world) spin bits = rank) while `dim` stays global.
    """

    def __init__(self, N: int, device=None):
        super().__init__()
        self.N = int(N)
        self.dim = 1 << self.N
        _lib.check(self.rt.lib.dsea_op_tfim(self.rt.handle, self.N, C.byref(self.handle)))
        self.n_loc = int(self.rt.lib.dsea_op_local_dim(self.handle))
        self.g: Optional[torch.Tensor] = None
        self.H = _BoundCallable(self, self._H, "H")
        self.pHpg = _BoundCallable(self, self._pHpg, "pHpg")
        self.Hadjoint_to_gadjoint = _BoundCallable(self, self._adjoint, "Hadjoint_to_gadjoint")

    def param_ptr(self, param):
        g = self.g if param is None else param
        if g is None:
            raise ValueError("TFIM.g is not set")
        g_ = dev_vec(g.reshape(-1), self.device)
        return g_.data_ptr(), g_          # keep the tensor alive alongside the pointer

    def _H(self, v):
        return _TFIMMatvec.apply(self, self.g, v)

    def _pHpg(self, v):
        return _PHpg.apply(self, v)

    def _adjoint(self, v1, v2):
        return _TFIMAdjoint.apply(self, v1, v2)

    def adjoint(self, v1, v2, param=None):
        out = _TFIMAdjoint.apply(self, v1, v2)
        return out if param is None else out.reshape(param.shape)

    # ---- dense builders used by the reference's cross-checks (small N only) ----
    def setHmatrix(self):
        """Dense H as a torch.Tensor in `self.Hmatrix`, differentiable in `self.g` (TFIM.py:67-89).

        Built from the same bit arithmetic as the kernels; like the reference it adds a symmetric 1e-12
        noise so that torch's full-spectrum AD does not divide by zero on exact degeneracies."""
        if self.rt.world != 1:
            raise RuntimeError("setHmatrix is a single-GPU, small-N cross-check")
        n, dev = self.dim, self.device
        s = torch.arange(n, device=dev)
        diag = torch.tensor([self.diagonal_element(self.N, int(x)) for x in range(n)], dtype=F64, device=dev)
        off = torch.zeros(n, n, dtype=F64, device=dev)
        for i in range(self.N):
            off[s ^ (1 << i), s] = 1.0
        noise = 1e-12 * torch.randn(n, n, dtype=F64, device=dev)
        self.Hmatrix = torch.diag(diag) - self.g.to(dev) * off + 0.5 * (noise + noise.T)
        return self.Hmatrix

    def setpHpg(self):
        """Dense dH/dg in `self.pHpgmatrix` (TFIM.py:53-56)."""
        n, dev = self.dim, self.device
        s = torch.arange(n, device=dev)
        m = torch.zeros(n, n, dtype=F64, device=dev)
        for i in range(self.N):
            m[s ^ (1 << i), s] = -1.0
        self.pHpgmatrix = m
        return m

    # bit maps, host side (bit-exact contract with TFIM.py:39-51)
    @staticmethod
    def flip_index(N: int, s: int, i: int) -> int:
        return int(_lib.load().dsea_tfim_flip_index(N, s, i))

    @staticmethod
    def diagonal_element(N: int, s: int) -> float:
        return float(_lib.load().dsea_tfim_diag(N, s))


class SparseMatrixOperator(NativeOperator):
    """A(p) = CSR + diag(p) with p the trainable parameter (schrodinger1D.py:18-34 generalised).

    `H(v)` applies it with the current `self.potential`; the parameter adjoint of v1 v2^T is v1 o v2.
    """

    def __init__(self, rowptr: torch.Tensor, colidx: torch.Tensor, vals: torch.Tensor, n: int,
                 potential: Optional[torch.Tensor] = None):
        super().__init__()
        dev = self.device
        self.rowptr = rowptr.to(device=dev, dtype=torch.int64).contiguous()
        self.colidx = colidx.to(device=dev, dtype=torch.int64).contiguous()
        self.vals = vals.to(device=dev, dtype=F64).contiguous()
        self.n_loc = self.dim = int(n)
        _lib.check(self.rt.lib.dsea_op_csr(self.rt.handle, self.n_loc, self.vals.numel(), self.rowptr.data_ptr(),
                                           self.colidx.data_ptr(), self.vals.data_ptr(), C.byref(self.handle)))
        self.potential = potential
        self.H = _BoundCallable(self, self._H, "H")
        self.Hadjoint_to_padjoint = _BoundCallable(self, self._adjoint, "Hadjoint_to_padjoint")

    @classmethod
    def from_scipy(cls, m, potential=None):
        m = m.tocsr()
        return cls(torch.from_numpy(m.indptr.astype("int64")), torch.from_numpy(m.indices.astype("int64")),
                   torch.from_numpy(m.data.astype("float64")), m.shape[0], potential)

    def param_ptr(self, param):
        p = self.potential if param is None else param
        if p is None:
            return None, None
        p_ = dev_vec(p, self.device)
        return p_.data_ptr(), p_

    def _H(self, v):
        return self.matvec_raw(self.potential, v)

    @staticmethod
    def _adjoint(v1, v2):
        return v1 * v2

    def adjoint(self, v1, v2, param=None):
        return v1 * v2


class DenseOperator(NativeOperator):
    """Symmetric dense matrix held on the GPU (row-major)."""

    def __init__(self, A: torch.Tensor):
        super().__init__()
        self.A = dev_vec(A, self.device)
        assert self.A.dim() == 2 and self.A.shape[0] == self.A.shape[1]
        self.n_loc = self.dim = int(self.A.shape[0])
        _lib.check(self.rt.lib.dsea_op_dense(self.rt.handle, self.n_loc, self.A.stride(0), self.A.data_ptr(),
                                             C.byref(self.handle)))

    def param_ptr(self, param):
        return None, None

    def adjoint(self, v1, v2, param=None):
        return v1[:, None] * v2


# =================================================================================================
# Python-callable operators
# =================================================================================================
class CallbackOperator:
    """Adapter for a user callable A(v) (and optionally its parameter adjoint).

    The callable is invoked on `call_device`: CUDA tensors are passed straight through; when the user's
    closure lives on the CPU (e.g. schrodinger1D.py as shipped) vectors are staged through host memory
    for that one call.  Everything else stays on the GPU inside libdsea.
    """

    _dsea_native = False

    def __init__(self, A: Callable, n: int, call_device: torch.device, Aadjoint: Optional[Callable] = None):
        self.rt = context()
        self.device = self.rt.device
        self.A = A
        self.Aadjoint = Aadjoint
        self.n_loc = self.dim = int(n)
        self.call_device = torch.device(call_device)

    def _apply(self, v: torch.Tensor) -> torch.Tensor:
        with torch.no_grad():
            if self.call_device.type == "cuda":
                u = self.A(v)
            else:
                u = self.A(v.to(self.call_device))
        return dev_vec(u, self.device)

    def lanczos(self, param, k: int, which: int, want_info: bool = False):
        rt, lib = self.rt, self.rt.lib
        n, ldq = self.n_loc, self.rt.col_stride(self.n_loc)
        Q = empty(k * ldq, self.device)
        runtime.start_vector(n, "lanczos", out=Q[:n])
        alpha, beta = empty(k, self.device), empty(k, self.device)
        st = stream_ptr()
        _lib.check(lib.dsea_lanczos_start(rt.handle, n, ptr(Q), st))
        for i in range(k):
            u = self._apply(Q[i * ldq:i * ldq + n])                                   # Lanczos.py:54,71
            _lib.check(lib.dsea_lanczos_step(rt.handle, n, k, i, ptr(Q), ptr(u), ptr(alpha), ptr(beta), st))
        evals = empty(2, self.device)
        vmin = empty(n, self.device) if which in (_lib.DSEA_MIN, _lib.DSEA_BOTH) else None
        vmax = empty(n, self.device) if which in (_lib.DSEA_MAX, _lib.DSEA_BOTH) else None
        info = (C.c_int64 * 2)() if want_info else None
        _lib.check(lib.dsea_lanczos_ritz(rt.handle, n, k, which, ptr(Q), ptr(alpha), ptr(beta), ptr(evals), ptr(vmin),
                                         ptr(vmax), info, st))
        runtime.stats["lanczos_calls"] += 1
        return evals, vmin, vmax, {"Q": Q, "ldq": ldq, "alpha": alpha, "beta": beta,
                                   "k_eff": int(info[0]) if info else None}

    def cg(self, param, shift, b, x0, eps: float = CG_EPS, maxit: int = 0):
        rt, lib = self.rt, self.rt.lib
        n = self.n_loc
        b_ = dev_vec(b, self.device)
        x = dev_vec(x0, self.device).clone()
        sh = None if shift is None else shift.detach().to(self.device)
        op = (lambda v: self._apply(v)) if sh is None else (lambda v: self._apply(v) - sh * v)    # CG.py:120
        r, d = empty(n, self.device), empty(n, self.device)
        st = stream_ptr()
        state = (C.c_double * 3)(eps, float(maxit if maxit > 0 else n), 0.0)
        _lib.check(lib.dsea_cg_init(rt.handle, n, ptr(b_), ptr(op(x)), ptr(r), ptr(d), state, st))
        check_every = 8
        it = 0
        limit = maxit if maxit > 0 else n
        while state[2] == 0.0 and it < limit:
            for q in range(check_every):
                Ad = op(d)
                last = (q == check_every - 1)
                _lib.check(lib.dsea_cg_update(rt.handle, n, ptr(x), ptr(r), ptr(d), ptr(Ad), state if last else None, st))
            it += check_every
        runtime.stats["cg_calls"] += 1
        runtime.record_cg_iterations(int(state[1]))
        if state[2] != 1.0:
            import warnings
            warnings.warn(f"CG stopped after {int(state[1])} iterations with |r| = {state[0]:.3e} >= eps = {eps:.3e} "
                          "(CG.py:32 returns silently here)", _lib.ConvergenceWarning, stacklevel=2)
        return x

    def adjoint(self, v1, v2, param=None):
        if self.Aadjoint is None:
            raise ValueError("this operator has no parameter adjoint")
        if self.call_device.type == "cuda":
            return self.Aadjoint(v1, v2)
        return self.Aadjoint(v1.to(self.call_device), v2.to(self.call_device))


def as_operator(A, n: Optional[int], call_device, Aadjoint=None):
    """Resolves what the reference calls `A` into an operator object."""
    op = getattr(A, "_dsea_operator", None)
    if op is not None:
        return op
    if isinstance(A, NativeOperator) or isinstance(A, CallbackOperator):
        return A
    if isinstance(A, torch.Tensor):
        return DenseOperator(A)
    if callable(A):
        if n is None:
            raise ValueError("`dim` is required when A is a function (Lanczos.py:9-11)")
        return CallbackOperator(A, n, call_device, Aadjoint)
    raise TypeError(f"cannot interpret {type(A)} as a linear operator")
