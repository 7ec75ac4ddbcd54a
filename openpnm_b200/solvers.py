"""Drop-in solver objects mirroring `openpnm.solvers`.

`B200PCG` has the interface of the reference's iterative solvers
(openpnm/solvers/_base.py:21-63 `IterativeSolver(tol, maxiter)`;
openpnm/solvers/_scipy.py:18-26 `ScipyCG.solve(A, b, **kwargs)`): it is
passed as `alg.run(solver=B200PCG())` and is called as
`solver.solve(A=self.A, b=self.b, x0=x0) -> (x, exit_code)`
(openpnm/algorithms/_transport.py:213).  The arithmetic runs in
libb200pnm.so on the GPU; there is no CPU path.
"""
import numpy as np
from numpy.linalg import norm

from . import _lib

__all__ = ["BaseSolver", "IterativeSolver", "B200PCG"]


class BaseSolver:
    """Base class for all solvers (openpnm/solvers/_base.py:6-14)."""

    def __init__(self):
        ...

    def solve(self):
        raise NotImplementedError


class IterativeSolver(BaseSolver):
    """Base class for iterative solvers (openpnm/solvers/_base.py:21-63)."""

    def __init__(self, tol=1e-8, maxiter=1000):
        self.tol = tol
        self.maxiter = maxiter
        self.atol = None
        self.rtol = None

    def _get_atol(self, b):
        return norm(b) * self.tol

    def _get_rtol(self, A, b, x0):
        res0 = self._get_residual(A, b, x0)
        atol = self._get_atol(b)
        return atol / res0

    def _get_residual(self, A, b, x):
        return norm(A * x - b)


class B200PCG(IterativeSolver):
    """Jacobi-preconditioned conjugate gradients on one B200.

    Parameters
    ----------
    tol : float
        Stop when ``norm(b - A x) <= tol * norm(b)`` -- the rule `ScipyCG`
        applies (``tol=self.tol, atol=norm(b)*tol``).
    maxiter : int or None
        ``None`` means ``10 * n``, which is what `ScipyCG` effectively uses
        (it never forwards ``self.maxiter``; SciPy's default applies).
    device : int or None
        CUDA device ordinal; ``None`` means ``LOCAL_RANK`` (torchrun) or 0.
    distributed : bool
        With ``torch.distributed`` initialised (torchrun, one process per GPU),
        ``openpnm_b200`` algorithms split the network into slabs over the ranks.
    fmt : {'auto', 'csr', 'dia'}
        Operator storage; 'auto' picks symmetric-diagonal storage when the
        matrix has at most 13 distinct super-diagonals (lattice networks, up to
        Cubic(connectivity=26)).
    precond : {'jacobi', 'amg'}
        'amg' preconditions the CG with an aggregation-multigrid K-cycle built
        on the device (the role `PyamgRugeStubenSolver` plays in the reference,
        solvers/_pyamg.py:10-17): tens of iterations at any size instead of
        O(N) on an N^3 lattice.  Same stopping rule; symmetric systems with more
        than 2048 unknowns (slab runs under torchrun included: one hierarchy whose
        levels are row-partitioned by the slabs), Jacobi otherwise.

    Returns from ``solve``: ``(x, exit_code)`` with SciPy's convention --
    0 converged, >0 maxiter reached, <0 breakdown (matrix not SPD).
    """

    _b200_native = True

    def __init__(self, tol=1e-8, maxiter=None, device=None, fmt="auto", distributed=True,
                 precond="jacobi"):
        super().__init__(tol=tol, maxiter=maxiter)
        if precond not in _lib.PRECOND_CODES:
            raise ValueError("precond must be 'jacobi' or 'amg'")
        self.precond = precond
        if device is None:          # under torchrun: one process per GPU
            import os
            device = int(os.environ.get("LOCAL_RANK", "0"))
        self.device = device
        self.fmt = fmt
        self.distributed = distributed   # slab-parallel `alg.run()` when torch.distributed is up
        self._ctx = None
        self.info = {}

    # one context per solver object, created lazily and reused across the
    # thousands of calls of a Picard loop (_reactive_transport.py:196-210)
    @property
    def ctx(self):
        if self._ctx is None:
            from .device import Context
            self._ctx = Context(self.device)
            self._ctx.set_precond(self.precond)
        return self._ctx

    def _fmt_code(self):
        return {"auto": 0, "csr": _lib.FMT_CSR, "dia": _lib.FMT_DIA}[self.fmt]

    def solve(self, A, b, x0=None, **kwargs):
        """Solves the given linear system of equations Ax=b."""
        import scipy.sparse as sprs
        if not sprs.issparse(A):
            raise TypeError("A must be a scipy.sparse matrix")
        b = np.asarray(b, dtype=np.float64)
        n = b.size
        if A.shape != (n, n):
            raise ValueError("shape mismatch between A and b")
        if A.format != "coo":
            A = A.tocoo()
        if max(A.shape) >= 2**31 or A.nnz >= 2**31:
            raise ValueError("matrix too large for int32 indices")
        ctx = self.ctx
        maxiter = 0 if self.maxiter is None else int(self.maxiter)
        ctx.set_format(self._fmt_code())
        x, iters, relres, info = ctx.solve_coo_host(
            A.row, A.col, A.data, n, b, x0, self.tol, 0.0, maxiter)
        st = ctx.stats()
        self.info = dict(iters=iters, relres=relres, exit_code=info,
                         fmt=_lib.FMT_NAMES.get(st["fmt"], "?"), ndiags=st["ndiags"],
                         t_setup_ms=st["t_setup_ms"], t_solve_ms=st["t_solve_ms"])
        return x, info
