"""Integrator objects mirroring `openpnm.integrators`
(openpnm/integrators/_scipy.py:9-56 `ScipyRK45(atol, rtol)`): `B200RK45` runs
SciPy's RK45 scheme (Dormand-Prince 5(4), same step controller and dense
output) on the GPU inside libb200pnm.so."""

__all__ = ["Integrator", "B200RK45"]


class Integrator:
    """openpnm/integrators/_base.py"""

    def solve(self, rhs, x0, tspan, saveat, **kwargs):
        raise NotImplementedError


class B200RK45(Integrator):
    """Explicit Runge-Kutta 5(4) time stepping on one B200 -- or, under torchrun, on the
    slabs of all ranks (every right-hand side exchanges the halo of its operand, the step
    controller's error norms are all-reduced, so all ranks take the same steps).

    Parameters mirror `ScipyRK45`: ``atol``, ``rtol`` (defaults 1e-6).  The
    transient algorithms of `openpnm_b200.algorithms` hand it their assembled
    device problem; there is no host right-hand-side callback.
    """

    _b200_native = True

    def __init__(self, atol=1e-6, rtol=1e-6, verbose=False, device=None, fmt="auto", max_saved=None,
                 distributed=True):
        if device is None:          # under torchrun: one process per GPU
            import os
            device = int(os.environ.get("LOCAL_RANK", "0"))
        self.distributed = distributed   # slab-parallel run() when torch.distributed is up
        self.atol = atol
        self.rtol = rtol
        self.verbose = verbose
        self.device = device
        self.fmt = fmt
        self.max_saved = max_saved
        self._ctx = None

    @property
    def ctx(self):
        if self._ctx is None:
            from .device import Context
            self._ctx = Context(self.device)
        return self._ctx

    def _fmt_code(self):
        from . import _lib
        return {"auto": 0, "csr": _lib.FMT_CSR, "dia": _lib.FMT_DIA}[self.fmt]
