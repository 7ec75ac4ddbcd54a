"""openpnm_b200 -- B200-native drop-in for OpenPNM's transport-solve path.

    from openpnm_b200 import B200PCG
    alg.run(solver=B200PCG(tol=1e-10))          # real OpenPNM algorithm object

or, with assembly / boundary conditions / Picard loop on the device as well:

    import openpnm_b200 as b2
    sf = b2.algorithms.StokesFlow(network=pn, phase=ph)   # pn, ph: OpenPNM or b2 objects
    sf.set_value_BC(pores=pn.pores('left'), values=2e5)
    sf.run(solver=b2.B200PCG(tol=1e-10))

Only the hot path lives here; everything needs a CUDA device
and the in-tree libb200pnm.so -- there is no CPU fallback.
"""
from . import _lib, algorithms, integrators, network, solvers
from .algorithms import (AdvectionDiffusion, FickianDiffusion, FourierConduction, OhmicConduction,
                         ReactiveTransport, SolutionContainer, SteadyStateSolution, StokesFlow,
                         TransientAdvectionDiffusion, TransientFickianDiffusion,
                         TransientFourierConduction, TransientOhmicConduction,
                         TransientReactiveTransport, TransientSolution, Transport)
from .integrators import B200RK45
from .network import Cubic, Network, Phase
from .solvers import B200PCG

__version__ = "0.1.0"


def register_with_openpnm(op=None):
    """Make `openpnm.solvers.B200PCG` resolvable so that
    `op.Workspace().settings.default_solver = 'B200PCG'` works
    (openpnm/algorithms/_transport.py:191-192, utils/_workspace.py:44-56)."""
    if op is None:
        import openpnm as op
    op.solvers.B200PCG = B200PCG
    if hasattr(op.solvers, "__all__") and "B200PCG" not in op.solvers.__all__:
        op.solvers.__all__.append("B200PCG")
    return op
