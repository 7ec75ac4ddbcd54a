"""Import shim for the *reference* OpenPNM (TEST INFRASTRUCTURE ONLY).

The reference at /root/reference (or an offline copy under baseline/_ref) does
not import in this image out of the box: several of its plotting / docstring /
thermo dependencies are absent and it targets older numpy / scipy APIs.  This
module installs `sys.modules` stubs for the missing packages and re-creates the
few removed numpy/scipy names so the reference's own transport path
(Cubic, Phase, StokesFlow, FickianDiffusion, ReactiveTransport, ScipySpsolve,
ScipyCG) runs unmodified.  Nothing in `openpnm_b200/` imports this file; it is
used by `tests/golden/make_golden.py` (golden-vector generation in the build
container) and by the optional drop-in tests that drive the real reference
classes with the B200 solver.

Shim items:
  1. stubs: docrep, matplotlib.*, mpl_toolkits.*, transforms3d, pyamg, h5py,
     skimage, thermo, chemicals, pypardiso is left ABSENT on purpose.
  2. scipy.optimize.nonlin.TerminationCondition (moved to scipy.optimize._nonlin)
     needed by openpnm/algorithms/_reactive_transport.py:6
  3. np.in1d (removed in numpy 2.x) and np.reshape(newshape=...) keyword
     needed by openpnm/_skgraph/queries/_funcs.py:167
  4. scipy.sparse.linalg.cg(tol=...) -> rtol, used by openpnm/solvers/_scipy.py:26
"""
import collections
import importlib
import os
import sys
import types

_REF_CANDIDATES = [
    os.environ.get("OPENPNM_REFERENCE", ""),
    "/root/reference",
    os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                 "baseline", "_ref"),
]


def find_reference():
    for p in _REF_CANDIDATES:
        if p and os.path.isdir(os.path.join(p, "openpnm")):
            return p
    return None


class _Anything:
    """Callable/attr-able placeholder for stubbed third-party symbols."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        # used as a decorator -> identity
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __iter__(self):
        return iter(())


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


def _stub(name):
    if name in sys.modules:
        return sys.modules[name]
    try:
        return importlib.import_module(name)
    except Exception:
        pass
    m = _StubModule(name)
    m.__path__ = []
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent:
        setattr(_stub(parent), child, m)
    return m


def _install_docrep():
    if "docrep" in sys.modules:
        return
    try:
        import docrep  # noqa: F401
        return
    except Exception:
        pass
    m = types.ModuleType("docrep")

    class DocstringProcessor:
        param_like_sections = []
        text_sections = []
        all_sections = []

        def __init__(self, *a, **k):
            self.params = collections.defaultdict(str)

        def _ident(self, *a, **k):
            if len(a) == 1 and callable(a[0]) and not k:
                return a[0]
            return lambda f: f

        get_sections = _ident
        dedent = _ident
        with_indent = _ident
        get_full_description = _ident
        get_extended_summary = _ident
        get_summary = _ident

        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return self._ident

    m.DocstringProcessor = DocstringProcessor
    sys.modules["docrep"] = m


def _install_matplotlib():
    try:
        import matplotlib  # noqa: F401
        return
    except Exception:
        pass
    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.cm",
                 "matplotlib.colors", "matplotlib.collections",
                 "matplotlib.patches", "matplotlib._docstring",
                 "matplotlib.docstring", "matplotlib.pylab",
                 "mpl_toolkits", "mpl_toolkits.mplot3d",
                 "mpl_toolkits.mplot3d.art3d"]:
        _stub(name)

    class Substitution:
        def __init__(self, *a, **k):
            pass

        def __call__(self, f):
            return f

    sys.modules["matplotlib._docstring"].Substitution = Substitution
    sys.modules["matplotlib.docstring"].Substitution = Substitution


def _install_numpy_scipy_compat():
    import numpy as np
    if not hasattr(np, "in1d"):
        np.in1d = np.isin
    _orig_reshape = np.reshape
    if not getattr(_orig_reshape, "_b200_shim", False):
        def reshape(a, *args, **kwargs):
            if "newshape" in kwargs:
                kwargs["shape"] = kwargs.pop("newshape")
            return _orig_reshape(a, *args, **kwargs)
        reshape._b200_shim = True
        np.reshape = reshape
    import scipy.optimize
    if "scipy.optimize.nonlin" not in sys.modules:
        try:
            importlib.import_module("scipy.optimize.nonlin")
        except Exception:
            from scipy.optimize import _nonlin
            m = types.ModuleType("scipy.optimize.nonlin")
            m.TerminationCondition = _nonlin.TerminationCondition
            sys.modules["scipy.optimize.nonlin"] = m
            scipy.optimize.nonlin = m
    else:
        m = sys.modules["scipy.optimize.nonlin"]
        if not hasattr(m, "TerminationCondition"):
            from scipy.optimize import _nonlin
            m.TerminationCondition = _nonlin.TerminationCondition


def _patch_cg(op):
    """openpnm/solvers/_scipy.py:26 calls cg(A, b, tol=..., atol=...)."""
    import inspect
    import scipy.sparse.linalg as spla
    if "tol" in inspect.signature(spla.cg).parameters:
        return
    mod = op.solvers._scipy

    def cg(A, b, tol=None, **kw):
        if tol is not None:
            kw.setdefault("rtol", tol)
        return spla.cg(A, b, **kw)

    mod.cg = cg


_OP = None


def import_reference():
    """Return the reference `openpnm` module, or None when it is not present."""
    global _OP
    if _OP is not None:
        return _OP
    ref = find_reference()
    if ref is None:
        return None
    _install_docrep()
    _install_matplotlib()
    for name in ["transforms3d", "pyamg", "h5py", "skimage", "skimage.morphology",
                 "skimage.segmentation", "skimage.measure", "thermo", "chemicals",
                 "chemicals.viscosity", "chemicals.thermal_conductivity",
                 "chemicals.interface", "chemicals.vapor_pressure",
                 "chemicals.volume", "chemicals.heat_capacity",
                 "chemicals.elements", "chemicals.critical",
                 "chemicals.identifiers", "chemicals.utils",
                 "pandas", "tqdm", "tqdm.auto", "numba", "networkx", "sympy",
                 "jsonschema", "rich", "rich.logging", "flatdict", "pyevtk",
                 "traits", "porespy"]:
        _stub(name)
    _install_numpy_scipy_compat()
    if ref not in sys.path:
        sys.path.insert(0, ref)
    import logging
    op = importlib.import_module("openpnm")
    logging.getLogger().setLevel(logging.ERROR)
    _patch_cg(op)
    _OP = op
    return op
