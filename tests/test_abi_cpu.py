"""CPU: the C-ABI library loads, exports every symbol include/b200pnm.h
declares, and the product path fails loudly (never falls back) without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "b200pnm.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_loads():
    import __graft_entry__ as ge
    lib = ge.build()
    assert os.path.exists(lib)
    from openpnm_b200 import _lib
    h = _lib.load()
    assert h.b200_version() >= 100


def test_every_declared_symbol_is_exported_and_bound():
    from openpnm_b200 import _lib
    lib = C.CDLL(_lib.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b200pnm.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in openpnm_b200/_lib.py"
    assert set(_lib.SIGNATURES) <= set(names)


def test_sass_is_sm100a():
    import shutil
    import subprocess
    from openpnm_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


@pytest.mark.skipif("__import__('torch').cuda.is_available()")
def test_no_cpu_fallback():
    import openpnm_b200 as b2
    from openpnm_b200 import _lib
    h = C.c_void_p()
    assert _lib.load().b200_create(0, C.byref(h)) != 0      # error code, not a crash
    import scipy.sparse as sprs
    A = sprs.identity(4, format="coo")
    with pytest.raises(Exception):
        b2.B200PCG().solve(A, np.ones(4))
    pn = b2.Cubic([3, 3, 3])
    ph = b2.Phase(pn)
    ph["throat.diffusive_conductance"] = 1.0
    fd = b2.FickianDiffusion(network=pn, phase=ph)
    fd.set_value_BC(pores=pn.pores("left"), values=1.0)
    fd.set_value_BC(pores=pn.pores("right"), values=0.0)
    with pytest.raises(Exception):
        fd.run()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "openpnm_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
