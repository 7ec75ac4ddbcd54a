"""tests/unit/algorithms/BCTest.py of the reference restated against the host-side
mirror of `Algorithm.set_BC` (openpnm/algorithms/_algorithm.py:105-225).  Pure
bookkeeping: runs without a GPU."""
import numpy as np
import pytest

import openpnm_b200 as b2


@pytest.fixture()
def fd():
    pn = b2.Cubic(shape=[3, 3, 1], spacing=1e-4)
    air = b2.Phase(pn, name="air")
    return b2.FickianDiffusion(network=pn, phase=air)


def test_add(fd):
    """:18-32"""
    fd.set_value_BC(pores=[0, 1, 2], values=3.0, mode="add")
    mask = np.isfinite(fd["pore.bc.value"])
    assert mask.sum() == 3
    assert fd["pore.bc.value"][mask].sum() == 9.0
    assert np.isfinite(fd["pore.bc.rate"]).sum() == 0
    with pytest.raises(Exception):
        fd.set_value_BC(pores=[2, 3], values=3.0, mode="add")
    with pytest.raises(Exception):
        fd.set_rate_BC(pores=[0, 1, 2], rates=3.0, mode="add")
    fd.set_value_BC(mode="remove")
    fd.set_rate_BC(pores=[0, 1, 2], rates=3.0, mode="add")
    assert np.isfinite(fd["pore.bc.rate"]).sum() == 3


def test_overwrite(fd):
    """:34-54"""
    fd["pore.bc.rate"][1] = 1.0
    fd["pore.bc.value"][0] = 1.0
    fd.set_value_BC(pores=[1, 2], values=3.0, mode="overwrite")
    mask = np.isfinite(fd["pore.bc.value"])
    assert mask.sum() == 3
    assert fd["pore.bc.value"][mask].sum() == 7.0
    assert np.isfinite(fd["pore.bc.rate"]).sum() == 0
    fd.set_rate_BC(pores=[0, 1, 2], rates=5.0, mode="overwrite")
    mask = np.isfinite(fd["pore.bc.rate"])
    assert mask.sum() == 3
    assert fd["pore.bc.rate"][mask].sum() == 15.0
    assert np.isfinite(fd["pore.bc.value"]).sum() == 0
    fd["pore.bc.rate"][1] = 1.0
    fd["pore.bc.value"][0] = 1.0
    fd.set_BC(pores=None, bctype=[], bcvalues=np.nan, mode="overwrite")
    assert np.isfinite(fd["pore.bc.rate"]).sum() == 0
    assert np.isfinite(fd["pore.bc.value"]).sum() == 0


def _dirty(fd):
    fd["pore.bc.rate"][1] = 1.0
    fd["pore.bc.value"][0] = 1.0


def test_remove(fd):
    """:56-88"""
    _dirty(fd)
    fd.set_value_BC(pores=[0, 1], mode="remove")          # given type, given pores
    assert np.isfinite(fd["pore.bc.value"]).sum() == 0
    assert np.isfinite(fd["pore.bc.rate"]).sum() == 1
    _dirty(fd)
    fd.set_BC(pores=None, bctype="rate", mode="remove")   # given type, all pores
    assert np.isfinite(fd["pore.bc.value"]).sum() == 1
    assert np.isfinite(fd["pore.bc.rate"]).sum() == 0
    _dirty(fd)
    fd.set_BC(bctype=fd["pore.bc"].keys(), mode="remove")  # all types, all pores
    assert np.isfinite(fd["pore.bc.value"]).sum() == 0
    assert np.isfinite(fd["pore.bc.rate"]).sum() == 0
    _dirty(fd)
    fd.set_BC(mode="remove")                               # bctype == []
    assert np.isfinite(fd["pore.bc.value"]).sum() == 0
    assert np.isfinite(fd["pore.bc.rate"]).sum() == 0
    _dirty(fd)
    fd.set_BC(pores=[0, 1], mode="remove")                 # all types, given pores
    assert np.isfinite(fd["pore.bc.value"]).sum() == 0
    assert np.isfinite(fd["pore.bc.rate"]).sum() == 0


def test_modes_as_list(fd):
    """:124-131"""
    _dirty(fd)
    fd.set_BC(bctype=["value", "rate"], mode="remove")
    assert np.isfinite(fd["pore.bc.value"]).sum() == 0
    assert np.isfinite(fd["pore.bc.rate"]).sum() == 0
    fd.set_value_BC(pores=[0, 1], values=1.0, mode=["remove", "add"])
    assert np.isfinite(fd["pore.bc.value"]).sum() == 2
