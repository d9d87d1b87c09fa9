"""The C-ABI library loads and exports every symbol include/mpasb.h declares; without a
GPU it refuses to create a handle (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    with open(os.path.join(ROOT, "include", "mpasb.h")) as f:
        return sorted(set(re.findall(r"\b(mpasb_[a-z_0-9]+)\s*\(", f.read())))


def test_library_exports_every_declared_symbol():
    from mpas_model_b200 import dycore
    if not os.path.exists(dycore.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(dycore.LIB_PATH)
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), n


def test_field_table_is_shared_and_consistent():
    from mpas_model_b200.fields import FIELDS
    assert {"u", "w", "rho_zz", "theta_m", "scalars"} <= {n for n, f in FIELDS.items() if f.levels == 2}
    assert FIELDS["zb_cell"].inner == "NL1_ME" and FIELDS["cellsOnEdge"].target == "CELL"


def test_create_fails_loudly_without_gpu(tiny_case):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from mpas_model_b200.dycore import Dycore
    d, cfg = tiny_case
    with pytest.raises(RuntimeError):
        Dycore(d, cfg)
