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


def _struct_members(text, name):
    body = re.search(r"typedef struct " + name + r" \{(.*?)\} " + name + ";", text, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    out = []
    for decl in body.split(";"):
        decl = decl.strip()
        if decl:
            out += [m.strip() for m in decl.split(None, 1)[1].split(",")]
    return out


def test_bindings_agree_with_the_header():
    """The three descriptions of the boundary name the same things in the same order: include/mpasb.h, the ctypes
    mirror (dycore.Dims / dycore.Config) and the Fortran ISO_C_BINDING block of INTEGRATION.md."""
    from mpas_model_b200 import dycore
    with open(os.path.join(ROOT, "include", "mpasb.h")) as f:
        header = f.read()
    with open(os.path.join(ROOT, "INTEGRATION.md")) as f:
        integ = f.read()
    dims, cfg = _struct_members(header, "mpasb_dims"), _struct_members(header, "mpasb_config")
    assert [n for n, _ in dycore.Dims._fields_] == dims
    assert [n for n, _ in dycore.Config._fields_] == cfg
    # Fortran derived types: same members, same order
    for tname, members in (("mpasb_dims", dims), ("mpasb_config", cfg)):
        body = re.search(r"type, bind\(C\) :: " + tname + r"(.*?)end type", integ, re.S).group(1)
        body = re.sub(r"!.*", "", body)
        f_members = []
        for line in body.splitlines():
            if "::" in line:
                f_members += [m.strip() for m in line.split("::", 1)[1].split(",") if m.strip()]
        assert f_members == members, tname
    # every bound symbol exists in the header
    bound = set(re.findall(r"bind\(C, name='(mpasb_[a-z_0-9]+)'\)", integ))
    assert len(bound) >= 15 and bound <= set(_declared()), bound - set(_declared())
    # and the symbols the replacement bodies call are bound or declared
    used = set(re.findall(r"\b(mpasb_[a-z_0-9]+)\(", integ))
    assert used <= set(_declared()) | {"mpasb_binding"}, used - set(_declared())
