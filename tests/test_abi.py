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
    """The descriptions of the boundary name the same things in the same order: include/mpasb.h, the ctypes mirror
    (dycore.Dims / dycore.Config) and the committed Fortran shim (fortran/mpasb_binding.F90: ISO_C_BINDING interface of
    EVERY exported symbol; fortran/mpas_atm_dynamics_b200.F: the replacement bodies of mpas_atm_dynamics_init /
    _finalize / atm_timestep with every upload and download spelled out)."""
    from mpas_model_b200 import dycore
    from mpas_model_b200.fields import FIELDS
    with open(os.path.join(ROOT, "include", "mpasb.h")) as f:
        header = f.read()
    with open(os.path.join(ROOT, "fortran", "mpasb_binding.F90")) as f:
        binding = f.read()
    with open(os.path.join(ROOT, "fortran", "mpas_atm_dynamics_b200.F")) as f:
        shim = f.read()
    dims, cfg = _struct_members(header, "mpasb_dims"), _struct_members(header, "mpasb_config")
    assert [n for n, _ in dycore.Dims._fields_] == dims
    assert [n for n, _ in dycore.Config._fields_] == cfg
    # Fortran derived types: same members, same order
    for tname, members in (("mpasb_dims", dims), ("mpasb_config", cfg)):
        body = re.search(r"type, bind\(C\) :: " + tname + r"(.*?)end type", binding, re.S).group(1)
        body = re.sub(r"!.*", "", body)
        f_members = []
        for line in body.splitlines():
            if "::" in line:
                f_members += [m.strip() for m in line.split("::", 1)[1].split(",") if m.strip()]
        assert f_members == members, tname
    # every exported symbol is bound, nothing is bound that the header does not declare
    bound = set(re.findall(r"bind\(C, name='(mpasb_[a-z_0-9]+)'\)", binding))
    assert bound == set(_declared()), bound ^ set(_declared())
    # both RKIND widths
    assert "#ifdef SINGLE_PRECISION" in binding and "c_float" in binding and "c_double" in binding
    # the replacement bodies only call bound symbols, keep the reference's entry-point names ...
    used = set(re.findall(r"\b(mpasb_[a-z_0-9]+)\(", shim)) - {"mpasb_check", "mpasb_upload_state", "mpasb_upload_lbc", "mpasb_download_for_output", "mpasb_c_to_f"}
    assert used <= bound, used - bound
    # regional runs: the driving fields (both time levels) and the time to the end of the LBC interval reach the library
    assert "mpasb_set_lbc_time(mpasb_h" in shim and all(f"'lbc_{n}', 'lbc_{n}', {lev}," in shim
                                                        for n in ("u", "ru", "rho_zz", "rtheta_m", "scalars") for lev in (1, 2))
    for name in ("subroutine mpas_atm_dynamics_init(domain)", "subroutine mpas_atm_dynamics_finalize(domain)",
                 "subroutine atm_timestep(domain, dt, nowTime, itimestep, exchange_halo_group)"):
        assert name in shim, name
    # ... and spell out an upload for every mesh field of the table and for the prognostic state
    uploaded = set(re.findall(r"call put_[ri]\d\(\w+, '(\w+)'", shim))
    mesh_start = open(os.path.join(ROOT, "include", "mpasb_fields.def")).read().index("/* ---- mesh, real ---- */")
    mesh_keys = set(re.findall(r"^F\((\w+),", open(os.path.join(ROOT, "include", "mpasb_fields.def")).read()[mesh_start:], re.M))
    assert mesh_keys <= uploaded, mesh_keys - uploaded
    assert {"u", "w", "rho_zz", "theta_m", "scalars", "ru", "rw", "rtheta_p", "rho_p", "exner", "pressure_p"} <= uploaded
    downloaded = set(re.findall(r"call get_r\d\(\w+, '(\w+)'", shim))
    assert {"u", "w", "rho_zz", "theta_m", "scalars", "uReconstructZonal", "theta", "pressure"} <= downloaded
    assert (uploaded | downloaded) <= set(FIELDS), (uploaded | downloaded) - set(FIELDS)


def test_fortran_shim_is_generated_from_the_header_and_the_field_table():
    """fortran/*.F* are the output of tools/gen_fortran_shim.py: regenerating them changes nothing."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_fortran_shim", os.path.join(ROOT, "tools", "gen_fortran_shim.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    for name, text in {"mpasb_binding.F90": gen.wrap(gen.binding()), "mpas_atm_dynamics_b200.F": gen.wrap(gen.dynamics())}.items():
        with open(os.path.join(ROOT, "fortran", name)) as f:
            assert f.read() == text, name


def test_no_invariant_load_above_the_dependency_wait():
    """Programmatic dependent launch (mpasb_dev.cuh): before `griddepcontrol.wait` (SASS ACQBULK) a kernel may only read static
    mesh data.  A load through a `const __restrict__` pointer is an invariant load (LDG.CONSTANT) that the compiler may hoist
    above the wait -- which is how a field read once raced with the previous kernel.  No kernel of either build may have one
    there."""
    import re
    import shutil
    import subprocess
    from mpas_model_b200 import dycore
    if shutil.which("cuobjdump") is None:
        import pytest
        pytest.skip("cuobjdump not on PATH")
    for lib in (dycore.LIB_PATH, dycore.LIB_PATH_SINGLE):
        sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
        name, waited, hoisted, n_wait = None, False, {}, 0
        for line in sass.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                name, waited = m.group(1), False
            elif "ACQBULK" in line:
                n_wait += not waited
                waited = True
            elif not waited and re.search(r"LDG\S*CONSTANT", line):
                hoisted[name] = hoisted.get(name, 0) + 1
        with_wait = set()
        name = None
        for line in sass.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                name = m.group(1)
            elif "ACQBULK" in line:
                with_wait.add(name)
        assert n_wait >= 40, n_wait                                   # every column-warp kernel waits
        assert {k: v for k, v in hoisted.items() if k in with_wait} == {}
