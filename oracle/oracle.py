"""ctypes binding of the CPU parity oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY -- see the header of dycore_oracle.cpp.  Imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never by the package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from mpas_model_b200.dycore import Backend, make_config, make_dims

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
LIB_PATH_SINGLE = os.path.join(_HERE, "liboracle_sp.so")      # PRECISION=single restatement (RKIND = float)


def build(force=False):
    src = os.path.join(_HERE, "dycore_oracle.cpp")
    if force or any(not os.path.exists(p) or os.path.getmtime(p) < os.path.getmtime(src) for p in (LIB_PATH, LIB_PATH_SINGLE)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return LIB_PATH


_libs = {}


def lib(precision="double"):
    if precision not in _libs:
        build()
        l = C.CDLL(LIB_PATH_SINGLE if precision == "single" else LIB_PATH)
        l.oracle_create.restype = C.c_void_p
        l.oracle_field_count.restype = C.c_long
        assert l.oracle_real_bytes() == (4 if precision == "single" else 8)
        _libs[precision] = l
    return _libs[precision]


def set_threads(n: int) -> None:
    """OpenMP threads of the oracle from here on (omp_set_num_threads of the libgomp it is linked against).  torchrun
    exports OMP_NUM_THREADS=1 into every rank, which would turn the multi-core CPU baseline into a single-core one."""
    lib()
    try:
        C.CDLL("libgomp.so.1").omp_set_num_threads(C.c_int(max(1, int(n))))
    except OSError:
        pass


class OracleDycore(Backend):
    def __init__(self, block: dict, cfg: dict, rank: int = 0, precision: str = "double"):
        self.lib = lib(precision)
        self.precision = precision
        if precision == "single":
            self.rdtype, self.creal = np.float32, C.c_float
        self.dims = make_dims(block)
        self.config = make_config(cfg, block)
        self._h = C.c_void_p(self.lib.oracle_create(C.byref(self.dims), C.byref(self.config), C.c_int(rank)))
        self.load_block(block)

    def _set_real(self, name, lev, arr):
        rc = self.lib.oracle_set_field(self._h, name.encode(), C.c_int(lev), arr.ctypes.data_as(C.c_void_p), C.c_long(arr.size))
        assert rc == 0, name

    def _get_real(self, name, lev, out):
        rc = self.lib.oracle_get_field(self._h, name.encode(), C.c_int(lev), out.ctypes.data_as(C.c_void_p), C.c_long(out.size))
        assert rc == 0, name

    def _set_int(self, name, arr):
        rc = self.lib.oracle_set_field_int(self._h, name.encode(), arr.ctypes.data_as(C.c_void_p), C.c_long(arr.size))
        assert rc == 0, name

    def close(self):
        if self._h:
            self.lib.oracle_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_halo_lists(self, kind, lists):
        nbrs, nl, n_send, ss, n_recv, rr = self._flatten_halo_lists(lists)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        self.lib.oracle_set_halo_lists(self._h, C.c_int(kind), C.c_int(len(nbrs)), p(nbrs), C.c_int(nl),
                                       p(n_send), p(ss), p(n_recv), p(rr))

    # reference entry points
    def atm_init_coupled_diagnostics(self): self.lib.oracle_init_coupled_diagnostics(self._h)
    def atm_init_solve_diagnostics(self, dt): self.lib.oracle_init_solve_diagnostics(self._h, C.c_double(dt))
    def atm_srk3(self, dt, itimestep=1): step([self], dt)
    atm_timestep = atm_srk3
    def mpas_pool_shift_time_levels(self): self.lib.oracle_shift_time_levels(self._h)
    def mpas_reconstruct(self, time_level=1, include_halos=False):
        self.lib.oracle_reconstruct(self._h, C.c_int(time_level), C.c_int(int(include_halos)))
    def atm_compute_output_diagnostics(self, time_level=1):
        self.lib.oracle_compute_output_diagnostics(self._h, C.c_int(time_level))

    def summarize_timestep(self):
        out = (C.c_double * 4)()
        self.lib.oracle_minmax(self._h, out)
        return tuple(out)

    def k(self, routine, *args):
        fn = getattr(self.lib, "oracle_" + routine)
        fn(self._h, *[C.c_double(a) if isinstance(a, float) else C.c_int(a) for a in args])


def step(blocks, dt):
    """atm_srk3 over N in-process blocks ("virtual ranks") in lock step."""
    arr = (C.c_void_p * len(blocks))(*[b._h for b in blocks])
    blocks[0].lib.oracle_step(arr, C.c_int(len(blocks)), C.c_double(dt))


def exchange(blocks, group):
    arr = (C.c_void_p * len(blocks))(*[b._h for b in blocks])
    blocks[0].lib.oracle_exchange(arr, C.c_int(len(blocks)), group.encode())
