"""Host-side mirror of the reference's time-integration interface.

``Dycore`` binds the C ABI of ``csrc/libmpasb.so`` (include/mpasb.h) with ctypes
and exposes the reference's entry points under their own names:

  atm_timestep(dt, itimestep)          mpas_atm_time_integration.F:739
  atm_srk3(dt, itimestep)              ... :803
  atm_init_coupled_diagnostics()       ... :6776
  atm_compute_solve_diagnostics(...)   ... :6243
  mpas_pool_shift_time_levels()        mpas_atm_core.F:808
  exchange_halo_group(name)            mpas_atm_halos.F:29
  get_array(name, time_level) / set_array(...)   == mpas_pool_get_array

There is no CPU fallback: constructing a ``Dycore`` without the CUDA library
(or calling into it without a GPU) raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .fields import FIELDS, host_shape

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MPASB_LIB") or os.path.join(_HERE, "csrc", "libmpasb.so")
LIB_PATH_SINGLE = os.path.join(_HERE, "csrc", "libmpasb_sp.so")      # PRECISION=single build (RKIND = float)


class Dims(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "nCells", "nEdges", "nVertices", "nCellsSolve", "nEdgesSolve", "nVerticesSolve",
        "nVertLevels", "maxEdges", "maxEdges2", "vertexDegree",
        "num_scalars", "index_qv", "moist_start", "moist_end")]


_CFG_INT = ("config_time_integration_order", "config_number_of_sub_steps", "config_dynamics_split_steps",
            "config_split_dynamics_transport", "config_scalar_advection", "config_monotonic",
            "config_positive_definite", "config_horiz_mixing", "config_mix_full", "config_rayleigh_damp_u",
            "config_number_rayleigh_damp_u_levels", "config_number_cam_damping_levels", "config_apply_lbcs",
            "config_print_global_minmax_vel")
_CFG_REAL = ("config_epssm", "config_smdiv", "config_len_disp", "config_coef_3rd_order",
             "config_visc4_2dsmag", "config_smagorinsky_coef", "config_del4u_div_factor",
             "config_h_mom_eddy_visc2", "config_h_mom_eddy_visc4", "config_v_mom_eddy_visc2",
             "config_h_theta_eddy_visc2", "config_h_theta_eddy_visc4", "config_v_theta_eddy_visc2",
             "config_apvm_upwinding", "config_mpas_cam_coef", "config_rayleigh_damp_u_timescale_days",
             "config_relax_zone_divdamp_coef", "cf1", "cf2", "cf3", "sphere_radius")


class Config(C.Structure):
    _fields_ = [(n, C.c_int) for n in _CFG_INT] + [(n, C.c_double) for n in _CFG_REAL] + [("on_a_sphere", C.c_int)]


def make_dims(d: dict) -> Dims:
    """Pool dimensions of a block dict (0-based scalar indices -> 1-based)."""
    return Dims(
        nCells=d["nCells"], nEdges=d["nEdges"], nVertices=d["nVertices"],
        nCellsSolve=d.get("nCellsSolve", d["nCells"]), nEdgesSolve=d.get("nEdgesSolve", d["nEdges"]),
        nVerticesSolve=d.get("nVerticesSolve", d["nVertices"]),
        nVertLevels=d["nVertLevels"], maxEdges=d["maxEdges"], maxEdges2=d["maxEdges2"],
        vertexDegree=d["vertexDegree"], num_scalars=d["num_scalars"], index_qv=d["index_qv"] + 1,
        moist_start=d["moist_start"] + 1, moist_end=d["moist_end"] + 1)


def make_config(cfg: dict, d: dict) -> Config:
    c = Config()
    for n in _CFG_INT:
        v = cfg.get(n, 0)
        if n == "config_horiz_mixing":
            v = {"2d_smagorinsky": 0, "2d_fixed": 1}[v]
        setattr(c, n, int(v))
    for n in _CFG_REAL:
        setattr(c, n, float(d[n] if n in ("cf1", "cf2", "cf3", "sphere_radius") else cfg.get(n, 0.0)))
    c.config_print_global_minmax_vel = 1
    c.on_a_sphere = int(cfg.get("on_a_sphere", 1))
    return c


class Backend:
    """Common host logic of the CUDA library binding and the oracle binding:
    moving a block dict (numpy, 0-based) through the by-name field ABI."""

    dims: Dims
    rdtype = np.float64          # RKIND of the bound library
    creal = C.c_double

    # -- to be provided by the concrete binding
    def _set_real(self, name, lev, arr): raise NotImplementedError
    def _get_real(self, name, lev, out): raise NotImplementedError
    def _set_int(self, name, arr): raise NotImplementedError

    def shape(self, name):
        return host_shape(FIELDS[name], self.dims)

    def set_array(self, name, arr, time_level=1):
        fd = FIELDS[name]
        shp = self.shape(name)
        if fd.type == "INT":
            a = np.ascontiguousarray(arr, dtype=np.int32).reshape(shp)
            if fd.target != "NONE":
                a = a + 1                     # the ABI is 1-based, like the Fortran pools
            self._set_int(name, np.ascontiguousarray(a))
        else:
            a = np.ascontiguousarray(arr, dtype=self.rdtype).reshape(shp)
            self._set_real(name, time_level, a)

    def get_array(self, name, time_level=1):
        out = np.empty(self.shape(name), dtype=self.rdtype)
        self._get_real(name, time_level, out)
        return out

    def load_block(self, d: dict):
        """Import every table field present in ``d``.  State fields go to time
        level 1 (and, for ``u``/``w``/..., ``<name>_2`` to level 2 if given)."""
        for name, fd in FIELDS.items():
            if name in d and not np.isscalar(d[name]):
                self.set_array(name, d[name], 1)
            if fd.levels == 2 and (name + "_2") in d:
                self.set_array(name, d[name + "_2"], 2)

    @staticmethod
    def _flatten_halo_lists(lists):
        """decomp.exchange_lists()[rank][kind] -> the flat 1-based arrays of mpasb_set_halo_lists."""
        nbrs = np.asarray(lists["neighbors"], dtype=np.int32)
        nl = lists["n_layers"]
        n_send = np.array([[len(lists["send"][i][l]) for l in range(nl)] for i in range(len(nbrs))], dtype=np.int32).reshape(-1)
        n_recv = np.array([[len(lists["recv"][i][l]) for l in range(nl)] for i in range(len(nbrs))], dtype=np.int32).reshape(-1)

        def cat(key):
            parts = [np.asarray(a, dtype=np.int32) for per in lists[key] for a in per] + [np.zeros(0, np.int32)]
            return np.ascontiguousarray(np.concatenate(parts) + 1, dtype=np.int32)
        return nbrs, nl, np.ascontiguousarray(n_send), cat("send"), np.ascontiguousarray(n_recv), cat("recv")

    def state(self, time_level=1, names=("u", "w", "rho_zz", "theta_m", "scalars")):
        return {n: self.get_array(n, time_level) for n in names}


def _load_lib(path=None):
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError(
            f"CUDA library {path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the dycore step)")
    lib = C.CDLL(path)
    lib.mpasb_last_error.restype = C.c_char_p
    lib.mpasb_last_error.argtypes = [C.c_void_p]
    lib.mpasb_kernel_launch_count.restype = C.c_long
    lib.mpasb_kernel_launch_count.argtypes = [C.c_void_p]
    return lib


# ---- atm_mpas_init_block, mesh part, in the library (mpasb_init_block / mpasb_init_block_host) ----
INIT_BLOCK_IN_INT = ("nEdgesOnCell", "edgesOnCell", "cellsOnCell", "verticesOnCell", "cellsOnEdge", "verticesOnEdge",
                     "edgesOnVertex", "cellsOnVertex")
INIT_BLOCK_IN_REAL = ("zb", "zb3", "deriv_two", "dcEdge", "dvEdge", "areaCell", "areaTriangle", "meshDensity", "zgrid")
INIT_BLOCK_IN_COORDS = ("xCell", "yCell", "zCell", "xEdge", "yEdge", "zEdge")     # optional: with them coeffs_reconstruct is derived too
INIT_BLOCK_OUT_REAL = ("edgesOnVertex_sign", "edgesOnCell_sign", "zb_cell", "zb3_cell", "invAreaCell", "invDvEdge", "invDcEdge",
                       "invAreaTriangle", "adv_coefs", "adv_coefs_3rd", "meshScalingDel2", "meshScalingDel4",
                       "meshScalingRegionalCell", "meshScalingRegionalEdge", "dss")
INIT_BLOCK_OUT_INT = ("kiteForCell", "nAdvCellsForEdge", "advCellsForEdge")
_INIT_BLOCK_ONE_BASED = {"edgesOnCell", "cellsOnCell", "verticesOnCell", "cellsOnEdge", "verticesOnEdge", "edgesOnVertex",
                         "cellsOnVertex", "kiteForCell", "advCellsForEdge"}


def _init_block_inputs(d, rdtype):
    """The raw mesh fields of a block dict in the layout of the ABI: dense, 1-based connectivity (the dict is 0-based)."""
    keep, names, ptrs = [], [], []
    coords = INIT_BLOCK_IN_COORDS if all(n in d for n in INIT_BLOCK_IN_COORDS) else ()
    for n in INIT_BLOCK_IN_INT + INIT_BLOCK_IN_REAL + coords:
        if n in INIT_BLOCK_IN_INT:
            a = np.ascontiguousarray(d[n], dtype=np.int32)
            if n in _INIT_BLOCK_ONE_BASED:
                a = np.ascontiguousarray(a + 1)
        else:
            a = np.ascontiguousarray(d[n], dtype=rdtype)
        keep.append(a); names.append(n.encode()); ptrs.append(a.ctypes.data)
    k = len(names)
    return keep, k, (C.c_char_p * k)(*names), (C.c_void_p * k)(*ptrs)


def init_block_host(d: dict, cfg: dict, precision: str = "double") -> dict:
    """mpasb_init_block_host: the derived mesh fields of ``init_block.init_block`` computed by the library's C++ host code (no
    device, no handle); returns them 0-based, shaped like the block dict's."""
    lib = _load_lib(LIB_PATH_SINGLE if precision == "single" else None)
    rdtype = np.float32 if lib.mpasb_real_bytes() == 4 else np.float64
    dims, config = make_dims(d), make_config(cfg, d)
    keep, k, names, ptrs = _init_block_inputs(d, rdtype)
    nC, nE, nV, mx, vd, nz = d["nCells"], d["nEdges"], d["nVertices"], d["maxEdges"], d["vertexDegree"], d["nVertLevels"]
    shapes = dict(edgesOnVertex_sign=(nV + 1, vd), edgesOnCell_sign=(nC + 1, mx), zb_cell=(nC + 1, mx, nz + 1), zb3_cell=(nC + 1, mx, nz + 1),
                  invAreaCell=(nC + 1,), invDvEdge=(nE + 1,), invDcEdge=(nE + 1,), invAreaTriangle=(nV + 1,), adv_coefs=(nE + 1, 15),
                  adv_coefs_3rd=(nE + 1, 15), meshScalingDel2=(nE + 1,), meshScalingDel4=(nE + 1,), meshScalingRegionalCell=(nC + 1,),
                  meshScalingRegionalEdge=(nE + 1,), dss=(nC + 1, nz), kiteForCell=(nC + 1, mx), nAdvCellsForEdge=(nE + 1,),
                  advCellsForEdge=(nE + 1, 15), coeffs_reconstruct=(nC + 1, mx, 3))
    want = INIT_BLOCK_OUT_REAL + INIT_BLOCK_OUT_INT
    if all(n in d for n in INIT_BLOCK_IN_COORDS) and cfg.get("on_a_sphere", True):
        want = want + ("coeffs_reconstruct",)            # mpas_init_reconstruct (mpas_vector_reconstruction.F:60-177)
    out = {n: np.zeros(shapes[n], dtype=np.int32 if n in INIT_BLOCK_OUT_INT else rdtype) for n in want}
    m = len(out)
    onames = (C.c_char_p * m)(*[n.encode() for n in out])
    optrs = (C.c_void_p * m)(*[a.ctypes.data for a in out.values()])
    rc = lib.mpasb_init_block_host(C.byref(dims), C.byref(config), C.c_int(int(cfg["config_h_ScaleWithMesh"])), C.c_double(cfg["config_zd"]),
                                   C.c_double(cfg["config_xnutr"]), C.c_int(k), names, ptrs, C.c_int(m), onames, optrs)
    if rc != 0:
        raise RuntimeError(f"mpasb_init_block_host failed ({rc})")
    for n in out:
        if n in _INIT_BLOCK_ONE_BASED:
            out[n] = out[n] - 1
    return out


class Dycore(Backend):
    """One mesh block resident on one GPU."""

    def __init__(self, block: dict, cfg: dict, device: int = 0, precision: str = "double"):
        self.lib = _load_lib(LIB_PATH_SINGLE if precision == "single" else None)
        if self.lib.mpasb_real_bytes() == 4:
            self.rdtype, self.creal = np.float32, C.c_float
        self.dims = make_dims(block)
        self.config = make_config(cfg, block)
        self.cfg = dict(cfg)
        self._h = C.c_void_p()
        self._strict = bool(self.lib.mpasb_strict_arithmetic())      # MPASB_STRICT is read when the handle is created
        rc = self.lib.mpasb_create(C.byref(self.dims), C.byref(self.config), C.c_int(device), C.byref(self._h))
        if rc != 0:
            raise RuntimeError(f"mpasb_create failed ({rc}): no usable CUDA device or bad dimensions")
        self.load_block(block)

    # -- low level
    def _check(self, rc, what):
        if rc != 0:
            msg = self.lib.mpasb_last_error(self._h)
            raise RuntimeError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def _set_real(self, name, lev, arr):
        self._check(self.lib.mpasb_set_field(self._h, name.encode(), C.c_int(lev),
                                             arr.ctypes.data_as(C.c_void_p), C.c_long(arr.size)), f"set_field {name}")

    def _get_real(self, name, lev, out):
        self._check(self.lib.mpasb_get_field(self._h, name.encode(), C.c_int(lev),
                                             out.ctypes.data_as(C.c_void_p), C.c_long(out.size)), f"get_field {name}")

    def _set_int(self, name, arr):
        self._check(self.lib.mpasb_set_field_int(self._h, name.encode(),
                                                 arr.ctypes.data_as(C.c_void_p), C.c_long(arr.size)), f"set_field_int {name}")

    # -- batched, asynchronous transfers (mpasb_set_fields_async / mpasb_get_fields_async)
    def _batch(self, items):
        n = len(items)
        names = (C.c_char_p * n)(*[nm.encode() for nm, _, _ in items])
        levels = (C.c_int * n)(*[lev for _, lev, _ in items])
        ptrs = (C.c_void_p * n)(*[a.ctypes.data for _, _, a in items])
        counts = (C.c_long * n)(*[a.size for _, _, a in items])
        for _, _, a in items:
            assert a.dtype == self.rdtype and a.flags["C_CONTIGUOUS"]
        return n, names, levels, ptrs, counts

    def set_fields_async(self, items):
        """items: [(pool key, time level, host array)].  Enqueue the uploads; the arrays may be reused after wait_uploads()."""
        n, names, levels, ptrs, counts = self._batch(items)
        self._check(self.lib.mpasb_set_fields_async(self._h, C.c_int(n), names, levels, ptrs, counts), "set_fields_async")

    def get_fields_async(self, items):
        """Enqueue the downloads behind everything enqueued so far; the arrays are valid after wait_downloads()."""
        n, names, levels, ptrs, counts = self._batch(items)
        self._check(self.lib.mpasb_get_fields_async(self._h, C.c_int(n), names, levels, ptrs, counts), "get_fields_async")

    def wait_uploads(self, lag=0):
        self._check(self.lib.mpasb_wait_uploads_lag(self._h, C.c_int(lag)), "wait_uploads")

    def wait_downloads(self, lag=0):
        self._check(self.lib.mpasb_wait_downloads_lag(self._h, C.c_int(lag)), "wait_downloads")

    def atm_init_solve_diagnostics_async(self, dt):
        self._check(self.lib.mpasb_init_solve_diagnostics_async(self._h, self.creal(dt)), "init_solve_diagnostics_async")

    def close(self):
        if self._h:
            self.lib.mpasb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- reference entry points
    def atm_mpas_init_block(self, d: dict, cfg: dict):
        """mpas_atm_core.F:368-602, mesh part: the library derives the row-M fields from the raw mesh fields of ``d`` and keeps
        them in the handle (mpasb_init_block)."""
        keep, k, names, ptrs = _init_block_inputs(d, self.rdtype)
        self._check(self.lib.mpasb_init_block(self._h, C.c_int(int(cfg["config_h_ScaleWithMesh"])), C.c_double(cfg["config_zd"]),
                                              C.c_double(cfg["config_xnutr"]), C.c_int(k), names, ptrs), "mpasb_init_block")

    def atm_init_coupled_diagnostics(self):
        self._check(self.lib.mpasb_init_coupled_diagnostics(self._h), "init_coupled_diagnostics")

    def atm_init_solve_diagnostics(self, dt):
        self._check(self.lib.mpasb_init_solve_diagnostics(self._h, self.creal(dt)), "init_solve_diagnostics")

    def atm_srk3(self, dt, itimestep=1):
        self._check(self.lib.mpasb_step(self._h, self.creal(dt), C.c_int(itimestep)), "mpasb_step")

    atm_timestep = atm_srk3      # config_time_integration == 'SRK3' is the only integrator (TI:773-790)

    def mpas_pool_shift_time_levels(self):
        self._check(self.lib.mpasb_shift_time_levels(self._h), "shift_time_levels")

    def mpas_reconstruct(self, time_level=1, include_halos=False):
        """mpas_vector_reconstruction.F:205 on ``u`` of ``time_level`` -> ``uReconstruct*``."""
        self._check(self.lib.mpasb_reconstruct(self._h, C.c_int(time_level), C.c_int(int(include_halos))), "mpas_reconstruct")

    def atm_compute_output_diagnostics(self, time_level=1):
        """mpas_atm_core.F:901: ``theta``, ``rho``, ``pressure`` of ``time_level``."""
        self._check(self.lib.mpasb_compute_output_diagnostics(self._h, C.c_int(time_level)), "atm_compute_output_diagnostics")

    def summarize_timestep(self):
        out = (self.creal * 4)()
        self._check(self.lib.mpasb_minmax(self._h, out), "minmax")
        return tuple(out)

    def summarize_timestep_async(self):
        """Enqueue the step summary behind the step; the host does not wait."""
        self._check(self.lib.mpasb_summarize_timestep_async(self._h), "summarize_timestep_async")

    def summarize_timestep_fetch(self, scalars=True):
        """-> (minmax array {min w, max w, min u, max u[, min s1, max s1, ...]}, (NaNs in w, NaNs in u))."""
        n = 2 * (2 + (self.dims.num_scalars if scalars else 0))
        out = (self.creal * n)()
        nan = (C.c_long * 2)()
        self._check(self.lib.mpasb_summarize_timestep_fetch(self._h, out, C.c_long(n), nan), "summarize_timestep_fetch")
        return np.array(out[:], dtype=np.float64), (int(nan[0]), int(nan[1]))

    def synchronize(self):
        self._check(self.lib.mpasb_synchronize(self._h), "synchronize")

    def exchange_halo_group(self, name):
        self._check(self.lib.mpasb_exchange_halo_group(self._h, name.encode()), f"exchange {name}")

    def exchange_halo_group_async(self, name):
        """the same without waiting for it on the host"""
        self._check(self.lib.mpasb_exchange_halo_group_async(self._h, name.encode()), f"exchange {name}")

    def set_halo_lists(self, kind, lists):
        """kind: 0 cells, 1 edges, 2 vertices; lists: decomp.exchange_lists()[rank][kind name]."""
        nbrs, nl, n_send, ss, n_recv, rr = self._flatten_halo_lists(lists)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        self._check(self.lib.mpasb_set_halo_lists(self._h, C.c_int(kind), C.c_int(len(nbrs)), p(nbrs), C.c_int(nl),
                                                  p(n_send), p(ss), p(n_recv), p(rr)), "set_halo_lists")

    def comm_init(self, rank, world_size, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        self._check(self.lib.mpasb_comm_init(self._h, C.c_int(rank), C.c_int(world_size), buf), "comm_init")

    def p2p_init(self, dist) -> bool:
        """Switch the halo exchanges of this handle to direct NVLink stores (CUDA IPC); ``dist`` is the host process
        group used to agree on the mailbox size and to all-gather the IPC handles.  The switch is collective: if any
        rank cannot export or map the buffers, every rank stays on NCCL send/recv.  Returns whether it is on."""
        world = dist.get_world_size()

        def all_ok(ok):
            flags = [None] * world
            dist.all_gather_object(flags, bool(ok))
            return all(flags)

        self.lib.mpasb_p2p_max_message.restype = C.c_long
        n = int(self.lib.mpasb_p2p_max_message(self._h))
        sizes = [None] * world
        dist.all_gather_object(sizes, n)
        if min(sizes) < 0:
            return False
        buf = C.create_string_buffer(128)
        rc = self.lib.mpasb_p2p_prepare(self._h, C.c_long(max(max(sizes), 2)), buf)
        if not all_ok(rc == 0):
            return False
        handles = [None] * world
        dist.all_gather_object(handles, buf.raw)
        allh = C.create_string_buffer(b"".join(handles), 128 * world)
        rc = self.lib.mpasb_p2p_open(self._h, allh)
        if not all_ok(rc == 0):
            return False
        self._check(self.lib.mpasb_p2p_enable(self._h, C.c_int(1)), "p2p_enable")
        dist.barrier()
        return True

    def nccl_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        if self.lib.mpasb_get_nccl_unique_id(buf) != 0:
            raise RuntimeError("mpasb_get_nccl_unique_id failed (libnccl.so.2 not found?)")
        return buf.raw

    def timer_start(self):
        self._check(self.lib.mpasb_timer_start(self._h), "timer_start")

    def timer_stop(self):
        ms = C.c_double()
        self._check(self.lib.mpasb_timer_stop(self._h, C.byref(ms)), "timer_stop")
        return ms.value

    def strict_arithmetic(self) -> bool:
        """True if this handle keeps the reference's operation order everywhere (MPASB_STRICT=1 at creation: results are
        bit-identical to the fp64 CPU arithmetic); False for the default, relaxed mode (re-associated stencil sums)."""
        return self._strict

    def kernel_launch_count(self):
        return int(self.lib.mpasb_kernel_launch_count(self._h))

    def set_profile(self, on=True):
        self.lib.mpasb_set_profile(self._h, C.c_int(1 if on else 0))

    def get_profile(self):
        buf = C.create_string_buffer(1 << 16)
        self.lib.mpasb_get_profile(self._h, buf, C.c_long(len(buf)))
        rows = []
        for line in buf.value.decode().splitlines():
            n, ms, cnt = line.split()
            rows.append((n, float(ms), int(cnt)))
        return rows

    # -- one *_work routine at a time (parity tests)
    def k(self, routine, *args):
        if routine.startswith("lbc_"):            # regional path: mpasb_k_lbc(handle, routine, a, b)
            ra = [float(a) for a in args] + [0.0, 0.0]
            self._check(self.lib.mpasb_k_lbc(self._h, routine[4:].encode(), self.creal(ra[0]), self.creal(ra[1])), routine)
            return
        fn = getattr(self.lib, "mpasb_k_" + routine)
        cargs = [self.creal(a) if isinstance(a, float) else C.c_int(a) for a in args]
        self._check(fn(self._h, *cargs), routine)

    def set_lbc_time(self, seconds_to_interval_end: float):
        """Regional runs: LBC_intv_end - currTime at the start of the next step (mpas_atm_boundaries.F:497-503)."""
        self._check(self.lib.mpasb_set_lbc_time(self._h, self.creal(seconds_to_interval_end)), "set_lbc_time")
