"""ctypes binding of oracle/_ref/libtiref.so: the reference's own dycore routines, transliterated from
mpas_atm_time_integration.F to C++ by oracle/f2cpp.py and compiled here (see ref_harness.cpp).

TEST INFRASTRUCTURE ONLY.  ``RefDycore`` has the interface of ``OracleDycore`` / ``Dycore`` (set_array / get_array by
pool key, ``k(routine, ...)`` per *_work routine) so the parity helpers of tests/util.py drive it unchanged; its
arrays are numpy buffers in the reference's own host layout, (nVertLevels[+1], n+1) column-major with 1-based indices,
registered by address in the pools the translated wrappers look them up in.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from mpas_model_b200.dycore import Backend, make_dims
from mpas_model_b200.fields import FIELDS

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
LIB = {"double": os.path.join(REF_DIR, "libtiref.so"), "single": os.path.join(REF_DIR, "libtiref_sp.so")}
REFERENCE = os.environ.get("MPAS_REFERENCE", "/root/reference")

# Registry.xml var_struct "tend": name_in_code -> this repo's field key (src/core_atmosphere/Registry.xml:1795-1830)
TEND_KEYS = {"u": "tend_u", "w": "tend_w", "theta_m": "tend_theta", "rho_zz": "tend_rho", "rt_diabatic_tend": "rt_diabatic_tend",
             "u_euler": "tend_u_euler", "w_euler": "tend_w_euler", "theta_euler": "tend_theta_euler", "scalars_tend": "scalars_tend",
             "rthdynten": "rthdynten"}
# module variables of atm_time_integration (TI:90-140) -> field key
MODULE_KEYS = {"tend_ru_physics": "tend_ru_physics", "tend_rtheta_physics": "tend_rtheta_physics", "tend_rho_physics": "tend_rho_physics",
               "qtot": "qtot", "delsq_theta": "delsq_theta", "delsq_w": "delsq_w", "delsq_divergence": "delsq_divergence",
               "delsq_u": "delsq_u", "delsq_vorticity": "delsq_vorticity", "dpdz": "dpdz", "horiz_flux_array": "horiz_flux_arr",
               "scalar_old_arr": "scalar_old", "scalar_new_arr": "scalar_new", "s_max_arr": "s_max", "s_min_arr": "s_min",
               "flux_array": "flux_arr", "flux_upwind_tmp_arr": "flux_upwind_tmp", "flux_tmp_arr": "flux_tmp", "wdtn_arr": "wdtn",
               "rho_zz_int": "rho_zz_int", "ke_vertex": "ke_vertex", "ke_edge": "ke_edge", "bdyMaskEdge": "bdyMaskEdge"}


def available():
    return os.path.exists(LIB["double"])


def build(force=False):
    """Generate and compile oracle/_ref (only where the reference tree exists: the build container)."""
    if not os.path.isdir(REFERENCE):
        return available()
    srcs = [os.path.join(_HERE, n) for n in ("f2cpp.py", "f2cpp_rt.h", "ref_harness.cpp")] + \
           [os.path.join(REFERENCE, "src/core_atmosphere/dynamics/mpas_atm_time_integration.F")]
    newest = max(os.path.getmtime(p) for p in srcs)
    if force or any(not os.path.exists(p) or os.path.getmtime(p) < newest for p in LIB.values()):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
    return available()


class RefDycore(Backend):
    def __init__(self, block: dict, cfg: dict, precision: str = "double", threads: int = 1):
        build()
        if not os.path.exists(LIB[precision]):
            raise RuntimeError("oracle/_ref is not built (needs /root/reference at build time)")
        self.lib = C.CDLL(LIB[precision])
        self.lib.ref_create.restype = C.c_void_p
        self.lib.ref_exchange_log.restype = C.c_char_p
        if precision == "single":
            self.rdtype, self.creal = np.float32, C.c_float
        assert self.lib.ref_real_bytes() == np.dtype(self.rdtype).itemsize
        self.dims = make_dims(block)
        self.cfg = dict(cfg)
        self._h = C.c_void_p(self.lib.ref_create())
        d = self.dims
        for n in ("nCells", "nEdges", "nVertices", "nCellsSolve", "nEdgesSolve", "nVerticesSolve", "nVertLevels", "maxEdges",
                  "maxEdges2", "vertexDegree", "num_scalars", "index_qv", "moist_start", "moist_end"):
            self.lib.ref_set_dim(self._h, n.encode(), C.c_int(getattr(d, n)))
        self.lib.ref_set_dim(self._h, b"nVertLevelsP1", C.c_int(d.nVertLevels + 1))
        for k, v in cfg.items():
            if isinstance(v, str):
                self.lib.ref_set_cfg_str(self._h, k.encode(), v.encode())
            elif isinstance(v, (bool, int, np.integer)):
                self.lib.ref_set_cfg_int(self._h, k.encode(), C.c_int(int(v)))
            else:
                self.lib.ref_set_cfg_real(self._h, k.encode(), C.c_double(float(v)))
        self.lib.ref_set_cfg_real(self._h, b"sphere_radius", C.c_double(float(block["sphere_radius"])))
        self.lib.ref_set_cfg_int(self._h, b"config_apply_lbcs", C.c_int(int(cfg.get("config_apply_lbcs", 0))))
        self.set_lbc_time(0.0)
        # host arrays in the reference's layout
        self.a = {}
        for name, fd in FIELDS.items():
            dt = np.int32 if fd.type == "INT" else self.rdtype
            for lev in range(1, fd.levels + 1):
                self.a[(name, lev)] = np.zeros(self.shape(name), dtype=dt)
        nl, nC, nE = d.nVertLevels, d.nCells, d.nEdges
        self.extra = {                         # looked up by the wrappers, read only by code paths that are compiled out or off
            ("mesh", "cf1"): np.array([block["cf1"]], dtype=self.rdtype), ("mesh", "cf2"): np.array([block["cf2"]], dtype=self.rdtype),
            ("mesh", "cf3"): np.array([block["cf3"]], dtype=self.rdtype),
            ("mesh", "deriv_two"): np.zeros((nE + 1, 2, 15), dtype=self.rdtype), ("mesh", "latEdge"): np.zeros(nE + 1, dtype=self.rdtype),
            ("mesh", "qv_init"): np.zeros(nl, dtype=self.rdtype),
            ("mesh", "zb"): np.zeros((nE + 1, 2, nl + 1), dtype=self.rdtype), ("mesh", "zb3"): np.zeros((nE + 1, 2, nl + 1), dtype=self.rdtype),
            ("tend", "w_pgf"): np.zeros((nC + 1, nl + 1), dtype=self.rdtype), ("tend", "w_buoy"): np.zeros((nC + 1, nl + 1), dtype=self.rdtype),
            # inputs / outputs of the init-time routines of mpas_atm_core.F (atm_compute_mesh_scaling, ...)
            ("mesh", "meshDensity"): np.ones(nC + 1, dtype=self.rdtype),
            ("mesh", "nearestRelaxationCell"): np.zeros(nC + 1, dtype=np.int32),
        }
        for k in ("zb", "zb3", "deriv_two", "meshDensity"):
            if k in block and np.shape(block[k]) == self.extra[("mesh", k)].shape:
                self.extra[("mesh", k)][...] = block[k]
        self._bind_all()
        self.load_block(block)
        self.set_threads(threads)

    def set_lbc_time(self, seconds_to_interval_end: float):
        """LBC_intv_end - currTime at the start of the next step (mpas_atm_boundaries.F:497-503)."""
        self.lib.ref_set_cfg_real(self._h, b"lbc_dt_end", C.c_double(float(seconds_to_interval_end)))

    def set_threads(self, n: int):
        """OpenMP threads, each with the index ranges of mpas_atm_threading.F:100-111 (the reference's MPAS_OPENMP build)."""
        self.lib.ref_set_threads(self._h, C.c_int(int(n)))

    def _bind(self, pool, key, lev, arr):
        shp = arr.shape[::-1]                  # numpy C order -> Fortran extents, fastest first
        n = list(shp) + [1] * (3 - len(shp))
        rc = self.lib.ref_bind_array(self._h, pool.encode(), key.encode(), C.c_int(lev), arr.ctypes.data_as(C.c_void_p),
                                     C.c_int(int(arr.dtype == np.int32)), C.c_int(len(shp)), C.c_long(n[0]), C.c_long(n[1]), C.c_long(n[2]))
        assert rc == 0, (pool, key)

    def _bind_all(self):
        for (name, lev), arr in self.a.items():
            for pool in ("state", "diag", "mesh", "tend_physics"):
                self._bind(pool, name, lev, arr)
        for key, name in TEND_KEYS.items():
            self._bind("tend", key, 1, self.a[(name, 1)])
        self._bind("halo_scratch", "scale", 1, self.a[("scale_arr", 1)])
        for key, name in MODULE_KEYS.items():
            self._bind("module", key, 1, self.a[(name, 1)])
        for (pool, key), arr in self.extra.items():
            self._bind(pool, key, 1, arr)

    # -- Backend
    def _set_real(self, name, lev, arr): self.a[(name, lev)][...] = arr
    def _get_real(self, name, lev, out): out[...] = self.a[(name, lev)]
    def _set_int(self, name, arr): self.a[(name, 1)][...] = arr          # 1-based already, as the Fortran pools hold them

    def close(self):
        if self._h:
            self.lib.ref_destroy(self._h)
            self._h = C.c_void_p()

    def mpas_pool_shift_time_levels(self):
        for name, fd in FIELDS.items():
            if fd.levels == 2 and not name.startswith("lbc_"):       # the state pool only (mpas_atm_core.F:808)
                self.a[(name, 1)], self.a[(name, 2)] = self.a[(name, 2)], self.a[(name, 1)]
        self._bind_all()

    def k(self, routine, *args):
        ia = (C.c_int * 4)(*([a for a in args if not isinstance(a, float)] + [0] * 4)[:4])
        ra = (C.c_double * 4)(*([a for a in args if isinstance(a, float)] + [0.0] * 4)[:4])
        rc = self.lib.ref_call(self._h, routine.encode(), ia, ra)
        assert rc == 0, routine

    def atm_srk3(self, dt, itimestep=1):
        """The routines of one step in atm_srk3's order (TI:1066-1611, single block: the halo exchanges are no-ops).  The
        order is the hand-written part; every routine is the reference's own.  The step's trailing mpas_reconstruct
        (TI:1606, another source file) is not part of the transliteration."""
        cfg = self.cfg
        split = cfg["config_dynamics_split_steps"] if cfg["config_split_dynamics_transport"] else 1
        dt_dyn = dt / float(split)
        nss = cfg["config_number_of_sub_steps"]
        order = cfg["config_time_integration_order"]
        if order == 3:                                                        # TI:1010-1038
            rk_t = [dt_dyn / 3.0, dt_dyn / 2.0, dt_dyn]; rk_s = [dt_dyn / 3.0, dt_dyn / float(nss), dt_dyn / float(nss)]
            n_sub = [1, max(1, nss // 2), nss]
        else:
            rk_t = [dt_dyn / 2.0, dt_dyn / 2.0, dt_dyn]; rk_s = [dt_dyn / float(nss)] * 3
            n_sub = [max(1, nss // 2), max(1, nss // 2), nss]
        coupled = cfg["config_scalar_advection"] and not cfg["config_split_dynamics_transport"]

        def scalars(rk, dt_rk):                                               # advance_scalars, TI:1730-1927
            if rk < 3 or not (cfg["config_monotonic"] or cfg["config_positive_definite"]):
                self.k("advance_scalars", dt_rk, rk)
            else:
                self.k("advance_scalars_mono", dt_rk)

        for n in ("tend_ru_physics", "tend_rtheta_physics", "tend_rho_physics"):     # TI:1091-1093 (no physics)
            self.a[(n, 1)][...] = 0.0
        lbcs = bool(cfg.get("config_apply_lbcs", False))
        self.k("rk_integration_setup"); self.k("compute_moist_coefficients")
        for ds in range(1, split + 1):
            self.k("compute_vert_imp_coefs", rk_s[0])
            for rk in (1, 2, 3):
                if order == 3 and rk == 2:
                    self.k("compute_vert_imp_coefs", rk_s[rk - 1])
                self.k("compute_dyn_tend", rk, float(dt))
                time_dyn_step = dt_dyn * float(ds - 1) + rk_t[rk - 1]             # TI:1246
                if lbcs:                                                          # TI:1218-1268
                    self.k("lbc_speczone_tend")
                    self.k("lbc_relaxzone_tend", float(time_dyn_step), float(dt))
                self.k("set_smlstep_pert_variables")
                for ss in range(1, n_sub[rk - 1] + 1):
                    self.k("advance_acoustic_step", rk_s[rk - 1], ss); self.k("divergence_damping_3d", rk_s[rk - 1])
                self.k("recover_large_step_variables", rk_t[rk - 1], n_sub[rk - 1], rk)
                if lbcs:                                                          # TI:1343-1388
                    self.k("lbc_reset_u_ru", float(time_dyn_step))
                if coupled:
                    scalars(rk, rk_t[rk - 1])
                    if lbcs:                                                      # TI:1409-1430
                        self.k("lbc_adjust_scalars", float(dt), float(rk_t[rk - 1]))
                self.k("compute_solve_diagnostics", float(dt), rk)
                if lbcs:                                                          # TI:1477-1484
                    self.k("lbc_zero_gradient_w")
            self.k("rk_dynamics_substep_finish", ds, split)
        if cfg["config_scalar_advection"] and not coupled:
            rk_t = [dt / 2.0 if order == 2 else dt / 3.0, dt / 2.0, float(dt)]
            for rk in (1, 2, 3):
                scalars(rk, rk_t[rk - 1])
                if lbcs:                                                          # TI:1560-1582
                    self.k("lbc_adjust_scalars", float(dt), float(rk_t[rk - 1]))
        if lbcs:                                                                  # TI:1676-1720
            self.k("lbc_reset_speczone_values", float(dt))
            self.k("lbc_set_scalars", float(dt))

    def atm_init_coupled_diagnostics(self): self.k("init_coupled_diagnostics")
    def atm_init_solve_diagnostics(self, dt): self.k("init_solve_diagnostics", float(dt))
    def exchange_log(self): return self.lib.ref_exchange_log(self._h).decode()
