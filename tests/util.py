"""Helpers shared by the parity tests."""
import numpy as np

from mpas_model_b200.fields import FIELDS

STATE = ("u", "w", "rho_zz", "theta_m", "scalars")
# module scratch that the CUDA path does not materialise the way the reference does:
# ke_edge is recomputed inline; the per-scalar work arrays of the monotonic transport
# are fused away or hold intermediate values (transport kernels)
NOT_COMPARED = {"ke_edge", "wdtn", "scalar_old", "scalar_new"}


def rel_l2(a, b):
    d = np.linalg.norm((a - b).ravel())
    n = np.linalg.norm(b.ravel())
    return d / n if n > 0 else d


def real_fields():
    for name, fd in FIELDS.items():
        if fd.type == "REAL":
            for lev in range(1, fd.levels + 1):
                yield name, lev


def sync_all(src, dst):
    """Copy every real field (all time levels) from one backend to another."""
    for name, lev in real_fields():
        dst.set_array(name, src.get_array(name, lev), lev)


def compare_all(ref, test, names=None):
    """{field: rel-L2 difference} over all real fields (or the given ones)."""
    out = {}
    for name, lev in real_fields():
        if (names is not None and name not in names) or name in NOT_COMPARED:
            continue
        a, b = test.get_array(name, lev), ref.get_array(name, lev)
        out[f"{name}@{lev}"] = rel_l2(a, b)
    return out


def srk3_stepwise(backends, cfg, dt, after=None, reconstruct=True):
    """atm_srk3 (mpas_atm_time_integration.F:803-1725) driven one *_work routine at a
    time on each backend; ``after(label)`` is called after every routine."""
    def call(routine, *args):
        for b in backends:
            b.k(routine, *args)
        if after:
            after(routine + str(args))

    split = cfg["config_dynamics_split_steps"] if cfg["config_split_dynamics_transport"] else 1
    dt_dyn = dt / float(split)
    nss = cfg["config_number_of_sub_steps"]
    if cfg["config_time_integration_order"] == 3:
        rk_t = [dt_dyn / 3.0, dt_dyn / 2.0, dt_dyn]
        rk_s = [dt_dyn / 3.0, dt_dyn / float(nss), dt_dyn / float(nss)]
        n_sub = [1, max(1, nss // 2), nss]
    else:
        rk_t = [dt_dyn / 2.0, dt_dyn / 2.0, dt_dyn]
        rk_s = [dt_dyn / float(nss)] * 3
        n_sub = [max(1, nss // 2), max(1, nss // 2), nss]
    coupled = cfg["config_scalar_advection"] and not cfg["config_split_dynamics_transport"]
    lbcs = bool(cfg.get("config_apply_lbcs", False))

    def advance_scalars(rk, dt_rk):
        if rk < 3 or not (cfg["config_monotonic"] or cfg["config_positive_definite"]):
            call("advance_scalars", dt_rk, rk)
        else:
            call("advance_scalars_mono", dt_rk)

    for b in backends:                      # TI:1091-1093 (no physics)
        for n in ("tend_ru_physics", "tend_rtheta_physics", "tend_rho_physics"):
            b.set_array(n, np.zeros(b.shape(n)))
    call("rk_integration_setup")
    call("compute_moist_coefficients")
    for ds in range(1, split + 1):
        call("compute_vert_imp_coefs", rk_s[0])
        for rk in (1, 2, 3):
            if cfg["config_time_integration_order"] == 3 and rk == 2:
                call("compute_vert_imp_coefs", rk_s[rk - 1])
            call("compute_dyn_tend", rk, float(dt))
            time_dyn_step = dt_dyn * float(ds - 1) + rk_t[rk - 1]      # TI:1246
            if lbcs:                                     # regional run, TI:1218-1268
                call("lbc_speczone_tend")
                call("lbc_relaxzone_tend", float(time_dyn_step), float(dt))
            call("set_smlstep_pert_variables")
            for ss in range(1, n_sub[rk - 1] + 1):
                call("advance_acoustic_step", rk_s[rk - 1], ss)
                call("divergence_damping_3d", rk_s[rk - 1])
            call("recover_large_step_variables", rk_t[rk - 1], n_sub[rk - 1], rk)
            if lbcs:                                     # TI:1343-1388
                call("lbc_reset_u_ru", float(time_dyn_step))
            if coupled:                                  # config_split_dynamics_transport = false, TI:1404-1407
                advance_scalars(rk, rk_t[rk - 1])
                if lbcs:                                 # TI:1409-1430
                    call("lbc_adjust_scalars", float(dt), float(rk_t[rk - 1]))
            call("compute_solve_diagnostics", float(dt), rk)
            if lbcs:                                     # TI:1477-1484
                call("lbc_zero_gradient_w")
        call("rk_dynamics_substep_finish", ds, split)
    if cfg["config_scalar_advection"] and not coupled:
        rk_t = [dt / 2.0 if cfg["config_time_integration_order"] == 2 else dt / 3.0, dt / 2.0, float(dt)]
        for rk in (1, 2, 3):
            advance_scalars(rk, rk_t[rk - 1])
            if lbcs:                                     # TI:1560-1582
                call("lbc_adjust_scalars", float(dt), float(rk_t[rk - 1]))
    if lbcs:                                             # TI:1676-1720
        call("lbc_reset_speczone_values", float(dt))
        call("lbc_set_scalars", float(dt))
    if not reconstruct:
        return
    for b in backends:                      # TI:1596-1611
        b.mpas_reconstruct(2, False)
    if after:
        after("mpas_reconstruct(2, False)")
