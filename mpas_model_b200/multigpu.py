"""One mesh block per GPU: host-side bootstrap of a decomposed run.

Mirrors what the reference does between reading the partition file and the first
time step (mpas_bootstrap_framework_phase1/2, src/framework/mpas_bootstrapping.F:106-500;
atm_mpas_init_block, mpas_atm_core.F:368-602):

  1. rank 0 builds the global case and the partition vector (``decomp.partition_rcb`` in
     place of METIS; any ``graph.info.part.N`` file can be used instead),
  2. every rank receives ITS block (owned + halo elements in the reference's local
     order) and its exchange lists (``decomp.decompose_case``),
  3. the lists go to the library (``mpasb_set_halo_lists``), NCCL is initialised
     from a unique id broadcast over ``torch.distributed`` (the only use of torch here),
  4. the init-time exchanges and diagnostics run in the reference's order
     (mpas_atm_core.F:250, 288, 515-527).

``srk3_host_exchange`` is the same step driven routine by routine with the halo
exchanges done on HOST arrays over a ``torch.distributed`` process group (gloo on CPU):
it is what the CPU tests use to cover the N > 1 host logic without a GPU.
"""
from __future__ import annotations

import os
import pickle
import time

import numpy as np

from . import decomp
from .fields import FIELDS

KINDS = (("cells", 0), ("edges", 1), ("vertices", 2))

# mpas_atm_halos.F:211-298: group -> ((field, time level, kind, layers), ...); kept in step with
# csrc/halo_host.inl (tests/test_multigpu.py compares the two through the known-answer exchange)
GROUPS = {
    "dynamics:theta_m,scalars,pressure_p,rtheta_p": (("theta_m", 1, "cells", (1, 2)), ("scalars", 1, "cells", (1, 2)),
                                                     ("pressure_p", 1, "cells", (1, 2)), ("rtheta_p", 1, "cells", (1, 2))),
    "dynamics:rw_p,ru_p,rho_pp,rtheta_pp": (("rw_p", 1, "cells", (1,)), ("ru_p", 1, "edges", (2,)),
                                            ("rho_pp", 1, "cells", (1, 2)), ("rtheta_pp", 1, "cells", (2,))),
    "dynamics:w,pv_edge,rho_edge": (("w", 2, "cells", (1, 2)), ("pv_edge", 1, "edges", (1, 2)), ("rho_edge", 1, "edges", (1, 2))),
    "dynamics:w,pv_edge,rho_edge,scalars": (("w", 2, "cells", (1, 2)), ("pv_edge", 1, "edges", (1, 2)), ("rho_edge", 1, "edges", (1, 2)),
                                            ("scalars", 2, "cells", (1, 2))),
    "dynamics:theta_m,pressure_p,rtheta_p": (("theta_m", 2, "cells", (1, 2)), ("pressure_p", 1, "cells", (1, 2)),
                                             ("rtheta_p", 1, "cells", (1, 2))),
    "dynamics:exner": (("exner", 1, "cells", (1, 2)),),
    "dynamics:tend_u": (("tend_u", 1, "edges", (1,)),),
    "dynamics:rho_pp": (("rho_pp", 1, "cells", (1,)),),
    "dynamics:rtheta_pp": (("rtheta_pp", 1, "cells", (1,)),),
    "dynamics:rtheta_pp,rho_pp": (("rtheta_pp", 1, "cells", (1,)), ("rho_pp", 1, "cells", (1,))),   # merged (library only)
    "dynamics:u_123": (("u", 2, "edges", (1, 2, 3)),),
    "dynamics:u_3": (("u", 2, "edges", (3,)),),
    "dynamics:scalars": (("scalars", 2, "cells", (1, 2)),),
    "dynamics:scalars_old": (("scalars", 1, "cells", (1, 2)),),
    "dynamics:w": (("w", 2, "cells", (1, 2)),),
    "dynamics:scale": (("scale_arr", 1, "cells", (1, 2)),),
    "dynamics:scale_all": (("scale_arr", 0, "cells", (1, 2)),),   # library only: the pairs of all scalars in one message (level 0)
    "initialization:u": (("u", 1, "edges", (1, 2, 3)),),
    "initialization:pv_edge,ru,rw": (("pv_edge", 1, "edges", (1, 2, 3)), ("ru", 1, "edges", (1, 2, 3)), ("rw", 1, "cells", (1, 2))),
}


# ---------------------------------------------------------------------------------- bootstrap
def _cache_dir():
    """Per-user, mode-0700 directory for the prepared blocks (they are pickles: never read one from a directory another
    user could have written).  MPASB_CACHE overrides the location, not the ownership check."""
    p = os.environ.get("MPASB_CACHE") or os.path.join(os.path.expanduser("~"), ".cache", "mpasb_cases")
    os.makedirs(p, mode=0o700, exist_ok=True)
    st = os.stat(p)
    if st.st_uid != os.getuid() or (st.st_mode & 0o022):
        raise RuntimeError(f"{p}: block cache must be owned by the current user and not group/world-writable")
    return p


def _source_key(n_cells, n_levels, num_scalars):
    """Hash of everything a prepared block depends on: the generating sources and the default namelist, so that an edit
    to the mesh generator, the JW initialisation, the init-time derivations or the decomposition never reuses stale blocks."""
    import hashlib
    from . import case, decomp as dc, init_block, jw_init, mesh, reconstruct
    hsh = hashlib.sha256()
    for mod in (case, dc, init_block, jw_init, mesh, reconstruct):
        with open(mod.__file__, "rb") as f:
            hsh.update(f.read())
    hsh.update(repr(sorted(init_block.default_config(1.0, 1.0).items())).encode())
    hsh.update(f"{n_cells}.{n_levels}.{num_scalars}".encode())
    return hsh.hexdigest()[:12]


def prepare_blocks(n_cells, n_levels, num_scalars, world, tag=""):
    """Rank-0 work: global case -> partition -> one pickle per rank.  Returns the path prefix.
    The pickles are reused by later runs of the same (mesh, levels, scalars, world)."""
    from .case import make_case
    key = _source_key(n_cells, n_levels, num_scalars)
    prefix = os.path.join(_cache_dir(), f"x1.{n_cells}.L{n_levels}.S{num_scalars}.{key}.part.{world}{tag}")
    if all(os.path.exists(f"{prefix}.{r}.pkl") for r in range(world)):
        return prefix
    gpath = os.path.join(_cache_dir(), f"x1.{n_cells}.L{n_levels}.S{num_scalars}.{key}.global.raw.pkl")
    if os.path.exists(gpath):                                   # the global case is shared by every partition count
        with open(gpath, "rb") as f:
            d, cfg = pickle.load(f)
    else:
        d, cfg = make_case(n_cells, n_levels, num_scalars=num_scalars, derive=False)     # derived fields are computed per block
        if n_cells <= 200000:                                   # 4 GB at x1.163842; not worth 17 GB at x1.655362
            with open(gpath + ".tmp", "wb") as f:
                pickle.dump((d, cfg), f, protocol=4)
            os.replace(gpath + ".tmp", gpath)
    part = decomp.partition_rcb(d, world)
    decomp.write_partition_file(prefix, part)               # the reference's "<prefix><N>" partition file format
    blocks, ex = decomp.decompose_case(d, cfg, part)
    for r in range(world):
        # keep what the library consumes (the field table) plus scalars, ids and small geometry; the init-only
        # intermediates (zb/zb3 per edge, deriv_two, *_init) are half of a block's bytes
        b = {k: v for k, v in blocks[r].items()
             if not isinstance(v, np.ndarray) or k in FIELDS or v.nbytes < (4 << 20) or k.startswith("indexTo")}
        with open(f"{prefix}.{r}.pkl.tmp", "wb") as f:
            pickle.dump({"block": b, "cfg": cfg, "ex": ex[r], "nCellsGlobal": d["nCells"]}, f, protocol=4)
        os.replace(f"{prefix}.{r}.pkl.tmp", f"{prefix}.{r}.pkl")
    return prefix


def load_block(prefix, rank):
    with open(f"{prefix}.{rank}.pkl", "rb") as f:
        return pickle.load(f)


def init_distributed(backend="gloo"):
    """A CPU-side process group for bootstrap data and barriers (the halo traffic itself goes
    through the library's own NCCL communicator)."""
    import torch.distributed as dist
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend)
    return dist


def setup_rank(n_cells, n_levels, num_scalars, rank, world, device, precision="double"):
    """Create this rank's ``Dycore`` with halo lists and NCCL communicator, run the init sequence."""
    from .dycore import Dycore
    dist = init_distributed()
    box = [None]
    if rank == 0:
        box[0] = prepare_blocks(n_cells, n_levels, num_scalars, world)
    dist.broadcast_object_list(box, src=0)
    rec = load_block(box[0], rank)
    block, cfg, ex = rec["block"], rec["cfg"], rec["ex"]
    g = Dycore(block, cfg, device=device, precision=precision)
    for kind, k in KINDS:
        g.set_halo_lists(k, ex[kind])
    uid = [g.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    g.comm_init(rank, world, uid[0])
    g.p2p_on = False
    if os.environ.get("MPASB_P2P", "1") != "0" and 2 <= world <= 9:
        g.p2p_on = g.p2p_init(dist)   # exchanges by direct NVLink stores (CUDA IPC) instead of NCCL send/recv
    dt = cfg["config_dt"]
    g.exchange_halo_group("initialization:u")                       # mpas_atm_core.F:250
    g.atm_init_coupled_diagnostics()
    g.atm_init_solve_diagnostics(dt)
    g.exchange_halo_group("initialization:pv_edge,ru,rw")           # mpas_atm_core.F:288
    return g, block, cfg, ex, rec["nCellsGlobal"]


# ---------------------------------------------------------------------------------- host exchange
class HostExchanger:
    """mpas_halo_exch_group_full_halo_exch (src/framework/mpas_halo.F:498-846) on host arrays over a
    torch.distributed group: one message per neighbour and group, field-major, then halo layer, then
    list element, vertical index fastest (MH:671,695) -- the layout the GPU pack kernel produces."""

    def __init__(self, dist, rank, ex):
        self.dist, self.rank, self.ex = dist, rank, ex

    def exchange(self, backend, group):
        import torch
        fields = GROUPS[group]
        peers = sorted({q for (_, _, kind, _) in fields for q in self.ex[kind]["neighbors"]})
        arrays = {(n, lev): backend.get_array(n, lev) for (n, lev, _, _) in fields}
        sends, recvs, reqs = {}, {}, []
        for q in peers:
            parts, nrecv = [], 0
            for (n, lev, kind, layers) in fields:
                e = self.ex[kind]
                if q not in e["neighbors"]:
                    continue
                qi = e["neighbors"].index(q)
                a = arrays[(n, lev)]
                inner = int(np.prod(a.shape[1:]))
                for l in layers:
                    parts.append(a[e["send"][qi][l - 1]].reshape(-1))
                    nrecv += len(e["recv"][qi][l - 1]) * inner
            sends[q] = torch.from_numpy(np.concatenate(parts + [np.zeros(0)]))
            recvs[q] = torch.empty(nrecv, dtype=torch.float64)
        for q in peers:                                           # MPI_Irecv all, then MPI_Isend all (MH:560-640)
            if recvs[q].numel():
                reqs.append(self.dist.irecv(recvs[q], src=q))
        for q in peers:
            if sends[q].numel():
                reqs.append(self.dist.isend(sends[q], dst=q))
        for r in reqs:
            r.wait()
        for q in peers:
            buf, off = recvs[q].numpy(), 0
            for (n, lev, kind, layers) in fields:
                e = self.ex[kind]
                if q not in e["neighbors"]:
                    continue
                qi = e["neighbors"].index(q)
                a = arrays[(n, lev)]
                inner = int(np.prod(a.shape[1:]))
                for l in layers:
                    idx = e["recv"][qi][l - 1]
                    cnt = len(idx) * inner
                    a[idx] = buf[off:off + cnt].reshape((len(idx),) + a.shape[1:])
                    off += cnt
        for (n, lev, _, _) in fields:
            backend.set_array(n, arrays[(n, lev)], lev)


def srk3_host_exchange(b, cfg, dt, xch):
    """atm_srk3 (mpas_atm_time_integration.F:803-1725) with the reference's exchange call sites,
    driven one *_work routine at a time on backend ``b``; ``xch(group)`` performs a halo exchange."""
    split = cfg["config_dynamics_split_steps"] if cfg["config_split_dynamics_transport"] else 1
    dt_dyn = dt / float(split)
    nss = cfg["config_number_of_sub_steps"]
    order = cfg["config_time_integration_order"]
    if order == 3:
        rk_t = [dt_dyn / 3.0, dt_dyn / 2.0, dt_dyn]
        rk_s = [dt_dyn / 3.0, dt_dyn / float(nss), dt_dyn / float(nss)]
        n_sub = [1, max(1, nss // 2), nss]
    else:
        rk_t = [dt_dyn / 2.0, dt_dyn / 2.0, dt_dyn]
        rk_s = [dt_dyn / float(nss)] * 3
        n_sub = [max(1, nss // 2), max(1, nss // 2), nss]
    coupled = cfg["config_scalar_advection"] and not cfg["config_split_dynamics_transport"]

    def advance_scalars(rk, dt_rk):                                        # advance_scalars, TI:1730-1927
        if rk < 3 or not (cfg["config_monotonic"] or cfg["config_positive_definite"]):
            b.k("advance_scalars", dt_rk, rk)
        else:
            b.k("advance_scalars_mono_pre", dt_rk)
            xch("dynamics:scalars_old")                                    # TI:4155
            for s in range(b.dims.num_scalars):
                b.k("advance_scalars_mono_a", dt_rk, s)
                xch("dynamics:scale")                                      # TI:4568
                b.k("advance_scalars_mono_b", dt_rk, s)

    for n in ("tend_ru_physics", "tend_rtheta_physics", "tend_rho_physics"):
        b.set_array(n, np.zeros(b.shape(n)))
    xch("dynamics:theta_m,scalars,pressure_p,rtheta_p")                    # TI:1066
    b.k("rk_integration_setup")
    b.k("compute_moist_coefficients")
    for ds in range(1, split + 1):
        b.k("compute_vert_imp_coefs", rk_s[0])
        xch("dynamics:exner")                                              # TI:1131
        for rk in (1, 2, 3):
            if order == 3 and rk == 2:
                b.k("compute_vert_imp_coefs", rk_s[rk - 1])
            b.k("compute_dyn_tend", rk, float(dt))
            xch("dynamics:tend_u")                                         # TI:1203
            b.k("set_smlstep_pert_variables")
            for ss in range(1, n_sub[rk - 1] + 1):
                xch("dynamics:rho_pp")                                     # TI:1279
                b.k("advance_acoustic_step", rk_s[rk - 1], ss)
                xch("dynamics:rtheta_pp")                                  # TI:1302
                b.k("divergence_damping_3d", rk_s[rk - 1])
            xch("dynamics:rw_p,ru_p,rho_pp,rtheta_pp")                     # TI:1322
            b.k("recover_large_step_variables", rk_t[rk - 1], n_sub[rk - 1], rk)
            xch("dynamics:u_3")                                            # TI:1398
            if coupled:
                advance_scalars(rk, rk_t[rk - 1])                          # TI:1404-1407
            b.k("compute_solve_diagnostics", float(dt), rk)
            xch("dynamics:w,pv_edge,rho_edge,scalars" if coupled else "dynamics:w,pv_edge,rho_edge")     # TI:1463-1473
        if ds < split:
            xch("dynamics:theta_m,pressure_p,rtheta_p")                    # TI:1493
        b.k("rk_dynamics_substep_finish", ds, split)
    if not cfg["config_scalar_advection"] or coupled:
        return
    rk_t = [dt / 2.0 if order == 2 else dt / 3.0, dt / 2.0, float(dt)]
    for rk in (1, 2, 3):
        advance_scalars(rk, rk_t[rk - 1])
        if rk < 3:
            xch("dynamics:scalars")                                        # TI:1588-1590


def gather_owned(dist, g, block, names=("u", "w", "rho_zz", "theta_m", "scalars"), time_level=1):
    """Rank 0 returns {name: global array} assembled from every rank's owned elements."""
    L = block["lists"]
    mine = {}
    for n in names:
        kind, key = (("edges", "edge_bounds") if FIELDS[n].loc == "EDGE" else ("cells", "cell_bounds"))
        cnt = int(L[key][0])
        mine[n] = (L[kind][:cnt], g.get_array(n, time_level)[:cnt])
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, mine)
    if dist.get_rank() != 0:
        return None
    glob = {}
    for n in names:
        total = sum(len(o[n][0]) for o in out)
        a = np.empty((total,) + out[0][n][1].shape[1:])
        for o in out:
            a[o[n][0]] = o[n][1]
        glob[n] = a
    return glob
