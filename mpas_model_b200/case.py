"""Build a complete JW baroclinic-wave case (mesh + initial state + derived
mesh fields + namelist) for one of the BASELINE.json configs."""
from __future__ import annotations

import numpy as np

from . import init_block, jw_init, mesh


def make_case(n_cells: int, n_levels: int, dt: float | None = None, init_case: int = 2,
              num_scalars: int = 1, lloyd_iters: int = 12, jitter: float = 0.0, derive: bool = True, **cfg_overrides):
    """Returns (block, cfg).  ``dt`` defaults to the reference's rule of thumb
    of 6 s per km of nominal grid distance.  ``derive=False`` leaves out the init-time derived
    fields (``init_block``): a case that is only going to be decomposed gets them per block."""
    m = mesh.generate(n_cells, lloyd_iters=lloyd_iters, jitter=jitter)
    d = jw_init.init_atm_case_jw(m, n_levels, init_case=init_case, num_scalars=num_scalars)
    if dt is None:
        dt = 6.0 * round(d["nominalMinDc"] / 1000.0)
    cfg = init_block.default_config(d["nominalMinDc"], dt)
    cfg.update(cfg_overrides)
    if derive:
        init_block.init_block(d, cfg)
    if num_scalars > 1:
        add_passive_tracers(d)
    return d, cfg


def add_passive_tracers(d):
    """Smooth positive analytic blobs (cosine bells) for scalars 2..S."""
    nC, S = d["nCells"], d["num_scalars"]
    lat, lon = d["latCell"][:nC], d["lonCell"][:nC]
    nz = d["nVertLevels"]
    prof = np.sin(np.pi * (np.arange(nz) + 0.5) / nz) ** 2
    for s in range(1, S):
        lat0 = np.pi / 3.0 * np.sin(1.7 * s)
        lon0 = 2.0 * np.pi * ((0.37 * s) % 1.0)
        r = jw_init.sphere_distance(lat, lon, lat0, lon0, 1.0)
        bell = np.where(r < 0.6, 0.5 * (1.0 + np.cos(np.pi * r / 0.6)), 0.0)
        d["scalars"][:nC, :, s] = 1.0e-3 * (0.1 + bell[:, None] * prof[None, :])
