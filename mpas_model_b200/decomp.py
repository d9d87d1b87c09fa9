"""Block decomposition and halo exchange lists, restating the reference's bootstrap.

  * partition file  "<prefix><nBlocks>": one 0-based block id per line, cell order
        src/framework/mpas_block_decomp.F:101-137
  * owned cells ascending by global id            mpas_block_creator.F:55-119 (mpas_quicksort at :102)
  * cell halo layer i = not-yet-present neighbours (cellsOnCell) of every cell in the block,
    sorted by global id                           mpas_block_creator.F:473-720, mpas_block_decomp.F:435-507
  * edges (vertices): all edges of the owned cells in first-touch order (mpas_block_decomp.F:370-422);
    owned iff cellsOnEdge(1,e) is owned, owned first in that order, the rest written from the end of
    the list backwards, i.e. in REVERSED first-touch order (mpas_block_decomp.F:302-357);
    layers 2 and 3 = new edges of the halo-1 / halo-2 cells in first-touch order
                                                  mpas_block_creator.F:269-456, 738-942
  * exchange lists per (neighbour, layer): the sender's srcList and the positions in the message are
    in ascending global id of the needed elements (mpas_dmpar.F:1893-1913); the receiver's list is in
    local halo order with srcList = position in the message (mpas_dmpar.F:2033-2043)
  * global -> local connectivity, anything outside the block -> n+1 (here: n)
                                                  mpas_block_creator.F:1399-1531
METIS is not available, so ``partition_rcb`` supplies a deterministic partition vector (recursive
coordinate bisection); any vector in the reference's file format can be used instead.
"""
from __future__ import annotations

import numpy as np

from .mesh import _CELL_FIELDS, _EDGE_FIELDS, _VERTEX_FIELDS, _GARBAGE_OF

N_CELL_HALOS = 2          # config_num_halos, Registry.xml:249


def partition_rcb(m: dict, nparts: int) -> np.ndarray:
    """Recursive coordinate bisection of the cell centres into ``nparts`` (power of two or not)."""
    nC = m["nCells"]
    xyz = np.stack([m["xCell"][:nC], m["yCell"][:nC], m["zCell"][:nC]], 1)
    part = np.zeros(nC, dtype=np.int32)

    def split(idx, lo, n):
        if n == 1:
            part[idx] = lo
            return
        nl = n // 2
        p = xyz[idx]
        ax = int(np.argmax(p.max(0) - p.min(0)))
        order = np.argsort(p[:, ax], kind="stable")
        cut = int(round(len(idx) * nl / n))
        split(idx[order[:cut]], lo, nl)
        split(idx[order[cut:]], lo + nl, n - nl)

    split(np.arange(nC), 0, nparts)
    return part


def write_partition_file(path: str, part: np.ndarray) -> None:
    np.savetxt(path, part, fmt="%d")


def read_partition_file(path: str) -> np.ndarray:
    return np.loadtxt(path, dtype=np.int32)


def _first_touch(conn, counts, cells):
    """Unique entries of conn[cells, :counts] in first-touch order (row-major)."""
    rows = conn[cells]
    live = np.arange(rows.shape[1])[None, :] < counts[cells][:, None]
    flat = rows[live]
    _, first = np.unique(flat, return_index=True)
    return flat[np.sort(first)]


def block_lists(g: dict, part: np.ndarray, rank: int) -> dict:
    """Local -> global index lists of one block in the reference's local order, plus layer bounds.
    ``g`` is the global mesh dict (0-based, garbage slot present)."""
    nC, nE, nV = g["nCells"], g["nEdges"], g["nVertices"]
    nec, coc = g["nEdgesOnCell"][:nC], g["cellsOnCell"][:nC]
    owned = np.nonzero(part == rank)[0]                      # ascending global id
    in_block = np.zeros(nC + 1, dtype=bool)
    in_block[owned] = True
    cells = [owned]
    cur = owned
    for _ in range(N_CELL_HALOS):
        allc = np.concatenate(cells)
        rows = coc[allc]
        live = np.arange(rows.shape[1])[None, :] < nec[allc][:, None]
        nb = np.unique(rows[live])
        nb = nb[(nb < nC) & ~in_block[nb]]                   # sorted by global id
        in_block[nb] = True
        cells.append(nb)
    cell_bounds = np.cumsum([len(c) for c in cells])
    cell_list = np.concatenate(cells)

    out = {"cells": cell_list, "cell_bounds": cell_bounds}
    for kind, conn, owner_conn, n_glob in (("edges", g["edgesOnCell"][:nC], g["cellsOnEdge"][:, 0], nE),
                                           ("vertices", g["verticesOnCell"][:nC], g["cellsOnVertex"][:, 0], nV)):
        present = np.zeros(n_glob + 1, dtype=bool)
        ft = _first_touch(conn, nec, cells[0])
        is_owned = part[owner_conn[ft]] == rank
        own, ghost = ft[is_owned], ft[~is_owned][::-1]       # ghosts were written from the end backwards
        layers = [own, ghost]
        present[ft] = True
        for l in range(N_CELL_HALOS):
            upto = np.concatenate(cells[: l + 2])
            ft = _first_touch(conn, nec, upto)
            new = ft[~present[ft]]
            present[new] = True
            layers.append(new)
        out[kind] = np.concatenate(layers)
        out[kind[:-1] + "_bounds"] = np.cumsum([len(x) for x in layers])
    return out


def _owner_of(g, part, kind, gid):
    if kind == "cells":
        return part[gid]
    if kind == "edges":
        return part[g["cellsOnEdge"][gid, 0]]
    return part[g["cellsOnVertex"][gid, 0]]


def exchange_lists(g: dict, part: np.ndarray, all_lists: list) -> list:
    """For every block: per kind the reference's exchange lists.

    Returns ``ex[rank][kind]`` = dict(neighbors=[...], n_layers=L,
        send=[[local idx array per layer] per neighbour]   # sendListSrc, message order = ascending gid
        recv=[[local idx array per layer] per neighbour]   # halo elements in MESSAGE order
        recv_ref=[[(srcList, destList) per layer] per neighbour])   # the reference's form, local halo order
    ``all_lists`` = [block_lists(g, part, r) for r in range(nranks)]."""
    nranks = len(all_lists)
    ex = []
    for r in range(nranks):
        ex.append({})
    for kind, bname, nlay in (("cells", "cell_bounds", N_CELL_HALOS), ("edges", "edge_bounds", N_CELL_HALOS + 1),
                              ("vertices", "vertice_bounds", N_CELL_HALOS + 1)):
        # needed[r][l] = (gids in local halo order, local indices)
        needed = []
        for r in range(nranks):
            L, b = all_lists[r][kind], all_lists[r][bname]
            needed.append([(L[b[l]:b[l + 1]], np.arange(b[l], b[l + 1])) for l in range(nlay)])
        for r in range(nranks):
            L, b = all_lists[r][kind], all_lists[r][bname]
            own_gid = L[:b[0]]
            order = np.argsort(own_gid, kind="stable")
            sorted_gid = own_gid[order]
            nbr_send, nbr_recv = {}, {}
            for q in range(nranks):
                if q == r:
                    continue
                # what q needs from r, layer by layer, ascending gid (mpas_dmpar.F:1893-1913)
                for l in range(nlay):
                    gids, _ = needed[q][l]
                    mine = np.sort(gids[_owner_of(g, part, kind, gids) == r])
                    if len(mine):
                        pos = np.searchsorted(sorted_gid, mine)
                        assert (sorted_gid[pos] == mine).all()
                        nbr_send.setdefault(q, [np.zeros(0, np.int64)] * nlay)
                        nbr_send[q] = list(nbr_send[q])
                        nbr_send[q][l] = order[pos]
                # what r needs from q
                for l in range(nlay):
                    gids, loc = needed[r][l]
                    sel = _owner_of(g, part, kind, gids) == q
                    if sel.any():
                        gq, lq = gids[sel], loc[sel]
                        rank_in_msg = np.argsort(np.argsort(gq, kind="stable"), kind="stable")    # position in the message (0-based)
                        nbr_recv.setdefault(q, [None] * nlay)
                        nbr_recv[q][l] = (rank_in_msg, lq)
            nbrs = sorted(set(nbr_send) | set(nbr_recv))
            send, recv, recv_ref = [], [], []
            for q in nbrs:
                s = nbr_send.get(q, [np.zeros(0, np.int64)] * nlay)
                send.append([np.asarray(a, dtype=np.int64) for a in s])
                rr = nbr_recv.get(q, [None] * nlay)
                rl, rref = [], []
                for l in range(nlay):
                    if rr[l] is None:
                        rl.append(np.zeros(0, np.int64)); rref.append((np.zeros(0, np.int64), np.zeros(0, np.int64)))
                    else:
                        pos, loc = rr[l]
                        by_msg = np.empty_like(loc)
                        by_msg[pos] = loc
                        rl.append(by_msg); rref.append((pos, loc))
                recv.append(rl); recv_ref.append(rref)
            # the reference appends send lists in ring order me-1, me-2, ... and recv lists by ascending rank
            ring = [(r - i) % nranks for i in range(1, nranks)]
            ex[r][kind] = dict(neighbors=nbrs, n_layers=nlay, send=send, recv=recv, recv_ref=recv_ref,
                               send_order=[q for q in ring if q in nbr_send], recv_order=sorted(nbr_recv))
    return ex


_LOCAL_KEYS = {"cells": ("nCells", _CELL_FIELDS), "edges": ("nEdges", _EDGE_FIELDS), "vertices": ("nVertices", _VERTEX_FIELDS)}
# per-element fields the JW init / init file adds (all carry the garbage row already)
_EXTRA = {
    "cells": ("hx", "zgrid", "zz", "rho", "theta", "scalars", "rho_base", "theta_base", "surface_pressure", "w",
              "defc_a", "defc_b", "t_init", "rho_zz_init", "theta_m_init", "rw_init", "pressure_p_init",
              "pressure_base_init", "exner_init", "exner_base_init"),
    "edges": ("zxu", "zb", "zb3", "fEdge", "u", "deriv_two", "ru_init"),
    "vertices": ("fVertex",),
}


def make_block(dglob: dict, lists: dict) -> dict:
    """Gather every per-cell/edge/vertex array of the global case dict onto one block and rewrite its
    connectivity to local indices (outside the block -> the garbage index n)."""
    d = {k: v for k, v in dglob.items() if np.isscalar(v) or isinstance(v, (bool, str))}
    for k in ("rdzw", "dzu", "rdzu", "fzm", "fzp", "u_init", "v_init", "qv_init"):
        if k in dglob:
            d[k] = dglob[k]
    maps = {}
    for kind, (nname, names) in _LOCAL_KEYS.items():
        L = lists[kind]
        n_glob = dglob[nname]
        g2l = np.full(n_glob + 1, len(L), dtype=np.int64)
        g2l[L] = np.arange(len(L))
        maps[kind] = g2l
        d[nname] = len(L)
        sel = np.concatenate([L, [n_glob]])                  # garbage row last
        for name in tuple(names) + _EXTRA[kind]:
            if name in dglob:
                d[name] = dglob[name][sel].copy()
    tgt = {"nCells": "cells", "nEdges": "edges", "nVertices": "vertices"}
    for name, nname in _GARBAGE_OF.items():
        d[name] = maps[tgt[nname]][d[name]].astype(np.int32)
    d["nCellsSolve"] = int(lists["cell_bounds"][0])
    d["nEdgesSolve"] = int(lists["edge_bounds"][0])
    d["nVerticesSolve"] = int(lists["vertice_bounds"][0])
    d["indexToCellID"] = np.concatenate([lists["cells"] + 1, [0]]).astype(np.int32)
    d["indexToEdgeID"] = np.concatenate([lists["edges"] + 1, [0]]).astype(np.int32)
    d["indexToVertexID"] = np.concatenate([lists["vertices"] + 1, [0]]).astype(np.int32)
    d["lists"] = lists
    return d


def decompose_case(dglob: dict, cfg: dict, part: np.ndarray, ranks=None):
    """Blocks (with init-time derived fields computed per block, as atm_mpas_init_block does) and
    exchange lists for the requested ranks."""
    from .init_block import init_block
    nranks = int(part.max()) + 1
    all_lists = [block_lists(dglob, part, r) for r in range(nranks)]
    ex = exchange_lists(dglob, part, all_lists)
    ranks = range(nranks) if ranks is None else ranks
    blocks = {}
    for r in ranks:
        b = make_block(dglob, all_lists[r])
        init_block(b, cfg)
        blocks[r] = b
    return blocks, ex
