// init_block_host.inl -- the mesh part of atm_mpas_init_block in C++.
//
// Reference: src/core_atmosphere/mpas_atm_core.F ("CORE")
//   atm_compute_signs           CORE:1151-1238     inverses                   CORE:456-470
//   atm_adv_coef_compression    CORE:1285-1430     atm_couple_coef_3rd_order  CORE:1433-1452
//   atm_compute_mesh_scaling    CORE:1091-1148     atm_compute_damping_coefs  CORE:1241-1282
//
// Host code: it runs once per mesh upload on the arrays a Fortran caller holds (dense, fastest index first, 1-based
// connectivity with the garbage slot n+1), so a run can start from the raw fields of an init file with neither Fortran nor
// Python deriving anything.  mpasb_init_block_host is stateless (no device needed); mpasb_init_block computes the same
// fields and stores them in the handle as mpasb_set_field would.  Values the reference leaves untouched (slots beyond
// nEdgesOnCell, the garbage row, edges without an owned cell) are zero -- or 1 for the LOCAL index kiteForCell, the
// garbage index for advCellsForEdge -- as in mpas_model_b200/init_block.py, which tests/test_reference_pin.py pins to the
// reference's own routines.
#pragma once
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

namespace initblk {

struct In {
    const int *nEdgesOnCell = nullptr, *edgesOnCell = nullptr, *cellsOnCell = nullptr, *verticesOnCell = nullptr, *cellsOnEdge = nullptr,
              *verticesOnEdge = nullptr, *edgesOnVertex = nullptr, *cellsOnVertex = nullptr;
    const real *zb = nullptr, *zb3 = nullptr, *deriv_two = nullptr, *dcEdge = nullptr, *dvEdge = nullptr, *areaCell = nullptr,
               *areaTriangle = nullptr, *meshDensity = nullptr, *zgrid = nullptr;
};
struct Out {
    std::vector<real> edgesOnVertex_sign, edgesOnCell_sign, zb_cell, zb3_cell, invAreaCell, invDvEdge, invDcEdge, invAreaTriangle,
                      adv_coefs, adv_coefs_3rd, meshScalingDel2, meshScalingDel4, meshScalingRegionalCell, meshScalingRegionalEdge, dss;
    std::vector<int> kiteForCell, nAdvCellsForEdge, advCellsForEdge;
};

static const char* const IN_INT[] = {"nEdgesOnCell", "edgesOnCell", "cellsOnCell", "verticesOnCell", "cellsOnEdge", "verticesOnEdge",
                                     "edgesOnVertex", "cellsOnVertex"};
static const char* const IN_REAL[] = {"zb", "zb3", "deriv_two", "dcEdge", "dvEdge", "areaCell", "areaTriangle", "meshDensity", "zgrid"};

// binds the (name, array) pairs of the caller to In; returns the name of a missing input or nullptr
static const char* bind(In& in, int n, const char* const* names, const void* const* arrays) {
    const int** ip[] = {&in.nEdgesOnCell, &in.edgesOnCell, &in.cellsOnCell, &in.verticesOnCell, &in.cellsOnEdge, &in.verticesOnEdge,
                        &in.edgesOnVertex, &in.cellsOnVertex};
    const real** rp[] = {&in.zb, &in.zb3, &in.deriv_two, &in.dcEdge, &in.dvEdge, &in.areaCell, &in.areaTriangle, &in.meshDensity, &in.zgrid};
    for (int k = 0; k < n; k++) {
        if (!names[k] || !arrays[k]) continue;
        for (size_t q = 0; q < sizeof(IN_INT) / sizeof(*IN_INT); q++) if (!strcmp(names[k], IN_INT[q])) *ip[q] = (const int*)arrays[k];
        for (size_t q = 0; q < sizeof(IN_REAL) / sizeof(*IN_REAL); q++) if (!strcmp(names[k], IN_REAL[q])) *rp[q] = (const real*)arrays[k];
    }
    for (size_t q = 0; q < sizeof(IN_INT) / sizeof(*IN_INT); q++) if (!*ip[q]) return IN_INT[q];
    for (size_t q = 0; q < sizeof(IN_REAL) / sizeof(*IN_REAL); q++) if (!*rp[q]) return IN_REAL[q];
    return nullptr;
}

// All indices below are the reference's: 1-based values, arrays addressed (fastest, ..., slowest) through these accessors.
static void compute(const mpasb_dims& dm, const mpasb_config& cf, int h_ScaleWithMesh, double config_zd, double config_xnutr,
                    const In& in, Out& o) {
    const int nC = dm.nCells, nE = dm.nEdges, nV = dm.nVertices, mx = dm.maxEdges, vd = dm.vertexDegree, nz = dm.nVertLevels, nz1 = nz + 1;
    auto eoc = [&](int i, int c) { return in.edgesOnCell[(size_t)(c - 1) * mx + (i - 1)]; };
    auto coc = [&](int i, int c) { return in.cellsOnCell[(size_t)(c - 1) * mx + (i - 1)]; };
    auto coe = [&](int i, int e) { return in.cellsOnEdge[(size_t)(e - 1) * 2 + (i - 1)]; };
    auto nec = [&](int c) { return in.nEdgesOnCell[c - 1]; };
    const real one = (real)1.0;

    // ---- atm_compute_signs, CORE:1151-1238
    o.edgesOnVertex_sign.assign((size_t)(nV + 1) * vd, (real)0);
    for (int v = 1; v <= nV; v++)
        for (int i = 1; i <= vd; i++) {
            const int e = in.edgesOnVertex[(size_t)(v - 1) * vd + (i - 1)];
            if (e <= nE) o.edgesOnVertex_sign[(size_t)(v - 1) * vd + (i - 1)] = (v == in.verticesOnEdge[(size_t)(e - 1) * 2 + 1]) ? one : -one;
        }
    o.edgesOnCell_sign.assign((size_t)(nC + 1) * mx, (real)0);
    o.zb_cell.assign((size_t)(nC + 1) * mx * nz1, (real)0);
    o.zb3_cell.assign((size_t)(nC + 1) * mx * nz1, (real)0);
    for (int c = 1; c <= nC; c++)
        for (int i = 1; i <= nec(c); i++) {
            const int e = eoc(i, c);
            if (e > nE) continue;
            const int side = (c == coe(1, e)) ? 0 : 1;
            o.edgesOnCell_sign[(size_t)(c - 1) * mx + (i - 1)] = side == 0 ? one : -one;
            const size_t src = ((size_t)(e - 1) * 2 + side) * nz1, dst = ((size_t)(c - 1) * mx + (i - 1)) * nz1;
            for (int k = 0; k < nz1; k++) {
                o.zb_cell[dst + k] = in.zb[src + k];
                o.zb3_cell[dst + k] = (real)cf.config_coef_3rd_order * in.zb3[src + k];          // atm_couple_coef_3rd_order, CORE:1450
            }
        }
    o.kiteForCell.assign((size_t)(nC + 1) * mx, 1);
    for (int c = 1; c <= nC; c++)
        for (int i = 1; i <= nec(c); i++) {
            const int v = in.verticesOnCell[(size_t)(c - 1) * mx + (i - 1)];
            if (v > nV) continue;
            for (int j = 1; j <= vd; j++)
                if (c == in.cellsOnVertex[(size_t)(v - 1) * vd + (j - 1)]) { o.kiteForCell[(size_t)(c - 1) * mx + (i - 1)] = j; break; }
        }

    // ---- inverses, CORE:456-470 (whole arrays, garbage slot included, as `invAreaCell = 1.0_RKIND / areaCell` does)
    o.invAreaCell.resize(nC + 1); o.invDvEdge.resize(nE + 1); o.invDcEdge.resize(nE + 1); o.invAreaTriangle.resize(nV + 1);
    for (int c = 0; c <= nC; c++) o.invAreaCell[c] = one / in.areaCell[c];
    for (int e = 0; e <= nE; e++) { o.invDvEdge[e] = one / in.dvEdge[e]; o.invDcEdge[e] = one / in.dcEdge[e]; }
    for (int v = 0; v <= nV; v++) o.invAreaTriangle[v] = one / in.areaTriangle[v];

    // ---- atm_adv_coef_compression, CORE:1285-1430 (+ the coupling of the 3rd-order part, CORE:1449)
    o.nAdvCellsForEdge.assign(nE + 1, 0);
    o.advCellsForEdge.assign((size_t)(nE + 1) * 15, nC + 1);
    o.adv_coefs.assign((size_t)(nE + 1) * 15, (real)0);
    o.adv_coefs_3rd.assign((size_t)(nE + 1) * 15, (real)0);
    for (int e = 1; e <= nE; e++) {
        const int cell1 = coe(1, e), cell2 = coe(2, e);
        if (!(cell1 <= nC || cell2 <= nC)) continue;              // only if this edge flux is needed to update owned cells
        int cell_list[20], n = 2;
        cell_list[0] = cell1; cell_list[1] = cell2;
        for (int i = 1; i <= nec(cell1); i++) if (coc(i, cell1) != cell2) cell_list[n++] = coc(i, cell1);
        for (int ic = 1; ic <= nec(cell2); ic++) {
            bool add = true;
            for (int i = 0; i < n; i++) if (cell_list[i] == coc(ic, cell2)) add = false;
            if (add) cell_list[n++] = coc(ic, cell2);
        }
        o.nAdvCellsForEdge[e - 1] = n;
        real a[20], b[20];
        for (int j = 0; j < 20; j++) { a[j] = 0; b[j] = 0; }
        auto pos = [&](int target) { int j_in = -1; for (int j = 0; j < n; j++) if (cell_list[j] == target) j_in = j; return j_in; };   // the LAST match, as the reference's loop
        auto d2 = [&](int i, int side) { return in.deriv_two[((size_t)(e - 1) * 2 + (side - 1)) * 15 + (i - 1)]; };
        int j = pos(cell1);
        a[j] = a[j] + d2(1, 1); b[j] = b[j] + d2(1, 1);
        for (int ic = 1; ic <= nec(cell1); ic++) { j = pos(coc(ic, cell1)); a[j] = a[j] + d2(ic + 1, 1); b[j] = b[j] + d2(ic + 1, 1); }
        j = pos(cell2);
        a[j] = a[j] + d2(1, 2); b[j] = b[j] - d2(1, 2);
        for (int ic = 1; ic <= nec(cell2); ic++) { j = pos(coc(ic, cell2)); a[j] = a[j] + d2(ic + 1, 2); b[j] = b[j] - d2(ic + 1, 2); }
        const real dc = in.dcEdge[e - 1], dv = in.dvEdge[e - 1];
        for (j = 0; j < n; j++) { a[j] = -((dc * dc) * a[j] / (real)12.); b[j] = -((dc * dc) * b[j] / (real)12.); }
        j = pos(cell1); a[j] = a[j] + (real)0.5;
        j = pos(cell2); a[j] = a[j] + (real)0.5;
        for (j = 0; j < n && j < 15; j++) {
            o.advCellsForEdge[(size_t)(e - 1) * 15 + j] = cell_list[j];
            o.adv_coefs[(size_t)(e - 1) * 15 + j] = dv * a[j];
            o.adv_coefs_3rd[(size_t)(e - 1) * 15 + j] = (real)cf.config_coef_3rd_order * (dv * b[j]);
        }
    }

    // ---- atm_compute_mesh_scaling, CORE:1091-1148
    o.meshScalingDel2.assign(nE + 1, one); o.meshScalingDel4.assign(nE + 1, one);
    o.meshScalingRegionalCell.assign(nC + 1, one); o.meshScalingRegionalEdge.assign(nE + 1, one);
    if (h_ScaleWithMesh) {
        for (int e = 1; e <= nE; e++) {
            const real m = (in.meshDensity[coe(1, e) - 1] + in.meshDensity[coe(2, e) - 1]) / (real)2.0;
            o.meshScalingDel2[e - 1] = one / std::pow(m, (real)0.25);
            o.meshScalingDel4[e - 1] = one / std::pow(m, (real)0.75);
            o.meshScalingRegionalEdge[e - 1] = one / std::pow(m, (real)0.25);
        }
        for (int c = 1; c <= nC; c++) o.meshScalingRegionalCell[c - 1] = one / std::pow(in.meshDensity[c - 1], (real)0.25);
    }

    // ---- atm_compute_damping_coefs, CORE:1241-1282
    o.dss.assign((size_t)(nC + 1) * nz, (real)0);
    const real pii = std::acos((real)-1.0), zd = (real)config_zd, xnutr = (real)config_xnutr;
    for (int c = 1; c <= nC; c++) {
        const real* zg = in.zgrid + (size_t)(c - 1) * nz1;
        const real zt = zg[nz];
        for (int k = 0; k < nz; k++) {
            const real z = (real)0.5 * (zg[k] + zg[k + 1]);
            if (z > zd) {
                const real s = std::sin((real)0.5 * pii * (z - zd) / (zt - zd));
                o.dss[(size_t)(c - 1) * nz + k] = xnutr * (s * s) / std::pow(in.meshDensity[c - 1], (real)0.25);
            }
        }
    }
}

}  // namespace initblk
