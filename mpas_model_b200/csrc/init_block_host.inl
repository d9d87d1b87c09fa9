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
    // optional (all six or none): with them, coeffs_reconstruct is derived as well (mpas_init_reconstruct)
    const real *xCell = nullptr, *yCell = nullptr, *zCell = nullptr, *xEdge = nullptr, *yEdge = nullptr, *zEdge = nullptr;
    bool coords() const { return xCell && yCell && zCell && xEdge && yEdge && zEdge; }
};
struct Out {
    std::vector<real> edgesOnVertex_sign, edgesOnCell_sign, zb_cell, zb3_cell, invAreaCell, invDvEdge, invDcEdge, invAreaTriangle,
                      adv_coefs, adv_coefs_3rd, meshScalingDel2, meshScalingDel4, meshScalingRegionalCell, meshScalingRegionalEdge, dss,
                      coeffs_reconstruct;
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
    static const char* const opt[] = {"xCell", "yCell", "zCell", "xEdge", "yEdge", "zEdge"};
    const real** op[] = {&in.xCell, &in.yCell, &in.zCell, &in.xEdge, &in.yEdge, &in.zEdge};
    for (int k = 0; k < n; k++)
        if (names[k] && arrays[k]) for (int q = 0; q < 6; q++) if (!strcmp(names[k], opt[q])) *op[q] = (const real*)arrays[k];
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
        // (cell_list(20) as in the reference; 2 + 7 + 8 entries at most with maxEdges = 8 -- a longer list is cut, not overrun)
        for (int i = 1; i <= nec(cell1); i++) if (coc(i, cell1) != cell2 && n < 20) cell_list[n++] = coc(i, cell1);
        for (int ic = 1; ic <= nec(cell2); ic++) {
            bool add = true;
            for (int i = 0; i < n; i++) if (cell_list[i] == coc(ic, cell2)) add = false;
            if (add && n < 20) cell_list[n++] = coc(ic, cell2);
        }
        o.nAdvCellsForEdge[e - 1] = n;
        real a[20], b[20];
        for (int j = 0; j < 20; j++) { a[j] = 0; b[j] = 0; }
        // the LAST match, as the reference's loop (a cell cut from an over-long list lands on slot 0 and is never stored wrong: n <= 20)
        auto pos = [&](int target) { int j_in = 0; for (int j = 0; j < n; j++) if (cell_list[j] == target) j_in = j; return j_in; };
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

// ---- mpas_initialize_vectors (src/operators/mpas_vector_operations.F:652-771, spherical non-periodic branch) and
// mpas_init_reconstruct (src/operators/mpas_vector_reconstruction.F:60-177) with the radial-basis-function fit of
// mpas_rbf_interp_func_3D_plane_vec_const_dir_comp_coeffs (src/operators/mpas_rbf_interpolation.F:1079-1145, matrix and right-hand
// sides :1527-1559, inverse multiquadric :1369-1376) solved by elgs / mpas_legs (:1782-1846, 1670-1698: Gaussian elimination
// with scaled partial pivoting): coeffs_reconstruct(3, maxEdges, nCells+1) of the owned cells.  Operation order as in
// mpas_model_b200/reconstruct.py, which the GPU tests of mpas_reconstruct are built on.
struct V3 { real x, y, z; };
static inline real dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline V3 unit3(V3 v) { const real n = std::sqrt((v.x * v.x + v.y * v.y) + v.z * v.z); return V3{v.x / n, v.y / n, v.z / n}; }
static void reconstruct_coeffs(const mpasb_dims& dm, const In& in, Out& o) {
    const int nC = dm.nCells, nE = dm.nEdges, nS = dm.nCellsSolve, mx = dm.maxEdges;
    o.coeffs_reconstruct.assign((size_t)(nC + 1) * mx * 3, (real)0);
    auto xc = [&](int c) { return V3{in.xCell[c], in.yCell[c], in.zCell[c]}; };          // 0-based; the garbage cell is index nC
    auto xe = [&](int e) { return V3{in.xEdge[e], in.yEdge[e], in.zEdge[e]}; };
    auto edge_normal = [&](int e) {                                                          // e 0-based, < nE
        const int c1 = in.cellsOnEdge[(size_t)e * 2] - 1, c2 = in.cellsOnEdge[(size_t)e * 2 + 1] - 1;
        V3 v;
        if (c1 == nC && c2 == nC) v = V3{(real)1, (real)0, (real)0};
        else if (c1 == nC) { const V3 a = xc(c2), b = xe(e); v = V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
        else if (c2 == nC) { const V3 a = xe(e), b = xc(c1); v = V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
        else { const V3 a = xc(c2), b = xc(c1); v = V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
        return unit3(v);
    };
    const int NMAX = 16 + 2;
    if (mx + 2 > NMAX) return;
    for (int c = 0; c < nS; c++) {
        const int n = in.nEdgesOnCell[c], N = n + 2;
        if (n < 1 || n > mx) continue;
        V3 src[16], uv[16];
        for (int i = 0; i < n; i++) { const int e = in.edgesOnCell[(size_t)c * mx + i] - 1; src[i] = xe(e); uv[i] = e < nE ? edge_normal(e) : V3{0, 0, 0}; }
        const V3 dest = xc(c), rhat = unit3(dest);
        const real ndr = dot3(uv[0], rhat);
        const V3 xhat = unit3(V3{uv[0].x - ndr * rhat.x, uv[0].y - ndr * rhat.y, uv[0].z - ndr * rhat.z});
        const V3 yhat = unit3(V3{rhat.y * xhat.z - rhat.z * xhat.y, rhat.z * xhat.x - rhat.x * xhat.z, rhat.x * xhat.y - rhat.y * xhat.x});
        real alpha = 0;
        for (int i = 0; i < n; i++) {
            const V3 d{dest.x - src[i].x, dest.y - src[i].y, dest.z - src[i].z};
            alpha = alpha + std::sqrt((d.x * d.x + d.y * d.y) + d.z * d.z);
        }
        alpha = alpha / (real)n;
        const real a2 = alpha * alpha;
        real ps[16][2], pu[16][2];
        for (int i = 0; i < n; i++) { ps[i][0] = dot3(src[i], xhat); ps[i][1] = dot3(src[i], yhat); pu[i][0] = dot3(uv[i], xhat); pu[i][1] = dot3(uv[i], yhat); }
        const real pd[2] = {dot3(dest, xhat), dot3(dest, yhat)};
        real M[NMAX][NMAX], rhs[NMAX][2];
        for (int i = 0; i < N; i++) { for (int j = 0; j < N; j++) M[i][j] = 0; rhs[i][0] = rhs[i][1] = 0; }
        for (int i = 0; i < n; i++) {
            for (int j = 0; j < n; j++) {
                const real dx = ps[i][0] - ps[j][0], dy = ps[i][1] - ps[j][1];
                const real rsq = (dx * dx + dy * dy) / a2;
                M[i][j] = ((real)1.0 / std::sqrt((real)1.0 + rsq)) * (pu[i][0] * pu[j][0] + pu[i][1] * pu[j][1]);
            }
            const real dx = pd[0] - ps[i][0], dy = pd[1] - ps[i][1];
            const real rb = (real)1.0 / std::sqrt((real)1.0 + (dx * dx + dy * dy) / a2);
            rhs[i][0] = rb * pu[i][0]; rhs[i][1] = rb * pu[i][1];
            M[i][n] = pu[i][0]; M[i][n + 1] = pu[i][1]; M[n][i] = pu[i][0]; M[n + 1][i] = pu[i][1];
        }
        rhs[n][0] = 1; rhs[n + 1][1] = 1;
        // elgs: scaled partial pivoting, factors stored in place
        int indx[NMAX]; real cs[NMAX];
        for (int i = 0; i < N; i++) { indx[i] = i; real m = 0; for (int j = 0; j < N; j++) m = std::max(m, std::fabs(M[i][j])); cs[i] = m; }
        for (int j = 0; j < N - 1; j++) {
            real pi1 = 0; int k = j;
            for (int i = j; i < N; i++) { const real pi = std::fabs(M[indx[i]][j]) / cs[indx[i]]; if (pi > pi1) { pi1 = pi; k = i; } }
            std::swap(indx[j], indx[k]);
            const int rj = indx[j];
            for (int i = j + 1; i < N; i++) {
                const int ri = indx[i];
                const real pj = M[ri][j] / M[rj][j];
                M[ri][j] = pj;
                for (int q = j + 1; q < N; q++) M[ri][q] = M[ri][q] - pj * M[rj][q];
            }
        }
        real co[2][NMAX];
        for (int q = 0; q < 2; q++) {                           // mpas_legs: forward elimination of the right-hand side, back substitution
            real b[NMAX];
            for (int i = 0; i < N; i++) b[i] = rhs[i][q];
            for (int i = 0; i < N - 1; i++) for (int j = i + 1; j < N; j++) b[indx[j]] = b[indx[j]] - M[indx[j]][i] * b[indx[i]];
            real* x = co[q];
            x[N - 1] = b[indx[N - 1]] / M[indx[N - 1]][N - 1];
            for (int i = N - 2; i >= 0; i--) {
                real xi = b[indx[i]];
                for (int j = i + 1; j < N; j++) xi = xi - M[indx[i]][j] * x[j];
                x[i] = xi / M[indx[i]][i];
            }
        }
        for (int i = 0; i < n; i++) {
            real* dst = &o.coeffs_reconstruct[((size_t)c * mx + i) * 3];
            dst[0] = xhat.x * co[0][i] + yhat.x * co[1][i];
            dst[1] = xhat.y * co[0][i] + yhat.y * co[1][i];
            dst[2] = xhat.z * co[0][i] + yhat.z * co[1][i];
        }
    }
}

}  // namespace initblk
