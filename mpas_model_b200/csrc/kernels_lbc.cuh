// kernels_lbc.cuh -- lateral boundary conditions of a regional (limited-area) run, config_apply_lbcs
// (mpas_atm_time_integration.F "TI":7198-7910 and the inline loops of atm_srk3 at TI:1343-1388).
//
// bdyMaskCell / bdyMaskEdge: 0 interior, 1..nRelaxZone relaxation zone (1 innermost), nRelaxZone+1..nRelaxZone+nSpecZone
// specified zone (mpas_atm_boundaries.F:35-38).  The driving fields live in lbc_<field>: time level 1 = tendency over the
// current LBC interval, time level 2 = state at its end; a driving VALUE at "now + delta_t" is
//     state - (seconds to the end of the interval - delta_t) * tendency          (mpas_atm_boundaries.F:497-549)
// and is formed on the fly (dtl = that bracket, computed by the host in RKIND exactly as the reference does).
// One thread per (level, column) like the other generic kernels: the zones are thin rings, interior columns leave at once.
// Operation order is the reference's, so results are bit-identical to it.
#pragma once
#include "kernels_dyn.cuh"

#define LBC_NRELAX 5
#define LBC_NSPEC 2

// atm_bdy_adjust_dynamics_speczone_tend, TI:7271-7342
__global__ void k_lbc_speczone_cell(const Dev D) {
    KI;
    if (i >= D.nCellsSolve || k >= nl) return;
    if (D.bdyMaskCell[i] > LBC_NRELAX) {
        AT(D.tend_rho, i, k) = AT(D.lbc_rho_zz, i, k);
        AT(D.tend_theta, i, k) = AT(D.lbc_rtheta_m, i, k);
        AT(D.tend_w, i, k) = 0.;
        AT(D.rt_diabatic_tend, i, k) = 0.;
    }
}
__global__ void k_lbc_speczone_edge(const Dev D) {
    KI;
    if (i >= D.nEdgesSolve || k >= nl) return;
    if (D.bdyMaskEdge[i] > LBC_NRELAX) AT(D.tend_u, i, k) = AT(D.lbc_ru, i, k);
}

// atm_bdy_adjust_dynamics_relaxzone_tend, TI:7346-7580: Rayleigh damping towards the driving values and a Laplacian filter
// of the departure from them, cells (TI:7449-7459, 7478-7500)
__global__ void k_lbc_relax_cell(const Dev D, real dt, real dtl) {
    KI;
    if (i >= D.nCellsSolve || k >= nl) return;
    const int mask = D.bdyMaskCell[i];
    if (!(mask > 1 && mask <= LBC_NRELAX)) return;
#define RT_DEP(c) (AT(D.rho_zz_2, (c), k) * AT(D.theta_m_2, (c), k) - (AT(D.lbc_rtheta_m_2, (c), k) - dtl * AT(D.lbc_rtheta_m, (c), k)))
#define RHO_DEP(c) (AT(D.rho_zz_2, (c), k) - (AT(D.lbc_rho_zz_2, (c), k) - dtl * AT(D.lbc_rho_zz, (c), k)))
    const real rayleigh_damping_coef = ((real)mask - 1.) / (real)LBC_NRELAX / (50. * dt * D.meshScalingRegionalCell[i]);
    real tend_rho = AT(D.tend_rho, i, k) - rayleigh_damping_coef * RHO_DEP(i);
    real tend_rt = AT(D.tend_theta, i, k) - rayleigh_damping_coef * RT_DEP(i);
    const real laplacian_filter_coef = ((real)mask - 1.) / (real)LBC_NRELAX / (10. * dt * D.meshScalingRegionalCell[i]);
    const int ne = D.nEdgesOnCell[i];
    for (int e = 0; e < ne; e++) {
        const size_t slot = (size_t)i * D.maxEdges + e;
        const int iEdge = D.edgesOnCell[slot];
        const real edge_sign = D.edgesOnCell_sign[slot] * D.dvEdge[iEdge] * D.invDcEdge[iEdge] * laplacian_filter_coef;
        const int cell1 = D.cellsOnEdge[2 * iEdge], cell2 = D.cellsOnEdge[2 * iEdge + 1];
        tend_rt = tend_rt + edge_sign * (RT_DEP(cell2) - RT_DEP(cell1));
        tend_rho = tend_rho + edge_sign * (RHO_DEP(cell2) - RHO_DEP(cell1));
    }
#undef RT_DEP
#undef RHO_DEP
    AT(D.tend_rho, i, k) = tend_rho;
    AT(D.tend_theta, i, k) = tend_rt;
}
// edges (TI:7461-7471, 7502-7576)
__global__ void k_lbc_relax_edge(const Dev D, real dt, real dtl, real divdamp_coef) {
    KI;
    if (i >= D.nEdges || k >= nl) return;
    const int mask = D.bdyMaskEdge[i];
    if (!(mask > 1 && mask <= LBC_NRELAX)) return;
#define RU_DEP(e) (AT(D.ru, (e), k) - (AT(D.lbc_ru_2, (e), k) - dtl * AT(D.lbc_ru, (e), k)))
    const real rayleigh_damping_coef = ((real)mask - 1.) / (real)LBC_NRELAX / (50. * dt * D.meshScalingRegionalEdge[i]);
    real tend_ru = AT(D.tend_u, i, k) - rayleigh_damping_coef * RU_DEP(i);
    const real dc = D.dcEdge[i];
    const real laplacian_filter_coef = dc * dc * ((real)mask - 1.) / (real)LBC_NRELAX / (10. * dt * D.meshScalingRegionalEdge[i]);
    const int cell1 = D.cellsOnEdge[2 * i], cell2 = D.cellsOnEdge[2 * i + 1];
    const int vertex1 = D.verticesOnEdge[2 * i], vertex2 = D.verticesOnEdge[2 * i + 1];
    const real r_dc = D.invDcEdge[i];
    const real r_dv = rmin(D.invDvEdge[i], 4 * D.invDcEdge[i]);
    real divergence1 = 0., divergence2 = 0., vorticity1 = 0., vorticity2 = 0.;
    {
        const real invArea = D.invAreaCell[cell1];
        const int ne = D.nEdgesOnCell[cell1];
        for (int e = 0; e < ne; e++) {
            const size_t slot = (size_t)cell1 * D.maxEdges + e;
            const int iEdge_div = D.edgesOnCell[slot];
            const real edge_sign = invArea * D.dvEdge[iEdge_div] * D.edgesOnCell_sign[slot];
            divergence1 = divergence1 + edge_sign * RU_DEP(iEdge_div);
        }
    }
    {
        const real invArea = D.invAreaCell[cell2];
        const int ne = D.nEdgesOnCell[cell2];
        for (int e = 0; e < ne; e++) {
            const size_t slot = (size_t)cell2 * D.maxEdges + e;
            const int iEdge_div = D.edgesOnCell[slot];
            const real edge_sign = invArea * D.dvEdge[iEdge_div] * D.edgesOnCell_sign[slot];
            divergence2 = divergence2 + edge_sign * RU_DEP(iEdge_div);
        }
    }
    for (int e = 0; e < 3; e++) {
        const int iEdge_vort = D.edgesOnVertex[3 * vertex1 + e];
        const real edge_sign = D.invAreaTriangle[vertex1] * D.dcEdge[iEdge_vort] * D.edgesOnVertex_sign[3 * vertex1 + e];
        vorticity1 = vorticity1 + edge_sign * RU_DEP(iEdge_vort);
    }
    for (int e = 0; e < 3; e++) {
        const int iEdge_vort = D.edgesOnVertex[3 * vertex2 + e];
        const real edge_sign = D.invAreaTriangle[vertex2] * D.dcEdge[iEdge_vort] * D.edgesOnVertex_sign[3 * vertex2 + e];
        vorticity2 = vorticity2 + edge_sign * RU_DEP(iEdge_vort);
    }
#undef RU_DEP
    AT(D.tend_u, i, k) = tend_ru + laplacian_filter_coef * (divdamp_coef * (divergence2 - divergence1) * r_dc
                                                          - (vorticity2 - vorticity1) * r_dv);
}

// the two loops of atm_srk3 after recover_large_step_variables, TI:1343-1388: u (owned edges) and ru (all edges) of the
// specified zone are the driving values
__global__ void k_lbc_reset_u_ru(const Dev D, real dtl) {
    KI;
    if (i >= D.nEdges || k >= nl) return;
    if (D.bdyMaskEdge[i] > LBC_NRELAX) {
        if (i < D.nEdgesSolve) AT(D.u_2, i, k) = AT(D.lbc_u_2, i, k) - dtl * AT(D.lbc_u, i, k);
        AT(D.ru, i, k) = AT(D.lbc_ru_2, i, k) - dtl * AT(D.lbc_ru, i, k);
    }
}

// atm_zero_gradient_w_bdy_work, TI:7227-7267
__global__ void k_lbc_zero_w(const Dev D) {
    KI;
    if (i >= D.nCellsSolve || k < 1 || k >= nl) return;
    if (D.bdyMaskCell[i] > LBC_NRELAX) AT(D.w_2, i, k) = 0.0;
}

// atm_bdy_reset_speczone_values, TI:7583-7635
__global__ void k_lbc_reset_speczone(const Dev D, real dtl) {
    KI;
    if (i >= D.nCellsSolve || k >= nl) return;
    if (D.bdyMaskCell[i] > LBC_NRELAX) {
        const real rt = AT(D.lbc_rtheta_m_2, i, k) - dtl * AT(D.lbc_rtheta_m, i, k);
        const real rho = AT(D.lbc_rho_zz_2, i, k) - dtl * AT(D.lbc_rho_zz, i, k);
        AT(D.theta_m_2, i, k) = rt / rho;
        AT(D.rtheta_p, i, k) = rt - AT(D.rtheta_base, i, k);
    }
}

// atm_bdy_adjust_scalars_work, TI:7696-7813, scalar s: (a) new values of the relaxation and specified zones into `tmp`
// (every neighbour is read before any cell is updated), (b) copy back
__global__ void k_lbc_adjust_scalars_a(const Dev D, int s, real dt, real dt_rk, real dtl, real* __restrict__ tmp) {
    KI;
    if (i >= D.nCellsSolve || k >= nl) return;
    const int mask = D.bdyMaskCell[i];
    if (mask <= 1) return;
    RP q = D.scalars_2 + (size_t)s * D.cellPlane;
    RP st = D.lbc_scalars_2 + (size_t)s * D.cellPlane;
    RP td = D.lbc_scalars + (size_t)s * D.cellPlane;
#define Q_DEP(c) (AT(q, (c), k) - (AT(st, (c), k) - dtl * AT(td, (c), k)))
    real v;
    if (mask <= LBC_NRELAX) {
        const real laplacian_filter_coef = dt_rk * ((real)mask - 1.) / (real)LBC_NRELAX / (10. * dt * D.meshScalingRegionalCell[i]);
        const real rayleigh_damping_coef = laplacian_filter_coef / 5.0;
        v = AT(q, i, k);
        const int ne = D.nEdgesOnCell[i];
        for (int e = 0; e < ne; e++) {
            const size_t slot = (size_t)i * D.maxEdges + e;
            const int iEdge = D.edgesOnCell[slot];
            const real edge_sign = D.edgesOnCell_sign[slot] * D.dvEdge[iEdge] * D.invDcEdge[iEdge] * laplacian_filter_coef;
            const int cell1 = D.cellsOnEdge[2 * iEdge], cell2 = D.cellsOnEdge[2 * iEdge + 1];
            const real filter_flux = edge_sign * (Q_DEP(cell2) - Q_DEP(cell1));
            v = v + filter_flux;
        }
        v = v - rayleigh_damping_coef * Q_DEP(i);
    } else {
        v = AT(st, i, k) - dtl * AT(td, i, k);
    }
#undef Q_DEP
    AT(tmp, i, k) = v;
}
__global__ void k_lbc_adjust_scalars_b(const Dev D, int s, const real* __restrict__ tmp) {
    KI;
    if (i >= D.nCellsSolve || k >= nl) return;
    if (D.bdyMaskCell[i] > 1) AT(D.scalars_2 + (size_t)s * D.cellPlane, i, k) = AT(tmp, i, k);
}
// atm_bdy_set_scalars_work, TI:7859-7910
__global__ void k_lbc_set_scalars(const Dev D, int s, real dtl) {
    KI;
    if (i >= D.nCellsSolve || k >= nl) return;
    if (D.bdyMaskCell[i] > LBC_NRELAX)
        AT(D.scalars_2 + (size_t)s * D.cellPlane, i, k) = AT(D.lbc_scalars_2 + (size_t)s * D.cellPlane, i, k) - dtl * AT(D.lbc_scalars + (size_t)s * D.cellPlane, i, k);
}

// atm_advance_acoustic_step_work, the specified-zone branch of the cell loop (TI:2848-2860, 2962-2971): no implicit solve,
// the perturbation variables just follow their (driving) tendencies
__global__ void k_lbc_acoustic_spec(const Dev D, real dts, int small_step, real epssm) {
    KI;
    if (i >= D.nCellsSolve || k > nl) return;
    if (D.specZoneMaskCell[i] == 0.0) return;
    const bool first = small_step == 1;
    if (k == nl) { if (first) { AT(D.wwAvg, i, k) = 0.0; AT(D.rw_p, i, k) = 0.0; } return; }
    const real rho_pp = first ? (real)0.0 : AT(D.rho_pp, i, k), rtheta_pp = first ? (real)0.0 : AT(D.rtheta_pp, i, k);
    const real rw_p0 = first ? (real)0.0 : AT(D.rw_p, i, k), wwAvg = first ? (real)0.0 : AT(D.wwAvg, i, k);
    AT(D.rho_pp, i, k) = rho_pp + dts * AT(D.tend_rho, i, k);
    AT(D.rtheta_pp, i, k) = rtheta_pp + dts * AT(D.tend_theta, i, k);
    const real rw_p = rw_p0 + dts * AT(D.tend_w, i, k);
    AT(D.rw_p, i, k) = rw_p;
    AT(D.wwAvg, i, k) = wwAvg + 0.5 * (1.0 + epssm) * rw_p;
}
