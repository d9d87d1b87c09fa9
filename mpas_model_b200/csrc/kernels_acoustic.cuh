// kernels_acoustic.cuh -- acoustic substep (edge update + per-column implicit solve),
// 3-D divergence damping and recovery of the large-step variables.
#pragma once
#include "kernels_dyn.cuh"

// ------------------------------------------------------------------ atm_advance_acoustic_step_work, edge part  TI:2751-2822
__global__ void k_acoustic_edge(const Dev D, real dts, int small_step, real c2) {
    KI;
    if (i >= D.nEdges || k >= nl) return;
    const int cell1 = D.cellsOnEdge[2 * i], cell2 = D.cellsOnEdge[2 * i + 1];
    if (!(cell1 < D.nCellsSolve || cell2 < D.nCellsSolve)) return;
    if (small_step != 1) {
        real pgrad = ((AT(D.rtheta_pp, cell2, k) - AT(D.rtheta_pp, cell1, k)) * D.invDcEdge[i]) / (.5 * (AT(D.zz, cell2, k) + AT(D.zz, cell1, k)));
        pgrad = AT(D.cqu, i, k) * 0.5 * c2 * (AT(D.exner, cell1, k) + AT(D.exner, cell2, k)) * pgrad;
        pgrad = pgrad + 0.5 * AT(D.zxu, i, k) * GRAVITY * (AT(D.rho_pp, cell1, k) + AT(D.rho_pp, cell2, k));
        const real rup = AT(D.ru_p, i, k) + dts * (AT(D.tend_u, i, k) - (1.0 - D.specZoneMaskEdge[i]) * pgrad);
        AT(D.ru_p, i, k) = rup;
        AT(D.ruAvg, i, k) = AT(D.ruAvg, i, k) + rup;
    } else {
        const real rup = dts * AT(D.tend_u, i, k);
        AT(D.ru_p, i, k) = rup;
        AT(D.ruAvg, i, k) = rup;
    }
}

// ------------------------------------------------------------------ atm_advance_acoustic_step_work, cell part  TI:2824-2973
// Block = (LDK, CPB).  All threads of a column assemble the right-hand sides in parallel
// (coalesced along k); one thread per column then runs the two tridiagonal sweeps out of
// shared memory in the reference's order, so the solve is bit-identical to the CPU's.
// Shared memory per column: ts, rs, rw (3 * LDK reals) + a_tri, alpha_tri, gamma_tri (3 * LDK).
__global__ void k_acoustic_cell(const Dev D, real dts, int small_step, real epssm, real resm) {
    KI;
    extern __shared__ real sm[];
    real* s_ts = sm + (size_t)threadIdx.y * 6 * LDK;
    real* s_rs = s_ts + LDK;
    real* s_rw = s_rs + LDK;
    real* s_a = s_rw + LDK;
    real* s_al = s_a + LDK;
    real* s_ga = s_al + LDK;
    const bool incell = i < D.nCells;
    // cells of the specified zone of a regional run take the other branch of TI:2862 (k_lbc_acoustic_spec)
    const bool solve = i < D.nCellsSolve && D.specZoneMaskCell[i] == 0.0;
    const bool first = small_step == 1;
    // old values of the perturbation variables (zero on the first small step, TI:2850-2860)
    real rho_pp_k = 0.0, rtheta_pp_k = 0.0, rw_p_k = 0.0, rw_p_k1 = 0.0, wwAvg_k = 0.0;
    if (incell && k <= nl) {
        if (!first) {
            if (k < nl) { rtheta_pp_k = AT(D.rtheta_pp, i, k); AT(D.rtheta_pp_old, i, k) = rtheta_pp_k; }
        } else if (k < nl) {
            AT(D.rtheta_pp_old, i, k) = 0.0;
        }
    }
    real ts = 0.0, rs = 0.0;
    if (solve && k <= nl) {
        if (!first) {
            rw_p_k = AT(D.rw_p, i, k);
            wwAvg_k = AT(D.wwAvg, i, k);
            if (k < nl) { rho_pp_k = AT(D.rho_pp, i, k); rw_p_k1 = AT(D.rw_p, i, k + 1); }
        }
        if (k < nl) {
            const int ne = D.nEdgesOnCell[i];
            const real invArea = D.invAreaCell[i];
            for (int e = 0; e < ne; e++) {
                const int iEdge = D.edgesOnCell[(size_t)i * D.maxEdges + e];
                const int cell1 = D.cellsOnEdge[2 * iEdge], cell2 = D.cellsOnEdge[2 * iEdge + 1];
                const real flux = D.edgesOnCell_sign[(size_t)i * D.maxEdges + e] * dts * D.dvEdge[iEdge] * AT(D.ru_p, iEdge, k) * invArea;
                rs = rs - flux;
                ts = ts - flux * 0.5 * (AT(D.theta_m, cell2, k) + AT(D.theta_m, cell1, k));
            }
            rs = rho_pp_k + dts * AT(D.tend_rho, i, k) + rs
                 - D.cofrz[k] * resm * (rw_p_k1 - rw_p_k);
            ts = rtheta_pp_k + dts * AT(D.tend_theta, i, k) + ts
                 - resm * D.rdzw[k] * (AT(D.coftz, i, k + 1) * rw_p_k1
                                       - AT(D.coftz, i, k) * rw_p_k);
            s_a[k] = AT(D.a_tri, i, k); s_al[k] = AT(D.alpha_tri, i, k); s_ga[k] = AT(D.gamma_tri, i, k);
        }
        if (k >= 1 && k < nl) wwAvg_k = wwAvg_k + 0.5 * (1.0 - epssm) * rw_p_k;
    }
    s_ts[k] = ts; s_rs[k] = rs;
    // rtheta_pp, rho_pp of level k-1 (old values) are needed by the rw_p right-hand side
    __syncthreads();
    if (solve && k <= nl) {
        real r = rw_p_k;
        if (k >= 1 && k < nl) {
            real rtheta_pp_m = 0.0, rho_pp_m = 0.0;
            if (!first) { rtheta_pp_m = AT(D.rtheta_pp_old, i, k - 1); rho_pp_m = AT(D.rho_pp, i, k - 1); }
            const real zzk = AT(D.zz, i, k), zzm = AT(D.zz, i, k - 1);
            const real cofwt_k = AT(D.cofwt, i, k), cofwt_m = AT(D.cofwt, i, k - 1);
            r = rw_p_k + dts * AT(D.tend_w, i, k)
                - AT(D.cofwz, i, k) * ((zzk * s_ts[k]
                                        - zzm * s_ts[k - 1])
                                       + resm * (zzk * rtheta_pp_k
                                                 - zzm * rtheta_pp_m))
                - AT(D.cofwr, i, k) * ((s_rs[k] + s_rs[k - 1])
                                       + resm * (rho_pp_k + rho_pp_m))
                + cofwt_k * (s_ts[k] + resm * rtheta_pp_k)
                + cofwt_m * (s_ts[k - 1] + resm * rtheta_pp_m);
        }
        s_rw[k] = r;
    }
    __syncthreads();
    if (solve && k == 0) {
        // tridiagonal solve sweeping up and then down the column, TI:2922-2930
        real prev = s_rw[0];
        for (int kk = 1; kk < nl; kk++) {
            prev = (s_rw[kk] - s_a[kk] * prev) * s_al[kk];
            s_rw[kk] = prev;
        }
        real next = s_rw[nl];
        for (int kk = nl - 1; kk >= 0; kk--) {
            next = s_rw[kk] - s_ga[kk] * next;
            s_rw[kk] = next;
        }
    }
    __syncthreads();
    if (solve && k <= nl) {
        real r = s_rw[k];
        if (k >= 1 && k < nl) {
            // implicit Rayleigh damping on w, TI:2936-2942
            const real dssk = AT(D.dss, i, k);
            const real dw = AT(D.rw_save, i, k) - AT(D.rw, i, k);
            r = (r + dw - dts * dssk *
                 (D.fzm[k] * AT(D.zz, i, k) + D.fzp[k] * AT(D.zz, i, k - 1))
                 * (D.fzm[k] * AT(D.rho_zz_2, i, k) + D.fzp[k] * AT(D.rho_zz_2, i, k - 1))
                 * AT(D.w_2, i, k)) / (1.0 + dts * dssk)
                - dw;
            wwAvg_k = wwAvg_k + 0.5 * (1.0 + epssm) * r;
        }
        AT(D.rw_p, i, k) = r;
        AT(D.wwAvg, i, k) = wwAvg_k;
        s_rw[k] = r;                                      // own slot only: no hazard before the barrier
    }
    __syncthreads();
    if (solve && k < nl) {
        AT(D.rho_pp, i, k) = s_rs[k] - D.cofrz[k] * (s_rw[k + 1] - s_rw[k]);
        AT(D.rtheta_pp, i, k) = s_ts[k] - D.rdzw[k] * (AT(D.coftz, i, k + 1) * s_rw[k + 1]
                                                       - AT(D.coftz, i, k) * s_rw[k]);
    }
}

// ------------------------------------------------------------------ atm_divergence_damping_3d  TI:2987-3075
__global__ void k_divergence_damping(const Dev D, real coef_divdamp) {
    KI;
    if (i >= D.nEdges || k >= nl) return;
    const int cell1 = D.cellsOnEdge[2 * i], cell2 = D.cellsOnEdge[2 * i + 1];
    if (!(cell1 < D.nCellsSolve || cell2 < D.nCellsSolve)) return;
    const real divCell1 = -(AT(D.rtheta_pp, cell1, k) - AT(D.rtheta_pp_old, cell1, k));
    const real divCell2 = -(AT(D.rtheta_pp, cell2, k) - AT(D.rtheta_pp_old, cell2, k));
    AT(D.ru_p, i, k) = AT(D.ru_p, i, k) + coef_divdamp * (divCell2 - divCell1) * (1.0 - D.specZoneMaskEdge[i])
                                          / (AT(D.theta_m, cell1, k) + AT(D.theta_m, cell2, k));
}

// ------------------------------------------------------------------ atm_recover_large_step_variables_work  TI:3189-3431
// (1) cell-all, TI:3294-3350
__global__ void k_recover_cell1(const Dev D, real dt, real invNs, int rk_step, real rcv, real rgas_p0) {
    KI;
    if (i > D.nCells || k > nl) return;
    if (i == D.nCells) { if (k < nl) AT(D.rho_zz_2, i, k) = 1.0; return; }      // garbage cell, TI:3282-3284
    if (k < nl) {
        const real rho_p = AT(D.rho_p_save, i, k) + AT(D.rho_pp, i, k);
        const real rho_zz = rho_p + AT(D.rho_base, i, k);
        AT(D.rho_p, i, k) = rho_p;
        AT(D.rho_zz_2, i, k) = rho_zz;
        const real rtb = AT(D.rtheta_base, i, k);
        if (rk_step == 3) {
            const real rtheta_p = AT(D.rtheta_p_save, i, k) + AT(D.rtheta_pp, i, k)
                                  - dt * rho_zz * AT(D.rt_diabatic_tend, i, k);
            AT(D.rtheta_p, i, k) = rtheta_p;
            AT(D.theta_m_2, i, k) = (rtheta_p + rtb) / rho_zz;
            const real zzk = AT(D.zz, i, k);
            const real ex = pow_cr(zzk * (rgas_p0) * (rtheta_p + rtb), rcv);
            AT(D.exner, i, k) = ex;
            AT(D.pressure_p, i, k) = zzk * RGAS * (ex * rtheta_p + rtb
                                                    * (ex - AT(D.exner_base, i, k)));
        } else {
            const real rtheta_p = AT(D.rtheta_p_save, i, k) + AT(D.rtheta_pp, i, k);
            AT(D.rtheta_p, i, k) = rtheta_p;
            AT(D.theta_m_2, i, k) = (rtheta_p + rtb) / rho_zz;
        }
    }
    if (k == 0 || k == nl) {
        AT(D.rw, i, k) = 0.0;
        AT(D.w_2, i, k) = 0.0;
    } else {
        AT(D.wwAvg, i, k) = AT(D.rw_save, i, k) + (AT(D.wwAvg, i, k) * invNs);
        const real rw = AT(D.rw_save, i, k) + AT(D.rw_p, i, k);
        AT(D.rw, i, k) = rw;
        AT(D.w_2, i, k) = rw / (D.fzm[k] * AT(D.zz, i, k) + D.fzp[k] * AT(D.zz, i, k - 1));
    }
}
// (2) edge-all, TI:3360-3372
__global__ void k_recover_edge(const Dev D, real invNs) {
    KI;
    if (i >= D.nEdges || k >= nl) return;
    const int cell1 = D.cellsOnEdge[2 * i], cell2 = D.cellsOnEdge[2 * i + 1];
    const real rus = AT(D.ru_save, i, k);
    AT(D.ruAvg, i, k) = rus + (AT(D.ruAvg, i, k) * invNs);
    const real ru = rus + AT(D.ru_p, i, k);
    AT(D.ru, i, k) = ru;
    AT(D.u_2, i, k) = 2. * ru / (AT(D.rho_zz_2, cell1, k) + AT(D.rho_zz_2, cell2, k));
}
// (3) cell-all, TI:3379-3416
__global__ void k_recover_cell2(const Dev D, real cf1, real cf2, real cf3) {
    KI;
    if (i >= D.nCells || k >= nl) return;
    if (D.bdyMaskCell[i] > 5) return;              // no update in the specified zone of a regional run, TI:3385
    const int ne = D.nEdgesOnCell[i];
    real w = AT(D.w_2, i, k);
    const real fm = D.fzm[k], fp = D.fzp[k];
    if (D.zb_any[i]) for (int e = 0; e < ne; e++) {
        const int iEdge = D.edgesOnCell[(size_t)i * D.maxEdges + e];
        real flux;
        if (k == 0) flux = (cf1 * AT(D.ru, iEdge, 0) + cf2 * AT(D.ru, iEdge, 1) + cf3 * AT(D.ru, iEdge, 2));
        else flux = (fm * AT(D.ru, iEdge, k) + fp * AT(D.ru, iEdge, k - 1));
        const size_t zi = ((size_t)i * D.maxEdges + e) * LDK + k;
        w = w + D.edgesOnCell_sign[(size_t)i * D.maxEdges + e] *
                (D.zb_cell[zi] + sign1(flux) * D.zb3_cell[zi]) * flux;
    }
    if (k == 0) w = w / (cf1 * AT(D.rho_zz_2, i, 0) + cf2 * AT(D.rho_zz_2, i, 1) + cf3 * AT(D.rho_zz_2, i, 2));
    else w = w / (fm * AT(D.rho_zz_2, i, k) + fp * AT(D.rho_zz_2, i, k - 1));
    AT(D.w_2, i, k) = w;
}
